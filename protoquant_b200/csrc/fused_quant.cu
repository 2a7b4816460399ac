// Producer-fused quantizers: the op in front of a dynamic-quant linear
// writes the int8 operand and its per-token scale directly, instead of a 16-bit tensor that the
// activation quantizer would read back.
//
//   pq_norm_quant     RMSNorm / LayerNorm  -> (xq, s_x [, y])      before q/k/v and gate/up
//   pq_act_mul_quant  act(gate) [* up]     -> (hq, s_h [, h])      before down_proj / FFN-down
//
// Semantics: y is exactly the tensor the unfused op would have stored (rounded to the input
// dtype T), and (xq, s_x) is exactly the row-wise quantisation of THAT tensor -- the same
// quant_math.cuh arithmetic as rowwise_quant.cu.  So the integer half of the result is bit-exact
// against the oracle applied to y, and y itself carries the usual floating-point tolerance
// against an fp32 reference of the norm / activation (reduction order, expf / erff).
//
// HBM bytes per row: norm  K*sizeof(T) + K + 4   (unfused: 2*K*sizeof(T) more);
//                    act*up  2*K*sizeof(T) + K + 4 (unfused: 2*K*sizeof(T) more).
// Same structure as rowwise_quant_vec_kernel: TPR threads own a row and keep it in registers
// (16-byte vectors) across the reductions, so global memory is read once and written once.
#include "common.cuh"
#include "ptx.cuh"
#include "quant_math.cuh"
#include <type_traits>

namespace pq {
namespace {

using namespace qmath;

template <typename T> __device__ __forceinline__ float round_to(float f);
template <> __device__ __forceinline__ float round_to<float>(float f) { return f; }
template <> __device__ __forceinline__ float round_to<__nv_bfloat16>(float f) { return __bfloat162float(__float2bfloat16_rn(f)); }
template <> __device__ __forceinline__ float round_to<__half>(float f) { return __half2float(__float2half_rn(f)); }

// pack EPV values that are already representable in T
template <typename T> __device__ __forceinline__ uint4 pack_vec(const float* f);
template <> __device__ __forceinline__ uint4 pack_vec<float>(const float* f) {
  return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
}
template <> __device__ __forceinline__ uint4 pack_vec<__nv_bfloat16>(const float* f) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = (__float_as_uint(f[2 * i]) >> 16) | (__float_as_uint(f[2 * i + 1]) & 0xffff0000u);
  return make_uint4(w[0], w[1], w[2], w[3]);
}
template <> __device__ __forceinline__ uint4 pack_vec<__half>(const float* f) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// Packed-pair arithmetic for the 16-bit dtypes: one cvt.rn.{bf16,f16}x2.f32 rounds two fp32 values to T, and the
// native HMUL2 gives T(a * b) with a single rounding -- identical to rounding the (exact) fp32 product of two
// T values, except for results below the smallest normal fp32 (documented in DESIGN.md).
template <typename T> struct Pk;
template <> struct Pk<__nv_bfloat16> {
  static __device__ __forceinline__ uint32_t cvt2(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  static __device__ __forceinline__ uint32_t mul2(uint32_t a, uint32_t b) {
    const __nv_bfloat162 r = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
};
template <> struct Pk<__half> {
  static __device__ __forceinline__ uint32_t cvt2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  static __device__ __forceinline__ uint32_t mul2(uint32_t a, uint32_t b) {
    const __half2 r = __hmul2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
};
template <> struct Pk<float> {   // never used (fp32 rows stay fp32); keeps the templates well-formed
  static __device__ __forceinline__ uint32_t cvt2(float a, float) { return __float_as_uint(a); }
  static __device__ __forceinline__ uint32_t mul2(uint32_t a, uint32_t) { return a; }
};

// reduction over the TPR threads of a row (ROWS rows per CTA); every thread gets the result
template <int TPR, int ROWS, bool IS_MAX>
__device__ __forceinline__ float row_reduce(float v, float* red /* [ROWS][WPR] */, int row_in_cta, int t) {
  constexpr int WPR = (TPR + 31) / 32;
#pragma unroll
  for (int o = (TPR < 32 ? TPR : 32) / 2; o > 0; o >>= 1) {
    const float w = __shfl_xor_sync(0xffffffffu, v, o);
    v = IS_MAX ? mag_max(v, w) : v + w;      // mag_max: integer max on the bits, NaN propagates
  }
  if (WPR > 1) {
    if ((t & 31) == 0) red[row_in_cta * WPR + (t >> 5)] = v;
    __syncthreads();
    v = red[row_in_cta * WPR];
#pragma unroll
    for (int w = 1; w < WPR; ++w) v = IS_MAX ? mag_max(v, red[row_in_cta * WPR + w]) : v + red[row_in_cta * WPR + w];
  }
  return v;
}

template <typename T, int VPT>
__device__ __forceinline__ float regs_absmax(const uint4 (&v)[VPT]) {
  float amax = 0.f;
  if (sizeof(T) == 2) {
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < VPT; ++i) m = absmax_u16x2(v[i], m);
    amax = u16_mag_to_float<T>(m);
  } else {
#pragma unroll
    for (int i = 0; i < VPT; ++i) amax = vec_absmax<float>(v[i], amax);
  }
  return amax;
}

// quantise the register-resident row `v` (values of type T) and store codes (+ optionally the row itself)
template <typename T, int TPR, int VPT, bool FULL = false>
__device__ __forceinline__ void emit_row(const uint4 (&v)[VPT], const RowQ& rq, int t, int nvec,
                                         int8_t* qr, T* yr) {
  constexpr int EPV = VecTraits<T>::EPV;
  auto emit = [&](auto path_tag) {
    constexpr int PATH = decltype(path_tag)::value;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int vi = t + i * TPR;
      if (FULL || vi < nvec) {
        float f[EPV];
        unpack<T>(v[i], f);
#pragma unroll
        for (int j = 0; j < EPV; ++j) f[j] = quant_one<PATH>(f[j], rq);
        if (EPV == 8) {
          uint2 o;
          o.x = pack4(f[0], f[1], f[2], f[3]);
          o.y = pack4(f[4 % EPV], f[5 % EPV], f[6 % EPV], f[7 % EPV]);
          *reinterpret_cast<uint2*>(qr + (int64_t)vi * 8) = o;
        } else {
          *reinterpret_cast<uint32_t*>(qr + (int64_t)vi * 4) = pack4(f[0], f[1], f[2], f[3]);
        }
        if (yr != nullptr) *reinterpret_cast<uint4*>(yr + (int64_t)vi * EPV) = v[i];
      }
    }
  };
  if (rq.path == 0) emit(std::integral_constant<int, 0>{});
  else if (rq.path == 2) emit(std::integral_constant<int, 2>{});
  else if (rq.path == 1) emit(std::integral_constant<int, 1>{});
  else emit(std::integral_constant<int, 3>{});
}

struct NormArgs {
  const void* x; const void* gamma; const void* beta;   // beta == null: RMSNorm, else LayerNorm
  int8_t* xq; float* s_out; void* y;
  long long M, ldx, ldq, ldy;
  int nvec; int K; float eps; int scale_mode; float qeps;
};

// FULL: rows fill every lane exactly (K == TPR * VPT vectors), so the per-vector bounds checks disappear.  The
// kernel is issue-bound (ncu: 83 % issue-slot utilisation at 0.78 of HBM peak), ~1.5 of ~18 slots per element.
template <typename T, int TPR, int VPT, bool FULL>
__device__ __forceinline__ void norm_quant_body(const NormArgs& a) {
  constexpr int EPV = VecTraits<T>::EPV;
  constexpr int THREADS = (TPR > 256 ? TPR : 256);
  constexpr int ROWS = THREADS / TPR;
  constexpr int WPR = (TPR + 31) / 32;
  __shared__ float red[3][ROWS * WPR];

  const int tid = threadIdx.x;
  const int row_in_cta = tid / TPR;
  const int t = tid % TPR;
  const long long row = (long long)blockIdx.x * ROWS + row_in_cta;
  const bool row_ok = row < a.M;
  const bool layer = a.beta != nullptr;
  constexpr bool full = FULL;

  ptx::griddep_launch_dependents();
  ptx::griddep_wait();
  const T* xr = reinterpret_cast<const T*>(a.x) + (row_ok ? row : 0) * a.ldx;
  uint4 v[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int vi = t + i * TPR;
    v[i] = make_uint4(0, 0, 0, 0);
    if (row_ok && (full || vi < a.nvec)) v[i] = ld_stream_16(xr + (long long)vi * EPV);
  }

  // ---- statistics (fp32) ----
  float mean = 0.f;
  if (layer) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      float f[EPV];
      unpack<T>(v[i], f);
#pragma unroll
      for (int j = 0; j < EPV; ++j) s += f[j];     // padding vectors are zero
    }
    s = row_reduce<TPR, ROWS, false>(s, red[0], row_in_cta, t);
    mean = __fdiv_rn(s, (float)a.K);
  }
  float ss = 0.f;
  if (layer) {
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int vi = t + i * TPR;
      if (full || vi < a.nvec) {
        float f[EPV];
        unpack<T>(v[i], f);
#pragma unroll
        for (int j = 0; j < EPV; ++j) { const float d = f[j] - mean; ss = fmaf(d, d, ss); }
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < VPT; ++i) {      // padding vectors are zero
      float f[EPV];
      unpack<T>(v[i], f);
#pragma unroll
      for (int j = 0; j < EPV; ++j) ss = fmaf(f[j], f[j], ss);
    }
  }
  ss = row_reduce<TPR, ROWS, false>(ss, red[1], row_in_cta, t);
  const float rstd = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fdiv_rn(ss, (float)a.K), a.eps)));

  // ---- normalise in place: v <- the tensor the unfused op would have stored ----
  const T* gr = reinterpret_cast<const T*>(a.gamma);
  const T* br = reinterpret_cast<const T*>(a.beta);
  if (layer) {
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int vi = t + i * TPR;
      if (full || vi < a.nvec) {
        float f[EPV], g[EPV], b[EPV];
        unpack<T>(v[i], f);
        unpack<T>(__ldg(reinterpret_cast<const uint4*>(gr + (long long)vi * EPV)), g);
        unpack<T>(__ldg(reinterpret_cast<const uint4*>(br + (long long)vi * EPV)), b);
#pragma unroll
        for (int j = 0; j < EPV; ++j) f[j] = fmaf(__fmul_rn(f[j] - mean, rstd), g[j], b[j]);
        if (sizeof(T) == 2) {        // one rounding to T, two elements per cvt
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = Pk<T>::cvt2(f[(2 * j) % EPV], f[(2 * j + 1) % EPV]);
          v[i] = make_uint4(o[0], o[1], o[2], o[3]);
        } else {
          v[i] = pack_vec<T>(f);
        }
      }
    }
  } else {
    // Llama-style RMSNorm: weight * (x * rstd).to(dtype), both products rounded to T
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int vi = t + i * TPR;
      if (full || vi < a.nvec) {
        float f[EPV];
        unpack<T>(v[i], f);
        const uint4 gv = __ldg(reinterpret_cast<const uint4*>(gr + (long long)vi * EPV));
        if (sizeof(T) == 2) {
          const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            o[j] = Pk<T>::mul2(gw[j], Pk<T>::cvt2(__fmul_rn(f[(2 * j) % EPV], rstd), __fmul_rn(f[(2 * j + 1) % EPV], rstd)));
          v[i] = make_uint4(o[0], o[1], o[2], o[3]);
        } else {
          float g[EPV];
          unpack<T>(gv, g);
#pragma unroll
          for (int j = 0; j < EPV; ++j) f[j] = __fmul_rn(g[j], __fmul_rn(f[j], rstd));
          v[i] = pack_vec<T>(f);
        }
      }
    }
  }

  float amax = regs_absmax<T, VPT>(v);
  amax = row_reduce<TPR, ROWS, true>(amax, red[2], row_in_cta, t);
  const RowQ rq = make_rowq(amax, a.scale_mode, a.qeps);
  if (row_ok && t == 0) a.s_out[row] = rq.s;
  if (!row_ok) return;
  T* yr = a.y ? reinterpret_cast<T*>(a.y) + row * a.ldy : nullptr;
  emit_row<T, TPR, VPT, FULL>(v, rq, t, a.nvec, a.xq + row * a.ldq, yr);
}

template <typename T, int TPR, int VPT>
__global__ void __launch_bounds__((TPR > 256 ? TPR : 256), (TPR > 256 ? 1024 / TPR : 4))   // <= 64 registers
norm_quant_kernel(const NormArgs a) {
  if (a.nvec == TPR * VPT) norm_quant_body<T, TPR, VPT, true>(a);
  else norm_quant_body<T, TPR, VPT, false>(a);
}

struct ActArgs {
  const void* gate; const void* up;    // up == null: h = act(gate)
  int8_t* hq; float* s_out; void* h;
  long long M, ldg, ldu, ldq, ldh;
  int nvec; int act; int scale_mode; float qeps;
  // row-parallel consumer: the row maximum of h spans every rank's column slice.
  //   n_amax_out > 0: "statistics" launch -- only the local row maxima are computed and stored to amax_out[d][row]
  //                   (this rank's slot in every rank's exchange buffer); no codes, no scales;
  //   amax_in != 0:   quantise with the maximum over `amax_slots` arrays (stride `amax_stride`) instead of the local one.
  const float* amax_in; int amax_slots; long long amax_stride;
  float* amax_out[8]; int n_amax_out;
};

// SiLU uses ex2.approx + rcp.approx (two MUFU ops, ~2^-21 relative error: below half an ulp of every storage
// dtype but fp32, where it is inside the stated 2e-6 tolerance); an IEEE division and expf() would make the kernel
// issue-bound at 40 % of HBM peak.
template <int ACT>
__device__ __forceinline__ float act_fn(float x) {
  if (ACT == PQ_ACT_SILU) return __fdividef(x, 1.0f + __expf(-x));
  if (ACT == PQ_ACT_GELU) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
  if (ACT == PQ_ACT_GELU_TANH) {
    const float u = 0.79788456080286535588f * (x + 0.044715f * x * x * x);
    return 0.5f * x * (1.0f + tanhf(u));
  }
  return x;   // PQ_ACT_IDENTITY
}

template <typename T, int TPR, int VPT>
__global__ void __launch_bounds__((TPR > 256 ? TPR : 256))
act_mul_quant_kernel(const ActArgs a) {
  constexpr int EPV = VecTraits<T>::EPV;
  constexpr int THREADS = (TPR > 256 ? TPR : 256);
  constexpr int ROWS = THREADS / TPR;
  constexpr int WPR = (TPR + 31) / 32;
  __shared__ float red[ROWS * WPR];

  const int tid = threadIdx.x;
  const int row_in_cta = tid / TPR;
  const int t = tid % TPR;
  const long long row = (long long)blockIdx.x * ROWS + row_in_cta;
  const bool row_ok = row < a.M;

  ptx::griddep_launch_dependents();
  ptx::griddep_wait();
  const T* gr = reinterpret_cast<const T*>(a.gate) + (row_ok ? row : 0) * a.ldg;
  const T* ur = a.up ? reinterpret_cast<const T*>(a.up) + (row_ok ? row : 0) * a.ldu : nullptr;
  uint4 v[VPT];
  uint4 uv[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {     // all loads first: 2 x VPT 16-byte requests in flight per thread
    const int vi = t + i * TPR;
    v[i] = make_uint4(0, 0, 0, 0);
    uv[i] = make_uint4(0, 0, 0, 0);
    if (row_ok && vi < a.nvec) {
      v[i] = ld_stream_16(gr + (long long)vi * EPV);
      if (ur != nullptr) uv[i] = ld_stream_16(ur + (long long)vi * EPV);
    }
  }
  auto body = [&](auto act_tag) {
    constexpr int ACT = decltype(act_tag)::value;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      float f[EPV];
      unpack<T>(v[i], f);
#pragma unroll
      for (int j = 0; j < EPV; ++j) f[j] = act_fn<ACT>(f[j]);
      // act(gate) is a tensor of dtype T in the unfused graph, so it is rounded before the product
      if (sizeof(T) == 2) {
        const uint32_t uw[4] = {uv[i].x, uv[i].y, uv[i].z, uv[i].w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          o[j] = Pk<T>::cvt2(f[(2 * j) % EPV], f[(2 * j + 1) % EPV]);
          if (ur != nullptr) o[j] = Pk<T>::mul2(o[j], uw[j]);
        }
        v[i] = make_uint4(o[0], o[1], o[2], o[3]);
      } else {
        if (ur != nullptr) {
          float u[EPV];
          unpack<T>(uv[i], u);
#pragma unroll
          for (int j = 0; j < EPV; ++j) f[j] = __fmul_rn(f[j], u[j]);
        }
        v[i] = pack_vec<T>(f);
      }
    }
  };
  if (a.act == PQ_ACT_SILU) body(std::integral_constant<int, PQ_ACT_SILU>{});
  else if (a.act == PQ_ACT_GELU) body(std::integral_constant<int, PQ_ACT_GELU>{});
  else if (a.act == PQ_ACT_GELU_TANH) body(std::integral_constant<int, PQ_ACT_GELU_TANH>{});
  else body(std::integral_constant<int, PQ_ACT_IDENTITY>{});
  float amax = regs_absmax<T, VPT>(v);
  amax = row_reduce<TPR, ROWS, true>(amax, red, row_in_cta, t);
  if (a.n_amax_out > 0) {
    if (row_ok && t == 0)
      for (int d = 0; d < a.n_amax_out; ++d) a.amax_out[d][row] = amax;
    return;
  }
  if (a.amax_in != nullptr && row_ok) amax = given_amax(a.amax_in, a.amax_slots, a.amax_stride, row);
  const RowQ rq = make_rowq(amax, a.scale_mode, a.qeps);
  if (row_ok && t == 0) a.s_out[row] = rq.s;
  if (!row_ok) return;
  T* hr = a.h ? reinterpret_cast<T*>(a.h) + row * a.ldh : nullptr;
  emit_row<T, TPR, VPT>(v, rq, t, a.nvec, a.hq + row * a.ldq, hr);
}

template <typename Args>
cudaError_t launch_pdl(void (*kern)(const Args), unsigned grid, unsigned block, cudaStream_t st, const Args& args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(block, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, args);
}

// (threads per row, 16-byte vectors per thread) covering nvec vectors with the fewest idle lanes;
// the same scoring as rowwise_quant.cu restricted to the shapes instantiated here
Knob g_fq_tpr{0}, g_fq_vpt{0};   // test hook (pq_debug_set_fused_quant_config): force (threads per row, vectors per thread)

void pick_config(int nvec, long long M, int* tpr_out, int* vpt_out) {
  const int ft = g_fq_tpr, fv = g_fq_vpt;
  if (ft > 0 && fv > 0 && ft * fv >= nvec) { *tpr_out = ft; *vpt_out = fv; return; }
  static const int kVpt[4] = {4, 3, 6, 8};
  int best_tpr = 1024, best_vpt = 8;
  double best = 1e30;
  for (int i = 0; i < 4; ++i) {
    const int vpt = kVpt[i];
    int tpr = 32;
    while (tpr < 1024 && tpr * vpt < nvec) tpr *= 2;
    if (tpr * vpt < nvec) continue;
    double score = (double)(tpr * vpt - nvec) / (double)(tpr * vpt);
    if (tpr >= 512) score += 0.04;
    if (tpr == 1024) score += 0.20;
    if (M >= 16384) score -= (vpt == 8 ? 0.02 : vpt == 6 ? 0.01 : 0.0);
    else if (vpt == 8) score += 0.03;
    if (score < best) { best = score; best_tpr = tpr; best_vpt = vpt; }
  }
  *tpr_out = best_tpr; *vpt_out = best_vpt;
}

template <typename T, template <typename, int, int> class Launcher, typename Args>
int dispatch_cfg(const Args& a, long long M, int nvec, cudaStream_t st) {
  int tpr, vpt;
  pick_config(nvec, M, &tpr, &vpt);
#define PQ_CASE_V(TPR, VPT) if (tpr == TPR && vpt == VPT) return Launcher<T, TPR, VPT>::run(a, M, st);
#define PQ_CASE_T(TPR) PQ_CASE_V(TPR, 3) PQ_CASE_V(TPR, 4) PQ_CASE_V(TPR, 6) PQ_CASE_V(TPR, 8)
  PQ_CASE_T(32) PQ_CASE_T(64) PQ_CASE_T(128) PQ_CASE_T(256) PQ_CASE_T(512) PQ_CASE_T(1024)
#undef PQ_CASE_T
#undef PQ_CASE_V
  PQ_FAIL(PQ_ERR_UNSUPPORTED, "fused quantizer: no kernel configuration for %d vectors per row", nvec);
}

template <typename T, int TPR, int VPT>
struct NormLauncher {
  static int run(const NormArgs& a, long long M, cudaStream_t st) {
    constexpr int THREADS = (TPR > 256 ? TPR : 256);
    constexpr int ROWS = THREADS / TPR;
    PQ_CUDA(launch_pdl<NormArgs>(norm_quant_kernel<T, TPR, VPT>, (unsigned)((M + ROWS - 1) / ROWS), THREADS, st, a));
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return PQ_OK;
  }
};
template <typename T, int TPR, int VPT>
struct ActLauncher {
  static int run(const ActArgs& a, long long M, cudaStream_t st) {
    constexpr int THREADS = (TPR > 256 ? TPR : 256);
    constexpr int ROWS = THREADS / TPR;
    PQ_CUDA(launch_pdl<ActArgs>(act_mul_quant_kernel<T, TPR, VPT>, (unsigned)((M + ROWS - 1) / ROWS), THREADS, st, a));
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return PQ_OK;
  }
};

bool aligned16(const void* p, long long ld_elems, int esz) {
  return (((uintptr_t)p & 15) == 0) && ((ld_elems * esz) % 16 == 0);
}

int check_spec(const pq_quant_spec& spec, const char* who) {
  if (spec.scale_mode < PQ_DIV || spec.scale_mode > PQ_INV_SCALE) PQ_FAIL(PQ_ERR_ARG, "%s: bad scale_mode %d", who, spec.scale_mode);
  if (spec.qmin != -128 && spec.qmin != -127) PQ_FAIL(PQ_ERR_ARG, "%s: qmin must be -128 or -127", who);
  return PQ_OK;
}

}  // namespace
}  // namespace pq

using namespace pq;

extern "C" int pq_norm_quant(const void* x, int x_dtype, int64_t M, int64_t K, int64_t ldx,
                             const void* gamma, const void* beta, float eps,
                             int8_t* xq, int64_t ldq, float* s_x, void* y, int64_t ldy,
                             const pq_quant_spec* spec_in, void* stream) {
  int rc = check_device(nullptr);
  if (rc) return rc;
  const pq_quant_spec spec = resolve_spec(spec_in);
  rc = check_spec(spec, "pq_norm_quant");
  if (rc) return rc;
  if (M < 0 || K < 1) PQ_FAIL(PQ_ERR_ARG, "pq_norm_quant: bad shape M=%lld K=%lld", (long long)M, (long long)K);
  if (M == 0) return PQ_OK;
  if (!x || !gamma || !xq || !s_x) PQ_FAIL(PQ_ERR_ARG, "pq_norm_quant: null pointer");
  if (ldx < K || ldq < K || (y && ldy < K)) PQ_FAIL(PQ_ERR_ARG, "pq_norm_quant: leading dimension smaller than K");
  if (M > 0x7fffffffLL) PQ_FAIL(PQ_ERR_ARG, "pq_norm_quant: M too large");
  const int esz = dtype_size(x_dtype);
  if (esz == 0 || x_dtype == PQ_I32) PQ_FAIL(PQ_ERR_ARG, "pq_norm_quant: unsupported dtype %d", x_dtype);
  const int epv = 16 / esz;
  if (K % epv != 0 || K / epv > 8192)
    PQ_FAIL(PQ_ERR_UNSUPPORTED, "pq_norm_quant: K=%lld must be a multiple of %d and at most %d", (long long)K, epv, 8192 * epv);
  if (!aligned16(x, ldx, esz) || !aligned16(gamma, 0, esz) || (beta && !aligned16(beta, 0, esz)) ||
      (y && !aligned16(y, ldy, esz)) || ((uintptr_t)xq % epv) || (ldq % epv))
    PQ_FAIL(PQ_ERR_ALIGN, "pq_norm_quant: x / gamma / beta / y rows must be 16-byte aligned, xq rows %d-byte aligned", epv);
  NormArgs a;
  a.x = x; a.gamma = gamma; a.beta = beta; a.xq = xq; a.s_out = s_x; a.y = y;
  a.M = M; a.ldx = ldx; a.ldq = ldq; a.ldy = ldy;
  a.nvec = (int)(K / epv); a.K = (int)K; a.eps = eps; a.scale_mode = mode_bits(spec); a.qeps = spec.eps;
  cudaStream_t st = (cudaStream_t)stream;
  switch (x_dtype) {
    case PQ_F32: return dispatch_cfg<float, NormLauncher>(a, M, a.nvec, st);
    case PQ_F16: return dispatch_cfg<__half, NormLauncher>(a, M, a.nvec, st);
    default: return dispatch_cfg<__nv_bfloat16, NormLauncher>(a, M, a.nvec, st);
  }
}

extern "C" int pq_act_mul_quant(const void* gate, const void* up, int dtype, int act,
                                int64_t M, int64_t K, int64_t ldg, int64_t ldu,
                                int8_t* hq, int64_t ldq, float* s_h, void* h, int64_t ldh,
                                const pq_quant_spec* spec_in, void* stream) {
  return pq::launch_act_mul_quant(gate, up, dtype, act, M, K, ldg, ldu, hq, ldq, s_h, h, ldh, resolve_spec(spec_in),
                                  (cudaStream_t)stream, nullptr, 0, 0, nullptr, 0);
}

int pq::launch_act_mul_quant(const void* gate, const void* up, int dtype, int act,
                             int64_t M, int64_t K, int64_t ldg, int64_t ldu,
                             int8_t* hq, int64_t ldq, float* s_h, void* h, int64_t ldh,
                             const pq_quant_spec& spec, cudaStream_t stream,
                             const float* amax_in, int amax_slots, long long amax_stride,
                             float* const* amax_out, int n_amax_out) {
  int rc = check_device(nullptr);
  if (rc) return rc;
  rc = check_spec(spec, "pq_act_mul_quant");
  if (rc) return rc;
  if (M < 0 || K < 1) PQ_FAIL(PQ_ERR_ARG, "pq_act_mul_quant: bad shape M=%lld K=%lld", (long long)M, (long long)K);
  if (act < PQ_ACT_IDENTITY || act > PQ_ACT_GELU_TANH) PQ_FAIL(PQ_ERR_ARG, "pq_act_mul_quant: unknown activation %d", act);
  if (M == 0) return PQ_OK;
  const bool stats_only = n_amax_out > 0;
  if (n_amax_out < 0 || n_amax_out > 8 || (stats_only && !amax_out)) PQ_FAIL(PQ_ERR_ARG, "pq_act_mul_quant: bad amax destination list");
  if (!gate || (!stats_only && (!hq || !s_h))) PQ_FAIL(PQ_ERR_ARG, "pq_act_mul_quant: null pointer");
  if (stats_only) { hq = nullptr; ldq = K; h = nullptr; }
  if (ldg < K || (up && ldu < K) || ldq < K || (h && ldh < K)) PQ_FAIL(PQ_ERR_ARG, "pq_act_mul_quant: leading dimension smaller than K");
  if (M > 0x7fffffffLL) PQ_FAIL(PQ_ERR_ARG, "pq_act_mul_quant: M too large");
  const int esz = dtype_size(dtype);
  if (esz == 0 || dtype == PQ_I32) PQ_FAIL(PQ_ERR_ARG, "pq_act_mul_quant: unsupported dtype %d", dtype);
  const int epv = 16 / esz;
  if (K % epv != 0 || K / epv > 8192)
    PQ_FAIL(PQ_ERR_UNSUPPORTED, "pq_act_mul_quant: K=%lld must be a multiple of %d and at most %d", (long long)K, epv, 8192 * epv);
  if (!aligned16(gate, ldg, esz) || (up && !aligned16(up, ldu, esz)) || (h && !aligned16(h, ldh, esz)) ||
      ((uintptr_t)hq % epv) || (ldq % epv))
    PQ_FAIL(PQ_ERR_ALIGN, "pq_act_mul_quant: gate / up / h rows must be 16-byte aligned, hq rows %d-byte aligned", epv);
  ActArgs a;
  a.gate = gate; a.up = up; a.hq = hq; a.s_out = s_h; a.h = h;
  a.M = M; a.ldg = ldg; a.ldu = ldu; a.ldq = ldq; a.ldh = ldh;
  a.nvec = (int)(K / epv); a.act = act; a.scale_mode = mode_bits(spec); a.qeps = spec.eps;
  a.amax_in = amax_in; a.amax_slots = amax_slots; a.amax_stride = amax_stride;
  a.n_amax_out = n_amax_out;
  for (int d = 0; d < 8; ++d) a.amax_out[d] = d < n_amax_out ? amax_out[d] : nullptr;
  cudaStream_t st = stream;
  switch (dtype) {
    case PQ_F32: return dispatch_cfg<float, ActLauncher>(a, M, a.nvec, st);
    case PQ_F16: return dispatch_cfg<__half, ActLauncher>(a, M, a.nvec, st);
    default: return dispatch_cfg<__nv_bfloat16, ActLauncher>(a, M, a.nvec, st);
  }
}

// Test/bench hook: force the launch shape of norm_quant_kernel / act_mul_quant_kernel; 0,0 = heuristic.
extern "C" void pq_debug_set_fused_quant_config(int tpr, int vpt) { pq::g_fq_tpr = tpr; pq::g_fq_vpt = vpt; }
