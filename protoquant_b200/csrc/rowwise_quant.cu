// Row-wise symmetric int8 quantizer.
//
//   amax[m] = max_k |x[m,k]|            (fp32, after an exact upcast of bf16/fp16)
//   s[m]    = amax/127                  (fp32 IEEE division; amax == 0 -> s = 1)
//   q[m,k]  = rne(x[m,k] / s[m])        (fp32 IEEE division, then round-half-even)
//
// The kernel is HBM-bound: 16-byte loads, the whole row kept in registers between
// the absmax pass and the quantize pass (one global read, one global write), warp
// shuffles + one smem hop for the row reduction.  Algorithmic bytes per row:
// K*sizeof(in) + K + 4.
//
// Exactness.  A true `div.rn.f32` per element costs ~10 issue slots plus a branch,
// which would make this kernel issue-bound at 6.5 TB/s.  Instead, per row we form
// y = RN(1/s) once and per element
//     q0 = RN(x*y);  r = fma(-q0, s, x);  q1 = fma(r, y, q0)
// which is the classic FMA division step: q1 == RN(x/s) whenever no intermediate
// underflows.  tools/check_fma_div.c proves rne(q1) == rne(x/s) exhaustively for
// every (amax, x) pair of bf16 and of fp16 inputs with s in [2^-100, 2^100]; rows
// whose scale falls outside [2^-60, 2^60] take the `div.rn.f32` path instead.
// The final round-half-even to integer is done with the 1.5*2^23 magic add, whose
// low mantissa byte is the two's complement int8.  On the fast paths |q1| <= 127.01 always, so the
// [-128,127] clamp of the reference formula cannot fire; rows with a NaN / inf / zero / denormal scale
// take the "careful" paths of quant_math.cuh, which evaluate the formula literally (clamp live, NaN -> 0).
#include "common.cuh"
#include "ptx.cuh"
#include "quant_math.cuh"
#include "gemm_common.cuh"   // PerDeviceOnce
#include <type_traits>

namespace pq {
Knob g_force_tpr{0}, g_force_vpt{0};   // test hook (pq_debug_set_quant_config)
Knob g_weight_prefetch{1};              // pq_qlinear: the act-quant kernel pulls the weights into L2
Knob g_quant_staged{0};                 // 0 = heuristic, 1 = always the shared-memory staged kernel, -1 = never
namespace {

using namespace qmath;

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), unsigned grid, unsigned block, cudaStream_t st, Args... args) {
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
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- vectorised, register-resident kernel ----------------------------------------
// TPR threads cooperate on one row; each holds up to VPT 16-byte vectors of it.
template <typename T, int TPR, int VPT>
__global__ void __launch_bounds__((TPR > 256 ? TPR : 256))
rowwise_quant_vec_kernel(const T* __restrict__ x, int64_t M, int nvec, int64_t ldx,
                         int8_t* __restrict__ xq, int64_t ldq, float* __restrict__ s_out,
                         int scale_mode, float eps, const uint8_t* __restrict__ pf, long long pf_bytes,
                         const float* __restrict__ amax_in, int amax_slots, long long amax_stride) {
  constexpr int EPV = VecTraits<T>::EPV;
  constexpr int THREADS = (TPR > 256 ? TPR : 256);
  constexpr int ROWS = THREADS / TPR;
  constexpr int WPR = (TPR + 31) / 32;  // warps per row
  __shared__ float red[ROWS][WPR > 1 ? WPR : 1];

  const int tid = threadIdx.x;
  const int row_in_cta = tid / TPR;
  const int t = tid % TPR;
  const int64_t row = (int64_t)blockIdx.x * ROWS + row_in_cta;
  const bool row_ok = row < M;

  ptx::griddep_launch_dependents();
  // Weight prefetch (pq_qlinear): the GEMM that follows streams `pf_bytes` of static int8 weights; this
  // kernel leaves most of the DRAM bandwidth idle at activation sizes, so every CTA pulls its slice
  // of them into L2 (no registers, no completion to wait for) before it waits for its own input.
  if (pf != nullptr && tid == 0) {
    const long long per = ((pf_bytes + gridDim.x - 1) / gridDim.x + 15) & ~15LL;
    long long off = (long long)blockIdx.x * per;
    const long long end = (off + per < pf_bytes) ? off + per : (pf_bytes & ~15LL);
    while (off < end) {
      const uint32_t n = (uint32_t)((end - off < 16384) ? end - off : 16384);
      ptx::prefetch_l2_bulk(pf + off, n);
      off += n;
    }
  }
  ptx::griddep_wait();   // x may be produced, and xq still be read, by the previous kernel
  const T* xr = x + (row_ok ? row : 0) * ldx;
  uint4 v[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int vi = t + i * TPR;
    v[i] = make_uint4(0, 0, 0, 0);
    if (row_ok && vi < nvec) v[i] = ld_stream_16(xr + (int64_t)vi * EPV);
  }

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
#pragma unroll
  for (int o = (TPR < 32 ? TPR : 32) / 2; o > 0; o >>= 1)
    amax = mag_max(amax, __shfl_xor_sync(0xffffffffu, amax, o));   // integer max on the bits: NaN propagates
  if (WPR > 1) {
    if ((t & 31) == 0) red[row_in_cta][t >> 5] = amax;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < WPR; ++w) amax = mag_max(amax, red[row_in_cta][w]);
  }
  // row-parallel shards quantise a K-slice with the |.|-max of the WHOLE row (pq_act_quant_amax)
  if (amax_in != nullptr && row_ok) amax = given_amax(amax_in, amax_slots, amax_stride, row);

  const RowQ rq = make_rowq(amax, scale_mode, eps);
  if (row_ok && t == 0) s_out[row] = rq.s;
  if (!row_ok) return;

  int8_t* qr = xq + row * ldq;
  auto emit = [&](auto path_tag) {
    constexpr int PATH = decltype(path_tag)::value;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int vi = t + i * TPR;
      if (vi < nvec) {
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
      }
    }
  };
  if (rq.path == 0) emit(std::integral_constant<int, 0>{});
  else if (rq.path == 2) emit(std::integral_constant<int, 2>{});
  else if (rq.path == 1) emit(std::integral_constant<int, 1>{});
  else emit(std::integral_constant<int, 3>{});
}

// ---- shared-memory staged kernel: one persistent CTA per SM --------------------------------
// At activation sizes (M x K of a few tens of MB) the register-resident kernel above is bound by latency, not
// bandwidth: its CTAs form 1.7-2 waves, every CTA runs load -> reduce -> quantise -> store in lock step, and only
// what fits in registers is in flight.  Here each SM gets ONE 1024-thread CTA that owns rows b, b + grid, ...; a
// single thread requests ALL of them up front with 1-D bulk copies (cp.async.bulk -> UBLKCP, up to ~220 KB in
// flight per SM, no registers involved), one mbarrier per slot, and eight 128-thread groups consume rows as they
// land: |.|-max pass and quantise pass both read the row from shared memory (128 B/clk -- free next to DRAM), codes
// go out with 8-byte stores.  Slots are refilled as soon as a group has finished with them, so any M works.
// Same arithmetic (quant_math.cuh) => bit-identical codes and scales.
constexpr int STG_THREADS = 1024, STG_TPR = 128, STG_GROUPS = STG_THREADS / STG_TPR;
constexpr int STG_HEADER = 1024;         // mbarriers (8 B x <= 64 slots) + reduction scratch
constexpr int STG_MAX_SLOTS = 64;

template <typename T>
__global__ void __launch_bounds__(STG_THREADS, 1)
rowwise_quant_staged_kernel(const T* __restrict__ x, int64_t M, int nvec, int64_t ldx,
                            int8_t* __restrict__ xq, int64_t ldq, float* __restrict__ s_out,
                            int scale_mode, float eps, const uint8_t* __restrict__ pf, long long pf_bytes,
                            const float* __restrict__ amax_in, int amax_slots, long long amax_stride, int slots) {
  constexpr int EPV = VecTraits<T>::EPV;
  extern __shared__ __align__(128) uint8_t stg_smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg_smem);                       // [slots]
  float* red = reinterpret_cast<float*>(stg_smem + 8 * STG_MAX_SLOTS);          // [2][GROUPS][4]
  uint8_t* buf = stg_smem + STG_HEADER;
  const uint32_t row_bytes = (uint32_t)nvec * 16u;
  const int tid = threadIdx.x, g = tid / STG_TPR, t = tid % STG_TPR;
  const int64_t nmine = (M - (int64_t)blockIdx.x + gridDim.x - 1) / gridDim.x;   // rows b, b + grid, ...
  // `slots` is a multiple of the number of ACTIVE groups (or smaller than STG_GROUPS, then it is that number): the row
  // that used a slot before row i, i - slots, then belongs to the same group as row i, so a group never waits for
  // phase p + 1 of a barrier whose phase p is still pending (an mbarrier parity wait cannot tell those apart).
  const int ngroups = slots < STG_GROUPS ? slots : STG_GROUPS;
  auto row_of = [&](int64_t i) { return (int64_t)blockIdx.x + i * gridDim.x; };

  if (tid == 0) {
    for (int s = 0; s < slots; ++s) ptx::mbar_init(ptx::smem_u32(bars + s), 1);
    ptx::fence_mbar_init();
  }
  ptx::griddep_launch_dependents();
  if (pf != nullptr && tid == 32) {      // weight prefetch for the GEMM that follows (see the kernel above)
    const long long per = ((pf_bytes + gridDim.x - 1) / gridDim.x + 15) & ~15LL;
    long long off = (long long)blockIdx.x * per;
    const long long end = (off + per < pf_bytes) ? off + per : (pf_bytes & ~15LL);
    while (off < end) {
      const uint32_t n = (uint32_t)((end - off < 16384) ? end - off : 16384);
      ptx::prefetch_l2_bulk(pf + off, n);
      off += n;
    }
  }
  __syncthreads();
  ptx::griddep_wait();   // x may be produced, and xq still be read, by the previous kernel
  if (tid == 0) {
    const int64_t n0 = nmine < slots ? nmine : slots;
    for (int64_t i = 0; i < n0; ++i) {
      const uint32_t bar = ptx::smem_u32(bars + i);
      ptx::mbar_arrive_expect_tx(bar, row_bytes);
      ptx::bulk_load_g2s(ptx::smem_u32(buf + (size_t)i * row_bytes), x + row_of(i) * ldx, row_bytes, bar);
    }
  }
  for (int64_t i = g; g < ngroups && i < nmine; i += ngroups) {
    const int slot = (int)(i % slots);
    const uint32_t parity = (uint32_t)((i / slots) & 1);
    const int64_t row = row_of(i);
    ptx::mbar_wait(ptx::smem_u32(bars + slot), parity);
    const uint4* rv = reinterpret_cast<const uint4*>(buf + (size_t)slot * row_bytes);
    // pass 1: |.|-max of the row
    float amax = 0.f;
    if (sizeof(T) == 2) {
      uint32_t m = 0;
      for (int v = t; v < nvec; v += STG_TPR) m = absmax_u16x2(rv[v], m);
      amax = u16_mag_to_float<T>(m);
    } else {
      for (int v = t; v < nvec; v += STG_TPR) amax = vec_absmax<float>(rv[v], amax);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = mag_max(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    float* rr = red + (((i / ngroups) & 1) * STG_GROUPS + g) * 4;        // double-buffered per group
    if ((t & 31) == 0) rr[t >> 5] = amax;
    ptx::named_bar_sync(1 + g, STG_TPR);
    amax = mag_max(mag_max(rr[0], rr[1]), mag_max(rr[2], rr[3]));
    if (amax_in != nullptr) amax = given_amax(amax_in, amax_slots, amax_stride, row);
    const RowQ rq = make_rowq(amax, scale_mode, eps);
    if (t == 0) s_out[row] = rq.s;
    // pass 2: quantise from shared memory
    int8_t* qr = xq + row * ldq;
    auto emit = [&](auto path_tag) {
      constexpr int PATH = decltype(path_tag)::value;
#pragma unroll 2
      for (int v = t; v < nvec; v += STG_TPR) {
        float f[EPV];
        unpack<T>(rv[v], f);
#pragma unroll
        for (int j = 0; j < EPV; ++j) f[j] = quant_one<PATH>(f[j], rq);
        if (EPV == 8) {
          uint2 o;
          o.x = pack4(f[0], f[1], f[2], f[3]);
          o.y = pack4(f[4 % EPV], f[5 % EPV], f[6 % EPV], f[7 % EPV]);
          *reinterpret_cast<uint2*>(qr + (int64_t)v * 8) = o;
        } else {
          *reinterpret_cast<uint32_t*>(qr + (int64_t)v * 4) = pack4(f[0], f[1], f[2], f[3]);
        }
      }
    };
    if (rq.path == 0) emit(std::integral_constant<int, 0>{});
    else if (rq.path == 2) emit(std::integral_constant<int, 2>{});
    else if (rq.path == 1) emit(std::integral_constant<int, 1>{});
    else emit(std::integral_constant<int, 3>{});
    // refill the slot with the row that will use it next (uniform across the group)
    if (i + slots < nmine) {
      ptx::named_bar_sync(1 + g, STG_TPR);            // every thread of the group is done reading the slot
      if (t == 0) {
        const uint32_t bar = ptx::smem_u32(bars + slot);
        ptx::fence_proxy_async_smem();                // generic-proxy reads before the async-proxy overwrite
        ptx::mbar_arrive_expect_tx(bar, row_bytes);
        ptx::bulk_load_g2s(ptx::smem_u32(buf + (size_t)slot * row_bytes), x + row_of(i + slots) * ldx, row_bytes, bar);
      }
    }
  }
}

// ---- generic kernel: any K / stride / alignment, optional transposed output --------
template <typename T> __device__ __forceinline__ float load_as_float(const T* p);
template <> __device__ __forceinline__ float load_as_float<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float load_as_float<__half>(const __half* p) { return __half2float(*p); }
template <> __device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename T>
__global__ void __launch_bounds__(256)
rowwise_quant_generic_kernel(const T* __restrict__ x, int64_t M, int64_t K, int64_t ldx,
                             int8_t* __restrict__ xq, int64_t ldq, float* __restrict__ s_out,
                             int transpose, int scale_mode, float eps, const float* __restrict__ amax_in,
                             int amax_slots, long long amax_stride) {
  __shared__ float red[8];
  ptx::griddep_launch_dependents();
  ptx::griddep_wait();
  const int64_t row = blockIdx.x;
  const T* xr = x + row * ldx;
  float amax = 0.f;
  for (int64_t k = threadIdx.x; k < K; k += 256) amax = mag_max(amax, mag_of(load_as_float<T>(xr + k)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = mag_max(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < 8; ++w) amax = mag_max(amax, red[w]);
  if (amax_in != nullptr) amax = given_amax(amax_in, amax_slots, amax_stride, row);
  const RowQ rq = make_rowq(amax, scale_mode, eps);
  if (threadIdx.x == 0) s_out[row] = rq.s;
  for (int64_t k = threadIdx.x; k < K; k += 256) {
    const float xv = load_as_float<T>(xr + k);
    const float m = quant_any(xv, rq);
    if (transpose) xq[k * ldq + row] = code_of(m);
    else xq[row * ldq + k] = code_of(m);
  }
}

// ---- transposed output in two launches: row scales, then [128 rows x 128 columns] tiles -----------------
// The scale of a row needs the whole row, the transposed output wants many rows per CTA (R contiguous bytes per
// output row) -- so the work is split: `row_scale_kernel` streams x once and writes s[m]; `transpose_tile_kernel`
// re-reads x tile by tile (from L2 at activation sizes), derives the row parameters from s[m] alone, quantises,
// turns 4 rows x 4 columns per lane into four words with PRMT, and writes 128 contiguous bytes per output row.
// Small independent CTAs (52 KB of shared memory, 4 per SM): loads, arithmetic and stores of different tiles
// overlap.  (An 8-CTA-cluster variant that kept an [R x K/8] slab resident and read x once was built and measured
// 3-5x SLOWER: one 168 KB CTA per SM runs its phases in lock step, and only ~8 such clusters are co-resident.)
// Same arithmetic (quant_math.cuh) => bit-identical codes and scales.
constexpr int TT_ROWS = 128, TT_COLS = 128, TT_THREADS = 512;
constexpr int TT_TP = TT_ROWS / 4 + 1;                // words per row of the transposed tile (odd)

template <typename T>
__global__ void __launch_bounds__(256)
row_scale_kernel(const T* __restrict__ x, int64_t K, int64_t ldx, float* __restrict__ s_out, int scale_mode, float eps, int vec) {
  constexpr int EPV = VecTraits<T>::EPV;
  __shared__ float red[8];
  ptx::griddep_launch_dependents();
  ptx::griddep_wait();
  const T* xr = x + (int64_t)blockIdx.x * ldx;
  float amax = 0.f;
  if (vec) {
    const int64_t nvec = K / EPV;
    if (sizeof(T) == 2) {
      uint32_t m = 0;
      for (int64_t v = threadIdx.x; v < nvec; v += 256) m = absmax_u16x2(ld_stream_16(xr + v * EPV), m);
      amax = u16_mag_to_float<T>(m);
    } else {
      for (int64_t v = threadIdx.x; v < nvec; v += 256) amax = vec_absmax<float>(ld_stream_16(xr + v * EPV), amax);
    }
  } else {
    for (int64_t k = threadIdx.x; k < K; k += 256) amax = mag_max(amax, mag_of((float)xr[k]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = mag_max(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) amax = mag_max(amax, red[w]);
    s_out[blockIdx.x] = make_rowq(amax, scale_mode, eps).s;
  }
}

// RowQ of the PQ_DIV / PQ_RCP_MUL modes from the stored scale alone (make_rowq derives everything but s from s there)
__device__ __forceinline__ RowQ rowq_from_scale(float s, int mode) {
  RowQ r;
  r.s = s;
  r.qmin = (mode & 0x100) ? -127.f : -128.f;
  const bool safe = (s >= 0x1p-60f) && (s <= 0x1p60f);
  r.mul = __frcp_rn(s);
  r.path = ((mode & 0xff) == PQ_DIV) ? (safe ? 0 : 1) : (safe ? 2 : 3);
  return r;
}

template <typename T>
__global__ void __launch_bounds__(TT_THREADS, 3)
transpose_tile_kernel(const T* __restrict__ x, int64_t M, int64_t K, int64_t ldx, int8_t* __restrict__ xq_t, int64_t ldq,
                      const float* __restrict__ s_in, int scale_mode, int vec) {
  constexpr int EPV = VecTraits<T>::EPV;
  constexpr int ESZ = (int)sizeof(T);
  constexpr int PITCH = TT_COLS * ESZ + 16;           // odd multiple of 16 bytes: row-strided 16-byte reads hit distinct banks
  constexpr int WARPS = TT_THREADS / 32, QUADS = TT_ROWS / 4;
  extern __shared__ __align__(16) uint8_t tt_smem[];       // slab [128][PITCH] | tile [128][33] u32 | rowq [128]
  uint8_t* slab = tt_smem;
  uint32_t* tile = reinterpret_cast<uint32_t*>(tt_smem + TT_ROWS * PITCH);
  RowQ* rowq = reinterpret_cast<RowQ*>(tt_smem + TT_ROWS * PITCH + TT_COLS * TT_TP * 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t row0 = (int64_t)blockIdx.y * TT_ROWS, k0 = (int64_t)blockIdx.x * TT_COLS;
  const int rows = (int)((M - row0 < TT_ROWS) ? M - row0 : TT_ROWS);
  const int kw = (int)((K - k0 < TT_COLS) ? K - k0 : TT_COLS);
  ptx::griddep_launch_dependents();
  ptx::griddep_wait();
  if (tid < TT_ROWS) rowq[tid] = rowq_from_scale(tid < rows ? __ldg(s_in + row0 + tid) : 1.0f, scale_mode);
  if (vec) {
    constexpr int VPR = TT_COLS / EPV;                // 16-byte vectors per tile row
    constexpr int PER = TT_ROWS * VPR / TT_THREADS;   // 4 (16-bit) or 8 (fp32) vectors per thread, all requested up front
    const int vw = kw / EPV;
    uint4 raw[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int v = tid + i * TT_THREADS, r = v / VPR, cv = v % VPR;
      raw[i] = make_uint4(0, 0, 0, 0);
      if (r < rows && cv < vw) raw[i] = ld_stream_16(x + (row0 + r) * ldx + k0 + cv * EPV);
    }
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int v = tid + i * TT_THREADS, r = v / VPR, cv = v % VPR;
      *reinterpret_cast<uint4*>(slab + r * PITCH + cv * 16) = raw[i];
    }
  } else {
    for (int e = tid; e < TT_ROWS * TT_COLS; e += TT_THREADS) {
      const int r = e / TT_COLS, cc = e % TT_COLS;
      T v = T(0.f);
      if (r < rows && cc < kw) v = x[(row0 + r) * ldx + k0 + cc];
      *reinterpret_cast<T*>(slab + r * PITCH + cc * ESZ) = v;
    }
  }
  __syncthreads();
  // quantise + transpose: warp task = one row quad x 128 columns; lane -> columns 4 lane .. 4 lane + 3
  for (int quad = warp; quad < QUADS; quad += WARPS) {
    const int kl = 4 * lane;
    float f[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = quad * 4 + j;
      const uint8_t* src = slab + r * PITCH + kl * ESZ;
      if (sizeof(T) == 2) {
        const uint2 w = *reinterpret_cast<const uint2*>(src);
        if (std::is_same<T, __nv_bfloat16>::value) {
          f[j][0] = __uint_as_float(w.x << 16); f[j][1] = __uint_as_float(w.x & 0xffff0000u);
          f[j][2] = __uint_as_float(w.y << 16); f[j][3] = __uint_as_float(w.y & 0xffff0000u);
        } else {
          const float2 a2 = __half22float2(*reinterpret_cast<const __half2*>(&w.x));
          const float2 b2 = __half22float2(*reinterpret_cast<const __half2*>(&w.y));
          f[j][0] = a2.x; f[j][1] = a2.y; f[j][2] = b2.x; f[j][3] = b2.y;
        }
      } else {
        const float4 t4 = *reinterpret_cast<const float4*>(src);
        f[j][0] = t4.x; f[j][1] = t4.y; f[j][2] = t4.z; f[j][3] = t4.w;
      }
      const RowQ rq = rowq[r];
      if (rq.path == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) f[j][q] = quant_fast(f[j][q], rq);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) f[j][q] = quant_any(f[j][q], rq);
      }
    }
    // word for column kl + q: the codes of rows quad*4 .. +3 (little endian = ascending row = ascending address)
#pragma unroll
    for (int q = 0; q < 4; ++q) tile[(kl + q) * TT_TP + quad] = pack4(f[0][q], f[1][q], f[2][q], f[3][q]);
  }
  __syncthreads();
  // write out: output row k0 + kk holds `rows` contiguous bytes at column row0
  const bool full_rows = (rows == TT_ROWS) && (((uintptr_t)(xq_t + row0) & 3) == 0) && (ldq % 4 == 0);
  if (full_rows) {
    if (lane < QUADS) {
      uint32_t* dst = reinterpret_cast<uint32_t*>(xq_t + (k0 + warp) * ldq + row0) + lane;
      const int64_t step = (ldq / 4) * WARPS;
      const uint32_t* src = tile + warp * TT_TP + lane;
      int kk = warp;
      for (; kk + 3 * WARPS < kw; kk += 4 * WARPS) {
        const uint32_t w0 = src[0], w1 = src[WARPS * TT_TP], w2 = src[2 * WARPS * TT_TP], w3 = src[3 * WARPS * TT_TP];
        dst[0] = w0; dst[step] = w1; dst[2 * step] = w2; dst[3 * step] = w3;
        dst += 4 * step; src += 4 * WARPS * TT_TP;
      }
      for (; kk < kw; kk += WARPS) { *dst = *src; dst += step; src += WARPS * TT_TP; }
    }
  } else {
    for (int kk = warp; kk < kw; kk += WARPS) {
      int8_t* dst = xq_t + (k0 + kk) * ldq + row0;
      for (int r = lane; r < rows; r += 32) dst[r] = (int8_t)((tile[kk * TT_TP + (r >> 2)] >> (8 * (r & 3))) & 0xffu);
    }
  }
}

// ---- transposed output, tiled: 32 rows per CTA ---------------------------------------
// Pass 1 computes the 32 row scales (one warp per 4 rows, 16-byte loads when VEC); pass 2
// re-reads the rows (L2 hits: the CTA just streamed them), quantises 32 x 128 tiles into
// shared memory row-major and writes them out transposed: each thread gathers 16 codes of
// one k and stores them with one 16-byte store (32 contiguous bytes per k per CTA).
// VEC requires 16-byte aligned rows of x and of xq_t and K % (16/sizeof(T)) == 0.
template <typename T, bool VEC>
__global__ void __launch_bounds__(256)
rowwise_quant_transposed_kernel(const T* __restrict__ x, int64_t M, int64_t K, int64_t ldx,
                                int8_t* __restrict__ xq_t, int64_t ldq,
                                float* __restrict__ s_out, int scale_mode, float eps) {
  constexpr int EPV = VecTraits<T>::EPV;
  constexpr int PITCH = 144;                         // bytes per staged row (16-byte aligned)
  __shared__ RowQ rowq[32];
  __shared__ __align__(16) int8_t tile[32 * PITCH];
  ptx::griddep_launch_dependents();
  ptx::griddep_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row0 = (int64_t)blockIdx.x * 32;
  for (int r = warp; r < 32; r += 8) {
    const int64_t row = row0 + r;
    float amax = 0.f;
    if (row < M) {
      const T* xr = x + row * ldx;
      if (VEC) {
        const int64_t nvec = K / EPV;
        if (sizeof(T) == 2) {
          uint32_t m = 0;
          for (int64_t v = lane; v < nvec; v += 32) m = absmax_u16x2(ld_stream_16(xr + v * EPV), m);
          amax = u16_mag_to_float<T>(m);
        } else {
          for (int64_t v = lane; v < nvec; v += 32) amax = vec_absmax<float>(ld_stream_16(xr + v * EPV), amax);
        }
      } else {
        for (int64_t k = lane; k < K; k += 32) amax = mag_max(amax, mag_of(load_as_float<T>(xr + k)));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = mag_max(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0) {
      const RowQ rq = make_rowq(amax, scale_mode, eps);
      rowq[r] = rq;
      if (row < M) s_out[row] = rq.s;
    }
  }
  __syncthreads();
  const bool rows_full = row0 + 32 <= M;
  for (int64_t k0 = 0; k0 < K; k0 += 128) {
    // quantise: thread (r = tid/8, group = tid%8) handles 16 consecutive k of row r
    {
      const int r = threadIdx.x >> 3, c0 = (threadIdx.x & 7) * 16;
      const int64_t row = row0 + r;
      const RowQ rq = rowq[r];
      float f[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = 0.f;
      if (row < M) {
        const T* xr = x + row * ldx + k0 + c0;
        if (VEC) {
#pragma unroll
          for (int v = 0; v < 16 / EPV; ++v)
            if (k0 + c0 + (v + 1) * EPV <= K) {
              const uint4 raw = *reinterpret_cast<const uint4*>(xr + v * EPV);
              unpack<T>(raw, f + v * EPV);
            }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (k0 + c0 + j < K) f[j] = load_as_float<T>(xr + j);
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = quant_any(f[j], rq);
      uint4 o;
      o.x = pack4(f[0], f[1], f[2], f[3]);   o.y = pack4(f[4], f[5], f[6], f[7]);
      o.z = pack4(f[8], f[9], f[10], f[11]); o.w = pack4(f[12], f[13], f[14], f[15]);
      *reinterpret_cast<uint4*>(tile + r * PITCH + c0) = o;
    }
    __syncthreads();
    // write: 128 k-rows x 32 bytes; thread -> (k = tid/2, half = tid%2) gathers 16 rows of one k
    {
      const int kk = threadIdx.x >> 1, h = (threadIdx.x & 1) * 16;
      const int64_t k = k0 + kk;
      if (k < K) {
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 16; ++j)
          w[j >> 2] |= (uint32_t)(uint8_t)tile[(h + j) * PITCH + kk] << (8 * (j & 3));
        int8_t* dst = xq_t + k * ldq + row0 + h;
        if (VEC && rows_full) {
          *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (row0 + h + j < M) dst[j] = (int8_t)((w[j >> 2] >> (8 * (j & 3))) & 0xff);
        }
      }
    }
    __syncthreads();
  }
}

template <typename T, int TPR, int VPT>
int launch_vec(const void* x, int64_t M, int nvec, int64_t ldx, int8_t* xq, int64_t ldq,
               float* s, const pq_quant_spec& spec, cudaStream_t st, const void* pf, long long pf_bytes,
               const float* amax_in, int amax_slots, long long amax_stride) {
  constexpr int THREADS = (TPR > 256 ? TPR : 256);
  constexpr int ROWS = THREADS / TPR;
  const int64_t grid = (M + ROWS - 1) / ROWS;
  PQ_CUDA(launch_pdl(rowwise_quant_vec_kernel<T, TPR, VPT>, (unsigned)grid, THREADS, st,
                     (const T*)x, M, nvec, ldx, xq, ldq, s, mode_bits(spec), spec.eps,
                     (const uint8_t*)pf, pf_bytes, amax_in, amax_slots, amax_stride));
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return PQ_OK;
}

template <typename T>
int launch_staged(const void* x, int64_t M, int nvec, int64_t ldx, int8_t* xq, int64_t ldq, float* s,
                  const pq_quant_spec& spec, cudaStream_t st, const void* pf, long long pf_bytes,
                  const float* amax_in, int amax_slots, long long amax_stride, int num_sms, int slots) {
  const int dyn = STG_HEADER + slots * nvec * 16;
  static gemm::PerDeviceOnce once;
  const cudaError_t e = once.run([&](int*) {
    return cudaFuncSetAttribute(rowwise_quant_staged_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  if (e != cudaSuccess) PQ_FAIL(PQ_ERR_CUDA, "cudaFuncSetAttribute(staged quantizer) failed: %s", cudaGetErrorString(e));
  const unsigned grid = (unsigned)(M < num_sms ? M : num_sms);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(STG_THREADS, 1, 1);
  cfg.dynamicSmemBytes = (size_t)dyn;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  PQ_CUDA(cudaLaunchKernelEx(&cfg, rowwise_quant_staged_kernel<T>, (const T*)x, M, nvec, ldx, xq, ldq, s,
                             mode_bits(spec), spec.eps, (const uint8_t*)pf, pf_bytes, amax_in, amax_slots, amax_stride, slots));
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return PQ_OK;
}

template <typename T>
int launch_transposed_two_pass(const void* x, int64_t M, int64_t K, int64_t ldx, int8_t* xq, int64_t ldq, float* s,
                               const pq_quant_spec& spec, cudaStream_t st, bool vec) {
  constexpr int DYN = TT_ROWS * (TT_COLS * (int)sizeof(T) + 16) + TT_COLS * TT_TP * 4 + TT_ROWS * (int)sizeof(RowQ);
  static gemm::PerDeviceOnce once;
  const cudaError_t e = once.run([&](int*) {
    return cudaFuncSetAttribute(transpose_tile_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, DYN);
  });
  if (e != cudaSuccess) PQ_FAIL(PQ_ERR_CUDA, "cudaFuncSetAttribute(transposed quantizer) failed: %s", cudaGetErrorString(e));
  cudaLaunchAttribute attr[1];
  cudaLaunchConfig_t c1 = pdl_config(dim3((unsigned)M), dim3(256), 0, st, attr);
  PQ_CUDA(cudaLaunchKernelEx(&c1, row_scale_kernel<T>, (const T*)x, K, ldx, s, mode_bits(spec), spec.eps, vec ? 1 : 0));
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  const int64_t kblocks = (K + TT_COLS - 1) / TT_COLS, rblocks = (M + TT_ROWS - 1) / TT_ROWS;
  for (int64_t r0 = 0; r0 < rblocks; r0 += 65535) {          // grid.y limit
    const int64_t nr = (rblocks - r0 < 65535) ? rblocks - r0 : 65535;
    cudaLaunchAttribute attr2[1];
    cudaLaunchConfig_t c2 = pdl_config(dim3((unsigned)kblocks, (unsigned)nr), dim3(TT_THREADS), (size_t)DYN, st, attr2);
    const int64_t row_off = r0 * TT_ROWS;
    PQ_CUDA(cudaLaunchKernelEx(&c2, transpose_tile_kernel<T>, (const T*)x + row_off * ldx, M - row_off, K, ldx, xq + row_off, ldq,
                               (const float*)s + row_off, mode_bits(spec), vec ? 1 : 0));
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
  }
  return PQ_OK;
}

template <typename T>
int dispatch(const void* x, int64_t M, int64_t K, int64_t ldx, int8_t* xq, int64_t ldq,
             float* s, int transpose, const pq_quant_spec& spec, cudaStream_t st,
             const void* pf, long long pf_bytes, const float* amax_in, int amax_slots, long long amax_stride) {
  constexpr int EPV = VecTraits<T>::EPV;
  if (M == 0) return PQ_OK;
  if (M > 0x7fffffffLL) PQ_FAIL(PQ_ERR_ARG, "rowwise quant: M=%lld too large", (long long)M);
  if (transpose) {
    if (amax_in) PQ_FAIL(PQ_ERR_UNSUPPORTED, "rowwise quant: an external row maximum cannot be combined with transpose");
    // Two launches (row scales, then 128 x 128 tiles): every scale mode whose row parameters follow from the stored
    // scale alone; PQ_INV_SCALE (needs amax itself) and pq_debug_set_quant_staged(-1) keep the 32-rows-per-CTA kernel.
    if (spec.scale_mode != PQ_INV_SCALE && g_quant_staged >= 0 && K <= 0x7fffffffLL) {
      const bool vec = (K % EPV == 0) && (((uintptr_t)x & 15) == 0) && ((ldx * (int64_t)sizeof(T)) % 16 == 0);
      return launch_transposed_two_pass<T>(x, M, K, ldx, xq, ldq, s, spec, st, vec);
    }
    const int64_t grid = (M + 31) / 32;
    const bool tvec = (K % EPV == 0) && (((uintptr_t)x & 15) == 0) && ((ldx * (int64_t)sizeof(T)) % 16 == 0) &&
                      (((uintptr_t)xq & 15) == 0) && (ldq % 16 == 0);
    if (tvec)
      PQ_CUDA(launch_pdl(rowwise_quant_transposed_kernel<T, true>, (unsigned)grid, 256u, st,
                         (const T*)x, M, K, ldx, xq, ldq, s, mode_bits(spec), spec.eps));
    else
      PQ_CUDA(launch_pdl(rowwise_quant_transposed_kernel<T, false>, (unsigned)grid, 256u, st,
                         (const T*)x, M, K, ldx, xq, ldq, s, mode_bits(spec), spec.eps));
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return PQ_OK;
  }
  const bool vec_ok = (K % EPV == 0) && (((uintptr_t)x & 15) == 0) &&
                      ((ldx * (int64_t)sizeof(T)) % 16 == 0) &&
                      (((uintptr_t)xq % EPV) == 0) && (ldq % EPV == 0) &&
                      (K / EPV <= 8192);
  if (!vec_ok) {
    PQ_CUDA(launch_pdl(rowwise_quant_generic_kernel<T>, (unsigned)M, 256u, st,
                       (const T*)x, M, K, ldx, xq, ldq, s, 0, mode_bits(spec), spec.eps, amax_in, amax_slots, amax_stride));
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return PQ_OK;
  }
  const int nvec = (int)(K / EPV);
  // Activation-sized problems (more rows than one wave of the register kernel can hold, less than a few waves of
  // streaming) go to the persistent shared-memory staged kernel; see its header comment.  Measured in
  // tools/quant_cfg.py -> profiles/quant_staged_r2.log.
  {
    int num_sms = 0;
    const int staged = g_quant_staged;
    if (staged >= 0 && ((uintptr_t)xq % 8 == 0) && (ldq % 8 == 0) && check_device(&num_sms) == PQ_OK) {
      const long long row_bytes = (long long)nvec * 16;
      const long long room = 227 * 1024 - STG_HEADER;
      const long long rows_per_cta = (M + num_sms - 1) / num_sms;
      long long slots = room / row_bytes;
      if (slots > rows_per_cta) slots = rows_per_cta;
      if (slots > STG_MAX_SLOTS) slots = STG_MAX_SLOTS;
      if (slots > STG_GROUPS) slots -= slots % STG_GROUPS;      // see the kernel: a multiple of the active groups
      // measured (profiles/quant_staged_r2.log, bf16): wins for long rows at activation sizes (2048 x 11008: 12.5 ->
      // 11.2 us, 2048 x 8192: 9.9 -> 9.4) and for a few hundred rows (512 x 4096: 3.4 -> 2.5); ties at 2048 x 4096;
      // loses for short rows (tiny bulk copies) and from ~100 MB on, where the register kernel streams at peak.
      const long long bytes = M * row_bytes;
      const bool auto_ok = row_bytes >= 8192 && M >= num_sms &&
                           ((row_bytes >= 16384 && bytes <= (64LL << 20)) || M <= 1024);
      if (slots >= 2 && (staged > 0 || auto_ok))
        return launch_staged<T>(x, M, nvec, ldx, xq, ldq, s, spec, st, pf, pf_bytes, amax_in, amax_slots, amax_stride, num_sms, (int)slots);
    }
  }
  // Pick (threads per row, vectors per thread): cover the row with as few idle lanes as possible,
  // preferring small thread groups (more rows in flight per SM) and <= 6 vectors per thread.
  static const int kVpt[5] = {4, 3, 6, 2, 8};
  int best_tpr = 1024, best_vpt = 8;
  double best_score = 1e30;
  for (int i = 0; i < 5; ++i) {
    const int vpt = kVpt[i];
    int tpr = 32;
    while (tpr < 1024 && tpr * vpt < nvec) tpr *= 2;
    if (tpr * vpt < nvec) continue;
    double score = (double)(tpr * vpt - nvec) / (double)(tpr * vpt);   // idle-lane fraction
    // measured on B200 (tools/sweep_quant.sh): 1024-thread groups lose ~35% (one row per CTA, two
    // CTAs per SM); with many rows, more bytes in flight per thread wins (64x8 beats 128x4 by 5%).
    if (vpt == 2) score += 0.03;
    if (tpr >= 512) score += 0.04;
    if (tpr == 1024) score += 0.20;
    if (M >= 16384) score -= (vpt == 8 ? 0.02 : vpt == 6 ? 0.01 : 0.0);
    else if (vpt == 8) score += 0.03;
    if (score < best_score) { best_score = score; best_tpr = tpr; best_vpt = vpt; }
  }
  if (g_force_tpr > 0 && g_force_vpt > 0 && g_force_tpr * g_force_vpt >= nvec) { best_tpr = g_force_tpr; best_vpt = g_force_vpt; }
#define PQ_CASE_V(TPR, VPT) if (best_tpr == TPR && best_vpt == VPT) return launch_vec<T, TPR, VPT>(x, M, nvec, ldx, xq, ldq, s, spec, st, pf, pf_bytes, amax_in, amax_slots, amax_stride);
#define PQ_CASE_T(TPR) PQ_CASE_V(TPR, 2) PQ_CASE_V(TPR, 3) PQ_CASE_V(TPR, 4) PQ_CASE_V(TPR, 6) PQ_CASE_V(TPR, 8)
  PQ_CASE_T(32) PQ_CASE_T(64) PQ_CASE_T(128) PQ_CASE_T(256) PQ_CASE_T(512) PQ_CASE_T(1024)
#undef PQ_CASE_T
#undef PQ_CASE_V
  PQ_FAIL(PQ_ERR_ARG, "rowwise quant: no kernel configuration for %d vectors per row", nvec);
}

}  // namespace

int launch_rowwise_quant(const void* x, int x_dtype, int64_t M, int64_t K, int64_t ldx,
                         int8_t* xq, int64_t ldq, float* s, int transpose,
                         const pq_quant_spec& spec, cudaStream_t stream,
                         const void* prefetch, long long prefetch_bytes, const float* amax_in,
                         int amax_slots, long long amax_stride) {
  if (((uintptr_t)prefetch & 15) || prefetch_bytes < 16 || !g_weight_prefetch) { prefetch = nullptr; prefetch_bytes = 0; }
  if (M < 0 || K < 1) PQ_FAIL(PQ_ERR_ARG, "rowwise quant: bad shape M=%lld K=%lld", (long long)M, (long long)K);
  if (M > 0 && (!x || !xq || !s)) PQ_FAIL(PQ_ERR_ARG, "rowwise quant: null pointer");
  if (ldx < K) PQ_FAIL(PQ_ERR_ARG, "rowwise quant: ldx=%lld < K=%lld", (long long)ldx, (long long)K);
  if (!transpose && ldq < K) PQ_FAIL(PQ_ERR_ARG, "rowwise quant: ldq=%lld < K=%lld", (long long)ldq, (long long)K);
  if (transpose && ldq < M) PQ_FAIL(PQ_ERR_ARG, "rowwise quant: transposed ldq=%lld < M=%lld", (long long)ldq, (long long)M);
  if (spec.scale_mode < PQ_DIV || spec.scale_mode > PQ_INV_SCALE)
    PQ_FAIL(PQ_ERR_ARG, "rowwise quant: bad scale_mode %d", spec.scale_mode);
  if (spec.qmin != -128 && spec.qmin != -127)
    PQ_FAIL(PQ_ERR_ARG, "rowwise quant: qmin must be -128 or -127");
  switch (x_dtype) {
    case PQ_F32: return dispatch<float>(x, M, K, ldx, xq, ldq, s, transpose, spec, stream, prefetch, prefetch_bytes, amax_in, amax_slots, amax_stride);
    case PQ_F16: return dispatch<__half>(x, M, K, ldx, xq, ldq, s, transpose, spec, stream, prefetch, prefetch_bytes, amax_in, amax_slots, amax_stride);
    case PQ_BF16: return dispatch<__nv_bfloat16>(x, M, K, ldx, xq, ldq, s, transpose, spec, stream, prefetch, prefetch_bytes, amax_in, amax_slots, amax_stride);
    default: PQ_FAIL(PQ_ERR_ARG, "rowwise quant: unsupported dtype %d", x_dtype);
  }
}

}  // namespace pq

// Test/bench hook: force (threads per row, 16-byte vectors per thread) of the vectorised kernel; 0,0 = heuristic.
extern "C" void pq_debug_set_quant_config(int tpr, int vpt) { pq::g_force_tpr = tpr; pq::g_force_vpt = vpt; }
extern "C" void pq_debug_set_weight_prefetch(int on) { pq::g_weight_prefetch = on; }
// 0 = heuristic, 1 = force the shared-memory staged quantizer wherever it is applicable, -1 = never use it
extern "C" void pq_debug_set_quant_staged(int mode) { pq::g_quant_staged = mode; }
