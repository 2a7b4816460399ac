// Pieces of the row-parallel (K-split) dynamic-quant linear:
//
//   pq_row_absmax      amax[m] = max_k |x[m,k]|  (fp32) -- the per-token scale needs the maximum over the
//                      WHOLE row, so K-shards take max over their slices and all-reduce(MAX) M floats;
//   pq_reduce_dequant  y[m,n] = cast(((float(sum_p part_p[m,n]) * s_x[m]) * s_w[n]) + bias[n]) written to one or
//                      more destinations: the second half of the fused GEMM + reduce-scatter (the first half is
//                      pq_qgemm_i32_scatter, whose epilogue stores each rank's exact int32 partial sums straight
//                      into the owner rank's inbox over NVLink).  int32 addition is associative, so the result is
//                      bit-identical to the unsharded GEMM for any number of K-shards.
// Both are HBM-bound streaming kernels: K*sizeof(T) bytes per row, and (4*P + sizeof(out)*D) bytes per element.
#include "common.cuh"
#include "ptx.cuh"
#include "quant_math.cuh"

namespace pq {
namespace {

using namespace qmath;

struct AmaxDst { float* p[8]; int n; };   // the row maxima are stored to every p[d][row] (exchange slots of the peers)

// one CTA per row, 16-byte loads when VEC
template <typename T, bool VEC>
__global__ void __launch_bounds__(256)
row_absmax_kernel(const T* __restrict__ x, int64_t K, int64_t ldx, const AmaxDst dst) {
  constexpr int EPV = VecTraits<T>::EPV;
  __shared__ float red[8];
  ptx::griddep_launch_dependents();
  ptx::griddep_wait();
  const T* xr = x + (int64_t)blockIdx.x * ldx;
  float amax = 0.f;
  if (VEC) {
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
    for (int d = 0; d < dst.n; ++d) dst.p[d][blockIdx.x] = amax;
  }
}

template <typename T>
int launch_absmax(const void* x, int64_t M, int64_t K, int64_t ldx, const AmaxDst& dst, cudaStream_t st) {
  constexpr int EPV = VecTraits<T>::EPV;
  const bool vec = (K % EPV == 0) && (((uintptr_t)x & 15) == 0) && ((ldx * (int64_t)sizeof(T)) % 16 == 0);
  cudaLaunchAttribute attr[1];
  cudaLaunchConfig_t cfg = pdl_config(dim3((unsigned)M), dim3(256), 0, st, attr);
  if (vec) PQ_CUDA(cudaLaunchKernelEx(&cfg, row_absmax_kernel<T, true>, (const T*)x, K, ldx, dst));
  else PQ_CUDA(cudaLaunchKernelEx(&cfg, row_absmax_kernel<T, false>, (const T*)x, K, ldx, dst));
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return PQ_OK;
}

struct ReduceArgs {
  const int32_t* parts[8];
  void* ys[8];
  const float* s_x; const float* s_w; const float* bias;
  long long ld_part, ldy, M, N;
  int n_parts, n_ys, vec;
};

template <typename O> struct Pack4;
template <> struct Pack4<float> {
  static __device__ __forceinline__ void store(float* dst, const float* f) { *reinterpret_cast<float4*>(dst) = make_float4(f[0], f[1], f[2], f[3]); }
  static __device__ __forceinline__ float one(float f) { return f; }
};
template <> struct Pack4<__nv_bfloat16> {
  static __device__ __forceinline__ void store(__nv_bfloat16* dst, const float* f) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
    *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
  }
  static __device__ __forceinline__ __nv_bfloat16 one(float f) { return __float2bfloat16_rn(f); }
};
template <> struct Pack4<__half> {
  static __device__ __forceinline__ void store(__half* dst, const float* f) {
    const __half2 a = __floats2half2_rn(f[0], f[1]), b = __floats2half2_rn(f[2], f[3]);
    *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
  }
  static __device__ __forceinline__ __half one(float f) { return __float2half_rn(f); }
};

// thread -> 4 consecutive columns of one row; grid.y = row, grid.x covers N/4
template <typename O>
__global__ void __launch_bounds__(256)
reduce_dequant_kernel(const ReduceArgs a) {
  ptx::griddep_launch_dependents();
  ptx::griddep_wait();
  const long long m = blockIdx.y;
  const long long n0 = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
  if (n0 >= a.N) return;
  const float sx = __ldg(a.s_x + m);
  int acc[4] = {0, 0, 0, 0};
  const bool full = a.vec && n0 + 4 <= a.N;
  if (full) {
    for (int p = 0; p < a.n_parts; ++p) {
      const int4 v = __ldcg(reinterpret_cast<const int4*>(a.parts[p] + m * a.ld_part + n0));
      acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
    }
  } else {
    for (int p = 0; p < a.n_parts; ++p)
      for (int j = 0; j < 4; ++j)
        if (n0 + j < a.N) acc[j] += __ldcg(a.parts[p] + m * a.ld_part + n0 + j);
  }
  float f[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long n = (n0 + j < a.N) ? n0 + j : a.N - 1;
    float v = __int2float_rn(acc[j]);
    v = __fmul_rn(v, sx);
    v = __fmul_rn(v, __ldg(a.s_w + n));
    if (a.bias != nullptr) v = __fadd_rn(v, __ldg(a.bias + n));
    f[j] = v;
  }
  for (int d = 0; d < a.n_ys; ++d) {
    O* dst = reinterpret_cast<O*>(a.ys[d]) + m * a.ldy + n0;
    if (full) Pack4<O>::store(dst, f);
    else
      for (int j = 0; j < 4; ++j)
        if (n0 + j < a.N) dst[j] = Pack4<O>::one(f[j]);
  }
}

template <typename O>
int launch_reduce(const ReduceArgs& a, cudaStream_t st) {
  for (long long r0 = 0; r0 < a.M; r0 += 65535) {
    ReduceArgs b = a;
    const long long nr = (a.M - r0 < 65535) ? a.M - r0 : 65535;
    for (int p = 0; p < a.n_parts; ++p) b.parts[p] = a.parts[p] + r0 * a.ld_part;
    for (int d = 0; d < a.n_ys; ++d) b.ys[d] = reinterpret_cast<O*>(a.ys[d]) + r0 * a.ldy;
    b.s_x = a.s_x + r0;
    dim3 grid((unsigned)((a.N + 1023) / 1024), (unsigned)nr);
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg = pdl_config(grid, dim3(256), 0, st, attr);
    PQ_CUDA(cudaLaunchKernelEx(&cfg, reduce_dequant_kernel<O>, b));
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
  }
  return PQ_OK;
}

}  // namespace
}  // namespace pq

using namespace pq;

int pq::launch_row_absmax(const void* x, int x_dtype, int64_t M, int64_t K, int64_t ldx, float* const* dsts, int n_dst,
                          cudaStream_t st) {
  int rc = check_device(nullptr);
  if (rc) return rc;
  if (M < 0 || K < 1 || ldx < K) PQ_FAIL(PQ_ERR_ARG, "pq_row_absmax: bad shape M=%lld K=%lld ldx=%lld", (long long)M, (long long)K, (long long)ldx);
  if (M == 0) return PQ_OK;
  if (!x || !dsts || n_dst < 1 || n_dst > 8) PQ_FAIL(PQ_ERR_ARG, "pq_row_absmax: null pointer or bad destination count");
  if (M > 0x7fffffffLL) PQ_FAIL(PQ_ERR_ARG, "pq_row_absmax: M too large");
  AmaxDst dst = {};
  dst.n = n_dst;
  for (int d = 0; d < n_dst; ++d) {
    if (!dsts[d]) PQ_FAIL(PQ_ERR_ARG, "pq_row_absmax: null destination %d", d);
    dst.p[d] = dsts[d];
  }
  switch (x_dtype) {
    case PQ_F32: return launch_absmax<float>(x, M, K, ldx, dst, st);
    case PQ_F16: return launch_absmax<__half>(x, M, K, ldx, dst, st);
    case PQ_BF16: return launch_absmax<__nv_bfloat16>(x, M, K, ldx, dst, st);
    default: PQ_FAIL(PQ_ERR_ARG, "pq_row_absmax: unsupported dtype %d", x_dtype);
  }
}

extern "C" int pq_row_absmax(const void* x, int x_dtype, int64_t M, int64_t K, int64_t ldx, float* amax, void* stream) {
  float* one[1] = {amax};
  if (!amax && M > 0) PQ_FAIL(PQ_ERR_ARG, "pq_row_absmax: null pointer");
  return launch_row_absmax(x, x_dtype, M, K, ldx, one, 1, (cudaStream_t)stream);
}

extern "C" int pq_act_quant_amax(const void* x, int x_dtype, int64_t M, int64_t K, int64_t ldx, const float* amax,
                                 int8_t* xq, int64_t ldq, float* s_x, const pq_quant_spec* spec, void* stream) {
  int rc = check_device(nullptr);
  if (rc) return rc;
  if (!amax && M > 0) PQ_FAIL(PQ_ERR_ARG, "pq_act_quant_amax: null amax");
  return launch_rowwise_quant(x, x_dtype, M, K, ldx, xq, ldq, s_x, 0, resolve_spec(spec), (cudaStream_t)stream,
                              nullptr, 0, amax);
}

extern "C" int pq_qgemm_i32_scatter(const int8_t* xq, int64_t lda, const int8_t* Wq, int64_t ldb,
                                    void* const* dests, int n_dests, int64_t ld_dest, int64_t cols_per_dest,
                                    int64_t M, int64_t N, int64_t K, void* stream) {
  if (cols_per_dest < 1) PQ_FAIL(PQ_ERR_ARG, "pq_qgemm_i32_scatter: cols_per_dest must be positive");
  return launch_qgemm(xq, lda, Wq, ldb, nullptr, nullptr, nullptr, dests, n_dests, PQ_I32, ld_dest, M, N, K,
                      (cudaStream_t)stream, cols_per_dest);
}

extern "C" int pq_reduce_dequant(const int32_t* const* parts, int n_parts, int64_t ld_part,
                                 const float* s_x, const float* s_w, const float* bias,
                                 void* const* ys, int n_ys, int y_dtype, int64_t ldy,
                                 int64_t M, int64_t N, void* stream) {
  return launch_reduce_dequant(parts, n_parts, ld_part, s_x, s_w, bias, ys, n_ys, y_dtype, ldy, M, N, (cudaStream_t)stream);
}

int pq::launch_reduce_dequant(const int32_t* const* parts, int n_parts, int64_t ld_part,
                              const float* s_x, const float* s_w, const float* bias,
                              void* const* ys, int n_ys, int y_dtype, int64_t ldy,
                              int64_t M, int64_t N, cudaStream_t st) {
  int rc = check_device(nullptr);
  if (rc) return rc;
  if (M < 0 || N < 0) PQ_FAIL(PQ_ERR_ARG, "pq_reduce_dequant: bad shape");
  if (M == 0 || N == 0) return PQ_OK;
  if (!parts || !ys || n_parts < 1 || n_parts > 8 || n_ys < 1 || n_ys > 8 || !s_x || !s_w)
    PQ_FAIL(PQ_ERR_ARG, "pq_reduce_dequant: null pointer or bad part / destination count (1..8)");
  if (ld_part < N || ldy < N) PQ_FAIL(PQ_ERR_ARG, "pq_reduce_dequant: leading dimension smaller than N");
  const int esz = dtype_size(y_dtype);
  if (esz == 0 || y_dtype == PQ_I32) PQ_FAIL(PQ_ERR_ARG, "pq_reduce_dequant: y_dtype must be PQ_BF16, PQ_F16 or PQ_F32");
  ReduceArgs a = {};
  a.vec = (ld_part % 4 == 0) && ((ldy * esz) % 16 == 0 || (esz == 2 && (ldy * esz) % 8 == 0));
  for (int p = 0; p < n_parts; ++p) {
    if (!parts[p]) PQ_FAIL(PQ_ERR_ARG, "pq_reduce_dequant: null part %d", p);
    a.parts[p] = parts[p];
    if ((uintptr_t)parts[p] & 15) a.vec = 0;
  }
  for (int d = 0; d < n_ys; ++d) {
    if (!ys[d]) PQ_FAIL(PQ_ERR_ARG, "pq_reduce_dequant: null destination %d", d);
    a.ys[d] = ys[d];
    if ((uintptr_t)ys[d] & (esz == 4 ? 15 : 7)) a.vec = 0;
  }
  a.s_x = s_x; a.s_w = s_w; a.bias = bias;
  a.ld_part = ld_part; a.ldy = ldy; a.M = M; a.N = N; a.n_parts = n_parts; a.n_ys = n_ys;
  switch (y_dtype) {
    case PQ_F32: return launch_reduce<float>(a, st);
    case PQ_F16: return launch_reduce<__half>(a, st);
    default: return launch_reduce<__nv_bfloat16>(a, st);
  }
}
