// QTensor.dequantize():  out[r,c] = float(q[r,c]) * s[axis==0 ? r : c]   
// One fp32 multiply, then a single RNE cast to the output dtype.  HBM-bound:
// cols + cols*sizeof(out) bytes per row.
#include "common.cuh"

namespace pq {
namespace {

template <typename O> __device__ __forceinline__ O cast_out(float f);
template <> __device__ __forceinline__ float cast_out<float>(float f) { return f; }
template <> __device__ __forceinline__ __half cast_out<__half>(float f) { return __float2half_rn(f); }
template <> __device__ __forceinline__ __nv_bfloat16 cast_out<__nv_bfloat16>(float f) { return __float2bfloat16_rn(f); }

// 16 codes per thread when everything is 16-byte aligned, scalar otherwise.
template <typename O, bool VEC>
__global__ void __launch_bounds__(256)
dequant_kernel(const int8_t* __restrict__ q, int64_t ldq, const float* __restrict__ s, int axis,
               O* __restrict__ out, int64_t ldo, int64_t rows, int64_t cols) {
  const int64_t r = blockIdx.y;
  const float* sp = s;
  const float srow = (axis == 0) ? s[r] : 0.f;
  if (VEC) {
    const int64_t c0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 16;
    if (c0 >= cols) return;
    const uint4 v = *reinterpret_cast<const uint4*>(q + r * ldq + c0);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    O o[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int8_t code = (int8_t)((w[i >> 2] >> (8 * (i & 3))) & 0xff);
      const float sc = (axis == 0) ? srow : __ldg(sp + c0 + i);
      o[i] = cast_out<O>(__fmul_rn((float)code, sc));
    }
    uint4* dst = reinterpret_cast<uint4*>(out + r * ldo + c0);
    const uint4* src = reinterpret_cast<const uint4*>(o);
#pragma unroll
    for (int i = 0; i < (int)(16 * sizeof(O) / 16); ++i) dst[i] = src[i];
  } else {
    for (int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x; c < cols; c += (int64_t)gridDim.x * 256) {
      const float sc = (axis == 0) ? srow : sp[c];
      out[r * ldo + c] = cast_out<O>(__fmul_rn((float)q[r * ldq + c], sc));
    }
  }
}

template <typename O>
int launch(const int8_t* q, int64_t ldq, const float* s, int axis, void* out, int64_t ldo,
           int64_t rows, int64_t cols, cudaStream_t st) {
  const bool vec = (cols % 16 == 0) && (ldq % 16 == 0) && (((uintptr_t)q & 15) == 0) &&
                   (((uintptr_t)out & 15) == 0) && ((ldo * (int64_t)sizeof(O)) % 16 == 0);
  for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
    const int64_t nr = (rows - r0 < 65535) ? rows - r0 : 65535;
    const int8_t* qp = q + r0 * ldq;
    O* op = (O*)out + r0 * ldo;
    const float* sp = (axis == 0) ? s + r0 : s;
    if (vec) {
      dim3 grid((unsigned)((cols / 16 + 255) / 256), (unsigned)nr);
      dequant_kernel<O, true><<<grid, 256, 0, st>>>(qp, ldq, sp, axis, op, ldo, nr, cols);
    } else {
      int64_t gx = (cols + 255) / 256;
      if (gx > 64) gx = 64;
      dim3 grid((unsigned)gx, (unsigned)nr);
      dequant_kernel<O, false><<<grid, 256, 0, st>>>(qp, ldq, sp, axis, op, ldo, nr, cols);
    }
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    PQ_CUDA(cudaGetLastError());
  }
  return PQ_OK;
}

}  // namespace
}  // namespace pq

extern "C" int pq_dequant(const int8_t* q, int64_t ldq, const float* s, int axis,
                          void* out, int out_dtype, int64_t ldo,
                          int64_t rows, int64_t cols, void* stream) {
  using namespace pq;
  if (rows < 0 || cols < 0) PQ_FAIL(PQ_ERR_ARG, "pq_dequant: bad shape");
  if (rows == 0 || cols == 0) return PQ_OK;
  if (!q || !s || !out) PQ_FAIL(PQ_ERR_ARG, "pq_dequant: null pointer");
  if (axis != 0 && axis != 1) PQ_FAIL(PQ_ERR_ARG, "pq_dequant: axis must be 0 or 1");
  if (ldq < cols || ldo < cols) PQ_FAIL(PQ_ERR_ARG, "pq_dequant: leading dimension < cols");
  int rc = check_device(nullptr);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  switch (out_dtype) {
    case PQ_F32: return launch<float>(q, ldq, s, axis, out, ldo, rows, cols, st);
    case PQ_F16: return launch<__half>(q, ldq, s, axis, out, ldo, rows, cols, st);
    case PQ_BF16: return launch<__nv_bfloat16>(q, ldq, s, axis, out, ldo, rows, cols, st);
    default: PQ_FAIL(PQ_ERR_ARG, "pq_dequant: unsupported out dtype %d", out_dtype);
  }
}
