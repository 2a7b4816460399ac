// Pieces shared by the tcgen05 GEMM kernels: UMMA descriptors, output packing, TMA tensor maps.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <mutex>
#include <type_traits>

namespace pq {
namespace gemm {

constexpr int BLOCK_K = 128;   // bytes (= int8 elements) per ring slot: one 128B swizzle atom
constexpr int UMMA_K = 32;     // int8 elements per tcgen05.mma

// ---- descriptors -----------------------------------------------------------------
// Shared-memory matrix descriptor, K-major, 128-byte swizzle: rows are 128 B apart,
// 8-row core groups are 1024 B apart (SBO); LBO is unused for swizzled K-major.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);        // start address  [0,14)
  d |= (uint64_t)0 << 16;                          // leading byte offset [16,30)
  d |= (uint64_t)(1024u >> 4) << 32;               // stride byte offset  [32,46)
  d |= (uint64_t)1 << 46;                          // descriptor version 1 (sm_100)
  d |= (uint64_t)2 << 61;                          // layout: SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::i8: D = S32, A = B = signed 8-bit, both K-major, no
// saturation (accumulators must match an exact int32 reference).
__host__ __device__ constexpr uint32_t make_idesc(int umma_m, int umma_n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(umma_n >> 3) << 17) |
         ((uint32_t)(umma_m >> 4) << 24);
}

template <typename OutT> struct OutPack;
template <> struct OutPack<__nv_bfloat16> {
  static constexpr int WORDS = 16;
  __device__ static __forceinline__ void pack(const float* f, uint32_t* o) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      o[i] = *reinterpret_cast<uint32_t*>(&h);
    }
  }
  __device__ static __forceinline__ __nv_bfloat16 one(float f) { return __float2bfloat16_rn(f); }
};
template <> struct OutPack<__half> {
  static constexpr int WORDS = 16;
  __device__ static __forceinline__ void pack(const float* f, uint32_t* o) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      o[i] = *reinterpret_cast<uint32_t*>(&h);
    }
  }
  __device__ static __forceinline__ __half one(float f) { return __float2half_rn(f); }
};
template <> struct OutPack<float> {
  static constexpr int WORDS = 32;
  __device__ static __forceinline__ void pack(const float* f, uint32_t* o) {
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(f[i]);
  }
  __device__ static __forceinline__ float one(float f) { return f; }
};

// ---- host side ---------------------------------------------------------------------
inline PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

// [rows, kbytes] int8 matrix, row stride `ld` bytes -> boxes of [box_rows x 128 B], 128B swizzle
inline int make_tmap(CUtensorMap* m, const void* base, int64_t rows, int64_t kbytes, int64_t ld, int box_rows) {
  auto fn = get_encode_fn();
  if (!fn) PQ_FAIL(PQ_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)kbytes, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) PQ_FAIL(PQ_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return PQ_OK;
}

}  // namespace gemm
}  // namespace pq
