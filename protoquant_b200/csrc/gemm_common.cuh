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

// ---- tensor-map cache ------------------------------------------------------------------
// A CUtensorMap is a pure function of (base address, dims, stride, box, dtype, swizzle), so encoded maps are kept
// in a small direct-mapped, mutex-protected table (per process; the address is a unified VA, so entries are valid
// on whichever device owns the buffer).  Saves 2-10 cuTensorMapEncodeTiled calls per launch on eager paths.
struct TmapKey {
  const void* base;
  uint64_t d0, d1, stride;
  uint32_t b0, b1, dtype, swizzle, promo;
  bool operator==(const TmapKey& o) const {
    return base == o.base && d0 == o.d0 && d1 == o.d1 && stride == o.stride && b0 == o.b0 && b1 == o.b1 &&
           dtype == o.dtype && swizzle == o.swizzle && promo == o.promo;
  }
};
struct TmapCache {
  static constexpr int SLOTS = 1024;
  std::mutex mu;
  TmapKey keys[SLOTS];
  CUtensorMap maps[SLOTS];
  bool used[SLOTS] = {};
  static size_t hash(const TmapKey& k) {
    uint64_t h = (uint64_t)(uintptr_t)k.base * 0x9E3779B97F4A7C15ull;
    h ^= (k.d0 + 0x7F4A7C15ull) * 0xBF58476D1CE4E5B9ull;
    h ^= (k.d1 + 0x1CE4E5B9ull) * 0x94D049BB133111EBull;
    h ^= (k.stride << 7) ^ ((uint64_t)k.b0 << 20) ^ ((uint64_t)k.b1 << 32) ^ ((uint64_t)k.dtype << 44) ^
         ((uint64_t)k.swizzle << 50) ^ ((uint64_t)k.promo << 56);
    h ^= h >> 29;
    return (size_t)(h % SLOTS);
  }
};
inline TmapCache& tmap_cache() {
  static TmapCache c;
  return c;
}

// 2-D tiled tensor map over a row-major matrix: dims {d0 (inner, elements), d1 (rows)}, row stride `stride` BYTES,
// box {b0, b1} elements.  Looks the map up in the cache first.
inline int encode_tmap_2d(CUtensorMap* m, const void* base, CUtensorMapDataType dt, uint64_t d0, uint64_t d1,
                          uint64_t stride, uint32_t b0, uint32_t b1, CUtensorMapSwizzle sw,
                          CUtensorMapL2promotion promo) {
  TmapKey key = {base, d0, d1, stride, b0, b1, (uint32_t)dt, (uint32_t)sw, (uint32_t)promo};
  TmapCache& c = tmap_cache();
  const size_t slot = TmapCache::hash(key);
  {
    std::lock_guard<std::mutex> lk(c.mu);
    if (c.used[slot] && c.keys[slot] == key) {
      *m = c.maps[slot];
      return PQ_OK;
    }
  }
  auto fn = get_encode_fn();
  if (!fn) PQ_FAIL(PQ_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)d0, (cuuint64_t)d1};
  cuuint64_t strides[1] = {(cuuint64_t)stride};
  cuuint32_t box[2] = {(cuuint32_t)b0, (cuuint32_t)b1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, promo,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) PQ_FAIL(PQ_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  {
    std::lock_guard<std::mutex> lk(c.mu);
    c.keys[slot] = key;
    c.maps[slot] = *m;
    c.used[slot] = true;
  }
  return PQ_OK;
}

// [rows, kbytes] int8 matrix, row stride `ld` bytes -> boxes of [box_rows x 128 B], 128B swizzle
inline int make_tmap(CUtensorMap* m, const void* base, int64_t rows, int64_t kbytes, int64_t ld, int box_rows) {
  return encode_tmap_2d(m, base, CU_TENSOR_MAP_DATA_TYPE_UINT8, (uint64_t)kbytes, (uint64_t)rows, (uint64_t)ld,
                        (uint32_t)BLOCK_K, (uint32_t)box_rows, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
}

// ---- per-device one-time kernel setup ------------------------------------------------------
// Function attributes (the opt-in to > 48 KB of dynamic shared memory, cluster occupancy) belong to a device's
// context, so a process that drives several GPUs must set them once per DEVICE, not once per process.
struct PerDeviceOnce {
  std::mutex mu;
  int state[64] = {};          // 0 = not run on this device yet, 1 = done
  cudaError_t err[64] = {};
  int value[64] = {};          // optional per-device result (e.g. co-resident clusters)
  template <typename F>
  cudaError_t run(F&& f, int* value_out = nullptr) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lk(mu);
    if (!state[dev]) {
      err[dev] = f(&value[dev]);
      state[dev] = 1;
    }
    if (value_out) *value_out = value[dev];
    return err[dev];
  }
};

}  // namespace gemm
}  // namespace pq
