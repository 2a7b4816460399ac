// NVLink store-bandwidth probe (measurement tool, not part of the product library).
// Every rank pushes its column slice [rows x seg_bytes] (row stride ld_bytes) of a gathered output from local memory
// into the same place of every peer's buffer -- the traffic pattern of the fused all-gather epilogue -- with
//   mode 0: LSU 16-byte stores, one (row, destination) item per warp
//   mode 1: multimem.st.v4 to the NVSwitch multicast address, one row per warp
//   mode 2: 1-D bulk copies (cp.async.bulk shared -> global) issued by one thread per CTA from a shared-memory ring
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC -o tools/libnvlink_probe.so tools/nvlink_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>

struct ProbeArgs {
  const uint8_t* src;
  uint8_t* dst[8];
  int n_dst;
  int rows;
  int seg_bytes;      // multiple of 512
  long long ld_bytes;
};

__device__ __forceinline__ uint4 ldg16(const void* p) {
  uint4 r;
  asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg16(void* p, uint4 v) {
  asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void mmst16(void* p, uint4 v) {
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(1024) probe_lsu(const ProbeArgs a) {
  const int warps_per_cta = blockDim.x >> 5;
  const long long gw = (long long)blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
  const long long nw = (long long)gridDim.x * warps_per_cta;
  const int lane = threadIdx.x & 31;
  const int nd = MODE == 1 ? 1 : a.n_dst;
  const long long items = (long long)a.rows * nd;
  const int iters = a.seg_bytes / 512;
  for (long long it = gw; it < items; it += nw) {
    const int row = (int)(it / nd), d = (int)(it % nd);
    const uint8_t* s = a.src + (long long)row * a.ld_bytes + lane * 16;
    uint8_t* t = a.dst[d] + (long long)row * a.ld_bytes + lane * 16;
    for (int i0 = 0; i0 < iters; i0 += 7) {
      uint4 v[7];
#pragma unroll
      for (int j = 0; j < 7; ++j)
        if (i0 + j < iters) v[j] = ldg16(s + (i0 + j) * 512);
#pragma unroll
      for (int j = 0; j < 7; ++j)
        if (i0 + j < iters) {
          if (MODE == 1) mmst16(t + (i0 + j) * 512, v[j]);
          else stg16(t + (i0 + j) * 512, v[j]);
        }
    }
  }
}

// mode 2: one thread per CTA drives bulk copies through a ring of SLOTS shared-memory buffers
constexpr int SLOTS = 8;
__global__ void __launch_bounds__(128) probe_bulk(const ProbeArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[SLOTS];
  if (threadIdx.x != 0) return;
  for (int s = 0; s < SLOTS; ++s) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bars + s);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const uint32_t seg = (uint32_t)a.seg_bytes;
  long long n = 0;
  for (long long row = blockIdx.x; row < a.rows; row += gridDim.x, ++n) {
    const int s = (int)(n % SLOTS);
    const uint32_t par = (uint32_t)((n / SLOTS) & 1);
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(bars + s);
    const uint32_t buf = (uint32_t)__cvta_generic_to_shared(smem + (size_t)s * seg);
    if (n >= SLOTS) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(SLOTS - 1) : "memory");   // slot's stores have read it
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(seg) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(buf),
                 "l"(a.src + row * a.ld_bytes), "r"(seg), "r"(bar)
                 : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar), "r"(par) : "memory");
    for (int d = 0; d < a.n_dst; ++d)
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a.dst[d] + row * a.ld_bytes), "r"(buf), "r"(seg)
                   : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

extern "C" int probe_launch(int mode, const void* src, void* const* dst, int n_dst, int rows, int seg_bytes, long long ld_bytes,
                            int ctas, int threads, void* stream) {
  ProbeArgs a;
  a.src = (const uint8_t*)src;
  for (int i = 0; i < 8; ++i) a.dst[i] = i < n_dst ? (uint8_t*)dst[i] : nullptr;
  a.n_dst = n_dst; a.rows = rows; a.seg_bytes = seg_bytes; a.ld_bytes = ld_bytes;
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0) probe_lsu<0><<<ctas, threads, 0, st>>>(a);
  else if (mode == 1) probe_lsu<1><<<ctas, threads, 0, st>>>(a);
  else {
    const int dyn = SLOTS * seg_bytes;
    cudaFuncSetAttribute(probe_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    probe_bulk<<<ctas, 128, dyn, st>>>(a);
  }
  return (int)cudaGetLastError();
}
