// int8 x int8 -> int32 GEMM on tcgen05 with the fused dequant epilogue
// (SURVEY.md §8 rows a3 + a4).
//
//   acc[m,n] = sum_k xq[m,k] * Wq[n,k]                       (exact int32)
//   y[m,n]   = cast( (float(acc) * s_x[m]) * s_w[n] + bias[n] )   (fp32, one RNE cast)
//
// Both operands are K-major (xq [M,K], Wq [N,K]), which is the native layout of
// `tcgen05.mma.kind::i8` with K-major shared-memory descriptors.
//
// Structure (one persistent CTA -- or CTA pair -- per SM, 256 threads):
//   warp 0   TMA producer: 128-byte-swizzled [rows x 128 B] boxes of A and B into a
//            STAGES-deep shared-memory ring, completion on `full` mbarriers.
//   warp 1   MMA issuer (one thread): 4 x tcgen05.mma (K=32 each) per ring slot into a
//            TMEM accumulator; tcgen05.commit releases the slot (`empty`) and, after the
//            last K block, publishes the accumulator (`tmem_full`).
//   warp 2   TMEM allocate / free (2 accumulator buffers of BLOCK_N columns).
//   warps 4-7 epilogue: tcgen05.ld 32 lanes x 32 columns, scale/bias/cast, 16-byte
//            global stores; then hands the accumulator back (`tmem_empty`) so the MMA
//            warp is already filling the other buffer meanwhile.
// CG == 2 runs the same roles on a 2-CTA cluster with `cta_group::2`: the pair shares a
// 256 x BLOCK_N tile, each CTA stages its own 128 rows of A and its half of B, the even
// CTA issues the MMAs for both and multicasts the commits.
#include "common.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <mutex>
#include <type_traits>

namespace pq {
namespace {

using namespace ptx;

constexpr int BLOCK_M = 128;   // rows of A per CTA (= TMEM lanes)
constexpr int BLOCK_K = 128;   // bytes (= int8 elements) per ring slot: one 128B swizzle atom
constexpr int UMMA_K = 32;     // int8 elements per tcgen05.mma
constexpr int NUM_THREADS = 256;
constexpr int EPI_WARP0 = 4;
constexpr int EPI_THREADS = 128;
constexpr int GROUP_M = 16;    // tile rasterisation: m-blocks per L2 swizzle group

struct GemmArgs {
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  const float* s_x;
  const float* s_w;
  const float* bias;
  void* out;
  long long ldo;  // elements
  int vec_ok;     // rows of `out` are 16-byte aligned
};

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

template <int CG, int BN, int STAGES>
struct SmemLayout {
  static constexpr int A_STAGE = BLOCK_M * BLOCK_K;
  static constexpr int B_ROWS = BN / CG;
  static constexpr int B_STAGE = B_ROWS * BLOCK_K;
  static constexpr int STAGE_BYTES = A_STAGE + B_STAGE;
  static constexpr int OFF_A = 0;
  static constexpr int OFF_B = OFF_A + STAGES * A_STAGE;
  static constexpr int OFF_SW = OFF_B + STAGES * B_STAGE;   // [2][BN] fp32
  static constexpr int OFF_BIAS = OFF_SW + 2 * BN * 4;      // [2][BN] fp32
  static constexpr int OFF_BAR = OFF_BIAS + 2 * BN * 4;     // full[S], empty[S], tfull[2], tempty[2]
  static constexpr int NUM_BARS = 2 * STAGES + 4;
  static constexpr int OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
  static constexpr int TOTAL = OFF_TMEM_PTR + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024-B alignment
  static_assert(DYN_BYTES <= 227 * 1024, "shared memory budget exceeded");
};

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

__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int& m_blk, int& n_blk) {
  const int per_group = GROUP_M * num_n;
  const int group = tile / per_group;
  const int first_m = group * GROUP_M;
  const int gsize = min(GROUP_M, num_m - first_m);
  const int within = tile - group * per_group;
  m_blk = first_m + within % gsize;
  n_blk = within / gsize;
}

template <int CG, int BN, int STAGES, typename OutT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
qgemm_kernel(const __grid_constant__ CUtensorMap tmap_a,
             const __grid_constant__ CUtensorMap tmap_b, const GemmArgs g) {
  using L = SmemLayout<CG, BN, STAGES>;
  constexpr bool RAW = std::is_same<OutT, int32_t>::value;
  constexpr int UMMA_M = BLOCK_M * CG;
  constexpr int UMMA_N = BN;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                 : (2 * BN <= 256) ? 256 : 512;
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BLOCK_N must be a multiple of 32 in [32,256]");
  static_assert(UMMA_N % 16 == 0, "invalid UMMA N");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = (cta_rank == 0);

  const uint32_t bar_full = smem_base + L::OFF_BAR;
  const uint32_t bar_empty = bar_full + STAGES * 8;
  const uint32_t bar_tfull = bar_empty + STAGES * 8;
  const uint32_t bar_tempty = bar_tfull + 2 * 8;
  const uint32_t tmem_ptr_smem = smem_base + L::OFF_TMEM_PTR;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + L::OFF_TMEM_PTR);
  float* sw_smem = reinterpret_cast<float*>(smem_gen + L::OFF_SW);
  float* bias_smem = reinterpret_cast<float*>(smem_gen + L::OFF_BIAS);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(bar_full + i * 8, CG);   // one producer arrive per CTA of the pair (leader's copy is used)
      mbar_init(bar_empty + i * 8, 1);   // one tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + i * 8, 1);                   // one tcgen05.commit
      mbar_init(bar_tempty + i * 8, CG * EPI_THREADS);   // every epilogue thread of the pair
    }
    fence_mbar_init();
  }
  if (CG == 2) cluster_sync();
  if (warp == 2) {
    tmem_alloc<CG>(tmem_ptr_smem, TMEM_COLS);
    tmem_relinquish<CG>();
  }
  tc_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  const int num_clusters = gridDim.x / CG;
  const int cluster_id = blockIdx.x / CG;
  const int num_tiles = g.num_m_blocks * g.num_n_blocks;
  const int num_kb = g.num_k_blocks;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        int m_blk, n_blk;
        tile_coords(tile, g.num_m_blocks, g.num_n_blocks, m_blk, n_blk);
        const int m_idx = m_blk * (BLOCK_M * CG) + (int)cta_rank * BLOCK_M;
        const int n_idx = n_blk * BN + (int)cta_rank * L::B_ROWS;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_empty + stage * 8, phase ^ 1);
          const uint32_t sa = smem_base + L::OFF_A + stage * L::A_STAGE;
          const uint32_t sb = smem_base + L::OFF_B + stage * L::B_STAGE;
          const uint32_t fb = bar_full + stage * 8;
          if (CG == 1) {
            mbar_arrive_expect_tx(fb, L::STAGE_BYTES);
            tma_load_2d(sa, &tmap_a, fb, kb * BLOCK_K, m_idx);
            tma_load_2d(sb, &tmap_b, fb, kb * BLOCK_K, n_idx);
          } else {
            tma_load_2d_2sm(sa, &tmap_a, fb, kb * BLOCK_K, m_idx);
            tma_load_2d_2sm(sb, &tmap_b, fb, kb * BLOCK_K, n_idx);
            if (leader) mbar_arrive_expect_tx(fb, L::STAGE_BYTES * 2);
            else mbar_arrive_remote(fb, 0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc(UMMA_M, UMMA_N);
      uint32_t stage = 0, phase = 0;
      int iter = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++iter) {
        const uint32_t as = iter & 1, aphase = (iter >> 1) & 1;
        mbar_wait(bar_tempty + as * 8, aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_full + stage * 8, phase);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(smem_base + L::OFF_A + stage * L::A_STAGE);
          const uint64_t bdesc = make_smem_desc(smem_base + L::OFF_B + stage * L::B_STAGE);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 32 bytes along K inside the 128B swizzle atom: +2 in the (addr>>4) field
            mma_i8<CG>(tmem_d, adesc + (uint64_t)(k * (UMMA_K >> 4)), bdesc + (uint64_t)(k * (UMMA_K >> 4)),
                       idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit<CG>(bar_empty + stage * 8);
          if (kb == num_kb - 1) tc_commit<CG>(bar_tfull + as * 8);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      // Drain: the peer CTA's epilogue arrives on OUR tmem_empty barriers; do not let this
      // CTA exit (and its shared memory be reclaimed) before those arrivals have landed.
      if (CG == 2 && iter > 0) {
        const int last = iter - 1;
        mbar_wait(bar_tempty + (last & 1) * 8, (last >> 1) & 1);
        if (iter > 1) {
          const int prev = iter - 2;
          mbar_wait(bar_tempty + (prev & 1) * 8, (prev >> 1) & 1);
        }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ================= epilogue =================
    const int ew = warp - EPI_WARP0;             // == warp % 4: TMEM lane quarter this warp may read
    const int et = ew * 32 + (int)lane;          // row inside the CTA's 128-row slab
    int iter = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++iter) {
      int m_blk, n_blk;
      tile_coords(tile, g.num_m_blocks, g.num_n_blocks, m_blk, n_blk);
      const uint32_t as = iter & 1, aphase = (iter >> 1) & 1;
      const int row = m_blk * (BLOCK_M * CG) + (int)cta_rank * BLOCK_M + et;
      const int col0 = n_blk * BN;
      float sx = 0.f;
      if constexpr (!RAW) {
        // stage this tile's column scales / bias (double-buffered by accumulator stage)
        float* sw = sw_smem + as * BN;
        float* bs = bias_smem + as * BN;
        for (int i = et; i < BN; i += EPI_THREADS) {
          const int c = col0 + i;
          sw[i] = (c < g.N) ? __ldg(g.s_w + c) : 0.f;
          bs[i] = (g.bias != nullptr && c < g.N) ? __ldg(g.bias + c) : 0.f;
        }
        if (row < g.M) sx = __ldg(g.s_x + row);
        named_bar_sync(1, EPI_THREADS);
      }
      mbar_wait(bar_tfull + as * 8, aphase);
      tc_fence_after();
      __syncwarp();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(ew * 32) << 16) + as * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(taddr0 + c * 32, r);
        tmem_ld_wait();
        const int col = col0 + c * 32;
        if (row < g.M && col < g.N) {
          if constexpr (RAW) {
            int32_t* dst = reinterpret_cast<int32_t*>(g.out) + (long long)row * g.ldo + col;
            if (g.vec_ok && col + 32 <= g.N) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                reinterpret_cast<uint4*>(dst)[i] = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col + j < g.N) dst[j] = (int32_t)r[j];
            }
          } else {
            using OT = typename std::conditional<RAW, float, OutT>::type;
            const float* sw = sw_smem + as * BN + c * 32;
            const float* bs = bias_smem + as * BN + c * 32;
            float f[32];
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 w4 = *reinterpret_cast<const float4*>(sw + 4 * j4);
              const float4 b4 = *reinterpret_cast<const float4*>(bs + 4 * j4);
              const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
              const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float v = __int2float_rn((int)r[4 * j4 + j]);
                v = __fmul_rn(v, sx);
                v = __fmul_rn(v, wv[j]);
                v = __fadd_rn(v, bv[j]);
                f[4 * j4 + j] = v;
              }
            }
            OT* dst = reinterpret_cast<OT*>(g.out) + (long long)row * g.ldo + col;
            if (g.vec_ok && col + 32 <= g.N) {
              uint32_t o[OutPack<OT>::WORDS];
              OutPack<OT>::pack(f, o);
#pragma unroll
              for (int i = 0; i < OutPack<OT>::WORDS / 4; ++i)
                reinterpret_cast<uint4*>(dst)[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col + j < g.N) dst[j] = OutPack<OT>::one(f[j]);
            }
          }
        }
      }
      // accumulator buffer fully read: hand it back to the MMA warp (leader CTA's barrier)
      tc_fence_before();
      if (CG == 1 || leader) mbar_arrive(bar_tempty + as * 8);
      else mbar_arrive_remote(bar_tempty + as * 8, 0);
    }
  }

  __syncwarp();
  tc_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();
  if (warp == 2) tmem_dealloc<CG>(tmem_base, TMEM_COLS);
}

// ---- host side ---------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
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
int make_tmap(CUtensorMap* m, const void* base, int64_t rows, int64_t kbytes, int64_t ld, int box_rows) {
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

template <int CG, int BN, int STAGES, typename OutT>
int launch_cfg(const int8_t* a, int64_t lda, const int8_t* b, int64_t ldb, const GemmArgs& g0,
               int num_sms, cudaStream_t st) {
  using L = SmemLayout<CG, BN, STAGES>;
  GemmArgs g = g0;
  g.num_m_blocks = (g.M + BLOCK_M * CG - 1) / (BLOCK_M * CG);
  g.num_n_blocks = (g.N + BN - 1) / BN;
  g.num_k_blocks = (g.K + BLOCK_K - 1) / BLOCK_K;
  CUtensorMap ta, tb;
  int rc = make_tmap(&ta, a, g.M, g.K, lda, BLOCK_M);
  if (rc) return rc;
  rc = make_tmap(&tb, b, g.N, g.K, ldb, L::B_ROWS);
  if (rc) return rc;

  auto kern = qgemm_kernel<CG, BN, STAGES, OutT>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES);
  });
  if (attr_err != cudaSuccess)
    PQ_FAIL(PQ_ERR_CUDA, "cudaFuncSetAttribute(smem=%d) failed: %s", L::DYN_BYTES, cudaGetErrorString(attr_err));

  const long long tiles = (long long)g.num_m_blocks * g.num_n_blocks;
  long long clusters = num_sms / CG;
  if (tiles < clusters) clusters = tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters * CG), 1, 1);
  cfg.blockDim = dim3(NUM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = L::DYN_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = CG;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  PQ_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, g));
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return PQ_OK;
}

int g_force_cfg = -1;  // test hook: see pq_debug_set_gemm_config

template <typename OutT>
int launch_typed(const int8_t* a, int64_t lda, const int8_t* b, int64_t ldb, const GemmArgs& g,
                 int num_sms, cudaStream_t st) {
  // Tile configuration heuristic.
  //   cfg 0: 1-CTA 128x256   cfg 1: 2-CTA 256x256   cfg 2: 1-CTA 128x128   cfg 3: 1-CTA 128x64
  //   cfg 4: 2-CTA 256x128
  int cfg = g_force_cfg;
  if (cfg < 0) {
    const long long t256 = (long long)((g.M + 255) / 256) * ((g.N + 255) / 256);
    if (g.M > 128 && t256 * 2 >= num_sms) cfg = 1;
    else if (g.M > 128) cfg = 4;
    else if ((long long)((g.N + 127) / 128) >= num_sms) cfg = 2;
    else cfg = 3;
  }
  switch (cfg) {
    case 0: return launch_cfg<1, 256, 4, OutT>(a, lda, b, ldb, g, num_sms, st);
    case 1: return launch_cfg<2, 256, 6, OutT>(a, lda, b, ldb, g, num_sms, st);
    case 2: return launch_cfg<1, 128, 6, OutT>(a, lda, b, ldb, g, num_sms, st);
    case 3: return launch_cfg<1, 64, 8, OutT>(a, lda, b, ldb, g, num_sms, st);
    case 4: return launch_cfg<2, 128, 8, OutT>(a, lda, b, ldb, g, num_sms, st);
    case 5: return launch_cfg<2, 256, 4, OutT>(a, lda, b, ldb, g, num_sms, st);
    case 6: return launch_cfg<2, 256, 3, OutT>(a, lda, b, ldb, g, num_sms, st);
    default: PQ_FAIL(PQ_ERR_ARG, "qgemm: unknown tile config %d", cfg);
  }
}

}  // namespace

int launch_qgemm(const int8_t* a, int64_t lda, const int8_t* b, int64_t ldb,
                 const float* s_x, const float* s_w, const float* bias,
                 void* out, int out_dtype, int64_t ldo,
                 int64_t M, int64_t N, int64_t K, cudaStream_t stream) {
  if (M < 0 || N < 0 || K < 1) PQ_FAIL(PQ_ERR_ARG, "qgemm: bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  if (M == 0 || N == 0) return PQ_OK;
  if (M > 0x7fffff00LL || N > 0x7fffff00LL || K > 0x7fffff00LL) PQ_FAIL(PQ_ERR_ARG, "qgemm: dimension too large");
  if (!a || !b || !out) PQ_FAIL(PQ_ERR_ARG, "qgemm: null pointer");
  if (out_dtype != PQ_I32 && (!s_x || !s_w)) PQ_FAIL(PQ_ERR_ARG, "qgemm: null scale pointer");
  if (lda < K || ldb < K || ldo < N) PQ_FAIL(PQ_ERR_ARG, "qgemm: leading dimension too small");
  if (((uintptr_t)a & 15) || ((uintptr_t)b & 15) || (lda & 15) || (ldb & 15))
    PQ_FAIL(PQ_ERR_ALIGN, "qgemm: xq/Wq base pointers and row strides must be multiples of 16 bytes (TMA)");
  int num_sms = 0;
  int rc = check_device(&num_sms);
  if (rc) return rc;
  GemmArgs g = {};
  g.M = (int)M; g.N = (int)N; g.K = (int)K;
  g.s_x = s_x; g.s_w = s_w; g.bias = bias;
  g.out = out; g.ldo = ldo;
  const int esz = dtype_size(out_dtype);
  g.vec_ok = (((uintptr_t)out & 15) == 0) && ((ldo * esz) % 16 == 0);
  switch (out_dtype) {
    case PQ_BF16: return launch_typed<__nv_bfloat16>(a, lda, b, ldb, g, num_sms, stream);
    case PQ_F16: return launch_typed<__half>(a, lda, b, ldb, g, num_sms, stream);
    case PQ_F32: return launch_typed<float>(a, lda, b, ldb, g, num_sms, stream);
    case PQ_I32: return launch_typed<int32_t>(a, lda, b, ldb, g, num_sms, stream);
    default: PQ_FAIL(PQ_ERR_ARG, "qgemm: unsupported output dtype %d", out_dtype);
  }
}

}  // namespace pq

// Test/bench hook (not part of the reference-facing API): force a tile configuration,
// -1 restores the heuristic.
extern "C" void pq_debug_set_gemm_config(int cfg) { pq::g_force_cfg = cfg; }
