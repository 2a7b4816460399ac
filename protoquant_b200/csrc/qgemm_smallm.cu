// Small-M (decode) int8 GEMM + fused dequant epilogue, M <= 64 tokens (SURVEY.md §8 rows a3+a4,
// "qgemv_i8_smallM").  HBM-bound: the job is to stream Wq[N,K] once at full bandwidth.
//
// Swap-AB: the weight tile is the MMA's M operand (128 output channels = 128 TMEM lanes) and the
// token block is the N operand (M_pad = 16/32/64 columns), so a CTA's ring slot is 16 KB of
// weights + only M_pad*128 B of activations; two CTAs share an SM (<= 104 KB each), which keeps
// ~180 KB of weight loads in flight per SM.
//
//   grid    = ceil(N/128) channel tiles  x  S K-splits, launched as clusters of S CTAs
//   per CTA : warp 0 TMA producer, warp 1 tcgen05.mma issuer (M=128, N=M_pad, K=32),
//             warp 2 TMEM alloc, warps 4-7 dump the int32 accumulator TMEM -> shared memory
//   reduce  : cluster barrier, then CTA r of the cluster sums channel slice r of all S partials
//             through distributed shared memory (exact: int32), applies
//             ((float(acc)*s_x[m])*s_w[n])+bias[n] and stores its slice of y.
// S is chosen so that tiles x S covers the SMs (N=4096 -> 32 tiles x 4).
#include "gemm_common.cuh"
#include "quant_math.cuh"

namespace pq {
Knob g_smallm_splits{0};  // pq_debug_set_smallm_splits: force the K-split (cluster size) of the decode GEMM; 0 = heuristic
Knob g_fused_decode{0};   // pq_debug_set_fused_decode; OFF: measured slower than quant kernel + PDL (profiles/README_r1.md)
namespace {

using namespace ptx;
using namespace gemm;

constexpr int SM_THREADS = 256;
constexpr int TILE_N = 128;   // output channels per CTA

struct SmallArgs {
  int M, N, K;
  int num_kb;       // ceil(K / 128)
  int splits;       // S = cluster size
  const float* s_x;
  const float* s_w;
  const float* bias;
  void* out[8];     // the [M,N] result goes to each of out[0..n_out): local buffer and NVLink peers (fused all-gather)
  int n_out;
  long long ldo;
};

template <int MP, int STAGES>
struct SmallLayout {
  static constexpr int W_STAGE = TILE_N * BLOCK_K;       // 16 KB
  static constexpr int X_STAGE = MP * BLOCK_K;
  static constexpr int X_STAGE_AL = (X_STAGE + 1023) / 1024 * 1024;
  static constexpr int STAGE_BYTES = W_STAGE + X_STAGE;
  static constexpr int OFF_W = 0;
  static constexpr int OFF_X = OFF_W + STAGES * W_STAGE;
  static constexpr int OFF_PART = OFF_X + STAGES * X_STAGE_AL;    // [MP][128] int32
  static constexpr int OFF_BAR = OFF_PART + MP * TILE_N * 4;       // full[S], empty[S], tfull
  static constexpr int OFF_TMEM_PTR = OFF_BAR + (2 * STAGES + 1) * 8;
  static constexpr int TOTAL = OFF_TMEM_PTR + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;
  static_assert(DYN_BYTES <= 113 * 1024, "two CTAs must fit per SM");
};

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ int ld_dsmem_s32(uint32_t local_addr, uint32_t cta) {
  int v;
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %1, %2;\n\t"
      "ld.shared::cluster.s32 %0, [ra];\n\t}"
      : "=r"(v)
      : "r"(local_addr), "r"(cta)
      : "memory");
  return v;
}

template <typename OutT> __device__ __forceinline__ void store_one(void* out, long long idx, float f, int raw);
template <> __device__ __forceinline__ void store_one<__nv_bfloat16>(void* out, long long idx, float f, int) {
  reinterpret_cast<__nv_bfloat16*>(out)[idx] = __float2bfloat16_rn(f);
}
template <> __device__ __forceinline__ void store_one<__half>(void* out, long long idx, float f, int) {
  reinterpret_cast<__half*>(out)[idx] = __float2half_rn(f);
}
template <> __device__ __forceinline__ void store_one<float>(void* out, long long idx, float f, int) {
  reinterpret_cast<float*>(out)[idx] = f;
}
template <> __device__ __forceinline__ void store_one<int32_t>(void* out, long long idx, float, int raw) {
  reinterpret_cast<int32_t*>(out)[idx] = raw;
}

template <int MP, int STAGES, typename OutT>
__global__ void __launch_bounds__(SM_THREADS, 2)
qgemm_smallm_kernel(const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ CUtensorMap tmap_x, const SmallArgs g) {
  using L = SmallLayout<MP, STAGES>;
  constexpr bool RAW = std::is_same<OutT, int32_t>::value;
  constexpr uint32_t TMEM_COLS = MP < 32 ? 32 : MP;
  constexpr uint32_t idesc = make_idesc(TILE_N, MP);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const int S = g.splits;
  const uint32_t split = (S > 1) ? cluster_ctarank() : 0u;
  const int n_tile = blockIdx.x / S;
  const int kb_begin = (int)((long long)g.num_kb * split / S);
  const int kb_end = (int)((long long)g.num_kb * (split + 1) / S);

  const uint32_t bar_full = smem_base + L::OFF_BAR;
  const uint32_t bar_empty = bar_full + STAGES * 8;
  const uint32_t bar_tfull = bar_empty + STAGES * 8;
  const uint32_t tmem_ptr_smem = smem_base + L::OFF_TMEM_PTR;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + L::OFF_TMEM_PTR);
  int32_t* part = reinterpret_cast<int32_t*>(smem_gen + L::OFF_PART);

  griddep_launch_dependents();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_x);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(bar_full + i * 8, 1);
      mbar_init(bar_empty + i * 8, 1);
    }
    mbar_init(bar_tfull, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<1>(tmem_ptr_smem, TMEM_COLS);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  // PDL: the weights do not depend on the previous kernel (the activation quantizer), so the
  // producer starts streaming them before it waits for that kernel; everyone else waits here.
  if (!(warp == 0 && lane == 0)) griddep_wait();

  if (warp == 0) {
    if (lane == 0) {
      const int n_pre = (kb_end - kb_begin) < STAGES ? (kb_end - kb_begin) : STAGES;
      for (int i = 0; i < n_pre; ++i) {      // ring is empty: no `empty` wait needed yet
        const uint32_t fb = bar_full + i * 8;
        mbar_arrive_expect_tx(fb, L::STAGE_BYTES);
        tma_load_2d(smem_base + L::OFF_W + i * L::W_STAGE, &tmap_w, fb, (kb_begin + i) * BLOCK_K, n_tile * TILE_N);
      }
      griddep_wait();                          // xq / s_x are produced by the previous kernel
      for (int i = 0; i < n_pre; ++i)
        tma_load_2d(smem_base + L::OFF_X + i * L::X_STAGE_AL, &tmap_x, bar_full + i * 8, (kb_begin + i) * BLOCK_K, 0);
      uint32_t stage = (n_pre == STAGES) ? 0 : (uint32_t)n_pre, phase = (n_pre == STAGES) ? 1 : 0;
      for (int kb = kb_begin + n_pre; kb < kb_end; ++kb) {
        mbar_wait(bar_empty + stage * 8, phase ^ 1);
        const uint32_t fb = bar_full + stage * 8;
        mbar_arrive_expect_tx(fb, L::STAGE_BYTES);
        tma_load_2d(smem_base + L::OFF_W + stage * L::W_STAGE, &tmap_w, fb, kb * BLOCK_K, n_tile * TILE_N);
        tma_load_2d(smem_base + L::OFF_X + stage * L::X_STAGE_AL, &tmap_x, fb, kb * BLOCK_K, 0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(bar_full + stage * 8, phase);
        tc_fence_after();
        const uint64_t wdesc = make_smem_desc(smem_base + L::OFF_W + stage * L::W_STAGE);
        const uint64_t xdesc = make_smem_desc(smem_base + L::OFF_X + stage * L::X_STAGE_AL);
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
          mma_i8<1>(tmem_base, wdesc + (uint64_t)(k * (UMMA_K >> 4)), xdesc + (uint64_t)(k * (UMMA_K >> 4)),
                    idesc, ((kb - kb_begin) | k) != 0 ? 1u : 0u);
        tc_commit<1>(bar_empty + stage * 8);
        if (kb == kb_end - 1) tc_commit<1>(bar_tfull);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // dump the accumulator: lane n of TMEM holds channel n, column m holds token m
    const int ew = warp & 3;
    const int n_local = ew * 32 + (int)lane;
    if (kb_end > kb_begin) {
      mbar_wait(bar_tfull, 0);
      tc_fence_after();
      __syncwarp();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16);
      if constexpr (MP == 16) {
        uint32_t r[16];
        tmem_ld_32x16(taddr, r);
        tmem_ld_wait();
#pragma unroll
        for (int m = 0; m < 16; ++m) part[m * TILE_N + n_local] = (int32_t)r[m];
      } else {
#pragma unroll 1
        for (int c = 0; c < MP / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int m = 0; m < 32; ++m) part[(c * 32 + m) * TILE_N + n_local] = (int32_t)r[m];
        }
      }
    } else {
      for (int m = 0; m < MP; ++m) part[m * TILE_N + n_local] = 0;   // empty K range (S > num_kb)
    }
  }

  __syncwarp();
  tc_fence_before();
  if (S > 1) cluster_sync(); else __syncthreads();

  // ---- reduce channel slice `split` of every partial and write y ----
  {
    const int slice = TILE_N / S;                       // channels this CTA finalises
    const int total = g.M * slice;
    const uint32_t part_addr = smem_base + L::OFF_PART;
    for (int idx = threadIdx.x; idx < total; idx += SM_THREADS) {
      const int m = idx / slice;
      const int nl = (int)split * slice + (idx - m * slice);
      const int n = n_tile * TILE_N + nl;
      if (n >= g.N) continue;
      // scales / bias first: their L2 round trip overlaps the distributed-shared-memory reads below instead of
      // following them (the decode launch is a chain of dependent latencies, profiles/decode_parts_r2.log)
      float sx = 0.f, sw = 0.f, bs = 0.f;
      if constexpr (!RAW) {
        sx = __ldg(g.s_x + m);
        sw = __ldg(g.s_w + n);
        if (g.bias != nullptr) bs = __ldg(g.bias + n);
      }
      int acc = 0;
      if (S > 1) {
        int pa[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) pa[p] = (p < S) ? ld_dsmem_s32(part_addr + (uint32_t)(m * TILE_N + nl) * 4u, (uint32_t)p) : 0;
#pragma unroll
        for (int p = 0; p < 8; ++p) acc += pa[p];
      } else {
        acc = part[m * TILE_N + nl];
      }
      float v = 0.f;
      if constexpr (!RAW) {
        v = __int2float_rn(acc);
        v = __fmul_rn(v, sx);
        v = __fmul_rn(v, sw);
        if (g.bias != nullptr) v = __fadd_rn(v, bs);   // no add at all without bias (-0.0 stays -0.0)
      }
      for (int d = 0; d < g.n_out; ++d) store_one<OutT>(g.out[d], (long long)m * g.ldo + n, v, acc);
    }
    if (g.n_out > 1) __threadfence_system();   // peer stores ordered before the kernel (and the caller's barrier) ends
  }

  if (S > 1) cluster_sync(); else __syncthreads();   // peers may still be reading our partial
  if (warp == 2) tmem_dealloc<1>(tmem_base, TMEM_COLS);
}

// ======================================================================================
// Fused decode linear: per-token activation quantisation inside the small-M
// GEMM.  The S CTAs of a cluster split K; each quantises only ITS K-slice of x:
//   A. warps 2-7 take the row |.|-max of the slice (16-byte loads, shared-memory atomicMax on the
//      float bits), push it to every CTA of the cluster through distributed shared memory and
//      wait on a cluster-scope mbarrier -> every CTA owns the full-row amax, hence the same scale;
//   B. the slice is re-read (L2 hits), quantised with the very same arithmetic as the stand-alone
//      quantizer (quant_math.cuh) and written straight into the 128B-swizzled K-major layout that
//      tcgen05.mma expects for its N operand (what TMA would have produced);
//   meanwhile warp 0 is already streaming the weights -- they do not depend on x.
// One launch instead of two, and xq / s_x never touch global memory.
// STATUS: bit-exact and tested, but OFF by default -- on B200 it measures 14.5 us vs 10.3 us for the
// two-kernel path at 16 x 4096 x 4096 (31 vs 17 us at N = 11008, 25 vs 12 us at 32 tokens): the cost scales with
// the number of CTAs and with M, i.e. it is the REDUNDANT quantisation -- each of the 32 channel-tile clusters
// re-quantises x, ~1000 instructions per quantising thread on every SM (round 2 replaced the shared-memory
// atomicMax reduction by one warp per row: no change, so it was never the reduction) -- whereas the separate 2 us
// quantizer kernel does the work once and overlaps the GEMM's weight prefetch through PDL.  A faster fusion has
// to quantise once and publish (flag + TMA loads by the consumers); that saves one kernel boundary, ~1 us.
struct FusedArgs {
  int M, N, K;
  int num_kb, splits;
  const void* x;
  long long ldx;     // elements
  const float* s_w;
  const float* bias;
  void* out;
  long long ldo;
  int scale_mode;
  float eps;
};

constexpr int FUSED_X_BYTES = 32768;    // quantised activation slice per CTA: M_pad x K_slice int8
constexpr int QTHREADS = 192;           // warps 2..7 quantise

template <int MP, int STAGES>
struct FusedLayout {
  static constexpr int W_STAGE = TILE_N * BLOCK_K;
  static constexpr int OFF_W = 0;
  static constexpr int OFF_X = OFF_W + STAGES * W_STAGE;
  static constexpr int OFF_PART = OFF_X + FUSED_X_BYTES;
  static constexpr int OFF_AMAX_ALL = OFF_PART + MP * TILE_N * 4;   // [8][MP] u32 (one row per peer)
  static constexpr int OFF_AMAX = OFF_AMAX_ALL + 8 * MP * 4;        // [MP] u32 local partial
  static constexpr int OFF_ROWQ = OFF_AMAX + MP * 4;                // [MP] RowQ (12 B)
  static constexpr int OFF_BAR = (OFF_ROWQ + MP * 16 + 7) / 8 * 8;  // full[S], empty[S], tfull, xbar, amaxbar
  static constexpr int OFF_TMEM_PTR = OFF_BAR + (2 * STAGES + 3) * 8;
  static constexpr int TOTAL = OFF_TMEM_PTR + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;
  static_assert(DYN_BYTES <= 113 * 1024, "two CTAs must fit per SM");
};

template <typename T, int MP, int STAGES, typename OutT>
__global__ void __launch_bounds__(SM_THREADS, 2)
qlinear_smallm_fused_kernel(const __grid_constant__ CUtensorMap tmap_w, const FusedArgs g) {
  using L = FusedLayout<MP, STAGES>;
  using namespace qmath;
  constexpr int EPV = VecTraits<T>::EPV;
  constexpr uint32_t TMEM_COLS = MP < 32 ? 32 : MP;
  constexpr uint32_t idesc = make_idesc(TILE_N, MP);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const int S = g.splits;
  const uint32_t split = (S > 1) ? cluster_ctarank() : 0u;
  const int n_tile = blockIdx.x / S;
  const int kb_begin = (int)((long long)g.num_kb * split / S);
  const int kb_end = (int)((long long)g.num_kb * (split + 1) / S);

  const uint32_t bar_full = smem_base + L::OFF_BAR;
  const uint32_t bar_empty = bar_full + STAGES * 8;
  const uint32_t bar_tfull = bar_empty + STAGES * 8;
  const uint32_t bar_x = bar_tfull + 8;
  const uint32_t bar_amax = bar_x + 8;
  const uint32_t tmem_ptr_smem = smem_base + L::OFF_TMEM_PTR;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + L::OFF_TMEM_PTR);
  int32_t* part = reinterpret_cast<int32_t*>(smem_gen + L::OFF_PART);
  uint32_t* amax_all = reinterpret_cast<uint32_t*>(smem_gen + L::OFF_AMAX_ALL);
  uint32_t* amax_loc = reinterpret_cast<uint32_t*>(smem_gen + L::OFF_AMAX);
  RowQ* rowq = reinterpret_cast<RowQ*>(smem_gen + L::OFF_ROWQ);
  uint8_t* xs = smem_gen + L::OFF_X;

  griddep_launch_dependents();
  if (warp == 0 && lane == 0) prefetch_tmap(&tmap_w);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(bar_full + i * 8, 1);
      mbar_init(bar_empty + i * 8, 1);
    }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_x, QTHREADS);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<1>(tmem_ptr_smem, TMEM_COLS);
    tmem_relinquish<1>();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + MP) amax_loc[threadIdx.x - 64] = 0u;
  tc_fence_before();
  if (S > 1) cluster_sync(); else __syncthreads();     // peers' barriers exist before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  // The row-amax exchange uses one split-phase cluster barrier: warps that have nothing to publish
  // arrive right away and only wait for it once their own work is done.
  if (warp == 0) {
    // ---- weight producer: independent of x, starts before the PDL wait ----
    if (S > 1) cluster_arrive();
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(bar_empty + stage * 8, phase ^ 1);
        const uint32_t fb = bar_full + stage * 8;
        mbar_arrive_expect_tx(fb, L::W_STAGE);
        tma_load_2d(smem_base + L::OFF_W + stage * L::W_STAGE, &tmap_w, fb, kb * BLOCK_K, n_tile * TILE_N);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    if (S > 1) cluster_wait();
  } else if (warp == 1) {
    if (S > 1) { cluster_arrive(); cluster_wait(); }
    if (lane == 0 && kb_end > kb_begin) {
      mbar_wait(bar_x, 0);                               // quantised slice is in shared memory
      tc_fence_after();
      uint32_t stage = 0, phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(bar_full + stage * 8, phase);
        tc_fence_after();
        const uint64_t wdesc = make_smem_desc(smem_base + L::OFF_W + stage * L::W_STAGE);
        const uint64_t xdesc = make_smem_desc(smem_base + L::OFF_X + (uint32_t)(kb - kb_begin) * (MP * BLOCK_K));
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
          mma_i8<1>(tmem_base, wdesc + (uint64_t)(k * (UMMA_K >> 4)), xdesc + (uint64_t)(k * (UMMA_K >> 4)),
                    idesc, ((kb - kb_begin) | k) != 0 ? 1u : 0u);
        tc_commit<1>(bar_empty + stage * 8);
        if (kb == kb_end - 1) tc_commit<1>(bar_tfull);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ---- warps 2..7: quantise this CTA's K-slice of x ----
    griddep_wait();
    const int qt = (int)threadIdx.x - 64;
    const T* x = reinterpret_cast<const T*>(g.x);
    const int k_lo = kb_begin * BLOCK_K;
    int k_hi = kb_end * BLOCK_K;
    if (k_hi > g.K) k_hi = g.K;
    const int vpr = (k_hi > k_lo) ? (k_hi - k_lo) / EPV : 0;    // 16-byte vectors per row in the slice
    const int nv = g.M * vpr;
    (void)nv;
    // One WARP per row (rows qw, qw + 6, ...): lane l holds vectors l, l + 32, ... of the row's slice, the row maximum
    // is a warp shuffle reduction -- the first version funnelled every vector through a shared-memory atomicMax on
    // one of 16 addresses (a 32-way conflict per warp instruction), which alone cost several microseconds.
    constexpr int QW = QTHREADS / 32;                            // quantising warps
    constexpr int VL = 4;                                        // 16-byte vectors per lane and row kept in registers
    constexpr int RPW = (MP + QW - 1) / QW;                      // rows per warp: 3 (MP = 16), 6, 11
    constexpr int KEEP_ROWS = RPW <= 3 ? RPW : 1;
    const int qw = warp - 2;
    const bool resident = (RPW <= 3) && vpr <= 32 * VL;          // the warp's rows fit in registers: read x once
    uint4 keep[KEEP_ROWS][VL];
    auto load_rv = [&](int r, int cv) -> uint4 {
      return *reinterpret_cast<const uint4*>(x + (long long)r * g.ldx + k_lo + cv * EPV);
    };
    auto vec_amax = [&](const uint4& raw, uint32_t m) -> uint32_t {
      if (sizeof(T) == 2) return absmax_u16x2(raw, m);           // packed u16 lanes, converted once per row
      return __float_as_uint(vec_absmax<float>(raw, __uint_as_float(m)));
    };
    auto quant_store = [&](int r, int cv, const uint4& raw, const RowQ& rq) {
      float f[EPV];
      unpack<T>(raw, f);
#pragma unroll
      for (int j = 0; j < EPV; ++j) f[j] = quant_any(f[j], rq);
      const int kl = cv * EPV;                       // k offset inside the slice
      const int kbl = kl >> 7, b = kl & 127;         // k-block, byte inside the 128-byte row
      uint8_t* dst = xs + kbl * (MP * BLOCK_K) + (r >> 3) * 1024 + (r & 7) * 128 + ((((b >> 4) ^ (r & 7))) << 4) + (b & 15);
      if (EPV == 8) {
        uint2 o;
        o.x = pack4(f[0], f[1], f[2], f[3]);
        o.y = pack4(f[4 % EPV], f[5 % EPV], f[6 % EPV], f[7 % EPV]);
        *reinterpret_cast<uint2*>(dst) = o;
      } else {
        *reinterpret_cast<uint32_t*>(dst) = pack4(f[0], f[1], f[2], f[3]);
      }
    };
    // A. row |.|-max of the slice (all loads of a batch are issued before any is used)
#pragma unroll
    for (int ri = 0; ri < RPW; ++ri) {
      const int r = qw + ri * QW;
      if (r < g.M) {
        uint32_t m = 0;
        for (int c0 = 0; c0 < vpr; c0 += 32 * VL) {
          uint4 raw[VL];
#pragma unroll
          for (int j = 0; j < VL; ++j) {
            const int cv = c0 + (int)lane + 32 * j;
            raw[j] = make_uint4(0, 0, 0, 0);
            if (cv < vpr) raw[j] = load_rv(r, cv);
          }
#pragma unroll
          for (int j = 0; j < VL; ++j) {
            m = vec_amax(raw[j], m);
            if (ri < KEEP_ROWS) keep[ri < KEEP_ROWS ? ri : 0][j] = raw[j];
          }
        }
        if (sizeof(T) == 2) m = __float_as_uint(u16_mag_to_float<T>(m));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) amax_loc[r] = m;
      }
    }
    named_bar_sync(1, QTHREADS);
    if (qt < g.M) {
      const uint32_t mine = amax_loc[qt];
      if (S > 1) {
        const uint32_t slot = smem_base + L::OFF_AMAX_ALL + (split * MP + (uint32_t)qt) * 4u;
        for (int p = 0; p < S; ++p) st_dsmem_u32(slot, (uint32_t)p, mine);
      } else {
        amax_all[qt] = mine;
      }
    }
    if (S > 1) {
      cluster_arrive();      // release: the remote stores above are visible to every CTA after the wait
      cluster_wait();
    } else {
      named_bar_sync(1, QTHREADS);
    }
    if (qt < g.M) {
      uint32_t m = 0;
      for (int p = 0; p < S; ++p) m = max(m, amax_all[p * MP + qt]);
      rowq[qt] = make_rowq(__uint_as_float(m), g.scale_mode, g.eps);
    }
    named_bar_sync(1, QTHREADS);
    // B. quantise into the swizzled K-major operand layout
#pragma unroll
    for (int ri = 0; ri < RPW; ++ri) {
      const int r = qw + ri * QW;
      if (r < g.M) {
        const RowQ rq = rowq[r];
        if (resident) {
#pragma unroll
          for (int j = 0; j < VL; ++j) {
            const int cv = (int)lane + 32 * j;
            if (cv < vpr) quant_store(r, cv, keep[ri < KEEP_ROWS ? ri : 0][j], rq);
          }
        } else {
          for (int c0 = 0; c0 < vpr; c0 += 32 * VL) {
            uint4 raw[VL];
#pragma unroll
            for (int j = 0; j < VL; ++j) {
              const int cv = c0 + (int)lane + 32 * j;
              raw[j] = make_uint4(0, 0, 0, 0);
              if (cv < vpr) raw[j] = load_rv(r, cv);
            }
#pragma unroll
            for (int j = 0; j < VL; ++j) {
              const int cv = c0 + (int)lane + 32 * j;
              if (cv < vpr) quant_store(r, cv, raw[j], rq);
            }
          }
        }
      }
    }
    fence_proxy_async_smem();                        // generic-proxy writes -> visible to tcgen05.mma
    mbar_arrive(bar_x);

    if (warp >= 4) {
      // dump the accumulator: lane n of TMEM holds channel n, column m holds token m
      const int ew = warp & 3;
      const int n_local = ew * 32 + (int)lane;
      if (kb_end > kb_begin) {
        mbar_wait(bar_tfull, 0);
        tc_fence_after();
        __syncwarp();
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16);
        if constexpr (MP == 16) {
          uint32_t r[16];
          tmem_ld_32x16(taddr, r);
          tmem_ld_wait();
#pragma unroll
          for (int m = 0; m < 16; ++m) part[m * TILE_N + n_local] = (int32_t)r[m];
        } else {
#pragma unroll 1
          for (int c = 0; c < MP / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(taddr + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int m = 0; m < 32; ++m) part[(c * 32 + m) * TILE_N + n_local] = (int32_t)r[m];
          }
        }
      } else {
        for (int m = 0; m < MP; ++m) part[m * TILE_N + n_local] = 0;
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  if (S > 1) cluster_sync(); else __syncthreads();
  if (warp < 2) griddep_wait();                        // these threads write y below

  {
    const int slice = TILE_N / S;
    const int total = g.M * slice;
    const uint32_t part_addr = smem_base + L::OFF_PART;
    for (int idx = threadIdx.x; idx < total; idx += SM_THREADS) {
      const int m = idx / slice;
      const int nl = (int)split * slice + (idx - m * slice);
      const int n = n_tile * TILE_N + nl;
      if (n >= g.N) continue;
      int acc = 0;
      if (S > 1) {
        for (int p = 0; p < S; ++p) acc += ld_dsmem_s32(part_addr + (uint32_t)(m * TILE_N + nl) * 4u, (uint32_t)p);
      } else {
        acc = part[m * TILE_N + nl];
      }
      float v = __int2float_rn(acc);
      v = __fmul_rn(v, rowq[m].s);
      v = __fmul_rn(v, __ldg(g.s_w + n));
      if (g.bias != nullptr) v = __fadd_rn(v, __ldg(g.bias + n));
      store_one<OutT>(g.out, (long long)m * g.ldo + n, v, acc);
    }
  }

  if (S > 1) cluster_sync(); else __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, TMEM_COLS);
}

template <typename T, int MP, int STAGES, typename OutT>
int launch_fused(const int8_t* b, int64_t ldb, FusedArgs g, int S, cudaStream_t st) {
  using L = FusedLayout<MP, STAGES>;
  CUtensorMap tw;
  int rc = make_tmap(&tw, b, g.N, g.K, ldb, TILE_N);
  if (rc) return rc;
  auto kern = qlinear_smallm_fused_kernel<T, MP, STAGES, OutT>;
  static PerDeviceOnce once;   // per device: a process may drive several GPUs
  const cudaError_t attr_err = once.run([&](int*) {
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES);
  });
  if (attr_err != cudaSuccess)
    PQ_FAIL(PQ_ERR_CUDA, "cudaFuncSetAttribute(smem=%d) failed: %s", L::DYN_BYTES, cudaGetErrorString(attr_err));
  const int n_tiles = (g.N + TILE_N - 1) / TILE_N;
  g.splits = S;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_tiles * S), 1, 1);
  cfg.blockDim = dim3(SM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = L::DYN_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = S;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = g_pdl ? 2 : 1;
  PQ_CUDA(cudaLaunchKernelEx(&cfg, kern, tw, g));
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return PQ_OK;
}

template <typename T, typename OutT>
int launch_fused_mp(const int8_t* b, int64_t ldb, const FusedArgs& g, int S, cudaStream_t st) {
  if (g.M <= 16) return launch_fused<T, 16, 4, OutT>(b, ldb, g, S, st);   // 64 + 32 + 8 KB
  return launch_fused<T, 32, 3, OutT>(b, ldb, g, S, st);                   // 48 + 32 + 16 KB
}

template <typename T>
int launch_fused_out(const int8_t* b, int64_t ldb, const FusedArgs& g, int S, int out_dtype, cudaStream_t st) {
  switch (out_dtype) {
    case PQ_BF16: return launch_fused_mp<T, __nv_bfloat16>(b, ldb, g, S, st);
    case PQ_F16: return launch_fused_mp<T, __half>(b, ldb, g, S, st);
    case PQ_F32: return launch_fused_mp<T, float>(b, ldb, g, S, st);
    default: PQ_FAIL(PQ_ERR_ARG, "fused decode linear: unsupported output dtype %d", out_dtype);
  }
}

template <int MP, int STAGES, typename OutT>
int launch_small(const int8_t* a, int64_t lda, const int8_t* b, int64_t ldb, SmallArgs g, int num_sms,
                 cudaStream_t st) {
  using L = SmallLayout<MP, STAGES>;
  CUtensorMap tw, tx;
  int rc = make_tmap(&tw, b, g.N, g.K, ldb, TILE_N);
  if (rc) return rc;
  rc = make_tmap(&tx, a, g.M, g.K, lda, MP);
  if (rc) return rc;
  auto kern = qgemm_smallm_kernel<MP, STAGES, OutT>;
  static PerDeviceOnce once;   // per device: a process may drive several GPUs
  const cudaError_t attr_err = once.run([&](int*) {
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES);
  });
  if (attr_err != cudaSuccess)
    PQ_FAIL(PQ_ERR_CUDA, "cudaFuncSetAttribute(smem=%d) failed: %s", L::DYN_BYTES, cudaGetErrorString(attr_err));
  const int n_tiles = (g.N + TILE_N - 1) / TILE_N;
  // Two CTAs fit per SM (<= 113 KB of shared memory each): split K until the grid fills those
  // 2 x num_sms slots, keeping at least 4 K blocks per CTA -- or 1 when K is so short (<= 1024) that the launch is
  // pure latency and more CTAs only shorten each one's chain (16 x 768 -> 3072: 6.9 -> 5.4 us, tools/decode_parts.py).
  int S = 1;
  const int min_kb = g.num_kb <= 8 ? 1 : 4;
  while (S < 8 && n_tiles * S * 2 <= 2 * num_sms && g.num_kb / (S * 2) >= min_kb) S *= 2;
  const int forced = g_smallm_splits;
  if (forced == 1 || forced == 2 || forced == 4 || forced == 8) S = (forced <= g.num_kb) ? forced : 1;
  g.splits = S;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_tiles * S), 1, 1);
  cfg.blockDim = dim3(SM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = L::DYN_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = S;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = g_pdl ? 2 : 1;
  PQ_CUDA(cudaLaunchKernelEx(&cfg, kern, tw, tx, g));
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return PQ_OK;
}

template <typename OutT>
int launch_small_typed(const int8_t* a, int64_t lda, const int8_t* b, int64_t ldb, const SmallArgs& g,
                       int num_sms, cudaStream_t st) {
  if (g.M <= 16) return launch_small<16, 5, OutT>(a, lda, b, ldb, g, num_sms, st);   //  98 KB smem
  if (g.M <= 32) return launch_small<32, 4, OutT>(a, lda, b, ldb, g, num_sms, st);   //  96 KB
  return launch_small<64, 3, OutT>(a, lda, b, ldb, g, num_sms, st);                  // 104 KB
}

}  // namespace

int launch_qgemm_smallm(const int8_t* a, int64_t lda, const int8_t* b, int64_t ldb,
                        const float* s_x, const float* s_w, const float* bias,
                        void* const* outs, int n_out, int out_dtype, int64_t ldo,
                        int64_t M, int64_t N, int64_t K, int num_sms, cudaStream_t stream) {
  if (M < 1 || M > 64) PQ_FAIL(PQ_ERR_ARG, "small-M GEMM needs 1 <= M <= 64");
  if (n_out < 1 || n_out > 8) PQ_FAIL(PQ_ERR_ARG, "small-M GEMM: 1..8 destinations");
  SmallArgs g = {};
  g.M = (int)M; g.N = (int)N; g.K = (int)K;
  g.num_kb = (int)((K + BLOCK_K - 1) / BLOCK_K);
  g.s_x = s_x; g.s_w = s_w; g.bias = bias;
  for (int d = 0; d < n_out; ++d) g.out[d] = outs[d];
  g.n_out = n_out;
  g.ldo = ldo;
  switch (out_dtype) {
    case PQ_BF16: return launch_small_typed<__nv_bfloat16>(a, lda, b, ldb, g, num_sms, stream);
    case PQ_F16: return launch_small_typed<__half>(a, lda, b, ldb, g, num_sms, stream);
    case PQ_F32: return launch_small_typed<float>(a, lda, b, ldb, g, num_sms, stream);
    case PQ_I32: return launch_small_typed<int32_t>(a, lda, b, ldb, g, num_sms, stream);
    default: PQ_FAIL(PQ_ERR_ARG, "small-M GEMM: unsupported output dtype %d", out_dtype);
  }
}

// Fused act-quant + GEMM for decode batches.  Returns 1 (and launches nothing) when the shape is not
// eligible -- the caller then runs the two-kernel path -- 0 on success, a PQ_ERR_* otherwise.
int launch_qlinear_smallm_fused(const void* x, int x_dtype, int64_t ldx,
                                const int8_t* b, int64_t ldb, const float* s_w, const float* bias,
                                void* out, int out_dtype, int64_t ldo,
                                int64_t M, int64_t N, int64_t K, const pq_quant_spec& spec,
                                int num_sms, cudaStream_t stream) {
  if (g_fused_decode == 0 || M < 1 || M > 32) return 1;
  const int esz = dtype_size(x_dtype);
  const int epv = 16 / esz;
  if (x_dtype != PQ_BF16 && x_dtype != PQ_F16 && x_dtype != PQ_F32) return 1;
  if ((K % 16) != 0 || (K % epv) != 0 || ((uintptr_t)x & 15) || ((ldx * esz) % 16) != 0) return 1;
  if (((uintptr_t)b & 15) || (ldb & 15)) return 1;
  const int num_kb = (int)((K + BLOCK_K - 1) / BLOCK_K);
  const int n_tiles = (int)((N + TILE_N - 1) / TILE_N);
  int S = 1;
  while (S < 8 && n_tiles * S * 2 <= 2 * num_sms && num_kb / (S * 2) >= 4) S *= 2;
  const int mp = M <= 16 ? 16 : 32;
  const long long kb_per_cta = (num_kb + S - 1) / S;
  if (kb_per_cta * BLOCK_K * mp > FUSED_X_BYTES) return 1;     // quantised slice must fit in shared memory
  FusedArgs g = {};
  g.M = (int)M; g.N = (int)N; g.K = (int)K;
  g.num_kb = num_kb;
  g.x = x; g.ldx = ldx;
  g.s_w = s_w; g.bias = bias;
  g.out = out; g.ldo = ldo;
  g.scale_mode = mode_bits(spec); g.eps = spec.eps;
  switch (x_dtype) {
    case PQ_BF16: return launch_fused_out<__nv_bfloat16>(b, ldb, g, S, out_dtype, stream);
    case PQ_F16: return launch_fused_out<__half>(b, ldb, g, S, out_dtype, stream);
    default: return launch_fused_out<float>(b, ldb, g, S, out_dtype, stream);
  }
}

}  // namespace pq

// Test/bench hook: 0 = never fuse the activation quantizer into the decode GEMM.
extern "C" void pq_debug_set_fused_decode(int on) { pq::g_fused_decode = on; }

extern "C" void pq_debug_set_smallm_splits(int s) { pq::g_smallm_splits = s; }
