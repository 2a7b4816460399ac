// int8 x int8 -> int32 GEMM on tcgen05 with the fused dequant epilogue
// (SURVEY.md §8 rows a3 + a4).
//
//   acc[m,n] = sum_k xq[m,k] * Wq[n,k]                       (exact int32)
//   y[m,n]   = cast( (float(acc) * s_x[m]) * s_w[n] + bias[n] )   (fp32, one RNE cast)
//
// Both operands are K-major (xq [M,K], Wq [N,K]), which is the native layout of
// `tcgen05.mma.kind::i8` with K-major shared-memory descriptors.
//
// Structure (one persistent CTA -- or CTA pair -- per SM, 384 threads):
//   warp 0   TMA producer: 128-byte-swizzled [rows x 128 B] boxes of A and B into a
//            STAGES-deep shared-memory ring, completion on `full` mbarriers.
//   warp 1   MMA issuer (one thread): 4 x tcgen05.mma (K=32 each) per ring slot into a
//            TMEM accumulator; tcgen05.commit releases the slot (`empty`) and, after the
//            last K block, publishes the accumulator (`tmem_full`).
//   warp 2   TMEM allocate / free (2 accumulator buffers of BLOCK_N columns).
//   warps 4-11 epilogue: tcgen05.ld 32 lanes x 32 columns, scale/bias/cast, 16-byte
//            global stores (warp w reads TMEM lane quarter w%4; warps 4-7 take the left
//            half of the tile's columns, 8-11 the right half); then hands the accumulator
//            back (`tmem_empty`) so the MMA warp is already filling the other buffer.
// CG == 2 runs the same roles on a 2-CTA cluster with `cta_group::2`: the pair shares a
// 256 x BLOCK_N tile, each CTA stages its own 128 rows of A and its half of B, the even
// CTA issues the MMAs for both and multicasts the commits.
#include "gemm_common.cuh"
#include <algorithm>
#include <string.h>

namespace pq {
int launch_qgemm_smallm(const int8_t* a, int64_t lda, const int8_t* b, int64_t ldb,
                        const float* s_x, const float* s_w, const float* bias,
                        void* const* outs, int n_out, int out_dtype, int64_t ldo,
                        int64_t M, int64_t N, int64_t K, int num_sms, cudaStream_t stream);
namespace {

using namespace ptx;
using namespace gemm;

constexpr int BLOCK_M = 128;   // rows of A per CTA (= TMEM lanes)
constexpr int NUM_THREADS = 384;   // 4 control warps + 8 epilogue warps
constexpr int EPI_WARP0 = 4;
constexpr int EPI_THREADS = 256;   // two warps per TMEM lane quarter, each takes half of the columns
constexpr int GROUP_M = 16;    // tile rasterisation: m-blocks per L2 swizzle group

struct GemmArgs {
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  const float* s_x;
  const float* s_w;
  const float* bias;
  void* out[8];   // the [M,N] result is stored into each of out[0..n_out): the local buffer, peer
  int n_out;      // buffers mapped over NVLink, or one NVSwitch multicast address (fused all-gather)
  long long ldo;  // elements
  int vec_ok;     // rows of `out` are 16-byte aligned
  // stream-K (exact: int32 partial sums are associative). Null workspace = data-parallel tiles.
  int32_t* sk_ws;     // [workers][2][CG][BN*128] int32 partial sums
  int* sk_flags;      // [workers*CG][2] ticket / done counters (self-cleaning, zero between launches)
  unsigned long long* tl;   // debug timeline (32 x u64 per CTA) or null
  int prefetch_b;           // 1 = warp 3 prefetches this CTA's weight boxes into L2 ahead of the ring
  int tma_store;            // staged epilogue hands its tiles to TMA bulk stores (single destination)
  int dbg;                  // profiling only (pq_debug_set_epilogue): bit 0 = the epilogue skips its global stores
  int multimem;             // out[0] is an NVSwitch multicast address: the LSU copy-out uses multimem.st
  int n_rot;                // first column block of the tile order (see tile_coords), 0 <= n_rot < num_n_blocks
  int scatter_cols;         // > 0 (staged epilogue): columns [d*scatter_cols, (d+1)*scatter_cols) go to out[d] ONLY, as a
                            // [M, scatter_cols] matrix with row stride ldo (fused GEMM + reduce-scatter: out[d] is this
                            // rank's inbox on the rank that owns those output columns)
};

// WS_BYTES: per-warp TMA-store staging of the direct epilogue: every epilogue warp owns WS_NBUF boxes of
// [32 rows x 32 columns] (64-byte rows for 16-bit outputs, 128-byte rows for fp32); 0 = no staging (int32
// outputs, or tile configurations whose operand ring leaves no room, keep per-lane global stores).
// OUT_BYTES: 2 = bf16 / fp16, 4 = fp32, 0 = int32 accumulators.
constexpr int NUM_BARS_C(int stages) { return 2 * stages + 4; }

// Output tensor maps: one per destination.  Both TMA-store epilogues (per-warp boxes and the CTA-staged tile) can
// write a finished tile to up to 8 destinations -- the local buffer and the peers' buffers over NVLink (fused
// all-gather), or one inbox per column block (reduce-scatter, staged epilogue only).
struct OutMaps { CUtensorMap m[8]; };
template <bool STAGED> struct YMap { using type = OutMaps; };
__device__ __forceinline__ const CUtensorMap* ymap(const OutMaps& t, int d) { return &t.m[d]; }
template <int CG, int BN, int STAGES, bool STAGED = false, int OUT_BYTES = 4>
struct SmemLayout {
  static constexpr int A_STAGE = BLOCK_M * BLOCK_K;
  static constexpr int B_ROWS = BN / CG;
  static constexpr int B_STAGE = B_ROWS * BLOCK_K;
  static constexpr int STAGE_BYTES = A_STAGE + B_STAGE;
  static constexpr int OFF_A = 0;
  static constexpr int OFF_B = OFF_A + STAGES * A_STAGE;
  static constexpr int STAGE_OUT = STAGED ? 2 * 32768 : 0;  // per column-half [128 rows][256 B] output staging
  static constexpr int OFF_STAGE_OUT = OFF_B + STAGES * B_STAGE;
  static constexpr int BNP = (BN + 31) / 32 * 32;           // columns rounded up to whole 32-column chunks
  static constexpr int WS_BOX = 32 * 32 * (OUT_BYTES == 4 ? 4 : 2);   // one warp's box: 32 rows x 32 columns
  static constexpr int WS_FIXED = STAGES * (A_STAGE + B_STAGE) + 4 * BNP * 4 + NUM_BARS_C(STAGES) * 8 + 16 + 1024;
  static constexpr int WS_NBUF = (OUT_BYTES == 0 || STAGED) ? 0 : (WS_FIXED + 2 * 8 * WS_BOX <= 227 * 1024) ? 2
                                 : (WS_FIXED + 8 * WS_BOX <= 227 * 1024) ? 1 : 0;
  static constexpr int WS_BYTES = WS_NBUF * 8 * WS_BOX;
  static constexpr int OFF_WS = OFF_STAGE_OUT + STAGE_OUT;  // 1024-aligned: every term before it is
  static constexpr int OFF_SW = OFF_WS + WS_BYTES;          // [2][BNP] fp32
  static constexpr int OFF_BIAS = OFF_SW + 2 * BNP * 4;     // [2][BNP] fp32
  static constexpr int OFF_BAR = OFF_BIAS + 2 * BNP * 4;    // full[S], empty[S], tfull[2], tempty[2]
  static constexpr int NUM_BARS = 2 * STAGES + 4;
  static constexpr int OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
  static constexpr int OFF_MISC = OFF_TMEM_PTR + 8;           // stream-K ticket broadcast
  static constexpr int TOTAL = OFF_TMEM_PTR + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024-B alignment
  static_assert(DYN_BYTES <= 227 * 1024, "shared memory budget exceeded");
};

// n_rot: the column blocks are visited starting at block n_rot (reduce-scatter: every rank starts at a different
// owner's columns, so at any moment the ranks store to DIFFERENT inboxes instead of all hammering one ingress port)
__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int n_rot, int& m_blk, int& n_blk) {
  const int per_group = GROUP_M * num_n;
  const int group = tile / per_group;
  const int first_m = group * GROUP_M;
  const int gsize = min(GROUP_M, num_m - first_m);
  const int within = tile - group * per_group;
  m_blk = first_m + within % gsize;
  n_blk = within / gsize + n_rot;
  if (n_blk >= num_n) n_blk -= num_n;
}

// Work scheduler shared by the three roles.  A "segment" is a K-range [kb0,kb1) of one tile.
//  data-parallel: worker w owns whole tiles w, w+W, w+2W, ...
//  stream-K:      the T*KB (tile, k-block) units are cut into W equal contiguous ranges.  A
//                 worker first does the (at most two) PARTIAL tiles at the ends of its range and
//                 only then its whole tiles, so every cross-worker fix-up happens early and is
//                 hidden behind the main loops of the whole tiles that follow.
// int32 accumulation is associative, so any split of K is bit-identical to the unsplit GEMM.
struct Sched {
  int KB, num_tiles, tile, stride;
  bool sk;
  long long U, u0, u1;
  int worker, workers, phase, t_full, t_full_end;
  __device__ __forceinline__ void init(const GemmArgs& g, int w, int W) {
    KB = g.num_k_blocks;
    num_tiles = g.num_m_blocks * g.num_n_blocks;
    sk = g.sk_ws != nullptr;
    worker = w; workers = W;
    tile = w; stride = W;
    U = 0; u0 = 0; u1 = 0; phase = 0; t_full = 0; t_full_end = 0;
    if (sk) {
      U = (long long)num_tiles * KB;
      u0 = U * w / W;
      u1 = U * (w + 1) / W;
      t_full = (int)((u0 + KB - 1) / KB);
      t_full_end = (int)(u1 / KB);
    }
  }
  __device__ __forceinline__ bool next(int& t, int& kb0, int& kb1) {
    if (!sk) {
      if (tile >= num_tiles) return false;
      t = tile; kb0 = 0; kb1 = KB; tile += stride;
      return true;
    }
    if (u0 >= u1) return false;
    if (phase == 0) {
      phase = 1;
      const long long tf = u0 / KB;
      if (u0 != tf * KB) {                       // tail (or an inner piece) of the first tile
        const long long te = (tf + 1) * KB;
        t = (int)tf; kb0 = (int)(u0 - tf * KB); kb1 = (int)((u1 < te ? u1 : te) - tf * KB);
        return true;
      }
    }
    if (phase == 1) {
      phase = 2;
      const long long tl = u1 / KB;
      if (u1 != tl * KB && tl * KB >= u0) {      // head of the last tile
        t = (int)tl; kb0 = 0; kb1 = (int)(u1 - tl * KB);
        return true;
      }
    }
    if (t_full < t_full_end) {
      t = t_full++; kb0 = 0; kb1 = KB;
      return true;
    }
    return false;
  }
  // worker whose range contains unit u
  __device__ __forceinline__ int worker_of_unit(long long u) const {
    return (int)(((u + 1) * workers - 1) / U);
  }
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

constexpr int TL_STRIDE = 40;   // u64 slots per CTA in the debug timeline: 0..31 %globaltimer stamps, 32/33 clock64 at start/end
#define PQ_TL(i) do { if (g.tl) g.tl[(size_t)blockIdx.x * TL_STRIDE + (i)] = globaltimer_ns(); } while (0)

// STAGED = true: the epilogue goes through shared memory so that every global store
// instruction writes whole 256-byte row segments (needed for NVLink peer / multicast
// destinations, where 16-byte scattered writes waste most of the link).
// MC == 2: two CTA pairs form a 4-CTA cluster working on a 512 x BN super-tile (pair p takes rows
// p*256..): both pairs need the same weight tile, so each CTA fetches only half of its B slice and
// TMA-multicasts it to the CTA with the same rank in the other pair.  L2 -> SM operand bytes per
// MMA drop by 25 % (the main loop is bound by L2 slice throughput, profiles/README_r1.md).
template <int CG, int BN, int STAGES, typename OutT, bool STAGED = false, int MC = 1>
__global__ void __launch_bounds__(NUM_THREADS, 1)
qgemm_kernel(const __grid_constant__ CUtensorMap tmap_a,
             const __grid_constant__ CUtensorMap tmap_b,
             const __grid_constant__ typename YMap<STAGED>::type tmap_y, const GemmArgs g) {
  using L = SmemLayout<CG, BN, STAGES, STAGED, std::is_same<OutT, int32_t>::value ? 0 : (int)sizeof(OutT)>;
  constexpr bool RAW = std::is_same<OutT, int32_t>::value;
  static_assert(!STAGED || (BN % 32 == 0 && BN >= 64), "staged epilogue works on whole 32-column chunks");
  static_assert(MC == 1 || (MC == 2 && CG == 2 && (BN / 4) % 8 == 0), "multicast clusters are pairs of CTA pairs");
  constexpr int SUPER_M = BLOCK_M * CG * MC;   // rows of one scheduled tile (all CTAs of the cluster)
  constexpr int UMMA_M = BLOCK_M * CG;
  constexpr int UMMA_N = BN;
  // accumulator buffer stride in TMEM columns: BN rounded up to a power of two (so BN = 240 keeps the
  // second buffer at column 256)
  constexpr uint32_t TSTRIDE = (BN <= 64) ? 64 : (BN <= 128) ? 128 : 256;
  constexpr uint32_t TMEM_COLS = 2 * TSTRIDE;
  constexpr int BNP = L::BNP;
  static_assert(BN % 16 == 0 && BN >= 64 && BN <= 256, "BLOCK_N must be a multiple of 16 in [64,256]");
  static_assert(UMMA_N % 16 == 0, "invalid UMMA N");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  // warp index through a shuffle: ptxas then knows it is warp-uniform (role dispatch on uniform branches)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const uint32_t lane = lane_id();
  const uint32_t cl_rank = (CG * MC > 1) ? cluster_ctarank() : 0u;   // rank in the cluster
  const uint32_t cta_rank = cl_rank & (CG - 1);                      // rank inside the MMA pair
  const uint32_t pair_id = cl_rank >> 1;                             // which pair of the cluster (MC == 2)
  const uint32_t leader_rank = pair_id * 2;                          // cluster rank of this pair's MMA-issuing CTA
  const bool leader = (cta_rank == 0);
  const int m_off = (int)pair_id * (BLOCK_M * CG) + (int)cta_rank * BLOCK_M;   // this CTA's rows inside the tile
  const uint16_t pair_mask = (uint16_t)(3u << leader_rank);
  const uint16_t all_mask = (uint16_t)((1u << (CG * MC)) - 1u);

  const uint32_t bar_full = smem_base + L::OFF_BAR;
  const uint32_t bar_empty = bar_full + STAGES * 8;
  const uint32_t bar_tfull = bar_empty + STAGES * 8;
  const uint32_t bar_tempty = bar_tfull + 2 * 8;
  const uint32_t tmem_ptr_smem = smem_base + L::OFF_TMEM_PTR;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + L::OFF_TMEM_PTR);
  float* sw_smem = reinterpret_cast<float*>(smem_gen + L::OFF_SW);
  float* bias_smem = reinterpret_cast<float*>(smem_gen + L::OFF_BIAS);
  volatile int* misc_smem = reinterpret_cast<volatile int*>(smem_gen + L::OFF_MISC);
  volatile int* prod_count = reinterpret_cast<volatile int*>(smem_gen + L::OFF_MISC + 4);   // k-blocks the producer has issued

  griddep_launch_dependents();   // PDL: the next kernel may begin its own prologue
  if (warp == 0 && lane == 0) {
    PQ_TL(0);
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    if (g.tma_store) {
      for (int d = 0; d < g.n_out; ++d) prefetch_tmap(ymap(tmap_y, d));
    }
  }
  if (warp == 1 && lane == 0) {
    *prod_count = 0;
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(bar_full + i * 8, CG);   // one producer arrive per CTA of the pair (leader's copy is used)
      mbar_init(bar_empty + i * 8, MC);  // one tcgen05.commit per pair that reads (or multicasts into) the slot
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
  // PDL: everything above overlapped the previous kernel's tail; from here on we touch global
  // memory that kernel may have produced (xq, s_x) or may still be reading (y).  The L2
  // prefetcher only touches the (static) weights, so it does not wait.
  if (!(warp == 3 && lane == 0)) griddep_wait();

  const int num_clusters = gridDim.x / (CG * MC);
  const int cluster_id = blockIdx.x / (CG * MC);
  Sched sched;
  sched.init(g, cluster_id, num_clusters);
  if (threadIdx.x == 0) { PQ_TL(1); if (g.tl) g.tl[(size_t)blockIdx.x * TL_STRIDE + 32] = (unsigned long long)clock64(); }

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one_sync()) {
      uint32_t stage = 0, phase = 0;
      bool first = true;
      int tile, kb0, kb1;
      while (sched.next(tile, kb0, kb1)) {
        int m_blk, n_blk;
        tile_coords(tile, g.num_m_blocks, g.num_n_blocks, g.n_rot, m_blk, n_blk);
        const int m_idx = m_blk * SUPER_M + m_off;
        const int n_idx = n_blk * BN + (int)cta_rank * L::B_ROWS;
        for (int kb = kb0; kb < kb1; ++kb) {
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
            if (MC == 1) {
              tma_load_2d_2sm(sb, &tmap_b, fb, kb * BLOCK_K, n_idx);
            } else {
              // my half of this rank's B slice, delivered to both pairs
              constexpr int HROWS = L::B_ROWS / 2;
              tma_load_2d_2sm_mcast(sb + pair_id * (HROWS * BLOCK_K), &tmap_b, fb, kb * BLOCK_K,
                                    n_idx + (int)pair_id * HROWS, (uint16_t)(5u << cta_rank));
            }
            if (leader) mbar_arrive_expect_tx(fb, L::STAGE_BYTES * 2);
            else mbar_arrive_remote(fb, leader_rank);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          if (g.prefetch_b) *prod_count = *prod_count + 1;   // only the (off by default) L2 prefetcher reads it
          if (first) { first = false; PQ_TL(2); }
        }
      }
      PQ_TL(3);
    }
  } else if (warp == 3) {
    // ================= L2 prefetcher =================
    // Experiment, OFF by default (pq_debug_set_prefetch): one thread walks the same schedule up to
    // PF_AHEAD k-blocks ahead of the producer and pulls this CTA's B boxes into L2, on the theory
    // that cold weights miss to DRAM for longer than the shared-memory ring can cover.  Measured:
    // the step gets 6-9 % SLOWER (profiles/README_r1.md) -- the kernel is bound by L2 request
    // bandwidth, and the redundant prefetches (8 m-blocks share a B box) add to exactly that.
    if (lane == 0 && g.prefetch_b) {
      constexpr int PF_AHEAD = 24;
      int count = 0;
      int tile, kb0, kb1;
      Sched pf = sched;
      while (pf.next(tile, kb0, kb1)) {
        int m_blk, n_blk;
        tile_coords(tile, g.num_m_blocks, g.num_n_blocks, g.n_rot, m_blk, n_blk);
        const int n_idx = n_blk * BN + (int)cta_rank * L::B_ROWS;
        // mode 1: every CTA prefetches its own boxes; mode 2: one m-block per n-block does it
        const bool duty = (g.prefetch_b == 1) || (m_blk == n_blk % g.num_m_blocks);
        for (int kb = kb0; kb < kb1; ++kb, ++count) {
          while (count - *prod_count > PF_AHEAD) __nanosleep(200);
          if (duty) tma_prefetch_l2_2d(&tmap_b, kb * BLOCK_K, n_idx);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    // One ELECTED lane runs the whole loop.  With the loop under `lane == 0` instead, ptxas did not know the region
    // is single-threaded and sent every UTCIMMA / UTCBAR operand through an ELECT + R2UR.BROADCAST + BRA.U.ANY
    // waterfall (~150 SASS instructions per k-block for the 2-CTA kernel): the issuing thread, not the tensor
    // pipe, paced the main loop (ncu: tensor pipe 69 % active next to cuBLASLt's 84 % with the same tile shape
    // and operand traffic; 2048x11008x4096 66.6 -> 59.2 us, 8192^3 357 -> 323 us once fixed).
    if (leader && elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc(UMMA_M, UMMA_N);
      // descriptor words: the high word is constant; the low word is (address >> 4) of the slot (+2 per 32-byte K step)
      const uint64_t desc_hi = make_smem_desc(0) & 0xFFFFFFFF00000000ull;
      const uint32_t a_lo0 = ((smem_base + L::OFF_A) & 0x3FFFFu) >> 4;
      const uint32_t b_lo0 = ((smem_base + L::OFF_B) & 0x3FFFFu) >> 4;
      uint32_t stage = 0, phase = 0;
      int iter = 0;
      bool first = true;   // debug timeline: stamp of the first k-block whose operands have landed
      int tile, kb0, kb1;
      for (; sched.next(tile, kb0, kb1); ++iter) {
        const uint32_t as = iter & 1, aphase = (iter >> 1) & 1;
        mbar_wait(bar_tempty + as * 8, aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * TSTRIDE;
        uint32_t accumulate = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_full + stage * 8, phase);
          tc_fence_after();
          if (first) { first = false; PQ_TL(4); }
          const uint64_t adesc = desc_hi | (uint64_t)(a_lo0 + stage * (L::A_STAGE >> 4));
          const uint64_t bdesc = desc_hi | (uint64_t)(b_lo0 + stage * (L::B_STAGE >> 4));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 32 bytes along K inside the 128B swizzle atom: +2 in the (addr>>4) field
            mma_i8<CG>(tmem_d, adesc + (uint64_t)(k * (UMMA_K >> 4)), bdesc + (uint64_t)(k * (UMMA_K >> 4)), idesc, accumulate);
            accumulate = 1;
          }
          tc_commit<CG>(bar_empty + stage * 8, all_mask);
          if (kb == kb1 - 1) tc_commit<CG>(bar_tfull + as * 8, pair_mask);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      PQ_TL(5);
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
    const int ew = (warp - EPI_WARP0) & 3;       // == warp % 4: TMEM lane quarter this warp may read
    const int half = (warp - EPI_WARP0) >> 2;    // which half of the tile's columns this warp handles
    const int et = ew * 32 + (int)lane;          // row inside the CTA's 128-row slab
    const int etid = (int)threadIdx.x - EPI_WARP0 * 32;   // 0..255
    constexpr int NCH = BNP / 32;                // 32-column chunks per tile (the last may be half a chunk)
    constexpr int CH = (NCH + 1) / 2;            // chunks of the left-half warps
    const int c_lo = half * CH, c_hi = half ? NCH : CH;
    const bool has_bias = g.bias != nullptr;   // without bias there is no add at all (-0.0 stays -0.0)
    int iter = 0;
    int tile, kb0, kb1;
    for (; sched.next(tile, kb0, kb1); ++iter) {
      int m_blk, n_blk;
      tile_coords(tile, g.num_m_blocks, g.num_n_blocks, g.n_rot, m_blk, n_blk);
      const uint32_t as = iter & 1, aphase = (iter >> 1) & 1;
      const int row = m_blk * SUPER_M + m_off + et;
      const int col0 = n_blk * BN;
      const uint32_t taddr0 = tmem_base + ((uint32_t)(ew * 32) << 16) + as * TSTRIDE;
      const int n_end = min(g.N, col0 + BN);     // first column past this tile
      constexpr long long SLOT = (long long)BNP * BLOCK_M;   // int32 elements per CTA partial
      // ---- stream-K: a split tile is completed by whichever participant finishes last ----
      // Every participant takes a ticket once its own accumulator is complete.  All but the last
      // dump their raw int32 partial into their own workspace slot (plain coalesced stores) and
      // bump `done`; the last one folds the others' partials into its registers and runs the
      // normal epilogue.  A worker has at most two partial segments -> two slots per worker.
      int role = 0;   // 0 whole tile, 1 contributor, 2 combiner
      int* ctr = nullptr;   // ctr[0] = tickets taken, ctr[1] = contributions finished
      int fw = 0, lw = -1;
      const bool split = (kb0 != 0 || kb1 != sched.KB);
      if (split) {
        mbar_wait(bar_tfull + as * 8, aphase);
        if (etid == 0 && iter < 5) PQ_TL(8 + iter * 4 + 0);
        fw = sched.worker_of_unit((long long)tile * sched.KB);
        lw = sched.worker_of_unit((long long)(tile + 1) * sched.KB - 1);
        ctr = g.sk_flags + 2 * (fw * CG + (int)cta_rank);
        if (etid == 0) *misc_smem = atomicAdd(ctr, 1);
        named_bar_sync(1, EPI_THREADS);
        role = (*misc_smem == lw - fw) ? 2 : 1;
        if (etid == 0 && iter < 5) { PQ_TL(8 + iter * 4 + 1); if (g.tl) g.tl[(size_t)blockIdx.x * TL_STRIDE + 28 + (iter & 3)] = (unsigned long long)role * 1000 + (lw - fw + 1); }
      }
      if (role == 1) {
        tc_fence_after();
        __syncwarp();
        int32_t* slot = g.sk_ws + (((long long)sched.worker * 2 + (kb0 != 0 ? 0 : 1)) * CG + cta_rank) * SLOT;
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr0 + c * 32, r);
          tmem_ld_wait();
          if (row < g.M && col0 + c * 32 < n_end) {
#pragma unroll
            for (int j = 0; j < 32; ++j) __stcg(slot + (c * 32 + j) * BLOCK_M + et, (int)r[j]);
          }
        }
        tc_fence_before();
        if (CG == 1 || leader) mbar_arrive(bar_tempty + as * 8);
        else mbar_arrive_remote(bar_tempty + as * 8, leader_rank);
        __threadfence();
        named_bar_sync(1, EPI_THREADS);
        if (etid == 0) atomicAdd(ctr + 1, 1);
        if (etid == 0 && iter < 5) PQ_TL(8 + iter * 4 + 3);
        continue;
      }
      if (role == 2) {
        if (etid == 0) {
          uint32_t polls = 0;
          uint64_t t0 = 0;
          while (ld_acquire_gpu(ctr + 1) != lw - fw) {
            if ((++polls & 1023u) == 0) {
              const uint64_t now = globaltimer_ns();
              if (t0 == 0) t0 = now;
              else if (now - t0 > PQ_MBAR_TIMEOUT_NS) {
                printf("pq: stream-K wait timed out (block %d tile %d)\n", (int)blockIdx.x, tile);
                __trap();
              }
            }
          }
          ctr[0] = 0;   // self-cleaning: the next launch finds the counters at zero
          ctr[1] = 0;
          if (iter < 5) PQ_TL(8 + iter * 4 + 2);
        }
        named_bar_sync(1, EPI_THREADS);
      }
      float sx = 0.f;
      if constexpr (!RAW) {
        // stage this tile's column scales / bias (double-buffered by accumulator stage)
        float* sw = sw_smem + as * BNP;
        float* bs = bias_smem + as * BNP;
        for (int i = etid; i < BNP; i += EPI_THREADS) {
          const int c = col0 + i;
          sw[i] = (c < g.N) ? __ldg(g.s_w + c) : 0.f;
          bs[i] = (g.bias != nullptr && c < g.N) ? __ldg(g.bias + c) : 0.f;
        }
        if (row < g.M) sx = __ldg(g.s_x + row);
        named_bar_sync(1, EPI_THREADS);
      }
      // (Letting one lane per warp poll and the others pass the completed phase afterwards was measured
      // 2-3 % SLOWER under the power cap: 8192^3 2576 vs 2636 TOPS sustained -- all lanes wait.)
      mbar_wait(bar_tfull + as * 8, aphase);
      tc_fence_after();
      __syncwarp();
      if (etid == 0 && iter < 5 && !split) PQ_TL(8 + iter * 4 + 0);
      if constexpr (STAGED) {
        // Staging tile of one column half and one pass: two sub-boxes of [128 rows x 128 B], each
        // laid out exactly like a 128B-swizzled TMA box (16-byte unit u of row r sits at u ^ (r & 7)).
        using OT = typename std::conditional<RAW, float, OutT>::type;
        constexpr int ROWB = 256;                                  // staged bytes per row and pass
        constexpr int ESZ = (int)sizeof(OT);
        constexpr int COLS_PASS = ROWB / ESZ;                      // 128 (16-bit) or 64 (fp32)
        constexpr int CH_PASS = COLS_PASS / 32;
        // The left-half warps own chunks [0, CH), the right-half warps [CH, NCH) (CH = ceil(NCH / 2)): for BLOCK_N =
        // 256 that is 128 + 128 columns, for 224 it is 128 + 96, for 128 it is 64 + 64.  Chunks >= c_hi do not exist.
        constexpr int PASSES = (CH * 32 + COLS_PASS - 1) / COLS_PASS;
        constexpr int UPC = OutPack<OT>::WORDS / 4;                // 16-byte units per 32-column chunk
        constexpr int EPU = 16 / ESZ;                              // elements per 16-byte unit
        constexpr int SUB = 16384;                                 // bytes per sub-box
        uint8_t* stg = smem_gen + L::OFF_STAGE_OUT + half * 32768;
        const uint32_t stg_u32 = smem_base + L::OFF_STAGE_OUT + half * 32768;
        const int gt = etid & 127;                                 // thread index inside this column half
        const int m0 = m_blk * SUPER_M + m_off;
        const bool tma_out = g.tma_store != 0;
#pragma unroll 1
        for (int pass = 0; pass < PASSES; ++pass) {
          if (tma_out) {
            // the bulk stores issued from this staging tile must have finished READING it
            if (gt == 0) tma_store_wait_read<0>();
            if (half == 0) named_bar_sync(2, 128); else named_bar_sync(3, 128);
          }
#pragma unroll 1
          for (int cc = 0; cc < CH_PASS; ++cc) {
            const int c = c_lo + pass * CH_PASS + cc;
            if (c >= c_hi) break;                                  // narrower right half (warp-uniform)
            uint32_t r[32];
            tmem_ld_32x32(taddr0 + c * 32, r);
            tmem_ld_wait();
            uint32_t o[OutPack<OT>::WORDS];
            if constexpr (RAW) {
#pragma unroll
              for (int j = 0; j < 32; ++j) o[j] = r[j];     // exact int32 partial sums
            } else {
              const float* sw = sw_smem + as * BNP + c * 32;
              const float* bs = bias_smem + as * BNP + c * 32;
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
                  if (has_bias) v = __fadd_rn(v, bv[j]);
                  f[4 * j4 + j] = v;
                }
              }
              OutPack<OT>::pack(f, o);
            }
#pragma unroll
            for (int i = 0; i < UPC; ++i) {
              const int u = cc * UPC + i;                          // unit inside the 256-byte pass row
              *reinterpret_cast<uint4*>(stg + (u >> 3) * SUB + et * 128 + (((u & 7) ^ (et & 7)) << 4)) =
                  make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
            }
          }
          if (pass == PASSES - 1) {
            // accumulator fully read: give the TMEM buffer back before the copy-out
            tc_fence_before();
            if (CG == 1 || leader) mbar_arrive(bar_tempty + as * 8);
            else mbar_arrive_remote(bar_tempty + as * 8, leader_rank);
          }
          const int colp = col0 + c_lo * 32 + pass * COLS_PASS;
          // columns this half staged in this pass (fewer than COLS_PASS for the narrower halves of BLOCK_N < 256)
          const int cols_here = min((c_hi - c_lo) * 32 - pass * COLS_PASS, COLS_PASS);
          if (tma_out) {
            // one thread hands both sub-boxes to the TMA store engine: full-line writes, rows >= M and
            // columns >= N are clipped by the tensor map
            fence_proxy_async_smem();
            if (half == 0) named_bar_sync(2, 128); else named_bar_sync(3, 128);
            if (gt == 0 && m0 < g.M) {
              constexpr int SUBC = 128 / ESZ;                      // columns per sub-box
              if (g.scatter_cols > 0) {
                // reduce-scatter: a sub-box belongs to exactly one destination (scatter_cols % SUBC == 0)
#pragma unroll
                for (int sb = 0; sb < 2; ++sb) {
                  const int cs = colp + sb * SUBC;
                  if (cs < g.N) {
                    const int d = cs / g.scatter_cols;
                    tma_store_2d(ymap(tmap_y, d), stg_u32 + sb * SUB, cs - d * g.scatter_cols, m0);
                  }
                }
              } else {
                // all-gather: the same tile goes to every destination (local buffer + NVLink peers)
                for (int d = 0; d < g.n_out; ++d) {
                  if (colp < g.N) tma_store_2d(ymap(tmap_y, d), stg_u32, colp, m0);
                  if (colp + SUBC < g.N) tma_store_2d(ymap(tmap_y, d), stg_u32 + SUB, colp + SUBC, m0);
                }
              }
              tma_store_commit();
            }
          } else {
            if (half == 0) named_bar_sync(2, 128); else named_bar_sync(3, 128);
            // LSU copy-out: a warp instruction moves two whole 256-byte row segments per destination
#pragma unroll 4
            for (int i = 0; i < 16; ++i) {
              const int idx = i * 128 + gt;
              const int rr = idx >> 4, u = idx & 15;
              const int grow = m0 + rr;
              const int gcol = colp + u * EPU;
              if (grow < g.M && gcol < n_end && u * EPU < cols_here) {
                const uint4 v = *reinterpret_cast<const uint4*>(stg + (u >> 3) * SUB + rr * 128 + (((u & 7) ^ (rr & 7)) << 4));
                if (g.scatter_cols > 0) {
                  // reduce-scatter: this 16-byte unit belongs to exactly one destination (scatter_cols % EPU == 0)
                  const int d = gcol / g.scatter_cols;
                  OT* dst = reinterpret_cast<OT*>(g.out[d]) + (long long)grow * g.ldo + (gcol - d * g.scatter_cols);
                  if (g.vec_ok && gcol + EPU <= n_end) {
                    *reinterpret_cast<uint4*>(dst) = v;
                  } else {
                    const OT* ev = reinterpret_cast<const OT*>(&v);
#pragma unroll
                    for (int e = 0; e < EPU; ++e)
                      if (gcol + e < n_end) dst[e] = ev[e];
                  }
                } else if (g.vec_ok && gcol + EPU <= n_end) {
                  if (g.multimem) {
                    multimem_st_v4(reinterpret_cast<OT*>(g.out[0]) + (long long)grow * g.ldo + gcol, v.x, v.y, v.z, v.w);
                  } else {
                    for (int d = 0; d < g.n_out; ++d)
                      *reinterpret_cast<uint4*>(reinterpret_cast<OT*>(g.out[d]) + (long long)grow * g.ldo + gcol) = v;
                  }
                } else {
                  const OT* ev = reinterpret_cast<const OT*>(&v);
                  for (int d = 0; d < g.n_out; ++d) {
                    OT* dst = reinterpret_cast<OT*>(g.out[d]) + (long long)grow * g.ldo + gcol;
#pragma unroll
                    for (int e = 0; e < EPU; ++e)
                      if (gcol + e < n_end) dst[e] = ev[e];
                  }
                }
              }
            }
            if (half == 0) named_bar_sync(2, 128); else named_bar_sync(3, 128);
          }
        }
        continue;
      }
      if constexpr (L::WS_NBUF > 0) {
        if (g.tma_store) {
          // ---- 16-bit outputs: per-warp TMA-store epilogue ----
          // Every warp converts one 32-column chunk at a time into its own [32 rows x 64 B] box (laid out
          // like a 64B-swizzled TMA box, conflict-free 16-byte shared stores) and hands it to the TMA
          // store engine: full-sector writes and ~40x fewer LSU wavefronts than one-row-per-lane global
          // stores, which were measured to slow the concurrent main loop (profiles/README_r1.md).  M and
          // N tails are clipped by the tensor map; a tile whose width is not a multiple of 32 re-stores
          // the overlap of its last two chunks (same values).
          const int wslot = warp - EPI_WARP0;
          const uint32_t ws_u32 = smem_base + L::OFF_WS + wslot * (L::WS_NBUF * L::WS_BOX);
          uint8_t* ws_gen = smem_gen + L::OFF_WS + wslot * (L::WS_NBUF * L::WS_BOX);
          const int row0 = m_blk * SUPER_M + m_off + ew * 32;
          const float* swb = sw_smem + as * BNP;
          const float* bsb = bias_smem + as * BNP;
#pragma unroll 1
          for (int c = c_lo; c < c_hi; ++c) {
            const int ct = min(c * 32, BN - 32);   // tile-relative first column of this chunk
            uint32_t r[32];
            tmem_ld_32x32(taddr0 + ct, r);
            tmem_ld_wait();
            if (c == c_hi - 1) {                   // accumulator fully read by this thread
              tc_fence_before();
              if (CG == 1 || leader) mbar_arrive(bar_tempty + as * 8);
              else mbar_arrive_remote(bar_tempty + as * 8, leader_rank);
            }
            if (role == 2 && row < g.M && col0 + ct < n_end) {
              for (int q = fw; q <= lw; ++q) {
                if (q == sched.worker) continue;
                const int which = (sched.U * q / sched.workers > (long long)tile * sched.KB) ? 0 : 1;
                const int32_t* ps = g.sk_ws + (((long long)q * 2 + which) * CG + cta_rank) * SLOT + ct * BLOCK_M + et;
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] += (uint32_t)__ldcg(ps + j * BLOCK_M);
              }
            }
            float f[32];
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 w4 = *reinterpret_cast<const float4*>(swb + ct + 4 * j4);
              const float4 b4 = *reinterpret_cast<const float4*>(bsb + ct + 4 * j4);
              const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
              const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float v = __int2float_rn((int)r[4 * j4 + j]);
                v = __fmul_rn(v, sx);
                v = __fmul_rn(v, wv[j]);
                if (has_bias) v = __fadd_rn(v, bv[j]);
                f[4 * j4 + j] = v;
              }
            }
            using OT = typename std::conditional<RAW, float, OutT>::type;
            uint32_t o[OutPack<OT>::WORDS];
            OutPack<OT>::pack(f, o);
            const int buf = (L::WS_NBUF == 2) ? ((c - c_lo) & 1) : 0;
            if (lane == 0) tma_store_wait_read<(L::WS_NBUF > 1) ? L::WS_NBUF - 1 : 0>();   // the box is free again
            __syncwarp();
            // row pitch 64 B (16-bit, 64B swizzle: unit ^ ((row >> 1) & 3)) or 128 B (fp32, 128B swizzle: unit ^ (row & 7))
            constexpr int UNITS = OutPack<OT>::WORDS / 4;
            uint8_t* box = ws_gen + buf * L::WS_BOX + lane * (UNITS * 16);
            const int swz = (UNITS == 4) ? (((int)lane >> 1) & 3) : ((int)lane & 7);
#pragma unroll
            for (int i = 0; i < UNITS; ++i)
              *reinterpret_cast<uint4*>(box + ((i ^ swz) << 4)) =
                  make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (row0 < g.M && col0 + ct < g.N && !(g.dbg & 1)) {
                for (int d = 0; d < g.n_out; ++d)      // n_out > 1: the local buffer and every NVLink peer (fused all-gather)
                  tma_store_2d(ymap(tmap_y, d), ws_u32 + buf * L::WS_BOX, col0 + ct, row0);
              }
              tma_store_commit();
            }
          }
          if (etid == 0 && iter < 5) PQ_TL(8 + iter * 4 + 3);
          continue;
        }
      }
#pragma unroll 1
      for (int c = c_lo; c < c_hi; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(taddr0 + c * 32, r);
        tmem_ld_wait();
        const int col = col0 + c * 32;
        const int lim = (g.dbg & 1) ? 0 : n_end - col;   // valid columns of this chunk (>= 32: all of them)
        if (row < g.M && lim > 0) {
          if (role == 2) {
            // fold in the other participants' partial sums (slot 0 = their first-tile tail,
            // slot 1 = their last-tile head)
            for (int q = fw; q <= lw; ++q) {
              if (q == sched.worker) continue;
              const int which = (sched.U * q / sched.workers > (long long)tile * sched.KB) ? 0 : 1;
              const int32_t* ps = g.sk_ws + (((long long)q * 2 + which) * CG + cta_rank) * SLOT + (c * 32) * BLOCK_M + et;
#pragma unroll
              for (int j = 0; j < 32; ++j) r[j] += (uint32_t)__ldcg(ps + j * BLOCK_M);
            }
          }
          if constexpr (RAW) {
            for (int d = 0; d < g.n_out; ++d) {
              int32_t* dst = reinterpret_cast<int32_t*>(g.out[d]) + (long long)row * g.ldo + col;
              if (g.vec_ok && lim >= 32) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  reinterpret_cast<uint4*>(dst)[i] = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
              } else if (g.vec_ok && lim == 16) {   // half chunk at the end of a BN % 32 == 16 tile
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  reinterpret_cast<uint4*>(dst)[i] = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (j < lim) dst[j] = (int32_t)r[j];
              }
            }
          } else {
            using OT = typename std::conditional<RAW, float, OutT>::type;
            const float* sw = sw_smem + as * BNP + c * 32;
            const float* bs = bias_smem + as * BNP + c * 32;
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
                if (has_bias) v = __fadd_rn(v, bv[j]);
                f[4 * j4 + j] = v;
              }
            }
            if (g.vec_ok && (lim >= 32 || lim == 16)) {
              uint32_t o[OutPack<OT>::WORDS];
              OutPack<OT>::pack(f, o);
              for (int d = 0; d < g.n_out; ++d) {
                OT* dst = reinterpret_cast<OT*>(g.out[d]) + (long long)row * g.ldo + col;
#pragma unroll
                for (int i = 0; i < OutPack<OT>::WORDS / 4; ++i)
                  if (i < OutPack<OT>::WORDS / 8 || lim >= 32)   // lim == 16: first half of the chunk only
                    reinterpret_cast<uint4*>(dst)[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
              }
            } else {
              for (int d = 0; d < g.n_out; ++d) {
                OT* dst = reinterpret_cast<OT*>(g.out[d]) + (long long)row * g.ldo + col;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (j < lim) dst[j] = OutPack<OT>::one(f[j]);
              }
            }
          }
        }
      }
      if (etid == 0 && iter < 5) PQ_TL(8 + iter * 4 + 3);
      // accumulator buffer fully read: hand it back to the MMA warp (leader CTA's barrier)
      tc_fence_before();
      if (CG == 1 || leader) mbar_arrive(bar_tempty + as * 8);
      else mbar_arrive_remote(bar_tempty + as * 8, leader_rank);
    }
    if constexpr (STAGED) {
      if (g.tma_store && (etid & 127) == 0) {
        tma_store_wait<0>();   // staging smem must outlive the bulk stores; peer writes are complete after this
        if (g.n_out > 1 || g.scatter_cols > 0) __threadfence_system();
      }
    } else if constexpr (L::WS_NBUF > 0) {
      if (g.tma_store && lane == 0) {
        tma_store_wait<0>();
        if (g.n_out > 1) __threadfence_system();   // peer writes are complete and ordered before the kernel ends
      }
    }
  }

  __syncwarp();
  if (threadIdx.x == 0) { PQ_TL(6); if (g.tl) g.tl[(size_t)blockIdx.x * TL_STRIDE + 33] = (unsigned long long)clock64(); }
  tc_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();
  if (warp == 2) tmem_dealloc<CG>(tmem_base, TMEM_COLS);
  if (threadIdx.x == 0) PQ_TL(7);
}

// ---- stream-K workspace pool ----------------------------------------------------------
// One slot = partial-accumulator space for every worker CTA (num_sms x 2 x 128 x 256 int32) plus
// two counters per CTA.  A slot is bound to the first stream that uses it, so kernels that can run
// concurrently (different streams) never share partials.  Slots are allocated outside stream
// capture only; one spare slot is kept ready so that a capturing stream (CUDA graphs) can be
// bound without calling cudaMalloc.  If no slot can be had the launch silently uses the
// data-parallel schedule (same results, int32 accumulation is exact either way).
constexpr int SK_MAX_SLOTS = 8;
struct SkSlot { int32_t* ws; int* flags; cudaStream_t stream; bool bound; };
struct SkPool {
  std::mutex mu;
  SkSlot slots[SK_MAX_SLOTS];
  int count = 0;
};
SkPool g_sk_pool[64];
std::atomic<unsigned long long*> g_timeline{nullptr};
Knob g_tma_store{1};     // staged epilogue uses TMA bulk stores when it can (pq_debug_set_tma_store)
Knob g_multi_tma{0};     // multi-destination (NVLink) epilogues: 0 = CTA-staged tile + LSU 256-byte stores (default: measured
                         // fastest over NVLink, profiles/README_r2.md); 1 = CTA-staged tile + TMA stores; 2 = per-warp TMA
                         // boxes with the regular tile heuristic (pq_debug_set_multi_tma)
Knob g_sk_mode{-1};  // -1 heuristic (default: single-wave long-K problems only), 0 never, 1 whenever legal

bool sk_alloc_slot(SkPool& pool, int num_sms) {
  if (pool.count >= SK_MAX_SLOTS) return false;
  const size_t ws_bytes = (size_t)num_sms * 2 * BLOCK_M * 256 * sizeof(int32_t);
  void* ws = nullptr;
  void* fl = nullptr;
  if (cudaMalloc(&ws, ws_bytes) != cudaSuccess) { (void)cudaGetLastError(); return false; }
  if (cudaMalloc(&fl, (size_t)num_sms * 2 * sizeof(int)) != cudaSuccess ||
      cudaMemset(fl, 0, (size_t)num_sms * 2 * sizeof(int)) != cudaSuccess) {
    (void)cudaGetLastError();
    cudaFree(ws);
    if (fl) cudaFree(fl);
    return false;
  }
  SkSlot& sl = pool.slots[pool.count++];
  sl.ws = (int32_t*)ws; sl.flags = (int*)fl; sl.stream = nullptr; sl.bound = false;
  return true;
}

// returns the slot bound to `st` (binding / allocating if possible), or nullptr
SkSlot* sk_get_slot(int dev, int num_sms, cudaStream_t st) {
  SkPool& pool = g_sk_pool[dev];
  std::lock_guard<std::mutex> lk(pool.mu);
  for (int i = 0; i < pool.count; ++i)
    if (pool.slots[i].bound && pool.slots[i].stream == st) return &pool.slots[i];
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
  const bool capturing = cs != cudaStreamCaptureStatusNone;
  SkSlot* free_slot = nullptr;
  for (int i = 0; i < pool.count; ++i)
    if (!pool.slots[i].bound) { free_slot = &pool.slots[i]; break; }
  if (!free_slot && !capturing && sk_alloc_slot(pool, num_sms)) free_slot = &pool.slots[pool.count - 1];
  if (!free_slot) return nullptr;
  free_slot->bound = true;
  free_slot->stream = st;
  if (!capturing) {   // keep one spare ready for a future capturing stream
    bool spare = false;
    for (int i = 0; i < pool.count; ++i) spare |= !pool.slots[i].bound;
    if (!spare) sk_alloc_slot(pool, num_sms);
  }
  return free_slot;
}

template <int CG, int BN, int STAGES, typename OutT, bool STAGED = false, int MC = 1>
int launch_cfg(const int8_t* a, int64_t lda, const int8_t* b, int64_t ldb, const GemmArgs& g0,
               int num_sms, cudaStream_t st) {
  using L = SmemLayout<CG, BN, STAGES, STAGED, std::is_same<OutT, int32_t>::value ? 0 : (int)sizeof(OutT)>;
  GemmArgs g = g0;
  g.num_m_blocks = (g.M + BLOCK_M * CG * MC - 1) / (BLOCK_M * CG * MC);
  g.num_n_blocks = (g.N + BN - 1) / BN;
  g.n_rot = (g0.n_rot / BN) % g.num_n_blocks;          // columns -> column blocks of this tile shape
  g.num_k_blocks = (g.K + BLOCK_K - 1) / BLOCK_K;
  CUtensorMap ta, tb;
  int rc = make_tmap(&ta, a, g.M, g.K, lda, BLOCK_M);
  if (rc) return rc;
  rc = make_tmap(&tb, b, g.N, g.K, ldb, L::B_ROWS / MC);
  if (rc) return rc;

  typename YMap<STAGED>::type ty;
  memset(&ty, 0, sizeof(ty));
  g.tma_store = 0;
  const bool tma_ok = g.vec_ok && g_tma_store.load(std::memory_order_relaxed) != 0;
  if constexpr (STAGED) {
    // output(s) as [M rows] x [N elements], boxes of 128 bytes x 128 rows, same 128B swizzle as the staging tile.
    // One map per destination: the local buffer and (fused all-gather / reduce-scatter) the peers' buffers, which
    // TMA writes over NVLink in full lines.  A reduce-scatter destination is an [M, scatter_cols] matrix of its own.
    constexpr int esz = (int)sizeof(OutT);
    constexpr int subc = 128 / esz;
    const bool multi = g.n_out > 1 || g.scatter_cols > 0;
    // (TMA boxes are whole 128-byte sub-boxes: only 256-wide tiles fill both halves exactly)
    const bool want = BN == 256 && tma_ok && !g.multimem && (!multi || g_multi_tma.load(std::memory_order_relaxed) != 0) &&
                      (g.scatter_cols == 0 || g.scatter_cols % subc == 0);
    if (want) {
      bool ok = true;
      for (int d = 0; d < g.n_out && ok; ++d) {
        uint64_t cols = (uint64_t)g.N;
        if (g.scatter_cols > 0) {
          const long long left = (long long)g.N - (long long)d * g.scatter_cols;
          if (left <= 0) { ty.m[d] = ty.m[0]; continue; }      // destination past the last column: never addressed
          cols = (uint64_t)(left < g.scatter_cols ? left : g.scatter_cols);
        }
        ok = encode_tmap_2d(&ty.m[d], g.out[d], esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32,
                            cols, (uint64_t)g.M, (uint64_t)(g.ldo * esz), (uint32_t)subc, (uint32_t)BLOCK_M,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE) == PQ_OK;
      }
      if (ok) g.tma_store = 1;
    }
  } else if (L::WS_NBUF > 0 && tma_ok) {
    // per-warp epilogue boxes: [32 rows] x [32 columns]; 64B swizzle for 16-bit outputs, 128B swizzle for fp32
    constexpr int esz = (int)sizeof(OutT);
    bool ok = true;
    for (int d = 0; d < g.n_out && ok; ++d)
      ok = encode_tmap_2d(&ty.m[d], g.out[d], esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32,
                          (uint64_t)g.N, (uint64_t)g.M, (uint64_t)(g.ldo * esz), 32, 32,
                          esz == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_NONE) == PQ_OK;
    if (ok) g.tma_store = 1;
  }
  // several destinations without a TMA path would fall back to one-row-per-lane 16-byte peer stores (measured 4x
  // slower over NVLink than coalesced segments): the caller routes those launches to the staged epilogue instead
  if constexpr (!STAGED) {
    if (g.n_out > 1 && !g.tma_store) {
      if (g.M > 128) return launch_cfg<2, 256, 4, OutT, true>(a, lda, b, ldb, g0, num_sms, st);
      return launch_cfg<1, 256, 3, OutT, true>(a, lda, b, ldb, g0, num_sms, st);
    }
  }
  auto kern = qgemm_kernel<CG, BN, STAGES, OutT, STAGED, MC>;
  static PerDeviceOnce once;   // function attributes are per device: set them on every device this process drives
  int max_clusters = 0;        // co-resident clusters of this kernel (matters for 4-CTA clusters: 33, not 37)
  const cudaError_t attr_err = once.run([&](int* value) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES);
    *value = num_sms / (CG * MC);
    if (e == cudaSuccess && MC > 1) {
      cudaLaunchConfig_t oc = {};
      oc.gridDim = dim3((unsigned)(num_sms / (CG * MC) * (CG * MC)), 1, 1);
      oc.blockDim = dim3(NUM_THREADS, 1, 1);
      oc.dynamicSmemBytes = L::DYN_BYTES;
      cudaLaunchAttribute oa[1];
      oa[0].id = cudaLaunchAttributeClusterDimension;
      oa[0].val.clusterDim.x = CG * MC;
      oa[0].val.clusterDim.y = 1;
      oa[0].val.clusterDim.z = 1;
      oc.attrs = oa;
      oc.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &oc) == cudaSuccess && n > 0) *value = n;
      else (void)cudaGetLastError();
    }
    return e;
  }, &max_clusters);
  if (attr_err != cudaSuccess)
    PQ_FAIL(PQ_ERR_CUDA, "cudaFuncSetAttribute(smem=%d) failed: %s", L::DYN_BYTES, cudaGetErrorString(attr_err));

  const long long tiles = (long long)g.num_m_blocks * g.num_n_blocks;
  long long clusters = max_clusters;
  // Stream-K when whole tiles would leave SMs idle: fewer tiles than workers (small M: the
  // weight stream is spread over all SMs) or a ragged last wave.
  {
    const long long W = num_sms / CG;
    const long long units = tiles * g.num_k_blocks;
    const long long waves = (tiles + W - 1) / W;
    const double eff = (double)tiles / (double)(waves * W);
    // Heuristic (mode -1, default): stream-K pays only when the whole problem keeps at most HALF of the workers
    // busy with whole tiles and K is long (>= 8192) -- each worker then streams K/S of a tile and the int32 fix-up
    // (one partial write + read per worker) is small next to it.  Measured (gpurun_out/sweep_tune_r2.csv, round 2):
    // 128x1024x16384 29.1 -> 17.7 us, 256x2048x16384 26.5 -> 20.3 us, 256x1024x8192 14.9 -> 13.9 us; with more than
    // half a wave of tiles the fix-up traffic loses: 1024x4096x8192 25.1 -> 36.5 us, 512x8192x8192 26.0 -> 36.4 us.
    (void)eff;
    bool want = !STAGED && MC == 1 && ((g_sk_mode == 1) || (g_sk_mode < 0 && waves == 1 && tiles * 2 <= W && g.num_k_blocks >= 64));
    long long w_sk = W;
    if (units / 4 < w_sk) w_sk = units / 4;     // at least ~4 K blocks per worker
    if (w_sk < 2 || tiles % w_sk == 0) want = false;
    if (want && g_sk_mode != 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      SkSlot* sl = sk_get_slot(dev, num_sms, st);
      if (sl) {
        g.sk_ws = sl->ws;
        g.sk_flags = sl->flags;
        clusters = w_sk;
      }
    }
  }
  if (g.sk_ws == nullptr && tiles < clusters) clusters = tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters * CG * MC), 1, 1);
  cfg.blockDim = dim3(NUM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = L::DYN_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = CG * MC;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = g_pdl ? 2 : 1;
  PQ_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, ty, g));
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return PQ_OK;
}

Knob g_force_cfg{-1};  // test hook: see pq_debug_set_gemm_config
Knob g_dbg_rot_cols{0};   // pq_debug_set_tile_rotation: first column of the tile order when the caller passes none
Knob g_multi_bn{0};    // multi-destination (staged) epilogues: tile width 256 / 224 / 128, 0 = model (pq_debug_set_multi_bn)
Knob g_force_staged{0};  // test hook: staged epilogue even for a single destination
Knob g_epi_dbg{0};       // profiling only: see GemmArgs::dbg
Knob g_narrow_tiles{1};  // heuristic may pick BLOCK_N in {240, 224, 208} (pq_debug_set_narrow_tiles)
Knob g_prefetch_b{0};    // L2 prefetch of the weight operand (pq_debug_set_prefetch): measured SLOWER, off

template <typename OutT>
int launch_typed(const int8_t* a, int64_t lda, const int8_t* b, int64_t ldb, const GemmArgs& g,
                 int num_sms, cudaStream_t st) {
  // Tile configuration heuristic.
  //   cfg 0: 1-CTA 128x256   cfg 1: 2-CTA 256x256   cfg 2: 1-CTA 128x128   cfg 3: 1-CTA 128x64
  //   cfg 16: 4-CTA multicast cluster, 512x256 super-tiles (experiment: no faster, see README_r1.md)
  //   cfg 4: 2-CTA 256x128   cfg 8/9/10: 2-CTA 256x{240,224,208}   cfg 11/12/13: 1-CTA 128x{240,224,208}
  // Fused all-gather (n_out > 1) / reduce-scatter: the CTA-staged epilogue (256-column tiles) writes whole 256-byte row
  // segments to every destination with LSU stores.  Measured on 8 x B200 (2048 x 3584 x 8192 shard, 103 MB to 7 peers):
  // LSU 256-byte segments 208 us, multimem.st 210 us, CTA-staged TMA stores (128-byte rows) 305 us, per-warp TMA boxes
  // (64-byte rows, any tile shape: g_multi_tma == 2) 336 us -- TMA writes reach NVLink as small requests, so the TMA
  // variants stay behind the knob.
  const bool per_warp_multi = g.n_out > 1 && g.scatter_cols == 0 && !g.multimem && sizeof(OutT) == 2 && g.vec_ok &&
                              g_multi_tma == 2 && g_tma_store != 0 && g.M > 128 && g_force_cfg < 0;
  if ((g.n_out > 1 && !per_warp_multi) || g.scatter_cols > 0 || g.multimem ||
      (g_force_staged && !std::is_same<OutT, int32_t>::value)) {
    // fused all-gather / reduce-scatter: coalesced (shared-memory staged) stores to the peer destinations
    if (g.M <= 128) return launch_cfg<1, 256, 3, OutT, true>(a, lda, b, ldb, g, num_sms, st);
    // Tile width (pq_debug_set_multi_bn forces one).  Model per width: the slower of (whole waves of tile pairs) and
    // (first tile pair, before which nothing can be sent, + the bytes this rank stores to its peers at the measured
    // all-to-all NVLink rate).  Compute-bound shards (2 GPUs) want the width with the fewest ragged waves (14336
    // columns: 7 waves of 224 instead of 7 of 256), link-bound ones (>= 4 GPUs) the narrowest first tile.
    int bn = g_multi_bn;
    if (bn != 256 && bn != 224 && bn != 128) {
      const double pair_rate = 3.0e15 / 74.0;                                   // int8 ops/s of one CTA pair (measured ~3000 TOPS)
      const double esz = std::is_same<OutT, float>::value || std::is_same<OutT, int32_t>::value ? 4.0 : 2.0;
      const int peers = g.scatter_cols > 0 ? 0 : (g.multimem ? 1 : g.n_out - 1);
      const double egress = g.scatter_cols > 0 ? (double)g.M * g.N * esz * (g.n_out - 1) / (double)g.n_out   // all but this rank's own block
                                               : (double)peers * g.M * g.N * esz;
      const double t_link = egress / 645.0e9;
      static const int widths[3] = {256, 224, 128};
      static const double eff[3] = {1.0, 1.0, 1.15};                            // narrower tiles load more operand bytes per MMA
      double best = 1e30;
      bn = 256;
      for (int i = 0; i < 3; ++i) {
        const double t_tile = 2.0 * 256.0 * widths[i] * g.K / pair_rate * eff[i];
        const long long tiles = (long long)((g.M + 255) / 256) * ((g.N + widths[i] - 1) / widths[i]);
        const long long waves = (tiles + num_sms / 2 - 1) / (num_sms / 2);
        const double t = std::max((double)waves * t_tile, t_tile + t_link);
        if (t < best * 0.97) { best = t; bn = widths[i]; }
      }
    }
    if (bn == 224) return launch_cfg<2, 224, 5, OutT, true>(a, lda, b, ldb, g, num_sms, st);
    if (bn == 128) return launch_cfg<2, 128, 6, OutT, true>(a, lda, b, ldb, g, num_sms, st);
    return launch_cfg<2, 256, 4, OutT, true>(a, lda, b, ldb, g, num_sms, st);
  }
  int cfg = g_force_cfg;
  if (cfg < 0) {
    // Measured on B200 (tools/sweep_mid.sh, profiles/README_r1.md):
    //  * if some 1-CTA tile shape covers the problem in ONE wave, the shape with the most tiles wins
    //    (M=256 x 4096^2: 128x64 tiles 10.0 us vs 256x256 pairs 16.9 us; M=512: 128x128 12.0 vs 17.1);
    //  * otherwise 128x256 (1-CTA) vs 256x256 (CTA pair): the pair has the better steady state (fewer
    //    operand bytes per MMA) but ~1 us more fixed cost, so it needs enough K blocks per worker.
    const long long m128 = (g.M + 127) / 128;
    const long long t0 = m128 * ((g.N + 255) / 256), t2 = m128 * ((g.N + 127) / 128), t3 = m128 * ((g.N + 63) / 64);
    const long long kb_all = (g.K + BLOCK_K - 1) / BLOCK_K;
    const long long t4 = (long long)((g.M + 255) / 256) * ((g.N + 127) / 128);   // 256x128 pair tiles
    // One-wave regime (gpurun_out/sweep_tune_r2.csv): the smallest tile shape that still fits one wave wins; for long
    // K (>= 8192) and at least one full pair of row blocks the CTA-pair shapes win (64 instead of 96 operand bytes
    // per SM-cycle): 512x4096x8192 17.1 (256x128 pairs) vs 17.9 us (128x128); 512x8192x8192 26.0 (256x256 pairs) vs
    // 29.8 us (128x256); 256x16384x16384 56.8 vs 65.6 us.
    const bool long_k_pairs = kb_all >= 64 && g.M >= 256;
    if (t3 <= num_sms) {
      cfg = 3;
      // very long K, few tiles: 256x128 pair tiles + stream-K (1024x1024x16384: 23.2 vs 27.6 us)
      if (kb_all >= 128 && g.M >= 256 && t4 * 2 <= num_sms / 2) cfg = 4;
    }
    else if (t2 <= num_sms) cfg = long_k_pairs ? 4 : 2;
    else if (t0 <= num_sms) {
      cfg = long_k_pairs ? 1 : 0;
      // one short-K wave: the epilogue is the kernel.  16-bit outputs already leave through per-warp TMA
      // stores (faster still: 4096x3072x768 12.9 us vs 14.5 us), fp32 outputs take the CTA-staged TMA path.
      if constexpr (std::is_same<OutT, float>::value) {
        if ((g.K + BLOCK_K - 1) / BLOCK_K <= 16 && g.vec_ok && g_force_cfg < 0)
          return launch_cfg<1, 256, 3, OutT, true>(a, lda, b, ldb, g, num_sms, st);
      }
    } else {
      // More than one wave of 128x256 tiles.  Per-worker cost model: waves x BLOCK_N column-units, with a
      // penalty for 1-CTA tiles.  Since the MMA issue loop was fixed (round 2: one elected lane inside uniform
      // control flow) the CTA pair is the faster steady state -- half of the weight tile per SM, 64 instead of
      // 96 operand bytes per SM-cycle: 2048x11008x4096 59.2 us vs 67.4 us, 8192^3 323 vs 377 us -- so 1-CTA
      // tiles are taken only when their wave quantisation is better by more than that (or M <= 128).
      // Narrower tiles (BLOCK_N = 240 / 224 / 208) cut the ragged last wave, e.g. 2048 x 4096: 2 x 256 -> 2 x 240
      // and 2048 x 12288: 6 x 256 -> 6 x 224 column-units per worker; a narrower tile loads ~3 % more operand
      // bytes per MMA, so it has to win by more than that.
      static const int bns[4] = {256, 240, 224, 208};
      static const int cfg_pair[4] = {1, 8, 9, 10}, cfg_single[4] = {0, 11, 12, 13};
      double best = 1e30;
      for (int cg = 2; cg >= 1; --cg) {
        if (cg == 2 && g.M <= 128) continue;
        const long long mt = (g.M + 128 * cg - 1) / (128 * cg), W = num_sms / cg;
        for (int i = 0; i < (g_narrow_tiles ? 4 : 1); ++i) {
          const long long nt = (g.N + bns[i] - 1) / bns[i];
          double cost = (double)(((mt * nt + W - 1) / W) * bns[i]);
          if (i > 0) cost *= 1.03;
          if (cg == 1) cost *= 1.12;
          if (cost < best) { best = cost; cfg = cg == 2 ? cfg_pair[i] : cfg_single[i]; }
        }
      }
      const long long kb = (g.K + BLOCK_K - 1) / BLOCK_K;
      // Short-K, multi-wave problems are bound by the output write, not by the MMAs: the staged
      // epilogue (whole 256-byte row segments per store) is 9-14 % faster there (K <= 2048:
      // 4096x3072x768 21.0 -> 18.0 us, 8192x8192x1024 75.2 -> 68.9 us) and slower for long K.
      if constexpr (std::is_same<OutT, float>::value) {
        if (kb <= 16 && g.vec_ok && g_force_cfg < 0)
          return launch_cfg<2, 256, 4, OutT, true>(a, lda, b, ldb, g, num_sms, st);
      }
    }
  }
  switch (cfg) {
    case 0: return launch_cfg<1, 256, 4, OutT>(a, lda, b, ldb, g, num_sms, st);
    case 1: return launch_cfg<2, 256, 6, OutT>(a, lda, b, ldb, g, num_sms, st);
    case 2: return launch_cfg<1, 128, 6, OutT>(a, lda, b, ldb, g, num_sms, st);
    case 3: return launch_cfg<1, 64, 8, OutT>(a, lda, b, ldb, g, num_sms, st);
    case 4: return launch_cfg<2, 128, 8, OutT>(a, lda, b, ldb, g, num_sms, st);
    case 5: return launch_cfg<2, 256, 4, OutT>(a, lda, b, ldb, g, num_sms, st);
    case 6: return launch_cfg<2, 256, 3, OutT>(a, lda, b, ldb, g, num_sms, st);
    case 8: return launch_cfg<2, 240, 6, OutT>(a, lda, b, ldb, g, num_sms, st);
    case 9: return launch_cfg<2, 224, 6, OutT>(a, lda, b, ldb, g, num_sms, st);
    case 10: return launch_cfg<2, 208, 7, OutT>(a, lda, b, ldb, g, num_sms, st);
    case 11: return launch_cfg<1, 240, 4, OutT>(a, lda, b, ldb, g, num_sms, st);
    case 12: return launch_cfg<1, 224, 4, OutT>(a, lda, b, ldb, g, num_sms, st);
    case 13: return launch_cfg<1, 208, 5, OutT>(a, lda, b, ldb, g, num_sms, st);
    case 16: return launch_cfg<2, 256, 6, OutT, false, 2>(a, lda, b, ldb, g, num_sms, st);
    default: PQ_FAIL(PQ_ERR_ARG, "qgemm: unknown tile config %d", cfg);
  }
}

}  // namespace

int launch_qgemm(const int8_t* a, int64_t lda, const int8_t* b, int64_t ldb,
                 const float* s_x, const float* s_w, const float* bias,
                 void* const* outs, int n_out, int out_dtype, int64_t ldo,
                 int64_t M, int64_t N, int64_t K, cudaStream_t stream, int64_t scatter_cols, int multimem,
                 int64_t rot_cols) {
  if (M < 0 || N < 0 || K < 1) PQ_FAIL(PQ_ERR_ARG, "qgemm: bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  if (M == 0 || N == 0) return PQ_OK;
  if (M > 0x7fffff00LL || N > 0x7fffff00LL || K > 0x7fffff00LL) PQ_FAIL(PQ_ERR_ARG, "qgemm: dimension too large");
  if (!a || !b || !outs || n_out < 1 || n_out > 8) PQ_FAIL(PQ_ERR_ARG, "qgemm: null pointer or bad destination count");
  for (int d = 0; d < n_out; ++d)
    if (!outs[d]) PQ_FAIL(PQ_ERR_ARG, "qgemm: null destination %d", d);
  if (out_dtype != PQ_I32 && (!s_x || !s_w)) PQ_FAIL(PQ_ERR_ARG, "qgemm: null scale pointer");
  if (scatter_cols > 0) {
    if (scatter_cols % 4 != 0 || scatter_cols * n_out < N || ldo < scatter_cols)
      PQ_FAIL(PQ_ERR_ARG, "qgemm: scatter needs scatter_cols %% 4 == 0, n_out * scatter_cols >= N and ldo >= scatter_cols");
  } else if (ldo < N) PQ_FAIL(PQ_ERR_ARG, "qgemm: leading dimension too small");
  if (lda < K || ldb < K) PQ_FAIL(PQ_ERR_ARG, "qgemm: leading dimension too small");
  if (((uintptr_t)a & 15) || ((uintptr_t)b & 15) || (lda & 15) || (ldb & 15))
    PQ_FAIL(PQ_ERR_ALIGN, "qgemm: xq/Wq base pointers and row strides must be multiples of 16 bytes (TMA)");
  int num_sms = 0;
  int rc = check_device(&num_sms);
  if (rc) return rc;
  // Decode-sized batches take the swap-AB weight-streaming kernel (qgemm_smallm.cu).
  // ... unless K is short: the cluster split-K machinery then costs more than it saves (64 x 4096 x 1024:
  // 11.4 us vs 5.0 us with 128x64 tiles; 64 x 4096 x 2048: 9.2 vs 6.4 us; 48 x 4096 x 4096: 8.3 vs 9.4 us).
  const bool smallm_pays = !(M > 32 && K <= 2048);
  // (several destinations = fused all-gather of a decode batch: the weight-streaming kernel writes its few rows to every
  // peer itself -- 16 x 3584 shard on 8 GPUs: 9.6 us, against 31 us for 128-row tiles with the staged epilogue)
  if (M <= 64 && scatter_cols == 0 && !multimem && !g_force_staged && g_sk_mode != 1 && ((g_force_cfg < 0 && smallm_pays) || g_force_cfg == 7))
    return launch_qgemm_smallm(a, lda, b, ldb, s_x, s_w, bias, outs, n_out, out_dtype, ldo, M, N, K, num_sms, stream);
  if (g_force_cfg == 7) PQ_FAIL(PQ_ERR_ARG, "qgemm: config 7 (small-M kernel) needs M <= 64");
  GemmArgs g = {};
  g.M = (int)M; g.N = (int)N; g.K = (int)K;
  g.s_x = s_x; g.s_w = s_w; g.bias = bias;
  g.ldo = ldo;
  g.n_out = n_out;
  const int esz = dtype_size(out_dtype);
  g.vec_ok = ((ldo * esz) % 16 == 0);
  for (int d = 0; d < n_out; ++d) {
    g.out[d] = outs[d];
    if ((uintptr_t)outs[d] & 15) g.vec_ok = 0;
  }
  g.tl = g_timeline.load(std::memory_order_relaxed);
  g.prefetch_b = g_prefetch_b;
  g.dbg = g_epi_dbg;
  g.scatter_cols = (int)scatter_cols;
  if (rot_cols <= 0) rot_cols = g_dbg_rot_cols;       // test hook: exercise the rotated tile order on one GPU
  g.n_rot = (int)(rot_cols > 0 ? rot_cols : 0);      // in COLUMNS here; launch_cfg converts to blocks of its BLOCK_N
  if (multimem) {
    if (n_out != 1 || scatter_cols != 0 || !g.vec_ok || (N * esz) % 16 != 0)
      PQ_FAIL(PQ_ERR_ARG, "qgemm: a multicast destination needs n_ys == 1 and 16-byte aligned rows / row length");
    g.multimem = 1;
  }
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
// -1 = heuristic, 0 = never use stream-K, 1 = use stream-K whenever it is legal
extern "C" void pq_debug_set_streamk(int mode) { pq::g_sk_mode = mode; }
extern "C" void pq_debug_set_staged(int on) { pq::g_force_staged = on; }
// How many clusters of `cluster_size` CTAs of the main GEMM kernel fit on the device at once
// (cudaOccupancyMaxActiveClusters); used to judge whether 4-CTA TMA multicast could pay.
extern "C" int pq_debug_max_active_clusters(int cluster_size) {
  using namespace pq;
  using L = SmemLayout<2, 256, 6, false, 2>;
  auto kern = qgemm_kernel<2, 256, 6, __nv_bfloat16, false>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES) != cudaSuccess) return -1;
  if (cluster_size > 8)
    cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148 * 4 / cluster_size * cluster_size, 1, 1);
  cfg.blockDim = dim3(NUM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = L::DYN_BYTES;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = cluster_size;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  int n = -1;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { (void)cudaGetLastError(); return -2; }
  return n;
}
extern "C" void pq_debug_set_epilogue(int bits) { pq::g_epi_dbg = bits; }
extern "C" void pq_debug_set_narrow_tiles(int on) { pq::g_narrow_tiles = on; }
extern "C" void pq_debug_set_prefetch(int on) { pq::g_prefetch_b = on; }
extern "C" void pq_debug_set_tma_store(int on) { pq::g_tma_store = on; }
// device buffer of 40 x u64 per CTA (zeroed by the caller) receiving %globaltimer stamps, or null
extern "C" void pq_debug_set_timeline(unsigned long long* dev_buf) { pq::g_timeline.store(dev_buf, std::memory_order_relaxed); }
extern "C" void pq_debug_set_multi_tma(int on) { pq::g_multi_tma = on; }
extern "C" void pq_debug_set_multi_bn(int bn) { pq::g_multi_bn = bn; }
extern "C" void pq_debug_set_tile_rotation(int cols) { pq::g_dbg_rot_cols = cols; }
