// Row-parallel (K-split) dynamic-quant linear as ONE call (VERDICT r1 "next" #3):
//   row |.|-max [-> exchanged through symmetric memory] -> quantise with the GLOBAL row maximum -> int32 GEMM whose
//   epilogue scatters every output-column block into its owner's inbox over NVLink -> cross-rank barrier ->
//   reduce the `world` int32 partials + dequant epilogue [-> stored into every rank's output: fused all-gather]
//   [-> barrier].
// No NCCL call and no host round trip anywhere: the ranks synchronise through signal pads in symmetric memory
// (symm_barrier_kernel), so the whole forward is a fixed sequence of launches on one stream and can be captured in
// a CUDA graph.  The input may also be the gated product act(gate) * up (Llama MLP): its row maximum spans all
// ranks' column slices, so a statistics launch publishes the local maxima to every peer before the quantising one.
#include "common.cuh"
#include "ptx.cuh"

namespace pq {
namespace {

struct Pads { uint32_t* p[8]; };

__device__ __forceinline__ uint32_t cas_release_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.release.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t cas_acquire_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.acquire.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}

// One CTA, thread t talks to peer t.  Slot [channel][src] of a rank's pad is owned by the pair (src -> that rank):
// the sender flips it 0 -> 1 (waiting for 0 first, so back-to-back barriers cannot overrun each other), the owner
// flips it back 1 -> 0.  Self-resetting, hence safe under CUDA-graph replay; release / acquire at system scope
// order the peer stores of the kernels before the barrier against the loads of the kernels after it.
// Every spin is bounded (wall clock): a rank that never arrives becomes a trap, not a hung GPU.
__global__ void __launch_bounds__(32) symm_barrier_kernel(const Pads pads, int rank, int world, int channel) {
  // launched with the PDL attribute: it may be scheduled early, so the launch latency is off the critical path, but
  // it signals only after every earlier kernel of this stream has completed and its (peer) stores are visible
  ptx::griddep_launch_dependents();
  ptx::griddep_wait();
  const int t = threadIdx.x;
  if (t >= world || t == rank) return;
  uint32_t* remote = pads.p[t] + channel * 8 + rank;
  uint32_t* mine = pads.p[rank] + channel * 8 + t;
  uint64_t t0 = 0;
  uint32_t polls = 0;
  auto check = [&](const char* what) {
    if ((++polls & 255u) == 0) {
      const uint64_t now = ptx::globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > PQ_MBAR_TIMEOUT_NS) {
        printf("pq: cross-rank barrier timed out (%s, rank %d peer %d channel %d)\n", what, rank, t, channel);
        __trap();
      }
    }
  };
  while (cas_release_sys(remote, 0u, 1u) != 0u) check("signal");
  while (cas_acquire_sys(mine, 1u, 0u) != 1u) check("wait");
}

int launch_barrier(const pq_symm_group* sg, int channel, cudaStream_t st) {
  Pads p = {};
  for (int r = 0; r < sg->world; ++r) p.p[r] = sg->pads[r];
  cudaLaunchAttribute attr[1];
  cudaLaunchConfig_t cfg = pdl_config(dim3(1), dim3(32), 0, st, attr);
  PQ_CUDA(cudaLaunchKernelEx(&cfg, symm_barrier_kernel, p, sg->rank, sg->world, channel));
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return PQ_OK;
}

}  // namespace
}  // namespace pq

using namespace pq;

extern "C" int pq_symm_barrier(const pq_symm_group* sg, int channel, void* stream) {
  int rc = check_device(nullptr);
  if (rc) return rc;
  if (!sg || sg->world < 1 || sg->world > 8 || sg->rank < 0 || sg->rank >= sg->world || channel < 0 || channel > 3)
    PQ_FAIL(PQ_ERR_ARG, "pq_symm_barrier: bad group or channel");
  for (int r = 0; r < sg->world; ++r)
    if (!sg->pads[r]) PQ_FAIL(PQ_ERR_ARG, "pq_symm_barrier: null signal pad of rank %d", r);
  if (sg->world == 1) return PQ_OK;
  return launch_barrier(sg, channel, (cudaStream_t)stream);
}

extern "C" int pq_rowparallel_forward(const void* x, const void* up, int x_dtype, int act, int64_t ldx, int64_t ldu,
                                      int input_is_sharded, int64_t K_in, int64_t k_lo,
                                      const int8_t* Wq, int64_t ldb, const float* s_w, const float* bias,
                                      const pq_symm_group* sg, int gather_output, void* y_local, int y_dtype, int64_t ldy,
                                      int8_t* xq_ws, float* sx_ws, float* amax_ws,
                                      int64_t M, int64_t N, int64_t K_slice, int64_t per_n,
                                      const pq_quant_spec* spec_in, void* stream) {
  int rc = check_device(nullptr);
  if (rc) return rc;
  if (M == 0) return PQ_OK;
  if (!sg || sg->world < 2 || sg->world > 8 || sg->rank < 0 || sg->rank >= sg->world)
    PQ_FAIL(PQ_ERR_ARG, "pq_rowparallel_forward: needs a symmetric-memory group of 2..8 ranks");
  if (M < 0 || N < 1 || K_slice < 1 || per_n < 1 || M > sg->cap)
    PQ_FAIL(PQ_ERR_ARG, "pq_rowparallel_forward: bad shape (M=%lld exceeds the buffers' %lld rows?)", (long long)M, (long long)sg->cap);
  if (!x || !Wq || !s_w || !xq_ws || !sx_ws) PQ_FAIL(PQ_ERR_ARG, "pq_rowparallel_forward: null pointer");
  if (y_dtype != PQ_BF16 && y_dtype != PQ_F16 && y_dtype != PQ_F32)
    PQ_FAIL(PQ_ERR_ARG, "pq_rowparallel_forward: y_dtype must be PQ_BF16, PQ_F16 or PQ_F32");
  if ((long long)per_n * sg->world < N) PQ_FAIL(PQ_ERR_ARG, "pq_rowparallel_forward: per_n * world < N");
  if (up && !input_is_sharded) PQ_FAIL(PQ_ERR_ARG, "pq_rowparallel_forward: a gated input is a K-sharded input");
  if (input_is_sharded && K_in != K_slice) PQ_FAIL(PQ_ERR_ARG, "pq_rowparallel_forward: a K-sharded input has K_slice columns");
  if (!input_is_sharded && (k_lo < 0 || k_lo + K_slice > K_in || !amax_ws))
    PQ_FAIL(PQ_ERR_ARG, "pq_rowparallel_forward: bad K slice of the replicated input (or null amax workspace)");
  const int world = sg->world, rank = sg->rank;
  for (int r = 0; r < world; ++r)
    if (!sg->inbox[r] || !sg->pads[r] || (gather_output && !sg->out[r]) || (input_is_sharded && !sg->amax[r]))
      PQ_FAIL(PQ_ERR_ARG, "pq_rowparallel_forward: null symmetric buffer of rank %d", r);
  if (!gather_output && !y_local) PQ_FAIL(PQ_ERR_ARG, "pq_rowparallel_forward: null local output");
  const pq_quant_spec spec = resolve_spec(spec_in);
  cudaStream_t st = (cudaStream_t)stream;
  const int esz_x = dtype_size(x_dtype), esz_y = dtype_size(y_dtype);
  if (esz_x == 0 || x_dtype == PQ_I32) PQ_FAIL(PQ_ERR_ARG, "pq_rowparallel_forward: unsupported input dtype %d", x_dtype);
  const int64_t ldq = (K_slice + 15) / 16 * 16;
  const int64_t cap = sg->cap;

  // 1. row maxima over the WHOLE row
  if (input_is_sharded) {
    float* slots[8];
    for (int r = 0; r < world; ++r) slots[r] = sg->amax[r] + (int64_t)rank * cap;     // my slot in every rank's buffer
    if (up) rc = launch_act_mul_quant(x, up, x_dtype, act, M, K_slice, ldx, ldu, nullptr, K_slice, nullptr, nullptr, 0, spec, st,
                                      nullptr, 0, 0, slots, world);
    else rc = launch_row_absmax(x, x_dtype, M, K_slice, ldx, slots, world, st);
    if (rc) return rc;
    rc = launch_barrier(sg, 0, st);
    if (rc) return rc;
    // 2. quantise the slice with the maximum over the `world` slots
    if (up) rc = launch_act_mul_quant(x, up, x_dtype, act, M, K_slice, ldx, ldu, xq_ws, ldq, sx_ws, nullptr, 0, spec, st,
                                      sg->amax[rank], world, cap, nullptr, 0);
    else rc = launch_rowwise_quant(x, x_dtype, M, K_slice, ldx, xq_ws, ldq, sx_ws, 0, spec, st, nullptr, 0,
                                   sg->amax[rank], world, cap);
  } else {
    float* one[1] = {amax_ws};
    rc = launch_row_absmax(x, x_dtype, M, K_in, ldx, one, 1, st);
    if (rc) return rc;
    rc = launch_rowwise_quant((const char*)x + k_lo * esz_x, x_dtype, M, K_slice, ldx, xq_ws, ldq, sx_ws, 0, spec, st,
                              nullptr, 0, amax_ws, 1, 0);
  }
  if (rc) return rc;

  // 3. int32 GEMM on this K-slice; the epilogue stores column block d into rank d's inbox slot `rank`
  void* dests[8];
  const int64_t slot_bytes = cap * per_n * 4;
  const int n_dests = (int)((N + per_n - 1) / per_n);
  for (int d = 0; d < n_dests; ++d) dests[d] = (char*)sg->inbox[d] + (int64_t)rank * slot_bytes;
  // every rank starts its tile order at its OWN column block, so the ranks store to different inboxes at any moment
  rc = launch_qgemm(xq_ws, ldq, Wq, ldb, nullptr, nullptr, nullptr, dests, n_dests, PQ_I32, per_n, M, N, K_slice, st, per_n, 0,
                    (int64_t)rank * per_n);
  if (rc) return rc;
  rc = launch_barrier(sg, 1, st);          // every rank's partial sums have landed
  if (rc) return rc;

  // 4. reduce my column block over the `world` slots + dequant epilogue (+ all-gather store)
  const int64_t n_lo = (int64_t)rank * per_n;
  const int64_t n_mine = n_lo >= N ? 0 : (N - n_lo < per_n ? N - n_lo : per_n);
  if (n_mine > 0) {
    const int32_t* parts[8];
    for (int s = 0; s < world; ++s) parts[s] = (const int32_t*)((const char*)sg->inbox[rank] + (int64_t)s * slot_bytes);
    void* ys[8];
    int n_ys = 1;
    int64_t ld = ldy;
    if (gather_output) {
      n_ys = world;
      for (int r = 0; r < world; ++r) ys[r] = (char*)sg->out[r] + n_lo * esz_y;
    } else {
      ys[0] = y_local;
    }
    rc = launch_reduce_dequant(parts, world, per_n, sx_ws, s_w + n_lo, bias ? bias + n_lo : nullptr, ys, n_ys, y_dtype, ld,
                               M, n_mine, st);
    if (rc) return rc;
  }
  if (gather_output) rc = launch_barrier(sg, 2, st);    // every rank's slice of the output has landed everywhere
  return rc;
}
