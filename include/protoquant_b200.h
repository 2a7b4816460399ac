/*
 * protoquant_b200 — C ABI of the B200 (sm_100a) dynamic-quantized linear path.
 *
 * This is the drop-in boundary for protoquant's dynamic int8 linear path
 * (BASELINE.json north_star).  The reference checkout is ABSENT in this
 * environment (/root/reference holds only CODE_OF_CONDUCT.md),
 * so no reference file:line can be cited; each entry point instead cites the
 * SURVEY.md §8 row it implements.  The Python side (the protoquant_b200 package) binds
 * these symbols with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns all buffers; nothing is allocated or freed here, except
 *     the opaque pq_linear handle of the host-buffer convenience API and a
 *     per-stream int32 workspace of the stream-K schedule (first use, never
 *     during stream capture);
 *   - process-wide state is limited to caches that cannot change results: the
 *     per-device capability / kernel-attribute flags, encoded TMA descriptors
 *     (pure functions of address, shape and stride) and that workspace pool.
 *     Entry points may be called concurrently from several threads on distinct
 *     streams.  The exported pq_debug_* symbols are test / profiling hooks, not
 *     part of this API: relaxed atomics that select between code paths with
 *     bit-identical results;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); no
 *     function synchronises unless it says so;
 *   - return value 0 = ok, non-zero = PQ_ERR_*; pq_last_error() returns a
 *     thread-local message for the last non-zero return;
 *   - dtype codes: PQ_F32 / PQ_F16 / PQ_BF16 (/ PQ_I32 for raw accumulators).
 *   - no CPU fallback exists: on a machine without an sm_100 device the
 *     compute entry points return PQ_ERR_DEVICE.
 */
#ifndef PROTOQUANT_B200_H
#define PROTOQUANT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PQ_VERSION 200 /* 0.2.0: round 2 added pq_qlinear_multi, pq_symm_barrier, pq_rowparallel_forward (additive; every 0.1.0 entry point is unchanged) */

enum pq_dtype { PQ_F32 = 0, PQ_F16 = 1, PQ_BF16 = 2, PQ_I32 = 3 };

enum pq_status {
  PQ_OK = 0,
  PQ_ERR_ARG = 1,      /* bad shape / stride / dtype / null pointer            */
  PQ_ERR_ALIGN = 2,    /* pointer or leading dimension not aligned as required */
  PQ_ERR_DEVICE = 3,   /* no CUDA device, or device is not sm_100              */
  PQ_ERR_CUDA = 4,     /* a CUDA runtime/driver call failed                    */
  PQ_ERR_UNSUPPORTED = 5
};

/* Scale-computation variants (SURVEY.md §8c "knobs").  Default = PQ_DIV. */
enum pq_scale_mode {
  PQ_DIV = 0,      /* s = amax/127 ; q = rne(x / s)          (true IEEE division)  */
  PQ_RCP_MUL = 1,  /* s = amax/127 ; q = rne(x * (1/s))      (torch.ao per-token)  */
  PQ_INV_SCALE = 2 /* s = amax/127 ; q = rne(x * (127/amax))                       */
};

typedef struct pq_quant_spec {
  int32_t scale_mode; /* enum pq_scale_mode                                        */
  float eps;          /* 0 = none; else s = max(amax, eps)/127 (torch.ao: 1e-5)    */
  int32_t qmin;       /* -128 (default) or -127                                    */
} pq_quant_spec;

/* Non-finite and denormal input (every quantizing entry point, every dtype, every kernel).
 * The formula  amax = max|x|,  s = amax/127 (amax == 0 -> 1),  q = clamp(rne(x/s), qmin, 127)  is evaluated
 * literally in IEEE-754 binary32 arithmetic, with NaN -> 0 at the float -> int8 conversion:
 *   - amax PROPAGATES NaN and +inf, so a row (token / output channel) holding a NaN gets s = NaN, a row holding
 *     +-inf (and no NaN) gets s = +inf; in both cases every code of that row is 0 (finite/inf = 0,
 *     inf/inf = NaN -> 0, x/NaN = NaN -> 0).  Dequantising such a row, or running it through the GEMM epilogue,
 *     yields NaN (0 * s) -- the non-finite input stays visible downstream; other rows are unaffected;
 *   - fp32 rows whose amax is so small that s is denormal or underflows to 0 (amax < 127 * 2^-126) keep the clamp
 *     live: x/0 = +-inf -> 127 / qmin, 0/0 -> 0, and a coarsely rounded denormal s may give |x/s| > 127 -> clamp;
 *     bf16 / fp16 denormals always have a normal or exactly representable fp32 scale and need no special case;
 *   - the sign and payload of a NaN scale are unspecified.
 * tests/test_gpu_quant.py::test_nonfinite_and_denormal_rows_follow_the_policy and the exact-rational golden
 * vectors (tests/golden/exact_*.npz) pin this for all kernels. */

int pq_version(void);
const char* pq_last_error(void);

/* Number of CUDA kernels this library has launched in the calling process
 * (all threads).  bench.py reports it as "gpu_launches". */
uint64_t pq_launch_count(void);

/* SURVEY §8 row a1 — per-token (row-wise) symmetric int8 quantizer.
 *   x   [M,K]  x_dtype, row stride ldx (elements)
 *   xq  [M,K]  int8,    row stride ldq (bytes)      (transpose == 0)
 *       [K,M]  int8,    row stride ldq (bytes)      (transpose != 0)
 *   s_x [M]    fp32
 * spec may be NULL (= {PQ_DIV, 0, -128}).  Requires K >= 1, M >= 0. */
int pq_act_quant(const void* x, int x_dtype, int64_t M, int64_t K, int64_t ldx,
                 int8_t* xq, int64_t ldq, float* s_x, int transpose,
                 const pq_quant_spec* spec, void* stream);

/* SURVEY §8 row a2 — per-output-channel int8 weight quantizer (run once at load).
 *   W [N,K] w_dtype row stride ldw;  Wq [N,K] int8 row stride ldwq;  s_w [N] fp32. */
int pq_weight_quant(const void* W, int w_dtype, int64_t N, int64_t K, int64_t ldw,
                    int8_t* Wq, int64_t ldwq, float* s_w,
                    const pq_quant_spec* spec, void* stream);

/* SURVEY §8 rows a3+a4 — int8 x int8 -> int32 GEMM on tcgen05 with the fused
 * dequant epilogue  y[m,n] = cast( (float(acc[m,n]) * s_x[m]) * s_w[n] + bias[n] ).
 *   xq [M,K] int8 row stride lda (bytes, multiple of 16, base 16-B aligned)
 *   Wq [N,K] int8 row stride ldb (bytes, multiple of 16, base 16-B aligned)
 *   bias [N] fp32 or NULL;  y [M,N] y_dtype (PQ_BF16/PQ_F16/PQ_F32) row stride ldy (elements). */
int pq_qgemm(const int8_t* xq, int64_t lda, const int8_t* Wq, int64_t ldb,
             const float* s_x, const float* s_w, const float* bias,
             void* y, int y_dtype, int64_t ldy,
             int64_t M, int64_t N, int64_t K, void* stream);

/* SURVEY §8e — fused GEMM + all-gather store.  Same computation as pq_qgemm, but the [M,N]
 * result is written to each of ys[0..n_ys) (1 <= n_ys <= 8).  For a column-parallel layer,
 * rank r passes, for every peer p, the peer-mapped address of p's full output buffer offset to
 * r's first column (NVLink peer stores issued from the GEMM epilogue), or a single NVSwitch
 * multicast address; ldy is the row stride of the full buffer.  The caller provides the
 * cross-rank barrier that follows. */
int pq_qgemm_multi(const int8_t* xq, int64_t lda, const int8_t* Wq, int64_t ldb,
                   const float* s_x, const float* s_w, const float* bias,
                   void* const* ys, int n_ys, int y_dtype, int64_t ldy,
                   int64_t M, int64_t N, int64_t K, void* stream);

/* SURVEY §8e, one call per forward of a column-parallel shard: act-quant into the caller's workspace (as pq_qlinear)
 * followed by the fused GEMM + all-gather store (as pq_qgemm_multi).  flags: PQ_MULTI_MULTICAST = ys[0] is an NVSwitch
 * multicast address (n_ys must be 1): the epilogue then writes it with multimem.st and the switch replicates the
 * tile into every rank; without the flag every ys[d] is a unicast (local or NVLink peer-mapped) address and is
 * written from the CTA-staged tile with coalesced 256-byte peer stores (measured faster over NVLink than TMA
 * stores; the tile width 256 / 224 / 128 is chosen per launch).  Two kernel launches, no sync; the caller issues the cross-rank barrier. */
enum pq_multi_flags { PQ_MULTI_MULTICAST = 1 };
int pq_qlinear_multi(const void* x, int x_dtype, int64_t ldx,
                     const int8_t* Wq, int64_t ldb, const float* s_w, const float* bias,
                     void* const* ys, int n_ys, int y_dtype, int64_t ldy,
                     int8_t* xq_ws, float* sx_ws,
                     int64_t M, int64_t N, int64_t K,
                     const pq_quant_spec* spec, int flags, void* stream);

/* Parity hook for row a3: raw int32 accumulators acc[M,N], row stride ldc (elements). */
int pq_qgemm_i32(const int8_t* xq, int64_t lda, const int8_t* Wq, int64_t ldb,
                 int32_t* acc, int64_t ldc,
                 int64_t M, int64_t N, int64_t K, void* stream);

/* SURVEY §8 row a5 — QTensor.dequantize():  out[r,c] = q[r,c] * s[axis==0 ? r : c]. */
int pq_dequant(const int8_t* q, int64_t ldq, const float* s, int axis,
               void* out, int out_dtype, int64_t ldo,
               int64_t rows, int64_t cols, void* stream);

/* SURVEY §8 row a6 — one dynamic-quant linear forward on device buffers:
 * act-quant into caller workspace (xq_ws [M, round_up(K,16)] int8, sx_ws [M] fp32)
 * followed by pq_qgemm.  Two kernel launches, no sync. */
int pq_qlinear(const void* x, int x_dtype, int64_t ldx,
               const int8_t* Wq, int64_t ldb, const float* s_w, const float* bias,
               void* y, int y_dtype, int64_t ldy,
               int8_t* xq_ws, float* sx_ws,
               int64_t M, int64_t N, int64_t K,
               const pq_quant_spec* spec, void* stream);

/* SURVEY §8f-2 — producer-fused quantizers: the op in front of a dynamic-quant linear writes the int8
 * operand and its per-token scale directly.  `y` / `h` (nullable) additionally receives the tensor the
 * unfused op would have stored (dtype of the input); (xq, s_x) is exactly the row-wise quantisation
 * (pq_act_quant arithmetic, same `spec`) of that tensor.  Rows of x / gamma / beta / y / gate / up / h must be
 * 16-byte aligned and K a multiple of 16 / sizeof(dtype) (PQ_ERR_UNSUPPORTED / PQ_ERR_ALIGN otherwise).
 *
 * pq_norm_quant: beta == NULL  ->  RMSNorm   y = T(gamma * T(x * rsqrt(mean(x^2) + eps)))      (Llama)
 *                beta != NULL  ->  LayerNorm y = T((x - mean) * rsqrt(var + eps) * gamma + beta) (BERT)
 *   x [M,K] row stride ldx (elements); gamma, beta [K] of x_dtype; statistics in fp32. */
int pq_norm_quant(const void* x, int x_dtype, int64_t M, int64_t K, int64_t ldx,
                  const void* gamma, const void* beta, float eps,
                  int8_t* xq, int64_t ldq, float* s_x, void* y, int64_t ldy,
                  const pq_quant_spec* spec, void* stream);

enum pq_act { PQ_ACT_IDENTITY = 0, PQ_ACT_SILU = 1, PQ_ACT_GELU = 2, PQ_ACT_GELU_TANH = 3 };

/* pq_act_mul_quant: h = T(T(act(gate)) * up)   (up != NULL: Llama MLP, act = PQ_ACT_SILU)
 *                   h = T(act(gate))           (up == NULL: BERT FFN, act = PQ_ACT_GELU)
 *   gate, up [M,K] of `dtype`, row strides ldg / ldu (elements). */
int pq_act_mul_quant(const void* gate, const void* up, int dtype, int act,
                     int64_t M, int64_t K, int64_t ldg, int64_t ldu,
                     int8_t* hq, int64_t ldq, float* s_h, void* h, int64_t ldh,
                     const pq_quant_spec* spec, void* stream);

/* SURVEY §8f-3 — row-parallel (K-split) sharding with a fused GEMM + reduce-scatter.  Rank r holds columns
 * [k_lo, k_hi) of x and of Wq.  The per-token scale needs the maximum over the WHOLE row, so each rank takes
 * pq_row_absmax of its slice, the M floats are all-reduced with MAX, and pq_act_quant_amax quantises the slice
 * with that global maximum: codes and scales are then those of the unsharded quantizer.  pq_qgemm_i32_scatter
 * multiplies the slices and stores the exact int32 partial sums of output columns [d*cols_per_dest,
 * (d+1)*cols_per_dest) into dests[d] (row stride ld_dest elements) -- peer-mapped inboxes on the ranks that own
 * those columns, written from the GEMM epilogue over NVLink in coalesced 256-byte segments.  After a barrier
 * every owner runs pq_reduce_dequant over the n_parts inboxes: int32 addition is associative, so
 *   y[m,n] = cast(((float(sum_p part_p[m,n]) * s_x[m]) * s_w[n]) + bias[n])
 * is bit-identical to the unsharded pq_qgemm for any number of shards.  ys[0..n_ys) may again be peer-mapped
 * (the all-gather of the finished slices fused into the same kernel).  With n_parts == 1 and n_ys == 1 this is
 * simply the dequant epilogue applied to pq_qgemm_i32's accumulators. */
int pq_row_absmax(const void* x, int x_dtype, int64_t M, int64_t K, int64_t ldx, float* amax, void* stream);
int pq_act_quant_amax(const void* x, int x_dtype, int64_t M, int64_t K, int64_t ldx, const float* amax,
                      int8_t* xq, int64_t ldq, float* s_x, const pq_quant_spec* spec, void* stream);
int pq_qgemm_i32_scatter(const int8_t* xq, int64_t lda, const int8_t* Wq, int64_t ldb,
                         void* const* dests, int n_dests, int64_t ld_dest, int64_t cols_per_dest,
                         int64_t M, int64_t N, int64_t K, void* stream);
int pq_reduce_dequant(const int32_t* const* parts, int n_parts, int64_t ld_part,
                      const float* s_x, const float* s_w, const float* bias,
                      void* const* ys, int n_ys, int y_dtype, int64_t ldy,
                      int64_t M, int64_t N, void* stream);

/* Host-buffer convenience API: the whole path behind one call on HOST buffers (H2D copy, act-quant, GEMM, D2H
 * copy).  bench.py's "e2e" figure goes through the Python module API (DynamicQuantLinear.forward with pinned host
 * tensors and torch copies inside the timed region); this handle is the same thing for a C caller.
 * A pq_linear owns device copies of (Wq, s_w, bias), a device activation/output
 * workspace for up to max_tokens rows, and a private stream. */
typedef struct pq_linear pq_linear;

/* W_host [N,K] row-major w_dtype, bias_host [N] fp32 or NULL.  Quantizes on the GPU. */
int pq_linear_create(pq_linear** out, const void* W_host, int w_dtype,
                     int64_t N, int64_t K, const float* bias_host,
                     int64_t max_tokens, int act_dtype, int out_dtype,
                     const pq_quant_spec* spec);
/* x_host [M,K] act_dtype (pinned or pageable), y_host [M,N] out_dtype.
 * H2D copy + act-quant + qgemm + D2H copy on the handle's stream, then waits for it. */
int pq_linear_forward_host(pq_linear* h, const void* x_host, void* y_host, int64_t M);
void pq_linear_destroy(pq_linear* h);

/* SURVEY §8f-3 — the row-parallel (K-split) linear as ONE call, no NCCL on the data path.
 * A group of 2..8 ranks (one process per GPU) shares symmetric memory: every pointer array below holds, for each
 * rank r, the peer-mapped device address of rank r's buffer (NVLink P2P; e.g. torch symmetric memory `buffer_ptrs`). */
typedef struct pq_symm_group {
  int32_t rank, world;
  void* inbox[8];     /* int32 [world][cap][per_n]: slot s receives rank s's partial sums of this rank's columns */
  void* out[8];       /* y_dtype [cap][ldy]: the gathered output (only read when gather_output != 0)          */
  float* amax[8];     /* fp32 [world][cap]: row-maximum exchange (only read when the input is K-sharded)       */
  uint32_t* pads[8];  /* >= 32 uint32, zero-initialised once: cross-rank signal pads of pq_symm_barrier        */
  int64_t cap;        /* row capacity of inbox / out / amax                                                    */
} pq_symm_group;

/* Cross-rank barrier on `stream` through the signal pads (channel 0..3): returns once every rank's kernel has
 * arrived; peer stores issued by earlier kernels of any rank are visible to later kernels of every rank.
 * Self-resetting (CUDA-graph replay safe), bounded spins (a missing rank traps after 4 s instead of hanging). */
int pq_symm_barrier(const pq_symm_group* sg, int channel, void* stream);

/* y = linear(x) for a weight split along K over the group: this rank holds Wq [N, K_slice] (its K-slice of every
 * output channel), the full s_w [N] / bias [N].
 *   x  [M, K_in] x_dtype row stride ldx.  input_is_sharded == 0: x is the replicated activation, this rank uses
 *      columns [k_lo, k_lo + K_slice) and the row maximum is taken over all K_in columns locally.
 *      input_is_sharded != 0: x is this rank's K-slice (K_in == K_slice); the row maxima are exchanged through
 *      sg->amax.  up != NULL (K-sharded only): the input is act(x) * up (enum pq_act; x = gate), [M, K_slice] each.
 *   Launches: row-max (+ barrier) + quantise + int32 GEMM scattering into the owners' inboxes + barrier +
 *   reduce/dequant of this rank's per_n output columns (stored into every rank's sg->out at column rank * per_n when
 *   gather_output, followed by a barrier; else into y_local [M, n_mine] row stride ldy).
 *   per_n: output columns per rank (multiple of 8, per_n * world >= N); ldy: row stride of sg->out / y_local.
 *   xq_ws [M, ld16(K_slice)] int8, sx_ws [M] fp32, amax_ws [M] fp32 (replicated input only): caller-owned scratch.
 * Bit-identical to the unsplit pq_qlinear for any world size: the K-shards quantise with the global row maximum
 * and the reduction runs on the exact int32 partial sums before ONE dequant epilogue. */
int pq_rowparallel_forward(const void* x, const void* up, int x_dtype, int act, int64_t ldx, int64_t ldu,
                           int input_is_sharded, int64_t K_in, int64_t k_lo,
                           const int8_t* Wq, int64_t ldb, const float* s_w, const float* bias,
                           const pq_symm_group* sg, int gather_output, void* y_local, int y_dtype, int64_t ldy,
                           int8_t* xq_ws, float* sx_ws, float* amax_ws,
                           int64_t M, int64_t N, int64_t K_slice, int64_t per_n,
                           const pq_quant_spec* spec, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PROTOQUANT_B200_H */
