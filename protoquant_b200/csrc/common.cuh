// Shared host/device helpers for the protoquant_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/protoquant_b200.h"

namespace pq {

// ---- error plumbing (thread-local message, see pq_last_error) -------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launch_count;

// Test / profiling knobs behind the exported pq_debug_* hooks (not part of the API declared in the header).  They are
// process-wide relaxed atomics: safe to flip from any thread, and every value selects between code paths that
// produce bit-identical results (tile shapes, schedules, store instructions) or a profiling mode.
struct Knob {
  std::atomic<int> v;
  constexpr Knob(int x) : v(x) {}
  operator int() const { return v.load(std::memory_order_relaxed); }
  int load(std::memory_order = std::memory_order_relaxed) const { return v.load(std::memory_order_relaxed); }
  void operator=(int x) { v.store(x, std::memory_order_relaxed); }
};
extern Knob g_pdl;   // 1 = launch kernels with programmatic stream serialization (default)

#define PQ_FAIL(code, ...)        \
  do {                            \
    ::pq::set_error(__VA_ARGS__); \
    return (code);                \
  } while (0)

#define PQ_CUDA(expr)                                                                  \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess)                                                             \
      PQ_FAIL(PQ_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
              __FILE__, __LINE__);                                                     \
  } while (0)

// Launch configuration with the programmatic-dependent-launch attribute (when enabled): the kernel may be scheduled
// while its predecessor in the stream is still running and must execute griddepcontrol.wait before it touches
// anything the predecessor produces.  `attr` must outlive the launch call.
inline cudaLaunchConfig_t pdl_config(dim3 grid, dim3 block, size_t smem, cudaStream_t st, cudaLaunchAttribute* attr) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  return cfg;
}

// Checks that the current device is an sm_100 part; caches the answer per device.
int check_device(int* num_sms);

inline pq_quant_spec resolve_spec(const pq_quant_spec* s) {
  pq_quant_spec d;
  d.scale_mode = PQ_DIV;
  d.eps = 0.f;
  d.qmin = -128;
  return s ? *s : d;
}

// what the kernels receive as `scale_mode`: the mode in the low byte, bit 8 = "qmin is -127" (quant_math.cuh make_rowq)
inline int mode_bits(const pq_quant_spec& s) { return (s.scale_mode & 0xff) | (s.qmin == -127 ? 0x100 : 0); }

inline int dtype_size(int dt) {
  switch (dt) {
    case PQ_F32: return 4;
    case PQ_F16: return 2;
    case PQ_BF16: return 2;
    case PQ_I32: return 4;
    default: return 0;
  }
}

// ---- internal launchers shared between translation units ------------------------
int launch_rowwise_quant(const void* x, int x_dtype, int64_t M, int64_t K, int64_t ldx,
                         int8_t* xq, int64_t ldq, float* s, int transpose,
                         const pq_quant_spec& spec, cudaStream_t stream,
                         const void* prefetch = nullptr, long long prefetch_bytes = 0,
                         const float* amax_in = nullptr, int amax_slots = 1, long long amax_stride = 0);

// amax[m] = max_k |x[m,k]| stored to every dsts[d][m] (rowparallel.cu)
int launch_row_absmax(const void* x, int x_dtype, int64_t M, int64_t K, int64_t ldx, float* const* dsts, int n_dst,
                      cudaStream_t st);
// y = cast(((float(sum of the int32 parts) * s_x[m]) * s_w[n]) + bias[n]) written to every ys[d] (rowparallel.cu)
int launch_reduce_dequant(const int32_t* const* parts, int n_parts, int64_t ld_part,
                          const float* s_x, const float* s_w, const float* bias,
                          void* const* ys, int n_ys, int y_dtype, int64_t ldy, int64_t M, int64_t N, cudaStream_t st);

// act(gate) [* up] -> int8 (+ scale), or -- n_amax_out > 0 -- only the row maxima of that product into amax_out[d][row];
// amax_in: quantise with the maximum over `amax_slots` external arrays (fused_quant.cu)
int launch_act_mul_quant(const void* gate, const void* up, int dtype, int act,
                         int64_t M, int64_t K, int64_t ldg, int64_t ldu,
                         int8_t* hq, int64_t ldq, float* s_h, void* h, int64_t ldh,
                         const pq_quant_spec& spec, cudaStream_t stream,
                         const float* amax_in, int amax_slots, long long amax_stride,
                         float* const* amax_out, int n_amax_out);

int launch_qgemm(const int8_t* a, int64_t lda, const int8_t* b, int64_t ldb,
                 const float* s_x, const float* s_w, const float* bias,
                 void* const* outs, int n_out, int out_dtype, int64_t ldo,
                 int64_t M, int64_t N, int64_t K, cudaStream_t stream, int64_t scatter_cols = 0, int multimem = 0,
                 int64_t rot_cols = 0);

int launch_qlinear_smallm_fused(const void* x, int x_dtype, int64_t ldx,
                                const int8_t* b, int64_t ldb, const float* s_w, const float* bias,
                                void* out, int out_dtype, int64_t ldo,
                                int64_t M, int64_t N, int64_t K, const pq_quant_spec& spec,
                                int num_sms, cudaStream_t stream);

}  // namespace pq
