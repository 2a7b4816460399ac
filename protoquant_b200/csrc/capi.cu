// C ABI of libprotoquant_b200.so (declared in include/protoquant_b200.h).
#include "common.cuh"

#include <stdarg.h>
#include <string.h>
#include <mutex>
#include <new>

namespace pq {

static thread_local char tl_error[512] = "";
std::atomic<uint64_t> g_launch_count{0};
Knob g_pdl{1};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
  va_end(ap);
}

int check_device(int* num_sms) {
  // per-device cache: 0 = unknown, >0 = SM count of an sm_100 device, -1 = unsupported
  static int cache[64] = {0};
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    PQ_FAIL(PQ_ERR_DEVICE, "no usable CUDA device: %s (protoquant_b200 has no CPU fallback)",
            cudaGetErrorString(e));
  }
  if (dev < 0 || dev >= 64) PQ_FAIL(PQ_ERR_DEVICE, "device ordinal %d out of range", dev);
  if (cache[dev] == 0) {
    int major = 0, minor = 0, sms = 0;
    PQ_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    PQ_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    PQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cache[dev] = (major == 10 && minor == 0 && sms > 0) ? sms : -1;
    if (cache[dev] < 0)
      set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev, major, minor);
  }
  if (cache[dev] < 0) {
    set_error("device %d is not an sm_100 part; this library is built for sm_100a (B200) only", dev);
    return PQ_ERR_DEVICE;
  }
  if (num_sms) *num_sms = cache[dev];
  return PQ_OK;
}

}  // namespace pq

using namespace pq;

extern "C" {

int pq_version(void) { return PQ_VERSION; }
const char* pq_last_error(void) { return tl_error; }
uint64_t pq_launch_count(void) { return g_launch_count.load(std::memory_order_relaxed); }
/* test/bench hook: 0 disables programmatic dependent launch */
void pq_debug_set_pdl(int on) { g_pdl = on; }

int pq_act_quant(const void* x, int x_dtype, int64_t M, int64_t K, int64_t ldx,
                 int8_t* xq, int64_t ldq, float* s_x, int transpose,
                 const pq_quant_spec* spec, void* stream) {
  int rc = check_device(nullptr);
  if (rc) return rc;
  return launch_rowwise_quant(x, x_dtype, M, K, ldx, xq, ldq, s_x, transpose, resolve_spec(spec),
                              (cudaStream_t)stream);
}

int pq_weight_quant(const void* W, int w_dtype, int64_t N, int64_t K, int64_t ldw,
                    int8_t* Wq, int64_t ldwq, float* s_w,
                    const pq_quant_spec* spec, void* stream) {
  int rc = check_device(nullptr);
  if (rc) return rc;
  return launch_rowwise_quant(W, w_dtype, N, K, ldw, Wq, ldwq, s_w, 0, resolve_spec(spec),
                              (cudaStream_t)stream);
}

int pq_qgemm(const int8_t* xq, int64_t lda, const int8_t* Wq, int64_t ldb,
             const float* s_x, const float* s_w, const float* bias,
             void* y, int y_dtype, int64_t ldy,
             int64_t M, int64_t N, int64_t K, void* stream) {
  if (y_dtype != PQ_BF16 && y_dtype != PQ_F16 && y_dtype != PQ_F32)
    PQ_FAIL(PQ_ERR_ARG, "pq_qgemm: y_dtype must be PQ_BF16, PQ_F16 or PQ_F32");
  void* outs[1] = {y};
  return launch_qgemm(xq, lda, Wq, ldb, s_x, s_w, bias, outs, 1, y_dtype, ldy, M, N, K, (cudaStream_t)stream);
}

int pq_qgemm_multi(const int8_t* xq, int64_t lda, const int8_t* Wq, int64_t ldb,
                   const float* s_x, const float* s_w, const float* bias,
                   void* const* ys, int n_ys, int y_dtype, int64_t ldy,
                   int64_t M, int64_t N, int64_t K, void* stream) {
  if (y_dtype != PQ_BF16 && y_dtype != PQ_F16 && y_dtype != PQ_F32)
    PQ_FAIL(PQ_ERR_ARG, "pq_qgemm_multi: y_dtype must be PQ_BF16, PQ_F16 or PQ_F32");
  return launch_qgemm(xq, lda, Wq, ldb, s_x, s_w, bias, ys, n_ys, y_dtype, ldy, M, N, K, (cudaStream_t)stream);
}

int pq_qlinear_multi(const void* x, int x_dtype, int64_t ldx,
                     const int8_t* Wq, int64_t ldb, const float* s_w, const float* bias,
                     void* const* ys, int n_ys, int y_dtype, int64_t ldy,
                     int8_t* xq_ws, float* sx_ws,
                     int64_t M, int64_t N, int64_t K,
                     const pq_quant_spec* spec, int flags, void* stream) {
  if (M == 0) return PQ_OK;
  if (!xq_ws || !sx_ws) PQ_FAIL(PQ_ERR_ARG, "pq_qlinear_multi: null workspace");
  if (y_dtype != PQ_BF16 && y_dtype != PQ_F16 && y_dtype != PQ_F32)
    PQ_FAIL(PQ_ERR_ARG, "pq_qlinear_multi: y_dtype must be PQ_BF16, PQ_F16 or PQ_F32");
  int rc = check_device(nullptr);
  if (rc) return rc;
  const int64_t ldq = (K + 15) / 16 * 16;
  const long long w_bytes = (M > 64 && Wq && N > 0 && ldb >= K) ? (long long)(N - 1) * ldb + K : 0;
  rc = launch_rowwise_quant(x, x_dtype, M, K, ldx, xq_ws, ldq, sx_ws, 0, resolve_spec(spec), (cudaStream_t)stream,
                            Wq, w_bytes < (64LL << 20) ? w_bytes : (64LL << 20));
  if (rc) return rc;
  return launch_qgemm(xq_ws, ldq, Wq, ldb, sx_ws, s_w, bias, ys, n_ys, y_dtype, ldy, M, N, K, (cudaStream_t)stream, 0,
                      (flags & PQ_MULTI_MULTICAST) ? 1 : 0);
}

int pq_qgemm_i32(const int8_t* xq, int64_t lda, const int8_t* Wq, int64_t ldb,
                 int32_t* acc, int64_t ldc, int64_t M, int64_t N, int64_t K, void* stream) {
  void* outs[1] = {acc};
  return launch_qgemm(xq, lda, Wq, ldb, nullptr, nullptr, nullptr, outs, 1, PQ_I32, ldc, M, N, K,
                      (cudaStream_t)stream);
}

int pq_qlinear(const void* x, int x_dtype, int64_t ldx,
               const int8_t* Wq, int64_t ldb, const float* s_w, const float* bias,
               void* y, int y_dtype, int64_t ldy,
               int8_t* xq_ws, float* sx_ws,
               int64_t M, int64_t N, int64_t K,
               const pq_quant_spec* spec, void* stream) {
  if (M == 0) return PQ_OK;
  if (!xq_ws || !sx_ws) PQ_FAIL(PQ_ERR_ARG, "pq_qlinear: null workspace");
  if (M <= 32 && x && Wq && s_w && y && N > 0 && K > 0 && ldx >= K && ldy >= N &&
      (y_dtype == PQ_BF16 || y_dtype == PQ_F16 || y_dtype == PQ_F32)) {
    // decode batch: one fused launch (act-quant inside the weight-streaming GEMM) when eligible
    int num_sms = 0;
    int rc0 = check_device(&num_sms);
    if (rc0) return rc0;
    const pq_quant_spec sp = resolve_spec(spec);
    if (sp.scale_mode >= PQ_DIV && sp.scale_mode <= PQ_INV_SCALE) {
      rc0 = launch_qlinear_smallm_fused(x, x_dtype, ldx, Wq, ldb, s_w, bias, y, y_dtype, ldy, M, N, K, sp, num_sms,
                                        (cudaStream_t)stream);
      if (rc0 != 1) return rc0;
    }
  }
  const int64_t ldq = (K + 15) / 16 * 16;
  int rc = check_device(nullptr);
  if (rc) return rc;
  // M > 64: the tcgen05 GEMM re-reads each weight tile from L2 once per 256 tokens; have the quantizer
  // pull the (DRAM-cold) weights into L2 meanwhile.  Capped well below the 126 MB L2.
  // (bytes of the [N, K] view: the last row ends after K bytes even when its stride is larger)
  const long long w_bytes = (M > 64 && Wq && N > 0 && ldb >= K) ? (long long)(N - 1) * ldb + K : 0;
  rc = launch_rowwise_quant(x, x_dtype, M, K, ldx, xq_ws, ldq, sx_ws, 0, resolve_spec(spec), (cudaStream_t)stream,
                            Wq, w_bytes < (64LL << 20) ? w_bytes : (64LL << 20));
  if (rc) return rc;
  return pq_qgemm(xq_ws, ldq, Wq, ldb, sx_ws, s_w, bias, y, y_dtype, ldy, M, N, K, stream);
}

// ---- host-buffer convenience handle ------------------------------------------------
struct pq_linear {
  int64_t N, K, max_tokens, ldk;
  int act_dtype, out_dtype;
  pq_quant_spec spec;
  int8_t* Wq;
  float* s_w;
  float* bias;
  void* x_dev;
  void* y_dev;
  int8_t* xq;
  float* s_x;
  cudaStream_t stream;
};

void pq_linear_destroy(pq_linear* h) {
  if (!h) return;
  if (h->stream) cudaStreamSynchronize(h->stream);
  cudaFree(h->Wq); cudaFree(h->s_w); cudaFree(h->bias);
  cudaFree(h->x_dev); cudaFree(h->y_dev); cudaFree(h->xq); cudaFree(h->s_x);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int pq_linear_create(pq_linear** out, const void* W_host, int w_dtype,
                     int64_t N, int64_t K, const float* bias_host,
                     int64_t max_tokens, int act_dtype, int out_dtype,
                     const pq_quant_spec* spec) {
  if (!out || !W_host || N < 1 || K < 1 || max_tokens < 1) PQ_FAIL(PQ_ERR_ARG, "pq_linear_create: bad argument");
  if (dtype_size(w_dtype) == 0 || w_dtype == PQ_I32 || dtype_size(act_dtype) == 0 || act_dtype == PQ_I32)
    PQ_FAIL(PQ_ERR_ARG, "pq_linear_create: bad dtype");
  if (out_dtype != PQ_BF16 && out_dtype != PQ_F16 && out_dtype != PQ_F32)
    PQ_FAIL(PQ_ERR_ARG, "pq_linear_create: bad out dtype");
  int rc = check_device(nullptr);
  if (rc) return rc;
  pq_linear* h = new (std::nothrow) pq_linear();
  if (!h) PQ_FAIL(PQ_ERR_CUDA, "pq_linear_create: out of host memory");
  memset(h, 0, sizeof(*h));
  h->N = N; h->K = K; h->max_tokens = max_tokens;
  h->ldk = (K + 15) / 16 * 16;
  h->act_dtype = act_dtype; h->out_dtype = out_dtype;
  h->spec = resolve_spec(spec);
  void* W_dev = nullptr;
  cudaError_t e = cudaSuccess;
  auto ok = [&](cudaError_t err) { if (e == cudaSuccess) e = err; return e == cudaSuccess; };
  ok(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  ok(cudaMalloc(&W_dev, (size_t)N * K * dtype_size(w_dtype)));
  ok(cudaMalloc((void**)&h->Wq, (size_t)N * h->ldk));
  ok(cudaMalloc((void**)&h->s_w, (size_t)N * 4));
  if (bias_host) ok(cudaMalloc((void**)&h->bias, (size_t)N * 4));
  ok(cudaMalloc(&h->x_dev, (size_t)max_tokens * K * dtype_size(act_dtype)));
  ok(cudaMalloc(&h->y_dev, (size_t)max_tokens * N * dtype_size(out_dtype)));
  ok(cudaMalloc((void**)&h->xq, (size_t)max_tokens * h->ldk));
  ok(cudaMalloc((void**)&h->s_x, (size_t)max_tokens * 4));
  if (e == cudaSuccess) {
    ok(cudaMemsetAsync(h->Wq, 0, (size_t)N * h->ldk, h->stream));
    ok(cudaMemsetAsync(h->xq, 0, (size_t)max_tokens * h->ldk, h->stream));
    ok(cudaMemcpyAsync(W_dev, W_host, (size_t)N * K * dtype_size(w_dtype), cudaMemcpyHostToDevice, h->stream));
    if (bias_host) ok(cudaMemcpyAsync(h->bias, bias_host, (size_t)N * 4, cudaMemcpyHostToDevice, h->stream));
  }
  if (e == cudaSuccess) {
    rc = launch_rowwise_quant(W_dev, w_dtype, N, K, K, h->Wq, h->ldk, h->s_w, 0, h->spec, h->stream);
    if (rc == PQ_OK) ok(cudaStreamSynchronize(h->stream));
  }
  cudaFree(W_dev);
  if (e != cudaSuccess) {
    pq_linear_destroy(h);
    PQ_FAIL(PQ_ERR_CUDA, "pq_linear_create: %s", cudaGetErrorString(e));
  }
  if (rc) { pq_linear_destroy(h); return rc; }
  *out = h;
  return PQ_OK;
}

int pq_linear_forward_host(pq_linear* h, const void* x_host, void* y_host, int64_t M) {
  if (!h || !x_host || !y_host) PQ_FAIL(PQ_ERR_ARG, "pq_linear_forward_host: null argument");
  if (M < 0 || M > h->max_tokens) PQ_FAIL(PQ_ERR_ARG, "pq_linear_forward_host: M=%lld exceeds max_tokens=%lld", (long long)M, (long long)h->max_tokens);
  if (M == 0) return PQ_OK;
  PQ_CUDA(cudaMemcpyAsync(h->x_dev, x_host, (size_t)M * h->K * dtype_size(h->act_dtype), cudaMemcpyHostToDevice, h->stream));
  int rc = pq_qlinear(h->x_dev, h->act_dtype, h->K, h->Wq, h->ldk, h->s_w, h->bias, h->y_dev, h->out_dtype,
                      h->N, h->xq, h->s_x, M, h->N, h->K, &h->spec, h->stream);
  if (rc) return rc;
  PQ_CUDA(cudaMemcpyAsync(y_host, h->y_dev, (size_t)M * h->N * dtype_size(h->out_dtype), cudaMemcpyDeviceToHost, h->stream));
  PQ_CUDA(cudaStreamSynchronize(h->stream));
  return PQ_OK;
}

}  // extern "C"
