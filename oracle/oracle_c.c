/* Plain-C restatement of SPEC v0.  TEST INFRASTRUCTURE ONLY: linked or
 * loaded solely by tests/ (as an independent cross-check of oracle/protoquant_oracle.py)
 * and by tools/check_fma_div.c-style proofs.  PARITY UNPINNED: the reference checkout is
 * absent, so this follows BASELINE.json north_star + SPEC v0, not a
 * protoquant file:line.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no FMA contraction, every
 * operation is an individually rounded IEEE fp32 op). */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

/* scale_mode: 0 = x / s, 1 = x * (1/s), 2 = x * (127/amax) */
void pqo_quantize_rowwise_f32(const float* x, int64_t rows, int64_t cols, int64_t ldx,
                              int scale_mode, float eps, int qmin,
                              int8_t* q, int64_t ldq, float* s_out) {
  for (int64_t r = 0; r < rows; ++r) {
    const float* xr = x + r * ldx;
    float amax = 0.f;
    for (int64_t c = 0; c < cols; ++c) {
      float a = fabsf(xr[c]);
      if (amax == amax && (a != a || a > amax)) amax = a;   /* NaN propagates (and then stays) */
    }
    if (eps > 0.f && amax < eps) amax = eps;
    volatile float s = amax / 127.0f;
    if (amax == 0.f) s = 1.0f;
    s_out[r] = s;
    volatile float inv = (scale_mode == 2) ? (amax == 0.f ? 1.0f : 127.0f / amax) : 1.0f / s;
    for (int64_t c = 0; c < cols; ++c) {
      volatile float t = (scale_mode == 0) ? xr[c] / s : xr[c] * inv;
      float v = nearbyintf(t); /* default rounding mode: half to even */
      if (v != v) v = 0.f;     /* non-finite policy: NaN -> 0, everything else saturates through the clamp */
      if (v < (float)qmin) v = (float)qmin;
      if (v > 127.f) v = 127.f;
      q[r * ldq + c] = (int8_t)v;
    }
  }
}

/* acc[m,n] = sum_k a[m,k] * b[n,k], exact int32 */
void pqo_int_mm(const int8_t* a, const int8_t* b, int64_t M, int64_t N, int64_t K, int32_t* acc) {
  for (int64_t m = 0; m < M; ++m)
    for (int64_t n = 0; n < N; ++n) {
      int32_t s = 0;
      const int8_t* ar = a + m * K;
      const int8_t* br = b + n * K;
      for (int64_t k = 0; k < K; ++k) s += (int32_t)ar[k] * (int32_t)br[k];
      acc[m * N + n] = s;
    }
}

/* y = ((float(acc) * s_x[m]) * s_w[n]) + bias[n], fp32, each op rounded separately */
void pqo_epilogue_f32(const int32_t* acc, const float* s_x, const float* s_w, const float* bias,
                      int64_t M, int64_t N, float* y) {
  for (int64_t m = 0; m < M; ++m)
    for (int64_t n = 0; n < N; ++n) {
      volatile float v = (float)acc[m * N + n];
      v = v * s_x[m];
      v = v * s_w[n];
      if (bias) v = v + bias[n];
      y[m * N + n] = v;
    }
}

void pqo_dequant_f32(const int8_t* q, const float* s, int axis, int64_t rows, int64_t cols, float* out) {
  for (int64_t r = 0; r < rows; ++r)
    for (int64_t c = 0; c < cols; ++c) {
      volatile float v = (float)q[r * cols + c] * s[axis == 0 ? r : c];
      out[r * cols + c] = v;
    }
}

/* Row-parallel (K-split) pieces, SURVEY.md §8f-3.  A K-shard quantises its column slice with the maximum of
 * the WHOLE row (amax_in, after the cross-shard max), so codes and scale equal those of the unsplit quantizer. */
void pqo_quantize_rowwise_amax_f32(const float* x, int64_t rows, int64_t cols, int64_t ldx, const float* amax_in,
                                   int scale_mode, float eps, int qmin, int8_t* q, int64_t ldq, float* s_out) {
  for (int64_t r = 0; r < rows; ++r) {
    const float* xr = x + r * ldx;
    float amax = amax_in[r];
    if (eps > 0.f && amax < eps) amax = eps;
    volatile float s = amax / 127.0f;
    if (amax == 0.f) s = 1.0f;
    s_out[r] = s;
    volatile float inv = (scale_mode == 2) ? (amax == 0.f ? 1.0f : 127.0f / amax) : 1.0f / s;
    for (int64_t c = 0; c < cols; ++c) {
      volatile float t = (scale_mode == 0) ? xr[c] / s : xr[c] * inv;
      float v = nearbyintf(t);
      if (v != v) v = 0.f;
      if (v < (float)qmin) v = (float)qmin;
      if (v > 127.f) v = 127.f;
      q[r * ldq + c] = (int8_t)v;
    }
  }
}

/* y = ((float(sum_p part_p) * s_x[m]) * s_w[n]) + bias[n]: the reduce + dequant half of the fused
 * GEMM + reduce-scatter.  parts is [n_parts][M][N] int32; the sum is exact. */
void pqo_reduce_epilogue_f32(const int32_t* parts, int n_parts, const float* s_x, const float* s_w, const float* bias,
                             int64_t M, int64_t N, float* y) {
  for (int64_t m = 0; m < M; ++m)
    for (int64_t n = 0; n < N; ++n) {
      int32_t acc = 0;
      for (int p = 0; p < n_parts; ++p) acc += parts[((int64_t)p * M + m) * N + n];
      volatile float v = (float)acc;
      v = v * s_x[m];
      v = v * s_w[n];
      if (bias) v = v + bias[n];
      y[m * N + n] = v;
    }
}
