// Per-row symmetric int8 quantisation arithmetic shared by the row-wise quantizer kernels and the
// fused decode kernel, so that every path produces bit-identical codes and scales
// (see rowwise_quant.cu for the derivation of the exact FMA division).
#pragma once
#include "common.cuh"

namespace pq {
namespace qmath {

constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23

template <typename T> struct VecTraits;
template <> struct VecTraits<float> { static constexpr int EPV = 4; };
template <> struct VecTraits<__half> { static constexpr int EPV = 8; };
template <> struct VecTraits<__nv_bfloat16> { static constexpr int EPV = 8; };

__device__ __forceinline__ uint4 ld_stream_16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// ---- unpack a 16-byte vector to fp32 -------------------------------------------
template <typename T> __device__ __forceinline__ void unpack(const uint4& v, float* f);
template <> __device__ __forceinline__ void unpack<float>(const uint4& v, float* f) {
  f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
  f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
}
template <> __device__ __forceinline__ void unpack<__nv_bfloat16>(const uint4& v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
template <> __device__ __forceinline__ void unpack<__half>(const uint4& v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
    const float2 t = __half22float2(h);
    f[2 * i] = t.x; f[2 * i + 1] = t.y;
  }
}

// ---- |.|-max of a 16-byte vector, returned as fp32 -----------------------------
template <typename T> __device__ __forceinline__ float vec_absmax(const uint4& v, float m);
template <> __device__ __forceinline__ float vec_absmax<float>(const uint4& v, float m) {
  m = fmaxf(m, fabsf(__uint_as_float(v.x))); m = fmaxf(m, fabsf(__uint_as_float(v.y)));
  m = fmaxf(m, fabsf(__uint_as_float(v.z))); m = fmaxf(m, fabsf(__uint_as_float(v.w)));
  return m;
}
// For 16-bit floats |x| ordering == ordering of the 15 magnitude bits as integers, so
// the max is taken on packed u16 lanes (exact) and converted once at the end.
__device__ __forceinline__ uint32_t absmax_u16x2(const uint4& v, uint32_t m) {
  m = __vmaxu2(m, v.x & 0x7fff7fffu); m = __vmaxu2(m, v.y & 0x7fff7fffu);
  m = __vmaxu2(m, v.z & 0x7fff7fffu); m = __vmaxu2(m, v.w & 0x7fff7fffu);
  return m;
}
template <typename T> __device__ __forceinline__ float u16_mag_to_float(uint32_t packed);
template <> __device__ __forceinline__ float u16_mag_to_float<__nv_bfloat16>(uint32_t p) {
  const uint32_t m = max(p & 0xffffu, p >> 16);
  return __uint_as_float(m << 16);
}
template <> __device__ __forceinline__ float u16_mag_to_float<__half>(uint32_t p) {
  const uint32_t m = max(p & 0xffffu, p >> 16);
  return __half2float(__ushort_as_half((unsigned short)m));
}
template <> __device__ __forceinline__ float u16_mag_to_float<float>(uint32_t) { return 0.f; }

// ---- per-row quantisation parameters --------------------------------------------
struct RowQ {
  float s;      // stored scale
  float mul;    // RN(1/s) (DIV fast path, RCP_MUL) or RN(127/amax) (INV_SCALE)
  int path;     // 0 = fma-division, 1 = div.rn, 2 = single multiply
};

__device__ __forceinline__ RowQ make_rowq(float amax, int scale_mode, float eps) {
  RowQ r;
  const float a = (eps > 0.f) ? fmaxf(amax, eps) : amax;
  float s = __fdiv_rn(a, 127.0f);
  if (a == 0.f) s = 1.0f;
  r.s = s;
  if (scale_mode == PQ_DIV) {
    const bool safe = (s >= 0x1p-60f) && (s <= 0x1p60f);
    r.path = safe ? 0 : 1;
    r.mul = __frcp_rn(s);
  } else if (scale_mode == PQ_RCP_MUL) {
    r.path = 2;
    r.mul = __frcp_rn(s);
  } else {
    r.path = 2;
    r.mul = (a == 0.f) ? 1.0f : __fdiv_rn(127.0f, a);
  }
  return r;
}

// returns a float whose low mantissa byte is the int8 code
__device__ __forceinline__ float quant_fast(float x, const RowQ& r) {
  const float q0 = __fmul_rn(x, r.mul);
  const float rem = __fmaf_rn(-q0, r.s, x);
  const float q1 = __fmaf_rn(rem, r.mul, q0);
  return __fadd_rn(q1, kMagic);
}
__device__ __forceinline__ float quant_div(float x, const RowQ& r) {
  return __fadd_rn(__fdiv_rn(x, r.s), kMagic);
}
__device__ __forceinline__ float quant_mul(float x, const RowQ& r) {
  return __fadd_rn(__fmul_rn(x, r.mul), kMagic);
}
__device__ __forceinline__ uint32_t pack4(float a, float b, float c, float d) {
  const uint32_t lo = __byte_perm(__float_as_uint(a), __float_as_uint(b), 0x0040);
  const uint32_t hi = __byte_perm(__float_as_uint(c), __float_as_uint(d), 0x0040);
  return __byte_perm(lo, hi, 0x5410);
}
__device__ __forceinline__ int8_t code_of(float magic_sum) {
  return (int8_t)(__float_as_uint(magic_sum) & 0xffu);
}

template <int PATH>
__device__ __forceinline__ float quant_one(float x, const RowQ& r) {
  if (PATH == 0) return quant_fast(x, r);
  if (PATH == 1) return quant_div(x, r);
  return quant_mul(x, r);
}


}  // namespace qmath
}  // namespace pq
