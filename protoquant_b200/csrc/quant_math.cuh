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

// ---- |.|-max of a 16-byte vector ------------------------------------------------
// The maximum is taken on the magnitude BITS as unsigned integers: for non-negative floats the integer order is
// the float order, +inf sorts above every finite value and every NaN above +inf -- so a NaN anywhere in the row
// PROPAGATES into amax (fmaxf would drop it) and the non-finite policy below is the same for every dtype.
template <typename T> __device__ __forceinline__ float vec_absmax(const uint4& v, float m);
template <> __device__ __forceinline__ float vec_absmax<float>(const uint4& v, float m) {
  uint32_t b = __float_as_uint(m);
  b = max(b, v.x & 0x7fffffffu); b = max(b, v.y & 0x7fffffffu);
  b = max(b, v.z & 0x7fffffffu); b = max(b, v.w & 0x7fffffffu);
  return __uint_as_float(b);
}
// max of two magnitudes (non-negative or NaN), NaN-propagating
__device__ __forceinline__ float mag_max(float a, float b) {
  return __uint_as_float(max(__float_as_uint(a), __float_as_uint(b)));
}
__device__ __forceinline__ float mag_of(float x) { return __uint_as_float(__float_as_uint(x) & 0x7fffffffu); }
// For 16-bit floats |x| ordering == ordering of the 15 magnitude bits as integers, so
// the max is taken on packed u16 lanes (exact, NaN on top) and converted once at the end.
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

// Row maximum supplied from outside (row-parallel shards): the maximum over `slots` arrays of M
// floats `stride` elements apart -- one per K-shard, written by the peers into this rank's symmetric memory -- read
// through L2 (ld.cg: the producers are other GPUs).  NaN propagates (integer max on the magnitude bits).
__device__ __forceinline__ float given_amax(const float* amax_in, int slots, long long stride, int64_t row) {
  float a = __ldcg(amax_in + row);
  for (int sl = 1; sl < slots; ++sl) a = mag_max(a, __ldcg(amax_in + sl * stride + row));
  return a;
}

// ---- per-row quantisation parameters --------------------------------------------
// Policy for rows the "fast" arithmetic cannot take (include/protoquant_b200.h "Non-finite and denormal input"):
//   * amax propagates NaN and inf (see above), s = amax/127 in IEEE arithmetic: NaN -> NaN, inf -> inf;
//   * q = clamp(rne(x / s), qmin, qmax) evaluated in IEEE arithmetic with NaN -> 0 (what a float -> int8
//     conversion of NaN gives on x86 and on the GPU): a row holding +-inf or NaN gets all-zero codes (finite/inf = 0,
//     inf/inf = NaN -> 0, anything/NaN = NaN -> 0);
//   * fp32 rows whose amax is so small that s is denormal or underflows to 0 (amax < 127 * 2^-126) quantise with
//     the clamp live: x/0 = +-inf -> qmax/qmin, 0/0 = NaN -> 0, and a coarsely rounded denormal s can push
//     |x/s| above 127.
// Those rows take the "careful" paths (1: division, 3: multiply); everything else the two fast ones.
struct RowQ {
  float s;      // stored scale
  float mul;    // RN(1/s) (DIV fast path, RCP_MUL) or RN(127/amax) (INV_SCALE)
  int path;     // 0 = fma-division, 1 = careful div.rn, 2 = single multiply, 3 = careful multiply
  float qmin;   // lower clamp of the careful paths (-128 or -127)
};

// `mode` = pq::mode_bits(spec): scale_mode in the low byte, bit 8 set for qmin = -127
__device__ __forceinline__ RowQ make_rowq(float amax, int mode, float eps) {
  RowQ r;
  const int scale_mode = mode & 0xff;
  const int qmin = (mode & 0x100) ? -127 : -128;
  // fmaxf drops a NaN operand; the clamp of the spec (torch.clamp / np.maximum) keeps it
  const float a = (eps > 0.f && amax == amax) ? fmaxf(amax, eps) : amax;
  float s = __fdiv_rn(a, 127.0f);
  if (a == 0.f) s = 1.0f;
  r.s = s;
  r.qmin = (float)qmin;
  const bool safe = (s >= 0x1p-60f) && (s <= 0x1p60f);       // false for NaN, inf, 0 and denormal scales
  if (scale_mode == PQ_DIV) {
    r.path = safe ? 0 : 1;
    r.mul = __frcp_rn(s);
  } else if (scale_mode == PQ_RCP_MUL) {
    r.path = safe ? 2 : 3;
    r.mul = __frcp_rn(s);
  } else {
    r.path = safe ? 2 : 3;
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
// careful paths: NaN -> 0, then clamp (the bounds are integers, so clamping before the rounding is the same)
__device__ __forceinline__ float clamp_nan0(float d, const RowQ& r) {
  d = (d != d) ? 0.f : d;
  return fminf(fmaxf(d, r.qmin), 127.0f);
}
__device__ __forceinline__ float quant_div(float x, const RowQ& r) {
  return __fadd_rn(clamp_nan0(__fdiv_rn(x, r.s), r), kMagic);
}
__device__ __forceinline__ float quant_mul(float x, const RowQ& r) {
  return __fadd_rn(__fmul_rn(x, r.mul), kMagic);
}
__device__ __forceinline__ float quant_mul_careful(float x, const RowQ& r) {
  return __fadd_rn(clamp_nan0(__fmul_rn(x, r.mul), r), kMagic);
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
  if (PATH == 2) return quant_mul(x, r);
  return quant_mul_careful(x, r);
}
// run-time path (kernels that do not specialise their inner loop)
__device__ __forceinline__ float quant_any(float x, const RowQ& r) {
  if (r.path == 0) return quant_fast(x, r);
  if (r.path == 1) return quant_div(x, r);
  if (r.path == 2) return quant_mul(x, r);
  return quant_mul_careful(x, r);
}


}  // namespace qmath
}  // namespace pq
