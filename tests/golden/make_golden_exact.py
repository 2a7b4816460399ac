"""Generates tests/golden/exact_*.npz: golden vectors for the PRODUCT DEFAULT spec (SPEC v0: s = fl(amax/127),
q = clamp(rne(fl(x/s)), qmin, 127), scale_mode DIV, no eps) and for the RCP_MUL / INV_SCALE knobs, from a THIRD,
independent producer.  Run once in the build container:  python tests/golden/make_golden_exact.py

Independence: nothing here uses floating-point division, multiplication or rounding of the host CPU, numpy or
torch.  Inputs are decoded from their bit patterns into exact rationals (`fractions.Fraction`); every IEEE
binary32 operation of the spec is performed as an exact rational operation followed by `round_to_f32`, an
integer-arithmetic implementation of round-to-nearest-even onto the binary32 grid (normal numbers, denormals,
overflow to infinity); the round-to-integer step is exact half-to-even on the rational.  It therefore checks the
two restatements in oracle/ (numpy fp32 and plain C) and the CUDA kernels against IEEE-754 semantics derived from
first principles, not against another run of the same hardware instructions.

It is NOT the reference (the protoquant checkout is absent): it pins the oracle to the written spec,
not to protoquant.

Row menu per file (M = 12 rows): N(0,1) rows, one x100 outlier, an all-zero row, a row of exact .5 ties (amax = 127
-> s = 1), a row whose maximum is negative, a tiny-magnitude row (1e-30: large 1/s), a huge-magnitude row (1e30),
fp32 only: a denormal row whose scale is denormal (clamp live) and a row whose scale underflows to 0
(x/0 = +-inf -> +-127 / qmin, 0/0 = NaN -> 0); every dtype: a row with +inf, a row with NaN (all-zero codes).
"""
import os
import struct
from fractions import Fraction

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
INF, NAN = "inf", "nan"          # non-finite markers in the rational domain (sign carried separately)


# ---- bit patterns -> exact values ----------------------------------------------------------------
def decode(bits: int, ebits: int, mbits: int):
    """(sign, value) of an IEEE-like pattern with `ebits` exponent and `mbits` mantissa bits; value is a
    non-negative Fraction, INF or NAN."""
    sign = (bits >> (ebits + mbits)) & 1
    e = (bits >> mbits) & ((1 << ebits) - 1)
    m = bits & ((1 << mbits) - 1)
    bias = (1 << (ebits - 1)) - 1
    if e == (1 << ebits) - 1:
        return sign, (NAN if m else INF)
    if e == 0:
        return sign, Fraction(m, 1 << mbits) * Fraction(2) ** (1 - bias)
    return sign, (1 + Fraction(m, 1 << mbits)) * Fraction(2) ** (e - bias)


FMT = {"f32": (8, 23), "f16": (5, 10), "bf16": (8, 7)}


def round_to_f32(v: Fraction):
    """Round a non-negative rational to the nearest binary32 value, ties to even.  Returns a Fraction or INF."""
    if v == 0:
        return Fraction(0)
    # find e with 2^e <= v < 2^(e+1)
    n, d = v.numerator, v.denominator
    e = n.bit_length() - d.bit_length()
    if Fraction(2) ** e > v:
        e -= 1
    elif Fraction(2) ** (e + 1) <= v:
        e += 1
    e = max(e, -126)                              # denormals share the exponent of the smallest normal
    ulp = Fraction(2) ** (e - 23)
    q, r = divmod(v, ulp)                         # q integer, 0 <= r < ulp
    q = int(q)
    if r * 2 > ulp or (r * 2 == ulp and (q & 1)):
        q += 1
    out = q * ulp
    if out >= Fraction(2) ** 128:
        return INF
    return out


def f32_bits(sign: int, v) -> int:
    """binary32 bit pattern of (sign, v) with v a Fraction already on the binary32 grid, INF or NAN."""
    if v == NAN:
        return 0x7fc00000
    if v == INF:
        return (sign << 31) | 0x7f800000
    if v == 0:
        return sign << 31
    n, d = v.numerator, v.denominator
    e = n.bit_length() - d.bit_length()
    if Fraction(2) ** e > v:
        e -= 1
    elif Fraction(2) ** (e + 1) <= v:
        e += 1
    if e < -126:
        m = v / Fraction(2) ** (-149)
        assert m.denominator == 1
        return (sign << 31) | int(m)
    m = (v / Fraction(2) ** e - 1) * (1 << 23)
    assert m.denominator == 1, "value is not on the binary32 grid"
    return (sign << 31) | ((e + 127) << 23) | int(m)


# ---- IEEE binary32 operations on (sign, magnitude) pairs -------------------------------------------
def f32_div(a, b):
    (sa, va), (sb, vb) = a, b
    s = sa ^ sb
    if va == NAN or vb == NAN:
        return 0, NAN
    if va == INF:
        return (0, NAN) if vb == INF else (s, INF)
    if vb == INF:
        return s, Fraction(0)
    if vb == 0:
        return (0, NAN) if va == 0 else (s, INF)
    return s, round_to_f32(va / vb)


def f32_mul(a, b):
    (sa, va), (sb, vb) = a, b
    s = sa ^ sb
    if va == NAN or vb == NAN:
        return 0, NAN
    if va == INF or vb == INF:
        other = vb if va == INF else va
        return (0, NAN) if other == 0 else (s, INF)
    return s, round_to_f32(va * vb)


def rne_int(sign: int, v) -> int:
    """clamp-free round-half-even of (sign, v) to a python int; INF -> a huge sentinel; NAN -> None."""
    if v == NAN:
        return None
    if v == INF:
        return -(10 ** 9) if sign else 10 ** 9
    fl = v.numerator // v.denominator
    r = v - fl
    if r > Fraction(1, 2) or (r == Fraction(1, 2) and (fl & 1)):
        fl += 1
    return -fl if sign else fl


def quantize_row(row, mode: int, eps_bits: int, qmin: int):
    """row: list of (sign, magnitude).  Returns (codes, scale bits)."""
    amax = Fraction(0)
    for _, v in row:
        if amax == NAN:
            break
        if v == NAN:
            amax = NAN
        elif v == INF:
            amax = INF
        elif amax != INF and v > amax:
            amax = v
    if eps_bits and amax != NAN:
        eps = decode(eps_bits, 8, 23)[1]
        if amax != INF and amax < eps:
            amax = eps
    c127 = (0, Fraction(127))
    if amax == 0:
        s = (0, Fraction(1))
    else:
        s = f32_div((0, amax), c127)
    if mode == 1:
        inv = f32_div((0, Fraction(1)), s)
    elif mode == 2:
        inv = (0, Fraction(1)) if amax == 0 else f32_div(c127, (0, amax))
    codes = []
    for x in row:
        t = f32_div(x, s) if mode == 0 else f32_mul(x, inv)
        q = rne_int(*t)
        q = 0 if q is None else max(qmin, min(127, q))
        codes.append(q)
    return codes, f32_bits(*s)


# ---- inputs -----------------------------------------------------------------------------------------
def to_bits(x32: np.ndarray, name: str) -> np.ndarray:
    """Round an fp32 array to `name` with torch (input construction only) and return the bit patterns."""
    import torch
    t = torch.from_numpy(x32.copy())
    if name == "f32":
        return x32.view(np.uint32).copy()
    t = t.to(torch.bfloat16 if name == "bf16" else torch.float16)
    return t.view(torch.int16).numpy().view(np.uint16).copy()


def make_rows(name: str, K: int, seed: int) -> np.ndarray:
    g = np.random.default_rng(seed)
    rows = [g.standard_normal(K).astype(np.float32) for _ in range(3)]
    r = g.standard_normal(K).astype(np.float32); r[K // 2] = 100.0; rows.append(r)          # outlier
    rows.append(np.zeros(K, np.float32))                                                     # all-zero row
    r = (g.integers(-253, 254, K).astype(np.float32)) / 2; r[0] = 127.0; rows.append(r)      # exact .5 ties, s = 1
    r = g.standard_normal(K).astype(np.float32); r[3] = -7.5; r = np.clip(r, -7.5, 5); rows.append(r)  # negative max
    big = 6.0e4 if name == "f16" else 1e30
    small = 6.0e-5 if name == "f16" else 1e-30
    rows.append(g.standard_normal(K).astype(np.float32) * np.float32(small))
    rows.append(np.clip(g.standard_normal(K), -1, 1).astype(np.float32) * np.float32(big))
    r = g.standard_normal(K).astype(np.float32); r[1] = np.inf; r[5] = -np.inf; rows.append(r)
    r = g.standard_normal(K).astype(np.float32); r[2] = np.nan; r[7] = np.inf; rows.append(r)
    bits = [to_bits(r, name) for r in rows]
    if name == "f32":
        # denormal scale (clamp live: s rounds coarsely) and scale underflowing to zero
        d1 = g.integers(0, 180, K).astype(np.uint32); d1[0] = 178; d1[1] |= 0x80000000
        d2 = g.integers(0, 4, K).astype(np.uint32); d2[0] = 3; d2[1] = 0x80000002; d2[2] = 0
        bits += [d1, d2]
    else:
        # smallest denormals of the 16-bit formats (their scales stay representable in fp32)
        d1 = g.integers(0, 64, K).astype(np.uint16); d1[0] = 63; d1[1] = 0x8001
        bits += [d1, to_bits(g.standard_normal(K).astype(np.float32), name)]
    return np.stack(bits)


# ---- dequant epilogue (row a4) and QTensor.dequantize (row a5), exact --------------------------------------
def f32_add(a, b):
    (sa, va), (sb, vb) = a, b
    if va == NAN or vb == NAN:
        return 0, NAN
    if va == INF or vb == INF:
        if va == INF and vb == INF:
            return (0, NAN) if sa != sb else (sa, INF)
        return (sa, INF) if va == INF else (sb, INF)
    x = (-va if sa else va) + (-vb if sb else vb)
    if x == 0:
        return (sa & sb), Fraction(0)          # (+0) + (-0) = +0 in round-to-nearest; (-0) + (-0) = -0
    return (1 if x < 0 else 0), round_to_f32(abs(x))


def round_to(v: Fraction, mbits: int, emin: int, emax: int):
    """Round a non-negative rational to a binary format with `mbits` mantissa bits (RNE); INF on overflow."""
    if v == 0:
        return Fraction(0)
    n, d = v.numerator, v.denominator
    e = n.bit_length() - d.bit_length()
    if Fraction(2) ** e > v:
        e -= 1
    elif Fraction(2) ** (e + 1) <= v:
        e += 1
    e = max(e, emin)
    ulp = Fraction(2) ** (e - mbits)
    q, r = divmod(v, ulp)
    q = int(q)
    if r * 2 > ulp or (r * 2 == ulp and (q & 1)):
        q += 1
    out = q * ulp
    return INF if out >= Fraction(2) ** (emax + 1) else out


def bits_of(sign, v, ebits, mbits):
    bias = (1 << (ebits - 1)) - 1
    top = sign << (ebits + mbits)
    if v == NAN:
        return ((1 << ebits) - 1) << mbits | (1 << (mbits - 1))
    if v == INF:
        return top | (((1 << ebits) - 1) << mbits)
    if v == 0:
        return top
    n, d = v.numerator, v.denominator
    e = n.bit_length() - d.bit_length()
    if Fraction(2) ** e > v:
        e -= 1
    elif Fraction(2) ** (e + 1) <= v:
        e += 1
    if e < 1 - bias:
        m = v / Fraction(2) ** (1 - bias - mbits)
        assert m.denominator == 1
        return top | int(m)
    m = (v / Fraction(2) ** e - 1) * (1 << mbits)
    assert m.denominator == 1
    return top | ((e + bias) << mbits) | int(m)


def epilogue_exact(acc: int, sx_bits: int, sw_bits: int, b_bits):
    """((float(acc) * s_x) * s_w) + bias, every step rounded to binary32; returns (sign, value)."""
    t = (1 if acc < 0 else 0, round_to_f32(Fraction(abs(acc))))          # int32 -> fp32, RNE
    t = f32_mul(t, decode(sx_bits, 8, 23))
    t = f32_mul(t, decode(sw_bits, 8, 23))
    if b_bits is not None:
        t = f32_add(t, decode(b_bits, 8, 23))
    return t


def make_epilogue_golden():
    g = np.random.default_rng(2024)
    M, N = 24, 40
    acc = g.integers(-(2 ** 26), 2 ** 26, (M, N), dtype=np.int64)
    acc[0, :8] = [0, 1, -1, 16777217, -16777217, 2 ** 31 - 1, -(2 ** 31), 33554433]     # not representable in fp32 / extremes
    acc[1] = g.integers(-300, 300, N)                                                  # small sums: results near the bias
    sx = (g.random(M) * 0.1 + 1e-3).astype(np.float32); sx[2] = np.float32(1e-30); sx[3] = np.float32(3e30)
    sw = (g.random(N) * 0.01 + 1e-4).astype(np.float32); sw[5] = np.float32(1e-20); sw[6] = np.float32(2e10)
    bias = g.standard_normal(N).astype(np.float32); bias[7] = np.float32(-0.0); bias[8] = np.float32(65504.0)
    out = {"acc": acc.astype(np.int32), "s_x": sx, "s_w": sw, "bias": bias}
    for tag, use_bias in (("bias", True), ("nobias", False)):
        y32 = np.zeros((M, N), np.uint32); y16 = np.zeros((M, N), np.uint16); yb16 = np.zeros((M, N), np.uint16)
        for m in range(M):
            for n in range(N):
                s_, v = epilogue_exact(int(acc[m, n]), int(sx[m].view(np.uint32)), int(sw[n].view(np.uint32)),
                                       int(bias[n].view(np.uint32)) if use_bias else None)
                y32[m, n] = f32_bits(s_, v)
                y16[m, n] = bits_of(s_, v if v in (INF, NAN) else round_to(v, 10, -14, 15), 5, 10)
                yb16[m, n] = bits_of(s_, v if v in (INF, NAN) else round_to(v, 7, -126, 127), 8, 7)
        out[f"y_f32_{tag}"], out[f"y_f16_{tag}"], out[f"y_bf16_{tag}"] = y32, y16, yb16
    # dequantize: q * s, one rounding
    q = g.integers(-128, 128, (M, N), dtype=np.int64).astype(np.int8)
    dq = np.zeros((M, N), np.uint32)
    for m in range(M):
        for n in range(N):
            s_, v = f32_mul((1 if q[m, n] < 0 else 0, Fraction(abs(int(q[m, n])))), decode(int(sx[m].view(np.uint32)), 8, 23))
            dq[m, n] = f32_bits(s_, v)
    out["q"], out["dequant_rows_f32"] = q, dq
    np.savez_compressed(os.path.join(HERE, "epilogue_exact_24x40.npz"), **out)
    print("wrote epilogue_exact_24x40.npz")


def main():
    make_epilogue_golden()
    for name in ("f32", "bf16", "f16"):
        ebits, mbits = FMT[name]
        for K in (40, 96):
            xb = make_rows(name, K, seed=99 + K)
            rows = [[decode(int(b), ebits, mbits) for b in r] for r in xb]
            out = {"x_bits": xb, "shape": np.array(xb.shape), "dtype": name}
            for label, mode, eps_bits, qmin in (("div", 0, 0, -128), ("div_qmin127", 0, 0, -127),
                                                ("rcp_mul_eps1e5", 1, struct.unpack("<I", struct.pack("<f", 1e-5))[0], -128),
                                                ("inv_scale", 2, 0, -128)):
                qs, ss = zip(*(quantize_row(r, mode, eps_bits, qmin) for r in rows))
                out[f"q_{label}"] = np.array(qs, dtype=np.int8)
                out[f"s_{label}"] = np.array(ss, dtype=np.uint32)
            np.savez_compressed(os.path.join(HERE, f"exact_{name}_{xb.shape[0]}x{K}.npz"), **out)
            print("wrote", f"exact_{name}_{xb.shape[0]}x{K}.npz")


if __name__ == "__main__":
    main()
