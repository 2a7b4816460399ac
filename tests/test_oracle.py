"""CPU tests of the oracle itself: golden vectors, hand-derived known answers, the plain-C
restatement as an independent cross-check, and the knob sensitivity SURVEY.md §8c reports."""
import ctypes
import glob
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import EXACT_SPECS, GOLDEN, ROOT, load_golden_x, same_scales
import protoquant_oracle as O


# ---- golden vectors produced by torch.ao's per-token ops (tests/golden/make_golden.py) ----
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "torch_ao_*.npz"))))
def test_oracle_matches_torch_ao_golden(path):
    d = np.load(path)
    x = load_golden_x(d)
    q, s = O.quantize_rowwise(x, O.QuantSpec.torch_ao())
    assert np.array_equal(q, d["q"])
    assert np.array_equal(s.view(np.uint32), d["s"].view(np.uint32))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "int_mm_*.npz"))))
def test_int_mm_golden(path):
    d = np.load(path)
    assert np.array_equal(O.int_mm(d["a"], d["b"]), d["acc"])


# ---- golden vectors of the DEFAULT spec from exact rational arithmetic (tests/golden/make_golden_exact.py) ----
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "exact_*.npz"))))
def test_oracle_matches_exact_rational_golden(path):
    """A third, independent producer (integer / Fraction arithmetic only, no hardware fp division or rounding) pins
    SPEC v0 -- true division, no eps, the product default -- and the other knob sets, including the NaN / inf /
    denormal-scale rows."""
    d = np.load(path)
    x = load_golden_x(d)
    for label, mode, eps, qmin in EXACT_SPECS:
        q, s = O.quantize_rowwise(x, O.QuantSpec(scale_mode=mode, eps=eps, qmin=qmin))
        assert np.array_equal(q, d["q_" + label]), label
        assert same_scales(s, d["s_" + label].view(np.float32)), label


def test_epilogue_and_dequantize_match_exact_rational_golden(clib):
    """Rows a4 / a5 against the exact-rational producer: ((float(acc) * s_x) * s_w) + bias with every step rounded to
    binary32, then ONE rounding to bf16 / fp16 (overflow -> inf), incl. accumulators that fp32 cannot represent, tiny and
    huge scales, a -0.0 bias; numpy oracle and C restatement."""
    d = np.load(os.path.join(GOLDEN, "epilogue_exact_24x40.npz"))
    acc, sx, sw = d["acc"], d["s_x"], d["s_w"]
    M, N = acc.shape
    with np.errstate(over="ignore"):
        for tag, b in (("bias", d["bias"]), ("nobias", None)):
            y = O.dequant_epilogue(acc, sx, sw, b)
            assert np.array_equal(y.view(np.uint32), d["y_f32_" + tag])
            assert np.array_equal(O.cast_out(y, "bf16").view(torch.int16).numpy().view(np.uint16), d["y_bf16_" + tag])
            assert np.array_equal(O.cast_out(y, "f16").view(torch.int16).numpy().view(np.uint16), d["y_f16_" + tag])
            yc = np.empty((M, N), np.float32)
            accc, bc = np.ascontiguousarray(acc), (np.ascontiguousarray(b) if b is not None else None)
            clib.pqo_epilogue_f32(accc.ctypes.data_as(ctypes.c_void_p), sx.ctypes.data_as(ctypes.c_void_p),
                                  sw.ctypes.data_as(ctypes.c_void_p), bc.ctypes.data_as(ctypes.c_void_p) if bc is not None else None,
                                  ctypes.c_int64(M), ctypes.c_int64(N), yc.ctypes.data_as(ctypes.c_void_p))
            assert np.array_equal(yc.view(np.uint32), d["y_f32_" + tag])
    assert np.array_equal(O.dequantize(d["q"], sx, 0).view(np.uint32), d["dequant_rows_f32"])


def test_nonfinite_policy_kat():
    """NaN / inf propagate into the scale; the row's codes are all zero.  A scale that underflows to 0 saturates."""
    x = np.array([[1.0, -2.0, np.inf, 0.5], [1.0, np.nan, -np.inf, 0.0], [3e-45, -3e-45, 0.0, 1.4e-45]], np.float32)
    for mode in (O.DIV, O.RCP_MUL, O.INV_SCALE):
        q, s = O.quantize_rowwise(x, O.QuantSpec(scale_mode=mode))
        assert np.isinf(s[0]) and s[0] > 0 and np.isnan(s[1])
        assert not q[0].any() and not q[1].any()
    q, s = O.quantize_rowwise(x)
    assert s[2] == 0.0 and q[2].tolist() == [127, -128, 0, 127]
    q, s = O.quantize_rowwise(x, O.QuantSpec(qmin=-127))
    assert q[2].tolist() == [127, -127, 0, 127]


# ---- hand-derived known answers for SPEC v0 ---------------------------------------------
def test_kat_round_half_even_and_scale():
    # amax = 127 -> s = 1 exactly, so q = rne(x): ties go to the even integer
    x = np.array([[127.0, 0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 3.49, -126.5, 126.5, 0.0, -127.0]], np.float32)
    q, s = O.quantize_rowwise(x)
    assert s[0] == np.float32(1.0)
    assert q[0].tolist() == [127, 0, 2, 2, 0, -2, -2, 3, -126, 126, 0, -127]


def test_kat_zero_row_and_symmetry():
    x = np.zeros((2, 8), np.float32)
    x[1] = [-254, 254, 2, -2, 1, -1, 3, -3]
    q, s = O.quantize_rowwise(x)
    assert s[0] == 1.0 and not q[0].any()
    assert s[1] == np.float32(2.0)
    assert q[1].tolist() == [-127, 127, 1, -1, 0, 0, 2, -2]   # 0.5 -> 0, 1.5 -> 2 (RNE)
    # -128 is never produced by a symmetric absmax scale
    g = np.random.default_rng(0)
    q2, _ = O.quantize_rowwise(g.standard_normal((64, 333)).astype(np.float32))
    assert q2.min() >= -127 and q2.max() <= 127
    assert (np.abs(q2).max(axis=1) == 127).all()


def test_kat_epilogue_association():
    # ((acc*s_x)*s_w)+bias evaluated left to right in fp32, NOT acc*(s_x*s_w)
    acc = np.array([[16777217]], np.int32)          # not representable in fp32 -> 16777216
    s_x = np.array([3.0], np.float32)
    s_w = np.array([1.0 / 3.0], np.float32)
    y = O.dequant_epilogue(acc, s_x, s_w, np.array([0.25], np.float32))
    t = np.float32(16777216.0) * np.float32(3.0)
    t = np.float32(t) * np.float32(1.0 / 3.0)
    assert y[0, 0] == np.float32(t + np.float32(0.25))


def test_qmin_knob_is_inert_for_finite_input():
    g = np.random.default_rng(1)
    x = g.standard_normal((32, 257)).astype(np.float32) * 50
    for mode in (O.DIV, O.RCP_MUL, O.INV_SCALE):
        a = O.quantize_rowwise(x, O.QuantSpec(scale_mode=mode, qmin=-128))
        b = O.quantize_rowwise(x, O.QuantSpec(scale_mode=mode, qmin=-127))
        assert np.array_equal(a[0], b[0])


def test_knobs_really_differ():
    """operation order decides bit-exactness (x/s vs x*(1/s) disagree on bf16 data)."""
    torch.manual_seed(0)
    x = torch.randn(2048, 4096).to(torch.bfloat16)
    qd, _ = O.quantize_rowwise(x, O.QuantSpec(scale_mode=O.DIV))
    qr, _ = O.quantize_rowwise(x, O.QuantSpec(scale_mode=O.RCP_MUL))
    n = int((qd != qr).sum())
    assert 100 < n < 20000, n


# ---- plain-C restatement as an independent implementation ----------------------------------
@pytest.fixture(scope="module")
def clib():
    so = os.path.join(ROOT, "oracle", "liboracle_c.so")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(so)
    return lib


def _c_quant(lib, x32, mode, eps, qmin=-128):
    R, C = x32.shape
    q = np.empty((R, C), np.int8)
    s = np.empty((R,), np.float32)
    lib.pqo_quantize_rowwise_f32(x32.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(R), ctypes.c_int64(C),
                                 ctypes.c_int64(C), ctypes.c_int(mode), ctypes.c_float(eps), ctypes.c_int(qmin),
                                 q.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(C), s.ctypes.data_as(ctypes.c_void_p))
    return q, s


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("mode,eps", [(O.DIV, 0.0), (O.RCP_MUL, 1e-5), (O.INV_SCALE, 0.0)])
def test_numpy_oracle_equals_c_restatement(clib, dtype, mode, eps):
    torch.manual_seed(3)
    x = torch.randn(257, 1031) * torch.logspace(-3, 3, 257)[:, None]
    x[5].zero_()
    x[6, 17] = 1e4
    x = x.to(dtype)
    x32 = np.ascontiguousarray(O.to_f32(x))
    q, s = O.quantize_rowwise(x, O.QuantSpec(scale_mode=mode, eps=eps))
    qc, sc = _c_quant(clib, x32, mode, eps)
    assert np.array_equal(q, qc)
    assert np.array_equal(s.view(np.uint32), sc.view(np.uint32))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "exact_*.npz"))))
def test_c_restatement_matches_exact_rational_golden(clib, path):
    d = np.load(path)
    x32 = np.ascontiguousarray(O.to_f32(load_golden_x(d)))
    for label, mode, eps, qmin in EXACT_SPECS:
        q, s = _c_quant(clib, x32, mode, eps, qmin)
        assert np.array_equal(q, d["q_" + label]), label
        assert same_scales(s, d["s_" + label].view(np.float32)), label


def test_c_int_mm_and_epilogue(clib):
    g = np.random.default_rng(5)
    M, N, K = 19, 23, 301
    a = g.integers(-128, 128, (M, K), dtype=np.int8)
    b = g.integers(-128, 128, (N, K), dtype=np.int8)
    acc = np.empty((M, N), np.int32)
    clib.pqo_int_mm(a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(M),
                    ctypes.c_int64(N), ctypes.c_int64(K), acc.ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(acc, O.int_mm(a, b))
    assert np.array_equal(acc, a.astype(np.int64) @ b.astype(np.int64).T)
    sx = g.random(M, dtype=np.float32)
    sw = g.random(N, dtype=np.float32)
    bias = g.standard_normal(N).astype(np.float32)
    y = np.empty((M, N), np.float32)
    clib.pqo_epilogue_f32(acc.ctypes.data_as(ctypes.c_void_p), sx.ctypes.data_as(ctypes.c_void_p),
                          sw.ctypes.data_as(ctypes.c_void_p), bias.ctypes.data_as(ctypes.c_void_p),
                          ctypes.c_int64(M), ctypes.c_int64(N), y.ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(y.view(np.uint32), O.dequant_epilogue(acc, sx, sw, bias).view(np.uint32))


def test_int_mm_large_matches_torch_and_int64():
    g = np.random.default_rng(6)
    a = g.integers(-128, 128, (64, 4096), dtype=np.int8)
    b = g.integers(-128, 128, (128, 4096), dtype=np.int8)
    assert np.array_equal(O.int_mm(a, b), (a.astype(np.int64) @ b.astype(np.int64).T).astype(np.int32))


def test_qlinear_close_to_float_linear():
    torch.manual_seed(7)
    x = torch.randn(48, 512)
    w = (torch.rand(256, 512) * 2 - 1) / 512 ** 0.5
    b = torch.randn(256)
    wq, sw = O.quantize_rowwise(w)
    y = O.qlinear(x, wq, sw, b.numpy(), out_dtype="f32")
    ref = torch.nn.functional.linear(x, w, b)
    err = (y - ref).abs().max().item()
    assert err < 0.05 * ref.abs().max().item()


def test_dequantize_roundtrip_error_bound():
    torch.manual_seed(8)
    x = torch.randn(33, 700)
    q, s = O.quantize_rowwise(x)
    xr = O.dequantize(q, s, axis=0)
    assert (np.abs(xr - x.numpy()) <= s[:, None] * 0.5 * (1 + 1e-6)).all()


def test_torch_threaded_baseline_matches_oracle():
    torch.manual_seed(9)
    x = torch.randn(64, 1024).to(torch.bfloat16)
    w = (torch.rand(512, 1024) * 2 - 1) / 32
    b = torch.randn(512)
    wq, sw = O.quantize_rowwise(w)
    y0 = O.qlinear(x, wq, sw, b.numpy(), out_dtype="bf16")
    y1 = O.qlinear_torch_cpu(x, torch.from_numpy(wq).t(), torch.from_numpy(sw), b, torch.bfloat16)
    assert torch.equal(y0, y1)


def test_producer_op_references_match_torch_float64():
    """oracle.rmsnorm_ref / layernorm_ref / act_mul_ref (the fp64 references of tests/test_gpu_fused.py)."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(6, 96, generator=g, dtype=torch.float64)
    w = torch.randn(96, generator=g, dtype=torch.float64)
    b = torch.randn(96, generator=g, dtype=torch.float64)
    ref = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6) * w
    assert np.allclose(O.rmsnorm_ref(x.float(), w.float(), 1e-6), ref.numpy(), rtol=1e-5, atol=1e-6)
    ln = torch.nn.functional.layer_norm(x, (96,), w, b, 1e-5)
    assert np.allclose(O.layernorm_ref(x.float(), w.float(), b.float(), 1e-5), ln.numpy(), rtol=1e-5, atol=1e-5)
    for act, fn in (("silu", torch.nn.functional.silu), ("gelu", torch.nn.functional.gelu),
                    ("gelu_tanh", lambda t: torch.nn.functional.gelu(t, approximate="tanh"))):
        assert np.allclose(O.act_mul_ref(x.float(), w.expand(6, 96).float(), act), (fn(x) * w).numpy(), rtol=1e-5, atol=1e-6)
        assert np.allclose(O.act_mul_ref(x.float(), None, act), fn(x).numpy(), rtol=1e-5, atol=1e-6)


def test_k_split_dataflow_equals_unsplit_linear_numpy_and_c(clib):
    """SURVEY.md §8f-3 on the CPU: K-shards quantise their slices with the global row maximum, multiply, and the
    int32 partial sums are reduced before ONE epilogue -- the result has the bits of the unsplit linear, in the numpy
    oracle and in the independent C restatement alike."""
    torch.manual_seed(11)
    M, N, K, G = 9, 21, 208, 4
    x = torch.randn(M, K).to(torch.bfloat16)
    x[0, K - 1] = 60.0                                   # row maximum in the last shard
    w = (torch.rand(N, K) * 2 - 1) / K ** 0.5
    b = torch.randn(N)
    wq, sw = O.quantize_rowwise(w)
    full = O.qlinear(x, wq, sw, b.numpy(), out_dtype="f32").numpy()
    x32 = np.ascontiguousarray(O.to_f32(x))
    amax = np.max(np.abs(x32), axis=-1).astype(np.float32)
    per = 64
    parts = np.zeros((G, M, N), np.int32)
    s_x = None
    for r in range(G):
        lo, hi = min(r * per, K), min((r + 1) * per, K)
        q_np, s_x = O.quantize_rowwise(x32[:, lo:hi], amax=amax)
        xs = np.ascontiguousarray(x32[:, lo:hi])
        q_c = np.empty((M, hi - lo), np.int8)
        s_c = np.empty((M,), np.float32)
        clib.pqo_quantize_rowwise_amax_f32(xs.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(M), ctypes.c_int64(hi - lo),
                                           ctypes.c_int64(hi - lo), amax.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(0),
                                           ctypes.c_float(0.0), ctypes.c_int(-128), q_c.ctypes.data_as(ctypes.c_void_p),
                                           ctypes.c_int64(hi - lo), s_c.ctypes.data_as(ctypes.c_void_p))
        assert np.array_equal(q_np, q_c) and np.array_equal(s_x, s_c)
        parts[r] = O.int_mm(q_np, np.ascontiguousarray(wq[:, lo:hi]))
    # the slices' codes are exactly the columns of the unsplit quantisation
    q_full, s_full = O.quantize_rowwise(x)
    assert np.array_equal(s_x, s_full)
    y_np = O.dequant_epilogue(parts.sum(0, dtype=np.int32), s_x, sw, b.numpy())
    y_c = np.empty((M, N), np.float32)
    bias = b.numpy().astype(np.float32)
    clib.pqo_reduce_epilogue_f32(parts.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(G), s_x.ctypes.data_as(ctypes.c_void_p),
                                 sw.ctypes.data_as(ctypes.c_void_p), bias.ctypes.data_as(ctypes.c_void_p),
                                 ctypes.c_int64(M), ctypes.c_int64(N), y_c.ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(y_np.view(np.uint32), full.view(np.uint32))
    assert np.array_equal(y_c.view(np.uint32), full.view(np.uint32))
