"""GPU parity tests of the row-wise int8 quantizer against the
CPU oracle: bit-exact int8 payloads and fp32 scales.  Everything goes through the C ABI
(protoquant_b200.functional is a ctypes shim over libprotoquant_b200.so)."""
import ctypes
import glob
import os

import numpy as np
import pytest
import torch

from conftest import EXACT_SPECS, GOLDEN, load_golden_x, same_scales
import protoquant_b200 as pq
from protoquant_b200 import functional as F
import protoquant_oracle as O

pytestmark = pytest.mark.gpu

DTYPES = [torch.bfloat16, torch.float16, torch.float32]
SPECS = [(pq.QuantSpec(), O.QuantSpec()),
         (pq.QuantSpec(scale_mode=1, eps=1e-5), O.QuantSpec.torch_ao()),
         (pq.QuantSpec(scale_mode=2), O.QuantSpec(scale_mode=O.INV_SCALE))]


def make_x(M, K, dtype, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(M, K, generator=g)
    if M > 0:
        x[0, K // 3] = 100.0                       # one heavy outlier
    if M > 1:
        x[1].zero_()                               # all-zero row
    if M > 2:
        x[2] = torch.round(x[2] * 2) / 2           # exact .5 ties once scaled by s = 1
        x[2, 0] = 127.0
    if M > 3:
        x[3] *= 1e-30                              # tiny scale -> div.rn slow path (fp32/bf16), flush for fp16
    if M > 4:
        x[4] *= 1e30 if dtype != torch.float16 else 1e3
    if M > 5:
        x[5, K - 1] = -x[5].abs().max() * 2        # amax on the negative side
    return x.to(dtype)


def check(x, spec, ospec, transpose=False):
    q, s = pq.quantize_act(x.cuda(), transpose=transpose, spec=spec)
    qo, so = O.quantize_rowwise(x, ospec)
    torch.cuda.synchronize()
    qn = q.cpu().numpy()
    assert np.array_equal(qn.T if transpose else qn, qo)
    assert np.array_equal(s.cpu().numpy().view(np.uint32), so.view(np.uint32))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("spec,ospec", SPECS)
@pytest.mark.parametrize("shape", [(1, 8), (7, 768), (9, 3072), (16, 4096), (12, 8192), (6, 11008), (6, 28672)])
def test_act_quant_bit_exact(dtype, spec, ospec, shape):
    check(make_x(*shape, dtype), spec, ospec)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(3, 1), (5, 7), (4, 100), (3, 4099), (2, 65536), (2, 70001)])
def test_act_quant_ragged_and_max_k(dtype, shape):
    """K not a multiple of the vector width, K at and beyond the register-resident limit."""
    check(make_x(*shape, dtype), *SPECS[0])


def test_act_quant_empty():
    q, s = pq.quantize_act(torch.empty(0, 64, dtype=torch.bfloat16, device="cuda"))
    assert q.shape == (0, 64) and s.shape == (0,)


@pytest.mark.parametrize("dtype", DTYPES)
def test_act_quant_strided_rows(dtype):
    big = make_x(33, 1024, dtype, seed=3)
    view = big[:, 128:128 + 512]                  # ldx = 1024 > K = 512
    q, s = pq.quantize_act(big.cuda()[:, 128:640])
    qo, so = O.quantize_rowwise(view)
    assert np.array_equal(q.cpu().numpy(), qo) and np.array_equal(s.cpu().numpy(), so)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("kernel", ["two_pass", "tiled"])
@pytest.mark.parametrize("shape", [(5, 64), (70, 520), (33, 4096), (300, 100), (2048, 4096), (1000, 11008), (129, 8192),
                                   (130, 1000), (64, 28672), (4100, 768)])
def test_act_quant_transposed_output(dtype, shape, kernel):
    """transpose=1: codes written as [K, M].  "two_pass" (default): a row-scale launch, then [128 x 128] tiles quantised
    and transposed through shared memory (128 contiguous bytes per output row; aligned or not, ragged M / K);
    "tiled": the single-launch 32-rows-per-CTA kernel (kept for PQ_INV_SCALE, whose row parameters need amax itself)."""
    pq.lib().pq_debug_set_quant_staged(0 if kernel == "two_pass" else -1)
    try:
        for spec, ospec in SPECS if shape[0] <= 300 else SPECS[:1]:
            check(make_x(*shape, dtype, seed=5), spec, ospec, transpose=True)
    finally:
        pq.lib().pq_debug_set_quant_staged(0)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "torch_ao_*.npz"))))
def test_act_quant_matches_committed_golden(path):
    d = np.load(path)
    x = load_golden_x(d)
    q, s = pq.quantize_act(x.cuda(), spec=pq.QuantSpec(scale_mode=1, eps=1e-5))
    assert np.array_equal(q.cpu().numpy(), d["q"])
    assert np.array_equal(s.cpu().numpy().view(np.uint32), d["s"].view(np.uint32))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "exact_*.npz"))))
@pytest.mark.parametrize("kernel", ["vec", "staged", "generic", "transposed", "transposed_tiled", "given_amax"])
def test_act_quant_matches_exact_rational_golden(path, kernel):
    """Default spec (and the other knob sets) against the exact-rational producer, including the NaN / inf /
    denormal-scale rows, through each quantizer kernel: the register-resident vector kernel, the generic one
    (misaligned rows), the transposed one and the given-row-maximum variant of the row-parallel path."""
    d = np.load(path)
    x = load_golden_x(d)
    M, K = x.shape
    for label, mode, eps, qmin in EXACT_SPECS:
        spec = pq.QuantSpec(scale_mode=mode, eps=eps, qmin=qmin)
        if kernel in ("vec", "staged"):
            pq.lib().pq_debug_set_quant_staged(1 if kernel == "staged" else -1)
            try:
                q, s = pq.quantize_act(x.cuda(), spec=spec)
            finally:
                pq.lib().pq_debug_set_quant_staged(0)
        elif kernel == "generic":
            big = torch.zeros(M, K + 3, dtype=x.dtype)
            big[:, 1:K + 1] = x                        # rows start at an odd element: no 16-byte alignment
            q, s = pq.quantize_act(big.cuda()[:, 1:K + 1], spec=spec)
        elif kernel in ("transposed", "transposed_tiled"):
            pq.lib().pq_debug_set_quant_staged(0 if kernel == "transposed" else -1)
            try:
                q, s = pq.quantize_act(x.cuda(), transpose=True, spec=spec)
            finally:
                pq.lib().pq_debug_set_quant_staged(0)
            q = q.t()
        else:
            amax = F.row_absmax(x.cuda())
            q, s = F.quantize_act_with_amax(x.cuda(), amax, spec=spec)
        assert np.array_equal(q.cpu().numpy(), d["q_" + label]), (label, kernel)
        assert same_scales(s.cpu().numpy(), d["s_" + label].view(np.float32)), (label, kernel)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("spec,ospec", SPECS)
def test_nonfinite_and_denormal_rows_follow_the_policy(dtype, spec, ospec):
    """+-inf / NaN policy and denormals.  NaN and inf propagate into the row's scale and the row's
    codes are zero; other rows of the same launch are untouched; denormal rows (fp32: denormal or zero scale) take
    the literal clamp(rne(x/s)) with NaN -> 0."""
    g = torch.Generator().manual_seed(21)
    x = torch.randn(8, 4096, generator=g)
    x[1, 77] = float("inf")
    x[2, 4095] = float("-inf")
    x[3, 0] = float("nan")
    x[4, 100] = float("nan"); x[4, 200] = float("inf")
    x = x.to(dtype)
    if dtype == torch.float32:
        x[5] = torch.from_numpy((np.arange(4096) % 200).astype(np.uint32).view(np.float32))   # denormals: bit patterns 0..199
        x[5, 1] = -x[5, 1]
        x[6] = 0; x[6, 5] = 1.4e-45; x[6, 9] = -2.8e-45                                  # scale underflows to 0
    q, s = pq.quantize_act(x.cuda(), spec=spec)
    qo, so = O.quantize_rowwise(x, ospec)
    assert np.array_equal(q.cpu().numpy(), qo)
    assert same_scales(s.cpu().numpy(), so)
    sn = s.cpu().numpy()
    assert np.isinf(sn[1]) and np.isinf(sn[2]) and np.isnan(sn[3]) and np.isnan(sn[4])
    assert not q[1:5].any()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("spec,ospec", SPECS)
@pytest.mark.parametrize("shape", [(2, 8), (300, 768), (700, 4096), (333, 11008), (160, 28672), (1500, 64), (40, 57344)])
def test_staged_quantizer_bit_exact(dtype, spec, ospec, shape):
    """The persistent shared-memory staged kernel (one CTA per SM, bulk-copy ring, refilled slots) forced on for every
    shape it accepts: few rows per CTA, many rows per CTA (slot reuse), rows as large as half the shared memory."""
    x = make_x(*shape, dtype, seed=9)
    pq.lib().pq_debug_set_quant_staged(1)
    try:
        check(x, spec, ospec)
        amax = F.row_absmax(x.cuda())
        q, s = F.quantize_act_with_amax(x.cuda(), amax * 2, spec=spec)        # external row maximum
        qo, so = O.quantize_rowwise(x, ospec, amax=(amax * 2).cpu().numpy())
        assert np.array_equal(q.cpu().numpy(), qo) and np.array_equal(s.cpu().numpy(), so)
    finally:
        pq.lib().pq_debug_set_quant_staged(0)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "reference_*.npz"))) or [None])
def test_act_quant_matches_reference_golden(path):
    """Vectors produced by the REAL reference (tools/repin.py --write).  None exist while the checkout is absent."""
    if path is None:
        pytest.skip("no reference_*.npz: the reference checkout is absent (SURVEY.md §0); tools/repin.py creates them")
    d = np.load(path)
    mode, eps, qmin = d["spec"]
    q, s = pq.quantize_act(load_golden_x(d).cuda(), spec=pq.QuantSpec(int(mode), float(eps), int(qmin)))
    assert np.array_equal(q.cpu().numpy(), d["q"]) and same_scales(s.cpu().numpy(), d["s"])


@pytest.mark.parametrize("out,key", [(torch.float32, "f32"), (torch.bfloat16, "bf16"), (torch.float16, "f16")])
def test_epilogue_and_dequantize_match_exact_rational_golden(out, key):
    """The dequant epilogue as a stand-alone kernel (pq_reduce_dequant on one int32 part) and pq_dequant against the
    exact-rational producer: same bits for fp32, bf16 and fp16 outputs (overflow -> inf included)."""
    d = np.load(os.path.join(GOLDEN, "epilogue_exact_24x40.npz"))
    acc = torch.from_numpy(d["acc"]).cuda()
    sx, sw, bias = (torch.from_numpy(d[k]).cuda() for k in ("s_x", "s_w", "bias"))
    for tag, b in (("bias", bias), ("nobias", None)):
        y = F.dequant_accumulators(acc, sx, sw, b, out_dtype=out).cpu()
        got = y.view(torch.int32).numpy().view(np.uint32) if out == torch.float32 else y.view(torch.int16).numpy().view(np.uint16)
        assert np.array_equal(got, d[f"y_{key}_{tag}"]), tag
    if out == torch.float32:
        q = F.alloc_q(*d["q"].shape, "cuda")
        q.copy_(torch.from_numpy(d["q"]))
        dq = pq.dequantize_tensor(q, sx, axis=0, out_dtype=torch.float32).cpu().view(torch.int32).numpy().view(np.uint32)
        assert np.array_equal(dq, d["dequant_rows_f32"])


def test_kat_round_half_even():
    x = torch.tensor([[127.0, 0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 3.49, -126.5, 126.5, 0.0, -127.0, 0, 0, 0, 0]])
    for dt in DTYPES:
        q, s = pq.quantize_act(x.to(dt).cuda())
        assert s.item() == 1.0
        assert q.cpu()[0, :12].tolist() == [127, 0, 2, 2, 0, -2, -2, 3, -126, 126, 0, -127]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape", [(4096, 768), (2048, 4096), (2048, 11008), (512, 28672)])
def test_act_quant_full_size_vs_oracle(dtype, shape):
    """BASELINE.json config sizes (BERT 32x128 tokens, Llama-7B 2048 tokens, Llama-70B K)."""
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(*shape, generator=g).to(dtype)
    check(x, *SPECS[0])


def test_act_quant_large_properties():
    """Size-independent properties at a size the oracle is too slow for (1.6 GB of traffic):
    every row hits |q| == 127, reconstruction error <= s/2, scales == amax/127."""
    M, K = 131072, 4096
    x = torch.randn(M, K, dtype=torch.bfloat16, device="cuda")
    q, s = pq.quantize_act(x)
    xf = x.float()
    amax = xf.abs().amax(dim=1)
    assert torch.equal(s, torch.div(amax, torch.full_like(amax, 127.0)))   # tensor/tensor: IEEE div (x/scalar is x*(1/scalar) on CUDA)
    assert bool((q.abs().amax(dim=1) == 127).all()) and int(q.min()) >= -127
    err = (q.float() * s[:, None] - xf).abs()
    assert bool((err <= s[:, None] * 0.5001).all())      # 0.5 ulp of the int grid + fp32 rounding of x/s and q*s


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(768, 768), (3072, 768), (4096, 4096), (1000, 4096)])
def test_weight_quant_bit_exact(dtype, shape):
    N, K = shape
    g = torch.Generator().manual_seed(0)
    w = ((torch.rand(N, K, generator=g) * 2 - 1) / K ** 0.5).to(dtype)   # nn.Linear default init range
    wq, sw = pq.quantize_weight(w.cuda())
    wo, so = O.quantize_weight(w)
    assert np.array_equal(wq.cpu().numpy(), wo)
    assert np.array_equal(sw.cpu().numpy().view(np.uint32), so.view(np.uint32))
    assert wq.stride(0) % 16 == 0


@pytest.mark.parametrize("out_dtype", DTYPES)
@pytest.mark.parametrize("axis", [0, 1])
@pytest.mark.parametrize("shape", [(5, 7), (64, 4096), (33, 1000)])
def test_dequantize_bit_exact(out_dtype, axis, shape):
    g = np.random.default_rng(0)
    q = g.integers(-128, 128, shape, dtype=np.int8)
    s = g.random(shape[0] if axis == 0 else shape[1], dtype=np.float32) + 0.01
    got = pq.dequantize_tensor(torch.from_numpy(q).cuda(), torch.from_numpy(s).cuda(), axis=axis, out_dtype=out_dtype)
    ref = torch.from_numpy(O.dequantize(q, s, axis)).to(out_dtype)
    assert torch.equal(got.cpu(), ref)


def test_qtensor_quantize_dequantize():
    x = make_x(24, 512, torch.bfloat16, seed=9).reshape(2, 12, 512)
    qt = pq.quantize(x.cuda())
    assert qt.shape == (2, 12, 512) and qt.data.dtype == torch.int8
    qo, so = O.quantize_rowwise(x.reshape(-1, 512))
    assert np.array_equal(qt.int_repr().cpu().numpy(), qo)
    back = qt.dequantize()
    assert back.dtype == torch.bfloat16 and back.shape == x.shape
    assert torch.equal(back.cpu().reshape(-1, 512), torch.from_numpy(O.dequantize(qo, so)).to(torch.bfloat16))
    assert torch.equal(pq.dequantize(qt, torch.float32).cpu().reshape(-1, 512), torch.from_numpy(O.dequantize(qo, so)))


def test_c_abi_direct_call_and_errors():
    """Call the exported symbol with raw pointers (what a non-Python host would do)."""
    lib = pq.lib()
    x = make_x(8, 256, torch.float32).cuda()
    q = torch.empty(8, 256, dtype=torch.int8, device="cuda")
    s = torch.empty(8, dtype=torch.float32, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.pq_act_quant(x.data_ptr(), 0, 8, 256, 256, q.data_ptr(), 256, s.data_ptr(), 0, None, st)
    assert rc == 0
    qo, so = O.quantize_rowwise(x.cpu())
    assert np.array_equal(q.cpu().numpy(), qo) and np.array_equal(s.cpu().numpy(), so)
    assert lib.pq_act_quant(x.data_ptr(), 7, 8, 256, 256, q.data_ptr(), 256, s.data_ptr(), 0, None, st) == 1
    assert lib.pq_act_quant(x.data_ptr(), 0, 8, 256, 128, q.data_ptr(), 256, s.data_ptr(), 0, None, st) == 1
    assert b"ldx" in lib.pq_last_error()


def test_qtensor_axis0_per_column_scales():
    """quantize(t, axis=0): one scale per column; equals the oracle's row-wise quantisation of t^T."""
    g = torch.Generator().manual_seed(71)
    t = torch.randn(300, 96, generator=g).to(torch.bfloat16)
    qt = pq.quantize(t.cuda(), axis=0)
    q_o, s_o = O.quantize_rowwise(t.t().contiguous())
    assert qt.axis == 0 and qt.data.shape == (300, 96) and qt.scale.shape == (96,)
    assert np.array_equal(qt.data.cpu().numpy(), q_o.T) and np.array_equal(qt.scale.cpu().numpy(), s_o)
    ref = torch.from_numpy(O.dequantize(q_o.T, s_o, axis=1)).to(torch.bfloat16)
    assert torch.equal(qt.dequantize().cpu(), ref)
