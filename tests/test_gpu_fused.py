"""GPU parity tests of the producer-fused quantizers: RMSNorm / LayerNorm -> int8 and
act(gate) * up -> int8.  Two halves, as for the GEMM epilogue:
  * integer half, bit-exact: (xq, s_x) == oracle.quantize_rowwise(tensor the kernel emitted);
  * floating-point half, stated tolerance: the emitted tensor vs an fp64 reference of the op,
    |err| <= TOL[dtype] * |ref| + 1e-6 * max|ref|  (two roundings to the storage dtype for RMSNorm and
    act*up: TOL = 2^-7 for bf16, 2^-10 for fp16, 2e-6 for fp32)."""
import numpy as np
import pytest
import torch

import protoquant_b200 as pq
import protoquant_oracle as O

pytestmark = pytest.mark.gpu

TOL = {torch.bfloat16: 2.0 ** -7, torch.float16: 2.0 ** -10, torch.float32: 2e-6}
DTYPES = [torch.bfloat16, torch.float16, torch.float32]
SHAPES = [(5, 768), (17, 4096), (64, 11008), (3, 28672), (300, 3072), (2048, 4096), (1, 8), (33, 1000)]


def _rand(shape, dt, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dt)


def _close(y, ref64, dt):
    ref = torch.from_numpy(ref64)
    err = (y.cpu().double() - ref).abs()
    return bool((err <= TOL[dt] * ref.abs() + 1e-6 * ref.abs().max()).all())


def _int_half_exact(xq, s_x, emitted, spec=None):
    ospec = O.QuantSpec(scale_mode=spec.scale_mode, eps=spec.eps) if spec is not None else O.SPEC_V0
    q_o, s_o = O.quantize_rowwise(emitted.reshape(-1, emitted.shape[-1]).cpu(), ospec)
    return np.array_equal(xq.cpu().numpy(), q_o) and np.array_equal(s_x.cpu().numpy().view(np.uint32), s_o.view(np.uint32))


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", SHAPES)
def test_rmsnorm_quant(shape, dt):
    if shape[1] % (16 // torch.empty(0, dtype=dt).element_size()):
        pytest.skip("K not a multiple of the 16-byte vector")
    x = _rand(shape, dt, 1, 3.0)
    x[0, 0] = 50.0
    w = (_rand((shape[1],), dt, 2) * 0.1 + 1.0).to(dt)
    xq, s_x, y = pq.rmsnorm_quant(x.cuda(), w.cuda(), eps=1e-5, return_normed=True)
    assert _close(y, O.rmsnorm_ref(x, w, 1e-5), dt)
    assert _int_half_exact(xq, s_x, y)
    xq2, s2 = pq.rmsnorm_quant(x.cuda(), w.cuda(), eps=1e-5)
    assert torch.equal(xq, xq2) and torch.equal(s_x, s2)
    # and the fused result is what the unfused pair of OUR ops gives on the emitted tensor
    xq3, s3 = pq.quantize_act(y)
    assert torch.equal(xq, xq3) and torch.equal(s_x, s3)


def test_rmsnorm_matches_torch_llama_composition():
    """HF-Llama RMSNorm in bf16 (fp32 statistics, two roundings): identical up to the summation order of
    mean(x^2), i.e. at most one bf16 ulp on a handful of elements."""
    x = _rand((512, 4096), torch.bfloat16, 3, 2.0).cuda()
    w = (_rand((4096,), torch.bfloat16, 4) * 0.1 + 1.0).to(torch.bfloat16).cuda()
    _, _, y = pq.rmsnorm_quant(x, w, eps=1e-6, return_normed=True)
    xf = x.float()
    ref = w * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).to(torch.bfloat16)
    diff = (y.view(torch.int16).int() - ref.view(torch.int16).int()).abs()
    assert int(diff.max()) <= 1 and float((diff != 0).float().mean()) < 1e-2


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(5, 768), (4096, 768), (300, 3072), (17, 4096), (2, 1024)])
def test_layernorm_quant(shape, dt):
    x = _rand(shape, dt, 5, 2.0) + 0.5
    w = (_rand((shape[1],), dt, 6) * 0.1 + 1.0).to(dt)
    b = _rand((shape[1],), dt, 7, 0.1)
    xq, s_x, y = pq.layernorm_quant(x.cuda(), w.cuda(), b.cuda(), eps=1e-12, return_normed=True)
    ref = O.layernorm_ref(x, w, b, 1e-12)
    # LayerNorm rounds once, but |y| can be far below |x - mean| * gamma + |beta| (cancellation): absolute floor
    err = (y.cpu().double() - torch.from_numpy(ref)).abs()
    assert bool((err <= TOL[dt] * torch.from_numpy(ref).abs() + 4 * TOL[dt] * 1e-2 * float(np.abs(ref).max())).all())
    assert _int_half_exact(xq, s_x, y)
    tref = torch.nn.functional.layer_norm(x.cuda().float(), (shape[1],), w.cuda().float(), b.cuda().float(), 1e-12)
    assert float((y.float() - tref).abs().max()) <= 4 * TOL[dt] * float(tref.abs().max())


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("act", ["silu", "gelu", "gelu_tanh", "identity"])
@pytest.mark.parametrize("shape", [(7, 11008), (2048, 11008), (300, 3072), (16, 28672), (1, 64)])
def test_act_mul_quant(shape, act, dt):
    g = _rand(shape, dt, 8, 2.0)
    u = _rand(shape, dt, 9, 1.5)
    hq, s_h, h = pq.act_mul_quant(g.cuda(), u.cuda(), act=act, return_float=True)
    assert _close(h, O.act_mul_ref(g, u, act), dt)
    assert _int_half_exact(hq, s_h, h)
    hq1, s1, h1 = pq.act_mul_quant(g.cuda(), None, act=act, return_float=True)
    tol_ok = _close(h1, O.act_mul_ref(g, None, act), dt)
    assert tol_ok and _int_half_exact(hq1, s1, h1)


def test_act_mul_quant_on_column_slices_of_one_gemm_output():
    """gate and up as the two halves of a fused gate_up projection output [M, 2K] (row stride 2K)."""
    M, K = 100, 11008
    gu = _rand((M, 2 * K), torch.bfloat16, 10, 2.0).cuda()
    hq, s_h, h = pq.act_mul_quant(gu[:, :K], gu[:, K:], act="silu", return_float=True)
    hq2, s2, h2 = pq.act_mul_quant(gu[:, :K].contiguous(), gu[:, K:].contiguous(), act="silu", return_float=True)
    assert torch.equal(hq, hq2) and torch.equal(s_h, s2) and torch.equal(h, h2)
    ref = torch.nn.functional.silu(gu[:, :K]) * gu[:, K:]          # torch's own bf16 composition
    diff = (h.view(torch.int16).int() - ref.view(torch.int16).int()).abs()
    assert int(diff.max()) <= 1 and float((diff != 0).float().mean()) < 2e-2


def test_quant_spec_knobs_apply_to_the_fused_kernels():
    x = _rand((40, 4096), torch.bfloat16, 11).cuda()
    w = torch.ones(4096, dtype=torch.bfloat16).cuda()
    for spec in (pq.QuantSpec(scale_mode=1, eps=1e-5), pq.QuantSpec(scale_mode=2)):
        xq, s_x, y = pq.rmsnorm_quant(x, w, spec=spec, return_normed=True)
        assert _int_half_exact(xq, s_x, y, spec)
        hq, s_h, h = pq.act_mul_quant(x, x, spec=spec, return_float=True)
        assert _int_half_exact(hq, s_h, h, spec)


def test_llama_mlp_block_fused_chain_equals_unfused_chain():
    """rmsnorm_quant -> gate/up GEMMs -> act_mul_quant -> down GEMM (4 launches + 3 GEMMs) gives the same bits as
    norm -> DynamicQuantLinear x2 -> silu*up -> DynamicQuantLinear on the emitted intermediates."""
    torch.manual_seed(0)
    M, H, I = 256, 1024, 2816
    x = torch.randn(M, H).to(torch.bfloat16).cuda()
    w_norm = torch.ones(H, dtype=torch.bfloat16).cuda()
    gate = pq.DynamicQuantLinear.from_float(torch.nn.Linear(H, I, bias=False).to(torch.bfloat16).cuda())
    up = pq.DynamicQuantLinear.from_float(torch.nn.Linear(H, I, bias=False).to(torch.bfloat16).cuda())
    down = pq.DynamicQuantLinear.from_float(torch.nn.Linear(I, H, bias=False).to(torch.bfloat16).cuda())
    xq, s_x, xn = pq.rmsnorm_quant(x, w_norm, return_normed=True)
    g, u = gate((xq, s_x)), up((xq, s_x))
    hq, s_h, h = pq.act_mul_quant(g, u, act="silu", return_float=True)
    y_fused = down((hq, s_h))
    g2, u2 = gate(xn), up(xn)
    assert torch.equal(g, g2) and torch.equal(u, u2)
    y_unfused = down(h)
    assert torch.equal(y_fused, y_unfused)
    # QTensor input is the same thing
    assert torch.equal(gate(pq.QTensor(xq, s_x, orig_dtype=torch.bfloat16, orig_shape=(M, H))), g)


def test_unsupported_shapes_are_refused_not_approximated():
    lib = pq.lib()
    x = torch.randn(4, 100).to(torch.bfloat16).cuda()        # K = 100 is not a multiple of 8
    w = torch.ones(100, dtype=torch.bfloat16).cuda()
    with pytest.raises(pq.ProtoquantError, match="multiple"):
        pq.rmsnorm_quant(x, w)
    with pytest.raises(pq.ProtoquantError, match="multiple"):
        pq.act_mul_quant(x, x)
    assert lib.pq_act_mul_quant(x.data_ptr(), None, 2, 9, 4, 96, 100, 0, x.data_ptr(), 96, x.data_ptr(), None, 0, None, None) == 1


def test_empty_input():
    x = torch.empty(0, 64, dtype=torch.bfloat16, device="cuda")
    xq, s = pq.rmsnorm_quant(x, torch.ones(64, dtype=torch.bfloat16, device="cuda"))
    assert xq.shape == (0, 64) and s.shape == (0,)
