"""Hypothesis shape fuzzing of the HBM-bound kernels: any (rows, K, dtype, stride) the C ABI accepts
must give the oracle's bits -- vector / generic / transposed quantizer paths, dequantize, and the fused producers'
integer half."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

import protoquant_b200 as pq
import protoquant_oracle as O

pytestmark = pytest.mark.gpu

DTS = [torch.bfloat16, torch.float16, torch.float32]
COMMON = dict(max_examples=40, deadline=None, derandomize=True, database=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])


def _x(rows, K, dt, seed, pad=0):
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(rows, K + pad, generator=g)
    if rows:
        base[0, : min(K, 3)] = torch.tensor([100.0, -0.5, 0.5])[: min(K, 3)]
    if rows > 1:
        base[1].zero_()
    return base.to(dt)


@settings(**COMMON)
@given(rows=st.integers(0, 70), K=st.integers(1, 2500), dti=st.integers(0, 2), pad=st.sampled_from([0, 0, 8, 5]),
       transpose=st.booleans(), seed=st.integers(0, 10 ** 6))
def test_quantizer_any_shape(rows, K, dti, pad, transpose, seed):
    dt = DTS[dti]
    x = _x(rows, K, dt, seed, pad)[:, :K]          # pad > 0: a row-strided view
    q, s = pq.quantize_act(x.cuda(), transpose=transpose)
    q_o, s_o = O.quantize_rowwise(x)
    want = q_o.T if transpose else q_o
    assert np.array_equal(q.cpu().numpy(), want) and np.array_equal(s.cpu().numpy(), s_o)


@settings(**COMMON)
@given(rows=st.integers(1, 40), k8=st.integers(1, 600), dti=st.integers(0, 2), seed=st.integers(0, 10 ** 6),
       act=st.sampled_from(["silu", "gelu", "gelu_tanh", "identity"]))
def test_fused_producers_integer_half_any_shape(rows, k8, dti, seed, act):
    dt = DTS[dti]
    K = k8 * 8                                      # multiple of the 16-byte vector for every dtype
    x = _x(rows, K, dt, seed).cuda()
    w = torch.ones(K, dtype=dt, device="cuda")
    for xq, s_x, emitted in (pq.rmsnorm_quant(x, w, return_normed=True),
                             pq.layernorm_quant(x, w, torch.zeros_like(w), return_normed=True),
                             pq.act_mul_quant(x, x, act=act, return_float=True)):
        q_o, s_o = O.quantize_rowwise(emitted.cpu())
        assert np.array_equal(xq.cpu().numpy(), q_o) and np.array_equal(s_x.cpu().numpy(), s_o)


@settings(**COMMON)
@given(M=st.integers(1, 300), N=st.integers(1, 700), k16=st.integers(1, 40), seed=st.integers(0, 10 ** 6))
def test_gemm_any_shape(M, N, k16, seed):
    K = k16 * 16
    g = torch.Generator().manual_seed(seed)
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, generator=g)
    b = torch.randint(-128, 128, (N, K), dtype=torch.int8, generator=g)
    acc = pq.qgemm_i32(a.cuda(), b.cuda())
    ref = O.int_mm(a.numpy(), b.numpy())
    assert np.array_equal(acc.cpu().numpy(), ref)
    s_x = torch.rand(M, generator=g) * 0.1 + 1e-3
    s_w = torch.rand(N, generator=g) * 0.01 + 1e-4
    y = pq.qgemm(a.cuda(), s_x.cuda(), b.cuda(), s_w.cuda(), None, torch.bfloat16)
    want = O.cast_out(O.dequant_epilogue(ref, s_x.numpy(), s_w.numpy(), None), "bf16")
    assert torch.equal(y.cpu().view(torch.int16), want.view(torch.int16))


@settings(**COMMON)
@given(rows=st.integers(1, 700), k8=st.integers(1, 1500), dti=st.integers(0, 2), pad=st.sampled_from([0, 0, 8, 16]),
       seed=st.integers(0, 10 ** 6))
def test_staged_quantizer_any_shape(rows, k8, dti, pad, seed):
    """The persistent shared-memory staged quantizer forced on: any row count (slot refills, fewer groups than slots),
    any 16-byte-multiple row length, strided rows."""
    dt = DTS[dti]
    K = k8 * 8
    x = _x(rows, K, dt, seed, pad)[:, :K]
    pq.lib().pq_debug_set_quant_staged(1)
    try:
        q, s = pq.quantize_act(x.cuda())
    finally:
        pq.lib().pq_debug_set_quant_staged(0)
    q_o, s_o = O.quantize_rowwise(x)
    assert np.array_equal(q.cpu().numpy(), q_o) and np.array_equal(s.cpu().numpy(), s_o)


@settings(**COMMON)
@given(M=st.integers(65, 600), n8=st.integers(1, 120), k16=st.integers(1, 24), bn=st.sampled_from([0, 256, 224, 128]),
       ndest=st.integers(2, 4), rot=st.sampled_from([0, 0, 256, 520]), seed=st.integers(0, 10 ** 6))
def test_multi_destination_and_scatter_gemm_any_shape(M, n8, k16, bn, ndest, rot, seed):
    """Fused all-gather (every destination gets the tile) and reduce-scatter (column block d goes to destination d)
    epilogues on every staged tile width, with a rotated tile order: the bits of the plain GEMM, nothing else touched."""
    from protoquant_b200 import functional as F
    N, K = n8 * 8, k16 * 16
    g = torch.Generator().manual_seed(seed)
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, generator=g).cuda()
    b = torch.randint(-128, 128, (N, K), dtype=torch.int8, generator=g).cuda()
    s_x = (torch.rand(M, generator=g) * 0.1 + 1e-3).cuda()
    s_w = (torch.rand(N, generator=g) * 0.01 + 1e-4).cuda()
    ref16 = pq.qgemm(a, s_x, b, s_w, None, torch.bfloat16)
    ref32 = pq.qgemm_i32(a, b)
    per = (-(-N // ndest) + 7) // 8 * 8
    pq.lib().pq_debug_set_multi_bn(bn)
    pq.lib().pq_debug_set_tile_rotation(rot)
    try:
        bufs = [torch.zeros(M + 2, N + 16, dtype=torch.bfloat16, device="cuda") for _ in range(ndest)]
        F.qgemm_multi(a, s_x, b, s_w, None, [t.data_ptr() + 16 for t in bufs], N + 16, torch.bfloat16)
        inbox = torch.full((ndest, M, per), 7, dtype=torch.int32, device="cuda")
        F.qgemm_i32_scatter(a, b, [inbox[d].data_ptr() for d in range(ndest)], per, per)
    finally:
        pq.lib().pq_debug_set_multi_bn(0)
        pq.lib().pq_debug_set_tile_rotation(0)
    for t in bufs:
        assert torch.equal(t[:M, 8:8 + N], ref16) and not t[:, :8].any() and not t[:, 8 + N:].any() and not t[M:].any()
    for d in range(ndest):
        lo, hi = min(d * per, N), min((d + 1) * per, N)
        assert torch.equal(inbox[d][:, : hi - lo], ref32[:, lo:hi]) and bool((inbox[d][:, hi - lo:] == 7).all())
