"""GPU parity tests of the tcgen05 int8 GEMM + fused dequant epilogue:
int32 accumulators bit-exact against an exact CPU integer matmul for every tile configuration,
epilogue bit-exact against the oracle's fp32 operation sequence, and within the stated bf16/fp16
tolerance of the un-quantised float linear."""
import ctypes
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
import protoquant_b200 as pq
import protoquant_oracle as O

pytestmark = pytest.mark.gpu

CONFIGS = [-1, 0, 1, 2, 3, 4, 8, 9, 10, 11, 12, 13, 16]   # 8-13 = narrow tiles (BLOCK_N 240/224/208), 16 = 4-CTA multicast cluster; -1 = heuristic (small-M kernel for M <= 64); see launch_typed() in csrc/qgemm_tcgen05.cu


@pytest.fixture(autouse=True)
def _reset_cfg():
    yield
    pq.lib().pq_debug_set_gemm_config(-1)
    pq.lib().pq_debug_set_streamk(-1)
    pq.lib().pq_debug_set_staged(0)


def rand_i8(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(-128, 128, shape, dtype=torch.int8, generator=g)


def cpu_int_mm(a, b):
    return torch.from_numpy(O.int_mm(a.numpy(), b.numpy()))


def _bits(t):
    return t.view(torch.int32 if t.dtype == torch.float32 else torch.int16)


@pytest.mark.parametrize("sk", [0, 1])   # 0 = data-parallel tiles only, 1 = stream-K whenever legal
@pytest.mark.parametrize("cfg", CONFIGS)
@pytest.mark.parametrize("shape", [
    (1, 8, 16), (1, 64, 16), (16, 4096, 4096), (17, 40, 144), (128, 256, 128), (129, 257, 130),
    (256, 256, 256), (300, 520, 1040), (384, 768, 768), (255, 1000, 3072), (512, 3072, 768),
])
def test_int32_accumulators_bit_exact(cfg, shape, sk):
    M, N, K = shape
    pq.lib().pq_debug_set_gemm_config(cfg)
    pq.lib().pq_debug_set_streamk(sk)
    a, b = rand_i8((M, K), 1), rand_i8((N, K), 2)
    got = pq.qgemm_i32(a.cuda(), b.cuda())
    assert torch.equal(got.cpu(), cpu_int_mm(a, b))


@pytest.mark.parametrize("sk", [0, 1])
@pytest.mark.parametrize("cfg", [0, 1, 4, 8, 11, 16])
@pytest.mark.parametrize("shape", [(2048, 4096, 4096), (2048, 11008, 4096), (2048, 4096, 11008),
                                   (4096, 3072, 768), (1024, 3584, 8192), (256, 8192, 28672)])
def test_int32_full_size_shapes(cfg, shape, sk):
    pq.lib().pq_debug_set_streamk(sk)
    """BASELINE.json configs: Llama-7B (4096/11008, 2048 tokens), BERT (768/3072, 4096 tokens),
    Llama-70B column shard (28672/8 = 3584 out-channels, K=8192) and down-proj K=28672."""
    M, N, K = shape
    pq.lib().pq_debug_set_gemm_config(cfg)
    a, b = rand_i8((M, K), 3), rand_i8((N, K), 4)
    got = pq.qgemm_i32(a.cuda(), b.cuda())
    assert torch.equal(got.cpu(), cpu_int_mm(a, b))


@pytest.mark.parametrize("shape", [(128, 1024, 16384), (96, 4096, 11008), (256, 2048, 8192), (64, 4096, 1024),
                                   (40, 1000, 2048), (64, 4096, 4096), (128, 4096, 4096)])
def test_default_heuristics_bit_exact(shape):
    """What launch_qgemm picks by itself: stream-K for single-wave long-K problems, 128x64 tiles instead of
    the small-M kernel for short K, the small-M kernel otherwise -- accumulators and epilogue bit-exact."""
    M, N, K = shape
    g = torch.Generator().manual_seed(61)
    xq, wq = rand_i8((M, K), 62), rand_i8((N, K), 63)
    ref = cpu_int_mm(xq, wq)
    for _ in range(2):      # twice: the stream-K counters must be clean for the second launch
        assert torch.equal(pq.qgemm_i32(xq.cuda(), wq.cuda()).cpu(), ref)
    s_x = torch.rand(M, generator=g) * 0.1 + 1e-3
    s_w = torch.rand(N, generator=g) * 0.01 + 1e-4
    bias = torch.randn(N, generator=g)
    for dt, name in ((torch.bfloat16, "bf16"), (torch.float32, "f32")):
        y = pq.qgemm(xq.cuda(), s_x.cuda(), wq.cuda(), s_w.cuda(), bias.cuda(), dt)
        want = O.cast_out(O.dequant_epilogue(ref.numpy(), s_x.numpy(), s_w.numpy(), bias.numpy()), name)
        assert torch.equal(_bits(y.cpu()), _bits(want))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "int_mm_*.npz"))))
@pytest.mark.parametrize("cfg", [-1, 0, 1])
def test_int32_committed_golden(path, cfg):
    d = np.load(path)
    pq.lib().pq_debug_set_gemm_config(cfg)
    got = pq.qgemm_i32(torch.from_numpy(d["a"]).cuda(), torch.from_numpy(d["b"]).cuda())
    assert np.array_equal(got.cpu().numpy(), d["acc"])


def test_int32_properties_at_scale():
    """Properties that need no CPU matmul: linearity in W and column-shard concatenation."""
    M, N, K = 8192, 8192, 8192
    a = rand_i8((M, K), 5).cuda()
    b1 = torch.randint(-64, 64, (N, K), dtype=torch.int8, generator=torch.Generator().manual_seed(6)).cuda()
    b2 = torch.randint(-64, 64, (N, K), dtype=torch.int8, generator=torch.Generator().manual_seed(7)).cuda()
    full = pq.qgemm_i32(a, b1 + b2)
    assert torch.equal(full, pq.qgemm_i32(a, b1) + pq.qgemm_i32(a, b2))
    parts = [pq.qgemm_i32(a, (b1 + b2)[lo:lo + 1024]) for lo in range(0, N, 1024)]
    assert torch.equal(full, torch.cat(parts, dim=1))
    # spot-check 64 random rows against the exact CPU product
    rows = torch.randperm(M, generator=torch.Generator().manual_seed(8))[:64]
    ref = cpu_int_mm(a[rows.cuda()].cpu(), (b1 + b2).cpu())
    assert torch.equal(full[rows.cuda()].cpu(), ref)


def test_persistent_scheduler_many_tiles():
    """More tiles than CTAs in both dimensions, ragged on every edge -> exercises ring/phase wrap."""
    M, N, K = 1100, 5000, 400
    a, b = rand_i8((M, K), 9), rand_i8((N, K), 10)
    for sk in (0, 1):
        pq.lib().pq_debug_set_streamk(sk)
        for cfg in (0, 1, 2, 3, 4, 8, 9, 10, 11, 12, 13, 16):
            pq.lib().pq_debug_set_gemm_config(cfg)
            assert torch.equal(pq.qgemm_i32(a.cuda(), b.cuda()).cpu(), cpu_int_mm(a, b)), (cfg, sk)


def test_streamk_repeated_launches_and_cuda_graph_replay():
    """Self-cleaning flags: the same workspace slot must give identical results launch after launch,
    on a second stream, and when the launches are replayed from a CUDA graph."""
    pq.lib().pq_debug_set_streamk(1)
    M, N, K = 2048, 4096, 4096
    a, b = rand_i8((M, K), 21).cuda(), rand_i8((N, K), 22).cuda()
    ref = cpu_int_mm(a.cpu(), b.cpu()).cuda()
    for _ in range(5):
        assert torch.equal(pq.qgemm_i32(a, b), ref)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            assert torch.equal(pq.qgemm_i32(a, b), ref)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s_x = torch.ones(M, device="cuda")
    s_w = torch.ones(N, device="cuda")
    y = torch.empty(M, N, dtype=torch.float32, device="cuda")
    with torch.cuda.graph(g):
        for _ in range(3):
            pq.qgemm(a, s_x, b, s_w, None, torch.float32, out=y)
    for _ in range(4):
        y.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(y, ref.float())


@pytest.mark.parametrize("sk", [-1, 1])
@pytest.mark.parametrize("cfg", [-1, 0, 1, 3, 8, 13, 16])
@pytest.mark.parametrize("out", [(torch.bfloat16, "bf16"), (torch.float16, "f16"), (torch.float32, "f32")])
@pytest.mark.parametrize("shape,use_bias", [((256, 512, 512), True), ((100, 264, 272), False),
                                            ((16, 4096, 4096), True), ((700, 1000, 768), True)])
def test_fused_epilogue_bit_exact_vs_oracle(cfg, out, shape, use_bias, sk):
    pq.lib().pq_debug_set_streamk(sk)
    """y = cast(((float(acc)*s_x)*s_w)+bias): same fp32 op order as the oracle -> identical bits."""
    M, N, K = shape
    dt, name = out
    pq.lib().pq_debug_set_gemm_config(cfg)
    g = torch.Generator().manual_seed(11)
    xq, wq = rand_i8((M, K), 12), rand_i8((N, K), 13)
    s_x = torch.rand(M, generator=g) * 0.1 + 1e-3
    s_w = torch.rand(N, generator=g) * 0.01 + 1e-4
    bias = torch.randn(N, generator=g) if use_bias else None
    y = pq.qgemm(xq.cuda(), s_x.cuda(), wq.cuda(), s_w.cuda(), bias.cuda() if use_bias else None, dt)
    ref = O.cast_out(O.dequant_epilogue(O.int_mm(xq.numpy(), wq.numpy()), s_x.numpy(), s_w.numpy(),
                                        bias.numpy() if use_bias else None), name)
    assert torch.equal(_bits(y.cpu()), _bits(ref))


@pytest.mark.parametrize("out_dtype,rtol", [(torch.bfloat16, 2 ** -8), (torch.float16, 2 ** -10), (torch.float32, 1e-6)])
@pytest.mark.parametrize("shape", [(16, 4096, 4096), (2048, 11008, 4096), (2048, 4096, 11008), (4096, 3072, 768)])
def test_qlinear_within_tolerance_of_fp32_epilogue(out_dtype, rtol, shape):
    """Stated tolerance: |y - y_fp32| <= rtol*|y_fp32| + 1e-3*max|y_fp32| where y_fp32 is the
    oracle's fp32 epilogue before the cast.  (In fact the outputs are bit-identical to the cast oracle.)"""
    M, N, K = shape
    g = torch.Generator().manual_seed(14)
    x = torch.randn(M, K, generator=g).to(torch.bfloat16)
    w = (torch.rand(N, K, generator=g) * 2 - 1) / K ** 0.5
    bias = torch.randn(N, generator=g)
    wq, sw = pq.quantize_weight(w.cuda())
    y = pq.qlinear(x.cuda(), wq, sw, bias.cuda(), out_dtype)
    wq_o, sw_o = O.quantize_weight(w)
    xq_o, sx_o = O.quantize_rowwise(x)
    y32 = torch.from_numpy(O.dequant_epilogue(O.int_mm(xq_o, wq_o), sx_o, sw_o, bias.numpy()))
    err = (y.cpu().float() - y32).abs()
    tol = rtol * y32.abs() + 1e-3 * y32.abs().max()
    assert bool((err <= tol).all())
    name = {torch.bfloat16: "bf16", torch.float16: "f16", torch.float32: "f32"}[out_dtype]
    assert torch.equal(_bits(y.cpu()), _bits(O.cast_out(y32.numpy(), name)))
    # and the whole thing is a faithful int8 approximation of the float linear
    ref = torch.nn.functional.linear(x.float(), w, bias)
    assert (y.cpu().float() - ref).abs().max() < 0.05 * ref.abs().max()


def test_unaligned_operands_are_repacked_by_python_and_rejected_by_c_abi():
    a, b = rand_i8((40, 100), 15), rand_i8((24, 100), 16)      # K=100: row stride not a multiple of 16
    got = pq.qgemm_i32(a.cuda(), b.cuda())
    assert torch.equal(got.cpu(), cpu_int_mm(a, b))
    lib = pq.lib()
    ac, bc = a.cuda().contiguous(), b.cuda().contiguous()
    acc = torch.empty(40, 24, dtype=torch.int32, device="cuda")
    rc = lib.pq_qgemm_i32(ac.data_ptr(), 100, bc.data_ptr(), 100, acc.data_ptr(), 24, 40, 24, 100, None)
    assert rc == 2 and b"16 bytes" in lib.pq_last_error()      # PQ_ERR_ALIGN


def test_empty_gemm():
    assert pq.qgemm_i32(torch.empty(0, 64, dtype=torch.int8, device="cuda"), rand_i8((8, 64), 1).cuda()).shape == (0, 8)


@pytest.mark.parametrize("out", [(torch.bfloat16, "bf16"), (torch.float16, "f16"), (torch.float32, "f32")])
@pytest.mark.parametrize("shape,use_bias", [((256, 512, 512), True), ((100, 264, 272), False), ((16, 4096, 4096), True),
                                            ((700, 1000, 768), True), ((2048, 3584, 8192), True), ((130, 250, 144), True)])
def test_staged_coalesced_epilogue_bit_exact(out, shape, use_bias):
    """The shared-memory staged epilogue used for NVLink peer / multicast destinations."""
    M, N, K = shape
    dt, name = out
    pq.lib().pq_debug_set_staged(1)
    g = torch.Generator().manual_seed(31)
    xq, wq = rand_i8((M, K), 32), rand_i8((N, K), 33)
    s_x = torch.rand(M, generator=g) * 0.1 + 1e-3
    s_w = torch.rand(N, generator=g) * 0.01 + 1e-4
    bias = torch.randn(N, generator=g) if use_bias else None
    y = pq.qgemm(xq.cuda(), s_x.cuda(), wq.cuda(), s_w.cuda(), bias.cuda() if use_bias else None, dt)
    ref = O.cast_out(O.dequant_epilogue(O.int_mm(xq.numpy(), wq.numpy()), s_x.numpy(), s_w.numpy(),
                                        bias.numpy() if use_bias else None), name)
    assert torch.equal(_bits(y.cpu()), _bits(ref))


def test_qgemm_multi_writes_every_destination_slice():
    """pq_qgemm_multi on one GPU: three full-width buffers each receive this shard's column slice."""
    from protoquant_b200 import functional as F
    M, N, K, N_total, off = 300, 512, 256, 1536, 512
    g = torch.Generator().manual_seed(41)
    xq, wq = rand_i8((M, K), 42).cuda(), rand_i8((N, K), 43).cuda()
    s_x = (torch.rand(M, generator=g) * 0.1 + 1e-3).cuda()
    s_w = (torch.rand(N, generator=g) * 0.01 + 1e-4).cuda()
    bias = torch.randn(N, generator=g).cuda()
    ref = pq.qgemm(xq, s_x, wq, s_w, bias, torch.bfloat16)
    bufs = [torch.zeros(M, N_total, dtype=torch.bfloat16, device="cuda") for _ in range(3)]
    F.qgemm_multi(xq, s_x, wq, s_w, bias, [b.data_ptr() + off * 2 for b in bufs], N_total, torch.bfloat16)
    for b in bufs:
        assert torch.equal(b[:, off:off + N], ref)
        assert not b[:, :off].any() and not b[:, off + N:].any()


@pytest.mark.parametrize("tma", [2, 1, 0])
@pytest.mark.parametrize("shape", [(300, 512, 256), (2048, 3584, 1024), (129, 200, 384), (70, 72, 128), (1000, 1000, 512),
                                   (16, 3584, 1024), (1, 264, 4096), (64, 520, 136)])
def test_qgemm_multi_tma_and_lsu_destinations_agree(shape, tma):
    """The multi-destination epilogue (fused all-gather) in its three forms -- 2: per-warp TMA boxes with the regular
    tile heuristic, 1: CTA-staged tile + TMA stores, 0: CTA-staged tile + LSU copy-out (default: fastest over NVLink);
    M <= 64 takes the weight-streaming kernel, which writes its rows to every destination itself -- every destination
    gets the bits of the single-destination GEMM;
    rows past M and columns outside the slice stay untouched (ragged M / N exercise the tensor-map clipping)."""
    from protoquant_b200 import functional as F
    M, N, K = shape
    N_total, off = N + 2 * 264, 264
    g = torch.Generator().manual_seed(51)
    xq, wq = rand_i8((M, K), 52).cuda(), rand_i8((N, K), 53).cuda()
    s_x = (torch.rand(M, generator=g) * 0.1 + 1e-3).cuda()
    s_w = (torch.rand(N, generator=g) * 0.01 + 1e-4).cuda()
    ref = pq.qgemm(xq, s_x, wq, s_w, None, torch.bfloat16)
    bufs = [torch.zeros(M + 3, N_total, dtype=torch.bfloat16, device="cuda") for _ in range(4)]
    pq.lib().pq_debug_set_multi_tma(tma)
    try:
        F.qgemm_multi(xq, s_x, wq, s_w, None, [b.data_ptr() + off * 2 for b in bufs], N_total, torch.bfloat16)
    finally:
        pq.lib().pq_debug_set_multi_tma(0)
    for b in bufs:
        assert torch.equal(b[:M, off:off + N], ref)
        assert not b[:, :off].any() and not b[:, off + N:].any() and not b[M:].any()


@pytest.mark.parametrize("bn", [256, 224, 128])
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape", [(300, 512, 256), (2048, 3584, 1024), (129, 200, 384), (1000, 1000, 512), (513, 1400, 640)])
def test_qgemm_multi_every_staged_tile_width(shape, out_dtype, bn):
    """The multi-destination (CTA-staged, LSU copy-out) epilogue on its three tile widths -- 256, 224 (128 + 96 column
    halves) and 128 (64 + 64): the launcher's model picks one per problem; here each is forced.  Ragged M / N, columns
    outside the slice and rows past M stay untouched."""
    from protoquant_b200 import functional as F
    M, N, K = shape
    N_total, off = N + 2 * 264, 264
    esz = torch.empty(0, dtype=out_dtype).element_size()
    g = torch.Generator().manual_seed(71)
    xq, wq = rand_i8((M, K), 72).cuda(), rand_i8((N, K), 73).cuda()
    s_x = (torch.rand(M, generator=g) * 0.1 + 1e-3).cuda()
    s_w = (torch.rand(N, generator=g) * 0.01 + 1e-4).cuda()
    bias = torch.randn(N, generator=g).cuda()
    ref = pq.qgemm(xq, s_x, wq, s_w, bias, out_dtype)
    bufs = [torch.zeros(M + 3, N_total, dtype=out_dtype, device="cuda") for _ in range(3)]
    pq.lib().pq_debug_set_multi_bn(bn)
    try:
        F.qgemm_multi(xq, s_x, wq, s_w, bias, [b.data_ptr() + off * esz for b in bufs], N_total, out_dtype)
    finally:
        pq.lib().pq_debug_set_multi_bn(0)
    for b in bufs:
        assert torch.equal(b[:M, off:off + N], ref)
        assert not b[:, :off].any() and not b[:, off + N:].any() and not b[M:].any()


def test_qlinear_multi_one_call_equals_qlinear():
    """pq_qlinear_multi (act-quant + multi-destination GEMM behind one C call) == pq_qlinear, bit for bit."""
    from protoquant_b200 import functional as F
    M, N, K = 257, 1032, 520
    torch.manual_seed(61)
    lin = torch.nn.Linear(K, N).to(torch.bfloat16).cuda()
    m = pq.DynamicQuantLinear.from_float(lin)
    x = torch.randn(M, K, dtype=torch.bfloat16, device="cuda")
    ref = m(x)
    bufs = [torch.zeros(M, N, dtype=torch.bfloat16, device="cuda") for _ in range(2)]
    xq_ws = torch.empty(M, m.qweight_storage.shape[1], dtype=torch.int8, device="cuda")
    sx_ws = torch.empty(M, dtype=torch.float32, device="cuda")
    before = pq.launch_count()
    F.qlinear_multi_into(x, m.qweight_storage, K, m.weight_scale, m.bias, [b.data_ptr() for b in bufs], N, torch.bfloat16,
                         xq_ws, sx_ws)
    assert pq.launch_count() - before == 2
    for b in bufs:
        assert torch.equal(b, ref)


@pytest.mark.parametrize("shape", [(1, 4096, 4096), (2, 4096, 4096), (7, 11008, 4096), (16, 4096, 11008), (17, 768, 3072),
                                   (31, 3072, 768), (33, 1000, 144), (64, 4096, 4096), (64, 8192, 8192), (50, 264, 272),
                                   (48, 28672, 8192), (16, 8192, 28672), (5, 8, 16), (16, 136, 4096)])
def test_small_m_kernel_int32_and_epilogue(shape):
    """Decode path (swap-AB, K split across a cluster, DSMEM reduction): exact accumulators and
    bit-exact fused epilogue for every M_pad bucket, ragged N/K, 1..8 K-splits."""
    M, N, K = shape
    pq.lib().pq_debug_set_gemm_config(7)
    g = torch.Generator().manual_seed(51)
    xq, wq = rand_i8((M, K), 52), rand_i8((N, K), 53)
    acc = pq.qgemm_i32(xq.cuda(), wq.cuda())
    ref = cpu_int_mm(xq, wq)
    assert torch.equal(acc.cpu(), ref)
    s_x = torch.rand(M, generator=g) * 0.1 + 1e-3
    s_w = torch.rand(N, generator=g) * 0.01 + 1e-4
    bias = torch.randn(N, generator=g)
    for dt, name in ((torch.bfloat16, "bf16"), (torch.float16, "f16"), (torch.float32, "f32")):
        y = pq.qgemm(xq.cuda(), s_x.cuda(), wq.cuda(), s_w.cuda(), bias.cuda(), dt)
        want = O.cast_out(O.dequant_epilogue(ref.numpy(), s_x.numpy(), s_w.numpy(), bias.numpy()), name)
        assert torch.equal(_bits(y.cpu()), _bits(want))
    y = pq.qgemm(xq.cuda(), s_x.cuda(), wq.cuda(), s_w.cuda(), None, torch.bfloat16)
    want = O.cast_out(O.dequant_epilogue(ref.numpy(), s_x.numpy(), s_w.numpy(), None), "bf16")
    assert torch.equal(_bits(y.cpu()), _bits(want))
