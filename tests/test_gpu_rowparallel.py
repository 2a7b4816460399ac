"""GPU parity tests of the row-parallel (K-split) pieces: row maxima, quantisation with an
external maximum, the int32 scatter GEMM (fused GEMM + reduce-scatter, first half) and the reduce + dequant
kernel (second half).  Everything here is integer / index work or the fixed fp32 epilogue: bit-exact."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
import protoquant_b200 as pq
from protoquant_b200 import functional as F
import protoquant_oracle as O

pytestmark = pytest.mark.gpu


def rand_i8(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(-128, 128, shape, dtype=torch.int8, generator=g)


def _bits(t):
    return t.view(torch.int32 if t.dtype == torch.float32 else torch.int16)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("shape", [(1, 16), (17, 4096), (64, 11008), (5, 100), (2048, 3584), (3, 28672)])
def test_row_absmax_and_quantize_with_external_max(shape, dt):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(*shape, generator=g).to(dt)
    amax = F.row_absmax(x.cuda())
    want = np.max(np.abs(O.to_f32(x)), axis=-1).astype(np.float32)
    assert np.array_equal(amax.cpu().numpy(), want)
    big = (amax * 1.7 + 0.25).contiguous()               # as if another K-shard held the row maximum
    q, s = F.quantize_act_with_amax(x.cuda(), big)
    q_o, s_o = O.quantize_rowwise(x, amax=big.cpu().numpy())
    assert np.array_equal(q.cpu().numpy(), q_o) and np.array_equal(s.cpu().numpy(), s_o)
    q2, s2 = F.quantize_act_with_amax(x.cuda(), amax)    # own maximum == the plain quantizer
    q3, s3 = pq.quantize_act(x.cuda())
    assert torch.equal(q2, q3) and torch.equal(s2, s3)


@pytest.mark.parametrize("shape,ndest", [((300, 512, 256), 3), ((100, 1000, 272), 4), ((2048, 4096, 1024), 8),
                                         ((16, 1024, 512), 2), ((129, 264, 144), 2)])
@pytest.mark.parametrize("bn", [0, 256, 224, 128])
def test_scatter_gemm_writes_each_column_block_to_its_destination(shape, ndest, bn):
    """bn: tile width of the staged epilogue (0 = the launcher's model)."""
    M, N, K = shape
    a, b = rand_i8((M, K), 2).cuda(), rand_i8((N, K), 3).cuda()
    per = (-(-N // ndest) + 7) // 8 * 8
    ref = pq.qgemm_i32(a, b)
    inbox = torch.full((ndest, M, per), 7, dtype=torch.int32, device="cuda")
    pq.lib().pq_debug_set_multi_bn(bn)
    try:
        F.qgemm_i32_scatter(a, b, [inbox[d].data_ptr() for d in range(ndest)], per, per)
    finally:
        pq.lib().pq_debug_set_multi_bn(0)
    for d in range(ndest):
        lo, hi = min(d * per, N), min((d + 1) * per, N)
        assert torch.equal(inbox[d][:, : hi - lo], ref[:, lo:hi])
        assert bool((inbox[d][:, hi - lo:] == 7).all())          # nothing outside the block was touched


@pytest.mark.parametrize("out", [(torch.bfloat16, "bf16"), (torch.float16, "f16"), (torch.float32, "f32")])
@pytest.mark.parametrize("shape,nparts,ndest,use_bias", [((64, 512), 1, 1, True), ((300, 1000), 4, 2, True),
                                                          ((2048, 1024), 8, 8, False), ((5, 17), 3, 1, True)])
def test_reduce_dequant(shape, nparts, ndest, use_bias, out):
    M, N = shape
    dt, name = out
    g = torch.Generator().manual_seed(4)
    parts = torch.randint(-2 ** 20, 2 ** 20, (nparts, M, N), dtype=torch.int32, generator=g).cuda()
    s_x = (torch.rand(M, generator=g) * 0.1 + 1e-3).cuda()
    s_w = (torch.rand(N, generator=g) * 0.01 + 1e-4).cuda()
    bias = torch.randn(N, generator=g).cuda() if use_bias else None
    ys = [torch.zeros(M, N + 8, dtype=dt, device="cuda") for _ in range(ndest)]
    F.reduce_dequant([parts[p].data_ptr() for p in range(nparts)], N, s_x, s_w, bias, [y.data_ptr() for y in ys],
                     N + 8, dt, M, N)
    acc = parts.sum(0, dtype=torch.int32).cpu().numpy()
    want = O.cast_out(O.dequant_epilogue(acc, s_x.cpu().numpy(), s_w.cpu().numpy(),
                                         bias.cpu().numpy() if use_bias else None), name)
    for y in ys:
        assert torch.equal(_bits(y[:, :N].cpu()), _bits(want))
        assert not y[:, N:].any()


@pytest.mark.parametrize("shape,shards", [((256, 1024, 3072), 3), ((2048, 8192, 3584), 2), ((40, 520, 1040), 4)])
def test_k_split_on_one_gpu_equals_unsplit_linear(shape, shards):
    """The whole row-parallel dataflow emulated on one GPU: per-shard quantisation with the global row maximum,
    scatter GEMMs into per-owner inboxes, reduce + dequant -- same bits as the single dynamic-quant linear."""
    M, N, K = shape
    torch.manual_seed(5)
    lin = torch.nn.Linear(K, N).to(torch.bfloat16).cuda()
    m = pq.DynamicQuantLinear.from_float(lin)
    x = torch.randn(M, K, dtype=torch.bfloat16, device="cuda")
    x[0, K - 1] = 30.0
    full = m(x)
    amax = F.row_absmax(x)
    per_n = (-(-N // shards) + 7) // 8 * 8
    inbox = torch.zeros(shards, shards, M, per_n, dtype=torch.int32, device="cuda")     # [owner][source]
    s_x = None
    for r in range(shards):
        k_lo, k_hi = pq.shard_bounds(K, shards, r, align=16)
        xq, s_x = F.quantize_act_with_amax(x[:, k_lo:k_hi], amax)
        F.qgemm_i32_scatter(xq, m.qweight[:, k_lo:k_hi], [inbox[o, r].data_ptr() for o in range(shards)], per_n, per_n)
    y = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    for o in range(shards):
        lo, hi = min(o * per_n, N), min((o + 1) * per_n, N)
        if hi > lo:
            F.reduce_dequant([inbox[o, r].data_ptr() for r in range(shards)], per_n, s_x, m.weight_scale[lo:hi],
                             m.bias[lo:hi], [y.data_ptr() + lo * 2], N, torch.bfloat16, M, hi - lo)
    assert torch.equal(y, full)


def test_row_parallel_module_single_rank_equals_unsharded():
    torch.manual_seed(6)
    lin = torch.nn.Linear(3072, 768).to(torch.bfloat16).cuda()
    m = pq.DynamicQuantLinear.from_float(lin)
    rp = pq.RowParallelDynamicQuantLinear(m.qweight, m.weight_scale, m.bias)
    x = torch.randn(77, 3072, dtype=torch.bfloat16, device="cuda")
    assert torch.equal(rp(x), m(x))


def test_parallel_gated_mlp_single_rank_equals_chained_modules():
    torch.manual_seed(7)
    H, I, M = 1024, 2816, 100
    gate = pq.DynamicQuantLinear.from_float(torch.nn.Linear(H, I, bias=False).to(torch.bfloat16).cuda())
    up = pq.DynamicQuantLinear.from_float(torch.nn.Linear(H, I, bias=False).to(torch.bfloat16).cuda())
    down = pq.DynamicQuantLinear.from_float(torch.nn.Linear(I, H, bias=True).to(torch.bfloat16).cuda())
    x = torch.randn(M, H, dtype=torch.bfloat16, device="cuda")
    want = down(F.act_mul(gate(x), up(x), "silu"))
    mlp = pq.ParallelGatedMLP(gate, up, down)
    assert torch.equal(mlp(x), want)
    assert torch.equal(mlp(x.reshape(4, 25, H)).reshape(M, H), want)
    # the concatenated gate/up copy follows the shards: new weights loaded into the module are used by the next forward
    gate2 = pq.DynamicQuantLinear.from_float(torch.nn.Linear(H, I, bias=False).to(torch.bfloat16).cuda())
    mlp.gate.load_state_dict(pq.ShardedDynamicQuantLinear(gate2.qweight, gate2.weight_scale, None, gather_output=False, align=16).state_dict())
    assert torch.equal(mlp(x), down(F.act_mul(gate2(x), up(x), "silu")))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_row_parallel_module_nvlink_bit_identical():
    n = min(torch.cuda.device_count(), 8)
    n = 1 << (n.bit_length() - 1)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29519", os.path.join(ROOT, "tests", "_rowparallel_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert "ROWPARALLEL_OK" in p.stdout


@pytest.mark.parametrize("rot", [0, 512, 1024, 1536, 4000])
@pytest.mark.parametrize("shape", [(2048, 4096, 1024), (300, 1000, 272), (700, 2048, 512)])
def test_rotated_tile_order_changes_nothing_but_the_order(shape, rot):
    """The reduce-scatter GEMM of rank r visits the column blocks starting at ITS block (no ingress hot-spot on the
    owners); any rotation must give the same bits -- scatter GEMM, plain GEMM with the fused epilogue, int32 output."""
    M, N, K = shape
    a, b = rand_i8((M, K), 12).cuda(), rand_i8((N, K), 13).cuda()
    g = torch.Generator().manual_seed(14)
    s_x = (torch.rand(M, generator=g) * 0.1 + 1e-3).cuda()
    s_w = (torch.rand(N, generator=g) * 0.01 + 1e-4).cuda()
    ref32 = pq.qgemm_i32(a, b)
    ref16 = pq.qgemm(a, s_x, b, s_w, None, torch.bfloat16)
    ndest = 4
    per = (-(-N // ndest) + 7) // 8 * 8
    pq.lib().pq_debug_set_tile_rotation(rot)
    try:
        inbox = torch.full((ndest, M, per), 7, dtype=torch.int32, device="cuda")
        F.qgemm_i32_scatter(a, b, [inbox[d].data_ptr() for d in range(ndest)], per, per)
        got32 = pq.qgemm_i32(a, b)
        got16 = pq.qgemm(a, s_x, b, s_w, None, torch.bfloat16)
    finally:
        pq.lib().pq_debug_set_tile_rotation(0)
    assert torch.equal(got32, ref32) and torch.equal(got16, ref16)
    for d in range(ndest):
        lo, hi = min(d * per, N), min((d + 1) * per, N)
        assert torch.equal(inbox[d][:, : hi - lo], ref32[:, lo:hi])
        assert bool((inbox[d][:, hi - lo:] == 7).all())


def test_symm_entry_points_validate_their_arguments():
    """pq_symm_barrier / pq_rowparallel_forward refuse malformed groups instead of launching (world = 1 is a no-op
    barrier; the one-call forward needs 2..8 ranks -- the multi-rank behaviour is covered by _rowparallel_worker.py
    and by bench.py --gpus N, which fails the run on a mismatch)."""
    import ctypes
    from protoquant_b200 import _lib
    lib = pq.lib()
    pad = torch.zeros(64, dtype=torch.int32, device="cuda")
    sg = _lib.PQSymmGroup()
    sg.rank, sg.world, sg.cap = 0, 1, 16
    sg.pads[0] = pad.data_ptr()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    before = pq.launch_count()
    assert lib.pq_symm_barrier(ctypes.byref(sg), 0, st) == 0 and pq.launch_count() == before     # nothing to wait for
    assert lib.pq_symm_barrier(ctypes.byref(sg), 9, st) == 1                                      # PQ_ERR_ARG: channel
    sg.world = 3                                                                                   # pads of ranks 1, 2 missing
    assert lib.pq_symm_barrier(ctypes.byref(sg), 0, st) == 1 and b"null signal pad" in lib.pq_last_error()
    sg.world = 1
    x = torch.zeros(16, 64, dtype=torch.bfloat16, device="cuda")
    rc = lib.pq_rowparallel_forward(x.data_ptr(), None, 2, 0, 64, 0, 1, 64, 0, x.data_ptr(), 64, x.data_ptr(), None,
                                    ctypes.byref(sg), 0, x.data_ptr(), 2, 64, x.data_ptr(), x.data_ptr(), None, 16, 64, 64, 64,
                                    None, st)
    assert rc == 1 and b"2..8 ranks" in lib.pq_last_error()
    torch.cuda.synchronize()
