"""CPU tests of the host-side logic: shard arithmetic, padded int8 allocation, module
construction/state_dict, compat alias table, and the world_size-2 gloo run of the
column-parallel path (oracle standing in for the kernel, as SURVEY.md §4 prescribes)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
import protoquant_b200 as pq
from protoquant_b200 import compat, functional as F
import protoquant_oracle as O


def test_shard_bounds_cover_and_align():
    for n in (8192, 28672, 11008, 4096, 1000, 7):
        for world in (1, 2, 4, 8):
            prev = 0
            for r in range(world):
                lo, hi = pq.shard_bounds(n, world, r)
                assert lo == prev and lo <= hi <= n
                if hi < n:
                    assert (hi - lo) % 8 == 0
                prev = hi
            assert prev == n


def test_alloc_q_pads_row_stride_to_16_bytes():
    for k in (1, 15, 16, 100, 4096, 11008):
        t = F.alloc_q(3, k, "cpu")
        assert t.shape == (3, k) and t.stride(0) % 16 == 0 and t.stride(1) == 1


def test_module_construction_and_state_dict_roundtrip():
    m = pq.DynamicQuantLinear(100, 24, bias=True)
    assert m.qweight.shape == (24, 100) and m.qweight.stride(0) == 112
    sd = m.state_dict()
    assert set(sd) == {"qweight_storage", "weight_scale", "bias"}
    m2 = pq.DynamicQuantLinear(100, 24, bias=True)
    m.weight_scale.fill_(0.5)
    m2.load_state_dict(m.state_dict())
    assert torch.equal(m2.weight_scale, m.weight_scale)
    assert "int8" in repr(m)


def test_compat_alias_table():
    assert compat.QLinear is pq.DynamicQuantLinear
    assert compat.quantize_per_token is pq.quantize_act
    for name in compat.__all__:
        assert hasattr(compat, name)


def test_fuse_linears_concatenates_per_channel_parameters():
    a, b = pq.DynamicQuantLinear(64, 32), pq.DynamicQuantLinear(64, 48)
    a.qweight_storage.random_(-127, 128); b.qweight_storage.random_(-127, 128)
    a.weight_scale.uniform_(0.1, 1.0); b.weight_scale.uniform_(0.1, 1.0)
    a.bias.normal_(); b.bias.normal_()
    f = pq.fuse_linears([a, b])
    assert (f.in_features, f.out_features) == (64, 80)
    assert torch.equal(f.qweight[:32], a.qweight) and torch.equal(f.qweight[32:], b.qweight)
    assert torch.equal(f.weight_scale, torch.cat([a.weight_scale, b.weight_scale]))
    assert torch.equal(f.bias, torch.cat([a.bias, b.bias]))
    with pytest.raises(ValueError):
        pq.fuse_linears([a, pq.DynamicQuantLinear(32, 8)])


def test_parallel_mlp_shards_line_up_and_has_no_cpu_path():
    gate, up, down = pq.DynamicQuantLinear(64, 160, bias=False), pq.DynamicQuantLinear(64, 160, bias=False), pq.DynamicQuantLinear(160, 64)
    mlp = pq.ParallelGatedMLP(gate, up, down)
    assert (mlp.gate.lo, mlp.gate.hi) == (mlp.down.k_lo, mlp.down.k_hi) == (0, 160)
    assert mlp.gate.gather_output is False and mlp.down.input_is_sharded is True
    with pytest.raises(pq.ProtoquantError, match="no CPU fallback"):
        mlp(torch.randn(4, 64))
    with pytest.raises(ValueError):
        pq.ParallelGatedMLP(gate, up, pq.DynamicQuantLinear(128, 64))
    for world in (2, 4, 8):          # column shards of gate/up == K shards of down for every rank
        for n in (11008, 28672, 16000, 1000):
            for r in range(world):
                assert pq.shard_bounds(n, world, r, align=16) == pq.shard_bounds(n, world, r, 16)
                lo, hi = pq.shard_bounds(n, world, r, align=16)
                assert lo % 16 == 0 and 0 <= lo <= hi <= n


def test_qtensor_metadata():
    qt = pq.QTensor(torch.zeros(6, 8, dtype=torch.int8), torch.ones(6), orig_dtype=torch.bfloat16, orig_shape=(2, 3, 8))
    assert qt.shape == (2, 3, 8) and qt.axis == -1
    assert "QTensor" in repr(qt)
    with pytest.raises(NotImplementedError):
        pq.quantize(torch.zeros(4, 4, 4), axis=0)


# ---- serialisation: metadata and layout handling, CPU only ------------------------------
def test_qtensor_state_roundtrip_and_versioning(tmp_path):
    data = F.alloc_q(6, 20, "cpu")                      # padded rows (stride 32)
    data.copy_(torch.randint(-127, 128, (6, 20), dtype=torch.int8))
    qt = pq.QTensor(data, torch.rand(6), orig_dtype=torch.bfloat16, orig_shape=(2, 3, 20))
    st = qt.state()
    assert st["format"] == "protoquant_b200.QTensor" and st["format_version"] == 1
    assert st["data"].is_contiguous() and st["data"].shape == (6, 20) and st["orig_dtype"] == "bfloat16"
    path = tmp_path / "qt.pt"
    qt.save(path)
    back = pq.QTensor.load(path)                        # torch.load(weights_only=True): tensors and plain metadata only
    assert torch.equal(back.data, qt.data) and torch.equal(back.scale, qt.scale)
    assert back.data.stride(0) % 16 == 0 and back.orig_dtype == torch.bfloat16 and back.shape == (2, 3, 20) and back.axis == -1
    # a round-1 state (torch.dtype object, no version field) still loads; a newer or foreign format is refused
    old = {"data": st["data"], "scale": st["scale"], "axis": -1, "orig_dtype": torch.float16, "orig_shape": (6, 20)}
    assert pq.QTensor.from_state(old).orig_dtype == torch.float16
    with pytest.raises(ValueError, match="format_version"):
        pq.QTensor.from_state({**st, "format_version": 99})
    with pytest.raises(ValueError, match="not a"):
        pq.QTensor.from_state({**st, "format": "something.else"})
    with pytest.raises(TypeError):
        pq.QTensor.from_state({**st, "scale": st["scale"][:3]})
    # axis = 0: one scale per column
    qc = pq.QTensor(torch.zeros(4, 8, dtype=torch.int8), torch.ones(8), axis=0)
    assert pq.QTensor.from_state(qc.state()).scale.numel() == 8


def test_module_checkpoint_versions_layouts_and_qtensor_bridge(tmp_path):
    m = pq.DynamicQuantLinear(100, 24, bias=True)
    m.qweight_storage[:, :100].copy_(torch.randint(-127, 128, (24, 100), dtype=torch.int8))
    m.weight_scale.uniform_(0.1, 1.0)
    m.bias.normal_()
    path = tmp_path / "lin.pt"
    torch.save(m.state_dict(), path)
    sd = torch.load(path, weights_only=True)
    assert sd._metadata[""]["version"] == 1
    m2 = pq.DynamicQuantLinear(100, 24, bias=True)
    m2.load_state_dict(sd)
    assert torch.equal(m2.qweight_storage, m.qweight_storage) and torch.equal(m2.weight_scale, m.weight_scale)
    # unpadded payload under `qweight` (what a QTensor stores), scales saved in half precision: re-padded / widened
    sd_unpadded = {"qweight": m.qweight.contiguous(), "weight_scale": m.weight_scale.double(), "bias": m.bias}
    m3 = pq.DynamicQuantLinear(100, 24, bias=True)
    m3.load_state_dict(sd_unpadded)
    assert torch.equal(m3.qweight, m.qweight) and m3.qweight_storage.stride(0) == 112
    assert m3.weight_scale.dtype == torch.float32 and torch.equal(m3.weight_scale, m.weight_scale)
    # wrong shape and a newer format version are load errors, not silent garbage
    with pytest.raises(RuntimeError, match="qweight"):
        pq.DynamicQuantLinear(100, 24).load_state_dict({**sd_unpadded, "qweight": m.qweight[:10].contiguous()})
    newer = m.state_dict()
    newer._metadata[""]["version"] = 7
    with pytest.raises(RuntimeError, match="format 7"):
        pq.DynamicQuantLinear(100, 24).load_state_dict(newer)
    # QTensor bridge: weight -> QTensor file -> module
    m.weight_qtensor().save(tmp_path / "w.pt")
    m4 = pq.DynamicQuantLinear.from_qtensor(pq.QTensor.load(tmp_path / "w.pt"), bias=m.bias)
    assert torch.equal(m4.qweight, m.qweight) and torch.equal(m4.weight_scale, m.weight_scale) and torch.equal(m4.bias, m.bias)


def test_full_checkpoint_into_sharded_modules_single_rank():
    m = pq.DynamicQuantLinear(96, 40, bias=True)
    m.qweight_storage.random_(-127, 128)
    m.weight_scale.uniform_(0.1, 1.0)
    m.bias.normal_()
    model_sd = {"mlp.up." + k: v for k, v in m.state_dict().items()}
    col = pq.ShardedDynamicQuantLinear.from_full_state_dict(model_sd, "mlp.up.", in_features=96, device="cpu")
    assert torch.equal(col.qweight, m.qweight) and torch.equal(col.weight_scale, m.weight_scale) and torch.equal(col.bias, m.bias)
    row = pq.RowParallelDynamicQuantLinear.from_full_state_dict(model_sd, "mlp.up.", in_features=96, device="cpu")
    assert torch.equal(row.qweight, m.qweight) and (row.k_lo, row.k_hi) == (0, 96)
    with pytest.raises(KeyError):
        pq.ShardedDynamicQuantLinear.from_full_state_dict(model_sd, "mlp.down.", device="cpu")


# ---- world_size 2 over gloo ------------------------------------------------------------
def _oracle_local(x2, wq, s_w, bias, out_dtype):
    name = {torch.bfloat16: "bf16", torch.float16: "f16", torch.float32: "f32"}[out_dtype]
    return O.qlinear(x2, wq.numpy(), s_w.numpy(), None if bias is None else bias.numpy(), out_dtype=name)


def _worker(rank, world, port, N, K, M, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        w = (torch.rand(N, K) * 2 - 1) / K ** 0.5
        b = torch.randn(N)
        x = torch.randn(M, K).to(torch.bfloat16)
        wq, sw = O.quantize_rowwise(w)
        full = O.qlinear(x, wq, sw, b.numpy(), out_dtype="bf16")
        lin = pq.ShardedDynamicQuantLinear(torch.from_numpy(wq), torch.from_numpy(sw), b, local_forward=_oracle_local)
        y = lin(x)
        lo, hi = pq.shard_bounds(N, world, rank)
        ok = torch.equal(y, full) and (lin.lo, lin.hi) == (lo, hi) and y.shape == (M, N)
        y3 = lin(x.reshape(2, M // 2, K))
        ok = ok and torch.equal(y3.reshape(M, N), full)
        # every rank loads the same FULL checkpoint and keeps its slice
        ref = pq.DynamicQuantLinear(K, N, bias=True)
        ref.qweight_storage[:, :K].copy_(torch.from_numpy(wq)); ref.weight_scale.copy_(torch.from_numpy(sw)); ref.bias.copy_(b)
        ck = {"layer." + k: v for k, v in ref.state_dict().items()}
        lin2 = pq.ShardedDynamicQuantLinear.from_full_state_dict(ck, "layer.", in_features=K, device="cpu", local_forward=_oracle_local)
        ok = ok and torch.equal(lin2(x), full) and torch.equal(lin2.qweight_storage, lin.qweight_storage)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("N", [64, 100])   # 100: last shard is ragged (56 + 44)
def test_column_parallel_gloo_world2_bit_identical(N):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, N, 96, 6, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


class _OracleShardOps:
    """CPU stand-ins for the device ops of RowParallelDynamicQuantLinear (oracle arithmetic)."""
    @staticmethod
    def row_absmax(x2):
        return torch.from_numpy(np.max(np.abs(O.to_f32(x2)), axis=-1).astype(np.float32))

    @staticmethod
    def quantize_with_amax(xs, amax, spec=None):
        q, s = O.quantize_rowwise(xs, amax=amax.numpy())
        return torch.from_numpy(q), torch.from_numpy(s)

    @staticmethod
    def int_mm(xq, wq):
        return torch.from_numpy(O.int_mm(xq.numpy(), wq.numpy()))

    @staticmethod
    def act_mul(g, u, act):
        # T(T(act(gate)) * up): the definition of protoquant_b200.act_mul, in fp64 then rounded twice like the kernel
        a = torch.from_numpy(O.act_ref(O.to_f32(g).astype(np.float64), act)).to(g.dtype)
        return (a.double() * u.double()).to(g.dtype)

    @staticmethod
    def epilogue(acc, s_x, s_w, bias, out_dtype):
        name = {torch.bfloat16: "bf16", torch.float16: "f16", torch.float32: "f32"}[out_dtype]
        return O.cast_out(O.dequant_epilogue(acc.numpy(), s_x.numpy(), s_w.numpy(), None if bias is None else bias.numpy()), name)


def _row_worker(rank, world, port, N, K, M, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        w = (torch.rand(N, K) * 2 - 1) / K ** 0.5
        b = torch.randn(N)
        x = torch.randn(M, K).to(torch.bfloat16)
        x[0, K - 1] = 40.0            # the row maximum lives in the LAST shard: local maxima would differ
        wq, sw = O.quantize_rowwise(w)
        full = O.qlinear(x, wq, sw, b.numpy(), out_dtype="bf16")
        ok = True
        for sharded_in in (False, True):
            for gather in (True, False):
                lin = pq.RowParallelDynamicQuantLinear(torch.from_numpy(wq), torch.from_numpy(sw), b, ops=_OracleShardOps,
                                                       input_is_sharded=sharded_in, gather_output=gather)
                xin = x[:, lin.k_lo:lin.k_hi].contiguous() if sharded_in else x
                y = lin(xin)
                want = full if gather else full[:, lin.n_lo:lin.n_hi]
                ok = ok and torch.equal(y, want)
        # gated input (the MLP's down projection): y = down(act(gate) * up) with K-sharded gate / up slices
        gate = torch.randn(M, K).to(torch.bfloat16)
        up = torch.randn(M, K).to(torch.bfloat16)
        h = _OracleShardOps.act_mul(gate, up, "silu")
        want = O.qlinear(h, wq, sw, b.numpy(), out_dtype="bf16")
        lin = pq.RowParallelDynamicQuantLinear(torch.from_numpy(wq), torch.from_numpy(sw), b, ops=_OracleShardOps, input_is_sharded=True)
        y = lin(gate[:, lin.k_lo:lin.k_hi].contiguous(), up[:, lin.k_lo:lin.k_hi].contiguous(), "silu")
        ok = ok and torch.equal(y, want)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N,K", [(64, 96), (100, 200)])
def test_row_parallel_gloo_world2_bit_identical(N, K):
    """K-split shards + int32 all-reduce: same bits as the unsharded oracle, whether the input arrives replicated
    or already K-sharded (max all-reduce), gathered or scattered output."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_row_worker, args=(r, 2, port, N, K, 6, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_maybe_shard_keeps_small_layers_replicated():
    m = pq.DynamicQuantLinear(64, 32)
    assert pq.maybe_shard(m) is m


def test_shared_input_group_cache_logic_with_a_stub_kernel():
    """The bookkeeping of swap_linear's shared-input fusion, on CPU tensors with a stub in place of the fused module:
    one fused forward per distinct activation, slices by member, release after the last member, recompute on a new
    tensor, on an in-place update (version counter) and when one member is called twice."""
    from protoquant_b200.modules import _SharedInputGroup

    class Stub:
        calls = 0

        def __call__(self, x):
            Stub.calls += 1
            return torch.cat([x[..., :1] * 1.0, x[..., :2] * 2.0, x[..., :3] * 3.0], dim=-1)

    grp = _SharedInputGroup(Stub(), [1, 2, 3])
    x = torch.arange(12.0).reshape(3, 4)
    a, b, c = grp.output_for(x, 0), grp.output_for(x, 1), grp.output_for(x, 2)
    assert Stub.calls == 1 and a.shape == (3, 1) and b.shape == (3, 2) and c.shape == (3, 3)
    assert torch.equal(b, x[:, :2] * 2.0) and torch.equal(c, x[:, :3] * 3.0)
    assert grp._x is None and grp._y is None                     # released after the last member
    grp.output_for(x, 0)
    assert Stub.calls == 2                                        # a new round recomputes
    grp.output_for(x, 0)
    assert Stub.calls == 3                                        # the same member twice: not served from the cache
    grp.output_for(x, 1), grp.output_for(x, 2)
    assert Stub.calls == 3
    y = torch.ones(3, 4)
    grp.output_for(y, 0), grp.output_for(y, 1), grp.output_for(y, 2)
    assert Stub.calls == 4


def test_swap_linear_structure_is_checked_on_cpu_only_for_patterns():
    """Fusion candidates are found by name and shape; nothing is converted on a CPU-only box (from_float needs CUDA)."""
    from protoquant_b200.modules import SHARED_INPUT_PATTERNS
    assert ("q_proj", "k_proj", "v_proj") in SHARED_INPUT_PATTERNS and ("gate_proj", "up_proj") in SHARED_INPUT_PATTERNS
    blk = torch.nn.Module()
    blk.q_proj, blk.k_proj, blk.v_proj = (torch.nn.Linear(8, 8) for _ in range(3))
    with pytest.raises(RuntimeError, match="CUDA"):
        pq.swap_linear(blk)


def test_dtype_casts_never_touch_the_fp32_abi_buffers():
    m = pq.DynamicQuantLinear(64, 32)
    m.weight_scale.fill_(1.2345678e-6)          # not representable in fp16 / bf16: a cast round trip would change it
    net = torch.nn.Sequential(m, torch.nn.LayerNorm(32))
    net.half()
    assert torch.all(m.weight_scale == torch.tensor(1.2345678e-6))
    assert m.weight_scale.dtype == torch.float32 and m.bias.dtype == torch.float32 and m.qweight_storage.dtype == torch.int8
    assert net[1].weight.dtype == torch.float16
    net.to(torch.bfloat16)
    assert m.weight_scale.dtype == torch.float32
    net.double()
    assert m.bias.dtype == torch.float32


def test_token_adaptive_linear_dispatches_by_token_count():
    calls = []

    class Rec(torch.nn.Module):
        def __init__(self, tag):
            super().__init__()
            self.tag, self.in_features, self.out_features = tag, 8, 4

        def forward(self, x):
            calls.append(self.tag)
            return x[..., :4]

    m = pq.TokenAdaptiveLinear(Rec("replicated"), Rec("sharded"), min_tokens=128)
    m(torch.zeros(16, 8)); m(torch.zeros(4, 32, 8)); m(torch.zeros(2, 8, 8))
    assert calls == ["replicated", "sharded", "replicated"]


def test_parallelize_gated_mlps_finds_the_hf_layout_and_maps_the_activation():
    from torch import nn

    class MLP(nn.Module):
        def __init__(self, act):
            super().__init__()
            self.gate_proj, self.up_proj = pq.DynamicQuantLinear(64, 160, bias=False), pq.DynamicQuantLinear(64, 160, bias=False)
            self.down_proj = pq.DynamicQuantLinear(160, 64, bias=False)
            self.act_fn = act

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.mlp, self.other = MLP(nn.SiLU()), nn.Linear(8, 8)

    model = nn.ModuleList([Block(), Block()])
    assert pq.parallelize_gated_mlps(model) == 2
    assert all(isinstance(b.mlp, pq.ParallelGatedMLP) and b.mlp.act == "silu" and isinstance(b.other, nn.Linear) for b in model)
    tanh = nn.ModuleDict({"m": MLP(nn.GELU(approximate="tanh"))})
    assert pq.parallelize_gated_mlps(tanh) == 1 and tanh["m"].act == "gelu_tanh"
    with pytest.raises(ValueError, match="unsupported activation"):
        pq.parallelize_gated_mlps(nn.ModuleDict({"m": MLP(nn.Tanh())}))
