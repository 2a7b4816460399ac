"""GPU tests of the nn.Linear replacement, the swap helper, the host-buffer C API and the
sharded module (single rank; multi-rank when the box has more than one GPU)."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
from torch import nn

from conftest import ROOT
import protoquant_b200 as pq
import protoquant_oracle as O

pytestmark = pytest.mark.gpu


def _bits(t):
    return t.view(torch.int32 if t.dtype == torch.float32 else torch.int16)


@pytest.mark.parametrize("dtype,name", [(torch.bfloat16, "bf16"), (torch.float16, "f16")])
@pytest.mark.parametrize("features,tokens", [((4096, 4096), (16,)), ((768, 3072), (32, 128)), ((3072, 768), (32, 128)),
                                             ((4096, 11008), (2, 256)), ((520, 264), (3, 5))])
def test_dynamic_quant_linear_matches_oracle(dtype, name, features, tokens):
    """BASELINE.json configs[0] (4096x4096, 16 tokens) and configs[1] (BERT-base, 32x128) among others."""
    fin, fout = features
    torch.manual_seed(0)
    lin = nn.Linear(fin, fout).to(dtype)
    x = torch.randn(*tokens, fin).to(dtype)
    m = pq.DynamicQuantLinear.from_float(lin.cuda())
    before = pq.launch_count()
    y = m(x.cuda())
    assert pq.launch_count() - before == 2          # act-quant + GEMM, nothing else is launched
    assert y.shape == (*tokens, fout) and y.dtype == dtype
    lin = lin.cpu()
    wq, sw = O.quantize_weight(lin.weight.detach())
    ref = O.qlinear(x, wq, sw, lin.bias.detach().float().numpy(), out_dtype=name)
    assert torch.equal(_bits(y.cpu()), _bits(ref))
    fl = lin.float()(x.float())
    assert (y.cpu().float() - fl).abs().max() < 0.06 * fl.abs().max()


def test_swap_linear_on_an_mlp_block():
    torch.manual_seed(1)

    class MLP(nn.Module):
        def __init__(self):
            super().__init__()
            self.up = nn.Linear(256, 1024)
            self.act = nn.GELU()
            self.down = nn.Linear(1024, 256)
            self.tiny = nn.Linear(256, 8)

        def forward(self, x):
            return self.tiny(self.down(self.act(self.up(x))))

    net = MLP().to(torch.bfloat16).cuda()
    x = torch.randn(64, 256, dtype=torch.bfloat16, device="cuda")
    ref = net(x)
    pq.swap_linear(net, min_features=64)
    assert isinstance(net.up, pq.DynamicQuantLinear) and isinstance(net.down, pq.DynamicQuantLinear)
    assert isinstance(net.tiny, nn.Linear)              # below min_features -> untouched
    y = net(x)
    assert (y.float() - ref.float()).abs().max() < 0.1 * ref.float().abs().max() + 0.05
    sd = net.state_dict()
    assert sd["up.qweight_storage"].dtype == torch.int8 and sd["up.weight_scale"].dtype == torch.float32


def test_host_buffer_c_api_roundtrip():
    """pq_linear_create / pq_linear_forward_host: the call bench.py's e2e number goes through."""
    lib = pq.lib()
    N, K, M = 512, 768, 100
    g = torch.Generator().manual_seed(2)
    w = ((torch.rand(N, K, generator=g) * 2 - 1) / K ** 0.5).contiguous()
    b = torch.randn(N, generator=g).contiguous()
    x = torch.randn(M, K, generator=g).to(torch.bfloat16).contiguous().pin_memory()
    y = torch.empty(M, N, dtype=torch.bfloat16).pin_memory()
    h = ctypes.c_void_p()
    assert lib.pq_linear_create(ctypes.byref(h), w.data_ptr(), 0, N, K, b.data_ptr(), 128, 2, 2, None) == 0
    try:
        assert lib.pq_linear_forward_host(h, x.data_ptr(), y.data_ptr(), M) == 0
        wq, sw = O.quantize_weight(w)
        ref = O.qlinear(x, wq, sw, b.numpy(), out_dtype="bf16")
        assert torch.equal(_bits(y), _bits(ref))
        assert lib.pq_linear_forward_host(h, x.data_ptr(), y.data_ptr(), 129) == 1   # > max_tokens
    finally:
        lib.pq_linear_destroy(h)


def test_fused_linears_equal_the_separate_linears_bit_for_bit():
    torch.manual_seed(9)
    qkv = [pq.DynamicQuantLinear.from_float(nn.Linear(1024, n).to(torch.bfloat16).cuda()) for n in (1024, 256, 256)]
    fused = pq.fuse_linears(qkv)
    x = torch.randn(300, 1024, dtype=torch.bfloat16, device="cuda")
    parts = fused(x).split([1024, 256, 256], dim=-1)
    for m, y in zip(qkv, parts):
        assert torch.equal(m(x), y)


class _LlamaLikeBlock(nn.Module):
    def __init__(self, hidden=512, kv=256, inter=1376):
        super().__init__()
        self.q_proj = nn.Linear(hidden, hidden, bias=False)
        self.k_proj = nn.Linear(hidden, kv, bias=False)      # grouped-query attention: narrower k / v
        self.v_proj = nn.Linear(hidden, kv, bias=False)
        self.o_proj = nn.Linear(hidden, hidden, bias=False)
        self.gate_proj = nn.Linear(hidden, inter, bias=False)
        self.up_proj = nn.Linear(hidden, inter, bias=False)
        self.down_proj = nn.Linear(inter, hidden, bias=False)


def test_swap_linear_quantises_each_shared_activation_once():
    """swap_linear(fuse_shared_inputs=True) (the default): q/k/v and gate/up share ONE act-quant launch and ONE GEMM,
    and every member's output has the bits of the separately swapped module."""
    import copy
    torch.manual_seed(5)
    blk = _LlamaLikeBlock().to(torch.bfloat16).cuda()
    sep = pq.swap_linear(copy.deepcopy(blk), fuse_shared_inputs=False)
    fus = pq.swap_linear(copy.deepcopy(blk))
    assert isinstance(fus.q_proj, pq.SharedInputLinear) and isinstance(fus.up_proj, pq.SharedInputLinear)
    assert isinstance(fus.o_proj, pq.DynamicQuantLinear) and isinstance(sep.q_proj, pq.DynamicQuantLinear)
    x = torch.randn(3, 50, 512, dtype=torch.bfloat16, device="cuda")
    before = pq.launch_count()
    q, k, v = fus.q_proj(x), fus.k_proj(x), fus.v_proj(x)
    assert pq.launch_count() - before == 2                     # one act-quant + one GEMM for all three
    for name, y in (("q_proj", q), ("k_proj", k), ("v_proj", v)):
        assert torch.equal(y, getattr(sep, name)(x)), name
    before = pq.launch_count()
    g, u = fus.gate_proj(x), fus.up_proj(x)
    assert pq.launch_count() - before == 2
    assert torch.equal(g, sep.gate_proj(x)) and torch.equal(u, sep.up_proj(x))
    # a second round with a new tensor recomputes; an in-place update of the same tensor is noticed (version counter)
    x2 = torch.randn_like(x)
    assert torch.equal(fus.k_proj(x2), sep.k_proj(x2)) and torch.equal(fus.q_proj(x2), sep.q_proj(x2))
    assert torch.equal(fus.v_proj(x2), sep.v_proj(x2))
    q_a = fus.q_proj(x).clone()
    x.mul_(2)
    assert torch.equal(fus.k_proj(x), sep.k_proj(x))            # not the stale activation
    assert torch.equal(fus.v_proj(x), sep.v_proj(x)) and torch.equal(fus.q_proj(x), sep.q_proj(x))
    assert not torch.equal(fus.q_proj(x), q_a)
    fus.k_proj(x), fus.v_proj(x)
    # state_dict: the fused parameters are registered once, on the parent
    sd = fus.state_dict()
    assert sd["_pq_fused_q_proj.qweight_storage"].shape[0] == 512 + 256 + 256 and "q_proj.qweight_storage" not in sd
    fus2 = pq.swap_linear(copy.deepcopy(blk))
    fus2.load_state_dict(sd)
    assert torch.equal(fus2.v_proj(x2), sep.v_proj(x2))


def test_shared_input_group_with_different_inputs_falls_back_to_its_own_slice():
    """Cross-attention style use: q reads one tensor, k / v another -> every member still returns the right bits."""
    import copy
    torch.manual_seed(6)
    blk = _LlamaLikeBlock().to(torch.bfloat16).cuda()
    sep = pq.swap_linear(copy.deepcopy(blk), fuse_shared_inputs=False)
    fus = pq.swap_linear(copy.deepcopy(blk))
    xa = torch.randn(40, 512, dtype=torch.bfloat16, device="cuda")
    xb = torch.randn(24, 512, dtype=torch.bfloat16, device="cuda")
    for _ in range(6):      # more rounds than _SharedInputGroup.MAX_DIVERGENT: the group stops fusing, results stay exact
        assert torch.equal(fus.q_proj(xa), sep.q_proj(xa))
        assert torch.equal(fus.k_proj(xb), sep.k_proj(xb))
        assert torch.equal(fus.v_proj(xb), sep.v_proj(xb))


def test_module_dtype_casts_keep_the_fp32_abi_buffers():
    """module.half() / .to(bfloat16) must not cast weight_scale / bias (read as fp32 by the kernels)."""
    torch.manual_seed(7)
    lin = nn.Linear(256, 384).to(torch.bfloat16).cuda()
    m = pq.DynamicQuantLinear.from_float(lin)
    x = torch.randn(33, 256, dtype=torch.bfloat16, device="cuda")
    ref = m(x)
    net = nn.Sequential(m)
    net.half()
    assert m.weight_scale.dtype == torch.float32 and m.bias.dtype == torch.float32 and m.qweight_storage.dtype == torch.int8
    net.to(torch.bfloat16)
    assert torch.equal(m(x), ref)
    m.weight_scale = m.weight_scale.to(torch.float16)            # bypassing _apply: the lean path must refuse, not misread
    with pytest.raises(TypeError):
        m(x)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_module_on_a_non_current_device():
    """weights and input on cuda:1 while cuda:0 is current (single-process model parallelism)."""
    torch.manual_seed(8)
    lin = nn.Linear(512, 640).to(torch.bfloat16)
    x = torch.randn(300, 512, dtype=torch.bfloat16)
    torch.cuda.set_device(0)
    m0 = pq.DynamicQuantLinear.from_float(copy_to(lin, "cuda:0"))
    y0 = m0(x.to("cuda:0"))
    m1 = pq.DynamicQuantLinear.from_float(copy_to(lin, "cuda:1"))
    assert torch.cuda.current_device() == 0
    y1 = m1(x.to("cuda:1"))
    assert y1.device.index == 1 and torch.equal(y1.cpu(), y0.cpu())
    with pytest.raises(pq.ProtoquantError):
        pq.qgemm(*pq.quantize_act(x.to("cuda:1")), m0.qweight, m0.weight_scale)


def copy_to(lin, device):
    import copy
    return copy.deepcopy(lin).to(device)


def test_sharded_module_single_rank_equals_unsharded():
    torch.manual_seed(3)
    lin = nn.Linear(512, 1000).to(torch.bfloat16).cuda()
    m = pq.DynamicQuantLinear.from_float(lin)
    sh = pq.ShardedDynamicQuantLinear(m.qweight, m.weight_scale, m.bias)
    x = torch.randn(77, 512, dtype=torch.bfloat16, device="cuda")
    assert torch.equal(sh(x), m(x))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_module_nccl_bit_identical():
    n = min(torch.cuda.device_count(), 8)
    n = 1 << (n.bit_length() - 1)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "_sharded_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert "SHARDED_OK" in p.stdout


@pytest.mark.parametrize("dtype,name", [(torch.bfloat16, "bf16"), (torch.float16, "f16"), (torch.float32, "f32")])
@pytest.mark.parametrize("shape", [(1, 4096, 4096), (16, 4096, 4096), (7, 11008, 4096), (16, 4096, 11008), (32, 4096, 4096),
                                   (17, 768, 3072), (9, 3072, 768), (16, 136, 4096), (5, 264, 144), (16, 28672, 1024)])
def test_fused_decode_linear_bit_exact(dtype, name, shape):
    """activation quantisation fused into the small-M GEMM (one launch; experimental, off by
    default).  Must equal the two-kernel path and the oracle bit for bit, for every scale-arithmetic knob."""
    M, N, K = shape
    g = torch.Generator().manual_seed(61)
    x = torch.randn(M, K, generator=g)
    x[0, K // 2] = 50.0
    if M > 2:
        x[2].zero_()
    if M > 3:
        x[3] *= 1e-30 if dtype != torch.float16 else 1e-3
    x = x.to(dtype)
    w = (torch.rand(N, K, generator=g) * 2 - 1) / K ** 0.5
    bias = torch.randn(N, generator=g)
    wq_o, sw_o = O.quantize_weight(w)
    lin = pq.DynamicQuantLinear(K, N, bias=True, device="cuda")
    lin.qweight_storage[:, :K].copy_(torch.from_numpy(wq_o))
    lin.weight_scale.copy_(torch.from_numpy(sw_o))
    lin.bias.copy_(bias)
    for spec, ospec in ((None, O.QuantSpec()), (pq.QuantSpec(scale_mode=1, eps=1e-5), O.QuantSpec.torch_ao())):
        lin.spec = spec
        want = O.qlinear(x, wq_o, sw_o, bias.numpy(), spec=ospec, out_dtype=name)
        outs = []
        for fused in (1, 0):
            pq.lib().pq_debug_set_fused_decode(fused)
            before = pq.launch_count()
            outs.append(lin(x.cuda()))
            n = pq.launch_count() - before
            assert n == 2 if fused == 0 else n in (1, 2)
        pq.lib().pq_debug_set_fused_decode(0)      # default: off (measured slower, see profiles/README_r1.md)
        assert torch.equal(_bits(outs[0].cpu()), _bits(want))
        assert torch.equal(_bits(outs[1].cpu()), _bits(want))


def _same_float_bits(a: torch.Tensor, b: torch.Tensor) -> bool:
    """Bit equality except that a NaN equals any NaN (NaN payloads are not part of the contract)."""
    a, b = a.cpu(), b.cpu()
    nan = torch.isnan(a) & torch.isnan(b)
    return bool(torch.equal(torch.where(nan, torch.zeros_like(a), a).view(torch.int16 if a.element_size() == 2 else torch.int32),
                            torch.where(nan, torch.zeros_like(b), b).view(torch.int16 if b.element_size() == 2 else torch.int32)))


@pytest.mark.parametrize("dtype,name", [(torch.bfloat16, "bf16"), (torch.float16, "f16"), (torch.float32, "f32")])
@pytest.mark.parametrize("M", [8, 200])
def test_linear_with_nonfinite_tokens_follows_the_policy(dtype, name, M):
    """A token holding NaN / +-inf gets a NaN / inf scale and zero codes, so its output row is 0 * s_x = NaN (+ bias);
    every other token of the batch is unaffected and still bit-exact.  Decode kernel (M = 8, both the two-launch and
    the fused variant) and the large-M kernel (M = 200)."""
    K, N = 1024, 384
    g = torch.Generator().manual_seed(5)
    x = torch.randn(M, K, generator=g)
    x[1, 3] = float("inf")
    x[2, K - 1] = float("nan")
    x[5, 17] = float("-inf"); x[5, 18] = float("nan")
    x = x.to(dtype)
    w = (torch.rand(N, K, generator=g) * 2 - 1) / K ** 0.5
    bias = torch.randn(N, generator=g)
    wq_o, sw_o = O.quantize_weight(w)
    lin = pq.DynamicQuantLinear(K, N, bias=True, device="cuda")
    lin.qweight_storage[:, :K].copy_(torch.from_numpy(wq_o))
    lin.weight_scale.copy_(torch.from_numpy(sw_o))
    lin.bias.copy_(bias)
    with np.errstate(invalid="ignore"):
        want = O.qlinear(x, wq_o, sw_o, bias.numpy(), out_dtype=name)
    assert bool(torch.isnan(want[[1, 2, 5]].float()).all()) and not bool(torch.isnan(want[[0, 3, 4, 6, 7]].float()).any())
    for fused in ((1, 0) if M <= 64 else (0,)):
        pq.lib().pq_debug_set_fused_decode(fused)
        try:
            y = lin(x.cuda())
        finally:
            pq.lib().pq_debug_set_fused_decode(0)
        assert _same_float_bits(y, want), f"fused={fused}"


def test_serialisation_roundtrip_keeps_the_bits(tmp_path):
    """QTensor.save/load and the module's state_dict through torch.save/torch.load reproduce the
    dequantised tensor and the forward output bit for bit; a full checkpoint loads into the (single-rank) sharded
    modules with the same result."""
    torch.manual_seed(3)
    x = torch.randn(40, 520, dtype=torch.bfloat16, device="cuda")
    qt = pq.quantize(x)
    qt.save(tmp_path / "act.pt")
    back = pq.QTensor.load(tmp_path / "act.pt", device="cuda")
    assert torch.equal(back.dequantize(), qt.dequantize()) and back.data.is_cuda

    lin = nn.Linear(520, 264).to(torch.bfloat16).cuda()
    m = pq.DynamicQuantLinear.from_float(lin)
    y = m(x)
    torch.save(m.state_dict(), tmp_path / "lin.pt")
    sd = torch.load(tmp_path / "lin.pt", map_location="cpu", weights_only=True)
    m2 = pq.DynamicQuantLinear(520, 264, bias=True, device="cuda")
    m2.load_state_dict(sd)
    assert torch.equal(_bits(m2(x)), _bits(y))
    assert torch.equal(_bits(m2(back)), _bits(y))                     # a loaded QTensor is a valid pre-quantised input
    col = pq.ShardedDynamicQuantLinear.from_full_state_dict(sd, in_features=520)
    row = pq.RowParallelDynamicQuantLinear.from_full_state_dict(sd, in_features=520)
    assert torch.equal(_bits(col(x)), _bits(y)) and torch.equal(_bits(row(x)), _bits(y))
    m3 = pq.DynamicQuantLinear.from_qtensor(pq.QTensor.from_state(m.weight_qtensor().state(), device="cuda"), bias=m.bias)
    assert torch.equal(_bits(m3(x)), _bits(y))
