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
