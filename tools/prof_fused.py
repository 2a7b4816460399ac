"""Time the producer-fused quantizers against their unfused compositions (CUDA-graph timing, rotating inputs).
usage: prof_fused.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import protoquant_b200 as pq
from protoquant_b200 import functional as F

def timed(fn, iters):
    fn(0); torch.cuda.synchronize()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i)
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    return best * 1e3

for (M, K) in ((2048, 4096), (131072, 4096), (4096, 768), (65536, 8192)):
    nb = max(1, min(8, int(600e6 // (M * K * 2))))
    xs = [torch.randn(M, K, device="cuda").to(torch.bfloat16) for _ in range(nb)]
    w = torch.ones(K, dtype=torch.bfloat16, device="cuda")
    q = F.alloc_q(M, K, "cuda"); s = torch.empty(M, dtype=torch.float32, device="cuda")
    iters = 20 if M * K > 5e7 else 100
    t_f = timed(lambda i: F.rmsnorm_quant(xs[i % nb], w, out=(q, s)), iters)
    def unfused(i):
        y = torch.nn.functional.rms_norm(xs[i % nb], (K,), w, 1e-6)
        F.quantize_act(y, out=(q, s))
    t_u = timed(unfused, iters)
    t_q = timed(lambda i: F.quantize_act(xs[i % nb], out=(q, s)), iters)
    byts = M * (3 * K + 4)
    print(f"rmsnorm_quant {M}x{K} bf16: fused {t_f:.1f} us = {byts/t_f/1e3:.0f} GB/s | torch rms_norm + act-quant {t_u:.1f} us | act-quant alone {t_q:.1f} us")
for (M, K) in ((2048, 11008), (65536, 11008), (4096, 3072), (16384, 28672)):
    nb = max(1, min(8, int(600e6 // (M * K * 4))))
    gs = [torch.randn(M, 2 * K, device="cuda").to(torch.bfloat16) for _ in range(nb)]
    q = F.alloc_q(M, K, "cuda"); s = torch.empty(M, dtype=torch.float32, device="cuda")
    iters = 20 if M * K > 2e7 else 100
    t_f = timed(lambda i: F.act_mul_quant(gs[i % nb][:, :K], gs[i % nb][:, K:], act="silu", out=(q, s)), iters)
    def unfused(i):
        h = torch.nn.functional.silu(gs[i % nb][:, :K]) * gs[i % nb][:, K:]
        F.quantize_act(h, out=(q, s))
    t_u = timed(unfused, iters)
    byts = M * (5 * K + 4)
    print(f"silu*up quant {M}x{K} bf16: fused {t_f:.1f} us = {byts/t_f/1e3:.0f} GB/s | torch silu*mul + act-quant {t_u:.1f} us")
