"""RMSNorm -> int8 launch-shape experiment at activation sizes (gpurun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes, torch, protoquant_b200 as pq
from protoquant_b200 import functional as F
L = pq.lib(); L.pq_debug_set_fused_quant_config.argtypes = [ctypes.c_int, ctypes.c_int]; L.pq_debug_set_fused_quant_config.restype = None
dev = torch.device("cuda")


def timed(fn, iters=50):
    fn(0); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s): fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters): fn(i)
    g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (5 * iters)


for M, K in ((2048, 4096), (2048, 8192), (4096, 768), (512, 4096), (8192, 4096)):
    nb = max(1, min(8, int(600e6 // (M * K * 2))))
    xs = [torch.randn(M, K, device=dev).to(torch.bfloat16) for _ in range(nb)]
    w = torch.ones(K, dtype=torch.bfloat16, device=dev)
    out = (F.alloc_q(M, K, dev), torch.empty(M, dtype=torch.float32, device=dev))
    byt = M * (3 * K + 4)
    res = []
    nvec = K // 8
    for tpr, vpt in ((0, 0), (32, 8), (64, 4), (64, 6), (64, 8), (128, 3), (128, 4), (128, 6), (128, 8), (256, 3), (256, 4), (512, 3)):
        if tpr and (tpr * vpt < nvec or tpr * vpt > 2 * nvec):
            continue
        L.pq_debug_set_fused_quant_config(tpr, vpt)
        us = timed(lambda i: F.rmsnorm_quant(xs[i % nb], w, out=out))
        res.append(f"{tpr}x{vpt}={us:.2f}us({byt / us / 1e3 / 6552:.2f})")
    L.pq_debug_set_fused_quant_config(0, 0)
    gs = [torch.randn(M, 2 * K, device=dev).to(torch.bfloat16) for _ in range(max(1, nb // 2))]
    res2 = []
    for tpr, vpt in ((0, 0), (64, 8), (128, 4), (128, 8), (256, 4), (256, 6)):
        if tpr and (tpr * vpt < nvec or tpr * vpt > 2 * nvec):
            continue
        L.pq_debug_set_fused_quant_config(tpr, vpt)
        us = timed(lambda i: F.act_mul_quant(gs[i % len(gs)][:, :K], gs[i % len(gs)][:, K:], act="silu", out=out))
        res2.append(f"{tpr}x{vpt}={us:.2f}us({M * (5 * K + 4) / us / 1e3 / 6552:.2f})")
    L.pq_debug_set_fused_quant_config(0, 0)
    print(f"rmsnorm {M}x{K}: " + "  ".join(res), flush=True)
    print(f"silu_mul {M}x{K}: " + "  ".join(res2), flush=True)
