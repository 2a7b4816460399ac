"""Decode forward (16 tokens) split into its parts, PDL on / off: CUDA-graph replay over 8 rotating weight sets."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, protoquant_b200 as pq
from protoquant_b200 import functional as F
dev = torch.device("cuda", 0)


def timed(run, per):
    run(); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s): run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): run()
    for _ in range(3): g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): g.replay()
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) * 1e3 / (20 * per), 2)


for K, N, M in ((4096, 4096, 16), (4096, 11008, 16), (11008, 4096, 16), (8192, 8192, 16), (8192, 28672, 16), (4096, 4096, 1), (4096, 4096, 64), (768, 3072, 16)):
    mods = []
    for i in range(8):
        m = pq.DynamicQuantLinear(K, N, bias=True, device=dev); m.qweight_storage.random_(-127, 128); m.weight_scale.uniform_(1e-4, 1e-3); mods.append(m)
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    ws = (F.alloc_q(M, K, dev), torch.empty(M, dtype=torch.float32, device=dev))
    y = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    F.quantize_act(x, out=ws)
    for pdl, S in ((1, 0), (0, 0), (1, 1), (1, 2), (1, 4), (1, 8)):
        pq.lib().pq_debug_set_pdl(pdl)
        pq.lib().pq_debug_set_smallm_splits(S)
        res = {"S": S}
        res["gemm_only"] = timed(lambda: [F.qgemm(ws[0], ws[1], m.qweight, m.weight_scale, m.bias, torch.bfloat16, out=y) for m in mods], 8)
        res["quant_only"] = timed(lambda: [F.quantize_act(x, out=ws) for m in mods], 8)
        res["quant+gemm"] = timed(lambda: [(F.quantize_act(x, out=ws), F.qgemm(ws[0], ws[1], m.qweight, m.weight_scale, m.bias, torch.bfloat16, out=y)) for m in mods], 8)
        res["module"] = timed(lambda: [m(x) for m in mods], 8)
        print(f"M={M} K={K} N={N} pdl={pdl}: {res}  weights {N*K/1e6:.1f} MB = {N*K/6552e3:.2f} us at HBM peak", flush=True)
    pq.lib().pq_debug_set_pdl(1)
    pq.lib().pq_debug_set_smallm_splits(0)
