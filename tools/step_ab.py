"""A/B the headline step (7 Llama-7B linears at 2048 tokens, CUDA-graph replay) under debug switches.
usage: step_ab.py [reps]   prints ms/step for each switch setting."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import protoquant_b200 as pq
from protoquant_b200 import functional as F
import bench
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device("cuda", 0)
torch.manual_seed(0)
mods = {}
for name, k, n, src in bench.LINEARS:
    lin = torch.nn.Linear(k, n, bias=True).to(torch.bfloat16).to(dev)
    mods[name] = pq.DynamicQuantLinear.from_float(lin)
M = bench.M_TOKENS
acts = {a: torch.randn(M, k, device=dev).to(torch.bfloat16) for a, k in bench.ACTS.items()}
ws = {a: (F.alloc_q(M, k, dev), torch.empty(M, dtype=torch.float32, device=dev)) for a, k in bench.ACTS.items()}
outs = {name: torch.empty(M, n, dtype=torch.bfloat16, device=dev) for name, k, n, _ in bench.LINEARS}
def step():
    for name, k, n, src in bench.LINEARS:
        m = mods[name]
        F.qlinear_into(acts[src], m.qweight_storage, m.in_features, m.weight_scale, m.bias, outs[name], *ws[src])
def gemm_only():
    for name, k, n, src in bench.LINEARS:
        m = mods[name]
        F.qgemm(ws[src][0], ws[src][1], m.qweight, m.weight_scale, m.bias, torch.bfloat16, out=outs[name])
def timed(fn):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(5): g.replay()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): g.replay()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best
L = pq.lib()
def setall(pf=1, narrow=1, tma=1):
    L.pq_debug_set_weight_prefetch(pf); L.pq_debug_set_narrow_tiles(narrow); L.pq_debug_set_tma_store(tma)
for label, kw in (("default", {}), ("no weight prefetch", dict(pf=0)), ("no narrow tiles", dict(narrow=0)),
                  ("no tma store", dict(tma=0)), ("none of the three", dict(pf=0, narrow=0, tma=0))):
    setall(**kw)
    t = timed(step)
    tg = timed(gemm_only)
    print(f"{label:22s}: step {t*1e3:7.1f} us = {bench.OPS_PER_STEP/t/1e9:6.0f} TOPS   gemm-only {tg*1e3:7.1f} us = {bench.OPS_PER_STEP/tg/1e9:6.0f} TOPS")
setall()
