"""Bench step (graph replay) with the L2 weight prefetcher on/off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import protoquant_b200 as pq
from protoquant_b200 import functional as F
import bench
dev = torch.device("cuda", 0)
mods = {}
for name, k, n, src in bench.LINEARS:
    mods[name] = pq.DynamicQuantLinear.from_float(torch.nn.Linear(k, n).to(torch.bfloat16).to(dev))
acts = {a: torch.randn(2048, k, device=dev).to(torch.bfloat16) for a, k in bench.ACTS.items()}
ws = {a: (F.alloc_q(2048, k, dev), torch.empty(2048, dtype=torch.float32, device=dev)) for a, k in bench.ACTS.items()}
outs = {name: torch.empty(2048, n, dtype=torch.bfloat16, device=dev) for name, k, n, _ in bench.LINEARS}
def step(gemm_only=False):
    for name, k, n, src in bench.LINEARS:
        m = mods[name]
        if not gemm_only:
            F.quantize_act(acts[src], out=ws[src])
        F.qgemm(ws[src][0], ws[src][1], m.qweight, m.weight_scale, m.bias, torch.bfloat16, out=outs[name])
for pf in (0, 2, 0, 2, 1):
    pq.lib().pq_debug_set_prefetch(pf)
    for gemm_only in (False, True):
        step(gemm_only); torch.cuda.synchronize()
        ref = {k: v.clone() for k, v in outs.items()}
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            step(gemm_only)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step(gemm_only)
        for _ in range(5): g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200): g.replay()
        e1.record(); torch.cuda.synchronize()
        ok = all(torch.equal(ref[k], outs[k]) for k in outs)
        ms = e0.elapsed_time(e1) / 200
        print(f"prefetch={pf} gemm_only={gemm_only}: {ms*1e3:.1f} us/step  {bench.OPS_PER_STEP/ms/1e9:.0f} TOPS  outputs {'identical' if ok else 'DIFFER'}")
