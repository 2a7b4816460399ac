"""A/B of the headline step (4 act-quant + 4 GEMM launches, CUDA-graph replay, bench.py's own step) with the staged
quantizer on (heuristic) / off, interleaved so that clocks and box are the same.  usage: step_ab2.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import protoquant_b200 as pq
from protoquant_b200 import functional as F
import bench
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = torch.device("cuda", 0)
torch.manual_seed(0)
blk = bench.Llama7BBlockLinears(torch, pq, dev)
fused = blk.fused
M = bench.M_TOKENS
acts = {a: torch.randn(M, k, device=dev).to(torch.bfloat16) for a, k in bench.ACTS.items()}
ws = {a: (F.alloc_q(M, k, dev), torch.empty(M, dtype=torch.float32, device=dev)) for a, k in bench.ACTS.items()}
outs = {g: torch.empty(M, fused[g].out_features, dtype=torch.bfloat16, device=dev) for g, _, _ in bench.GROUPS}


def step():
    for g, members, src in bench.GROUPS:
        m = fused[g]
        F.qlinear_into(acts[src], m.qweight_storage, m.in_features, m.weight_scale, m.bias, outs[g], *ws[src])


def quant_only():
    for g, members, src in bench.GROUPS:
        F.quantize_act(acts[src], out=ws[src])


def graph_of(fn):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


def t(g):
    for _ in range(5): g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


graphs = {}
for mode, label in ((0, "staged_auto"), (-1, "staged_off")):
    pq.lib().pq_debug_set_quant_staged(mode)
    graphs[label] = (graph_of(step), graph_of(quant_only))
pq.lib().pq_debug_set_quant_staged(0)
for rnd in range(4):
    for label, (gs, gq) in graphs.items():
        print(f"round {rnd} {label:12s}: step {t(gs):7.1f} us   quant-only (4 launches) {t(gq):6.1f} us", flush=True)
