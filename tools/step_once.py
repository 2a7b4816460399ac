"""Run the headline step eagerly N times (for ncu captures; no graphs, no timing): 4 act-quant + 4 GEMM launches per step,
exactly bench.py's step (q/k/v and gate/up fused).  usage: step_once.py [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import protoquant_b200 as pq
from protoquant_b200 import functional as F
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
torch.manual_seed(0)
sizes = {"qkv_proj": (4096, 3 * 4096), "o_proj": (4096, 4096), "gate_up_proj": (4096, 2 * 11008), "down_proj": (11008, 4096)}
mods = {}
for g, (k, nn_) in sizes.items():
    m = pq.DynamicQuantLinear(k, nn_, bias=True, device=dev)
    m.qweight_storage.random_(-127, 128); m.weight_scale.uniform_(1e-4, 1e-3)
    mods[g] = m
M = bench.M_TOKENS
acts = {a: torch.randn(M, k, device=dev).to(torch.bfloat16) for a, k in bench.ACTS.items()}
ws = {a: (F.alloc_q(M, k, dev), torch.empty(M, dtype=torch.float32, device=dev)) for a, k in bench.ACTS.items()}
outs = {g: torch.empty(M, nn_, dtype=torch.bfloat16, device=dev) for g, (k, nn_) in sizes.items()}
for _ in range(n):
    for g, members, src in bench.GROUPS:
        m = mods[g]
        F.qlinear_into(acts[src], m.qweight_storage, m.in_features, m.weight_scale, m.bias, outs[g], *ws[src])
torch.cuda.synchronize()
