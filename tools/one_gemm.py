"""Launch one GEMM shape/config a few times (for ncu captures).  usage: one_gemm.py M N K cfg [launches]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import protoquant_b200 as pq
M, N, K, cfg = (int(v) for v in sys.argv[1:5])
n = int(sys.argv[5]) if len(sys.argv) > 5 else 3
pq.lib().pq_debug_set_gemm_config(cfg)
pq.lib().pq_debug_set_staged(int(os.environ.get("PQ_STAGED", "0")))
a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device="cuda")
b = torch.randint(-128, 128, (N, K), dtype=torch.int8, device="cuda")
sx = torch.rand(M, device="cuda"); sw = torch.rand(N, device="cuda")
y = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
for _ in range(n):
    pq.qgemm(a, sx, b, sw, None, torch.bfloat16, out=y)
torch.cuda.synchronize()
