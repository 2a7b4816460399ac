"""Run a few GEMM launches of a given shape/config (for ncu).  usage: prof_gemm.py M N K cfg [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import protoquant_b200 as pq
M, N, K, cfg = (int(v) for v in sys.argv[1:5])
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 3
pq.lib().pq_debug_set_gemm_config(cfg)
a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device="cuda")
b = torch.randint(-128, 128, (N, K), dtype=torch.int8, device="cuda")
sx = torch.rand(M, device="cuda"); sw = torch.rand(N, device="cuda")
y = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
for _ in range(iters):
    pq.qgemm(a, sx, b, sw, None, torch.bfloat16, out=y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    pq.qgemm(a, sx, b, sw, None, torch.bfloat16, out=y)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"M={M} N={N} K={K} cfg={cfg}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.0f} TOPS")
