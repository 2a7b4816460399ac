"""Time the transposed-output quantizer.  usage: prof_quant_t.py M K"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, protoquant_b200 as pq
from protoquant_b200 import functional as F
M, K = int(sys.argv[1]), int(sys.argv[2])
if len(sys.argv) > 4 and sys.argv[4] == "tiled":
    pq.lib().pq_debug_set_quant_staged(-1)                        # the 32-rows-per-CTA kernel instead of the two-launch path
x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
q = F.alloc_q(K, M, "cuda"); s = torch.empty(M, dtype=torch.float32, device="cuda")
def run():
    for _ in range(10): F.quantize_act(x, transpose=True, out=(q, s))
run(); torch.cuda.synchronize()
st = torch.cuda.Stream()
with torch.cuda.stream(st): run()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g): run()
g.replay(); torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) / 10)
print(f"transposed quant M={M} K={K} bf16: {best*1e3:.1f} us  {M*(3*K+4)/best/1e6:.0f} GB/s  ({M*(3*K+4)/best/1e6/6552:.2f} of the measured HBM peak)")
