"""Time the row-wise quantizer from a CUDA graph.  usage: prof_quant.py M K dtype [tpr vpt]  (dtype: bf16|f16|f32)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import protoquant_b200 as pq
from protoquant_b200 import functional as F
M, K = int(sys.argv[1]), int(sys.argv[2])
dt = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[sys.argv[3]]
cfgs = [(0, 0)]
if len(sys.argv) > 5:
    cfgs = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(4, len(sys.argv) - 1, 2)]
esz = torch.empty(0, dtype=dt).element_size()
nbuf = max(1, min(8, int(600e6 // (M * K * esz))))          # rotate inputs so small cases are not L2-resident
xs = [torch.randn(M, K, device="cuda").to(dt) for _ in range(nbuf)]
q = F.alloc_q(M, K, "cuda"); s = torch.empty(M, dtype=torch.float32, device="cuda")
for tpr, vpt in cfgs:
    pq.lib().pq_debug_set_quant_config(tpr, vpt)
    for mode, bufs in (("hot", xs[:1]), ("rot", xs)):
        iters = 20 if M * K > 5e7 else 100
        def run():
            for i in range(iters):
                F.quantize_act(bufs[i % len(bufs)], out=(q, s))
        run(); torch.cuda.synchronize()
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run()
        g.replay(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / iters)
        byts = M * (K * esz + K + 4)
        print(f"M={M} K={K} {sys.argv[3]} tpr={tpr} vpt={vpt} {mode}({len(bufs)} bufs): {best*1e3:.2f} us  {byts/best/1e6:.0f} GB/s")
