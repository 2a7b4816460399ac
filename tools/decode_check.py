"""Module forward at 16 tokens (graph replay, 8 rotating weight sets) with the fused decode kernel on/off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, protoquant_b200 as pq
dev = torch.device("cuda", 0)
for fused in (1, 0, 1, 0):
    pq.lib().pq_debug_set_fused_decode(fused)
    res = {}
    for name, K, N, M in (("4096x4096", 4096, 4096, 16), ("4096->11008", 4096, 11008, 16), ("11008->4096", 11008, 4096, 16), ("4096x4096 M=1", 4096, 4096, 1), ("4096x4096 M=32", 4096, 4096, 32)):
        mods = []
        for i in range(8):
            m = pq.DynamicQuantLinear(K, N, bias=True, device=dev); m.qweight_storage.random_(-127, 128); m.weight_scale.uniform_(1e-4, 1e-3); mods.append(m)
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        def run():
            for m in mods: m(x)
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
        res[name] = round(a.elapsed_time(b) * 1e3 / 160, 2)
    print("fused_decode", fused, res)
