"""Act-quant launch-shape experiment at activation sizes (gpurun): register-resident kernel under each forced
(threads per row, vectors per thread) vs the persistent shared-memory staged kernel, CUDA-graph timing over
rotating inputs (> L2 in total).   python tools/quant_cfg.py > gpurun_out/quant_cfg.log"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import protoquant_b200 as pq
from protoquant_b200 import functional as F

HBM = 6552.0
dev = torch.device("cuda")
L = pq.lib()


def timed(fn, iters=40, reps=5):
    fn(0)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn(0)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i)
    g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (reps * iters)


def main():
    shapes = [(2048, 4096), (2048, 11008), (4096, 768), (4096, 3072), (512, 4096), (8192, 4096), (2048, 8192), (131072, 4096)]
    for dt in (torch.bfloat16, torch.float32):
        for M, K in shapes:
            if dt == torch.float32 and M > 8192:
                continue
            esz = 2 if dt == torch.bfloat16 else 4
            nb = max(1, min(12, int(700e6 // (M * K * esz))))
            xs = [torch.randn(M, K, device=dev).to(dt) for _ in range(nb)]
            out = (F.alloc_q(M, K, dev), torch.empty(M, dtype=torch.float32, device=dev))
            byt = M * (K * esz + K + 4)
            res = []
            L.pq_debug_set_quant_staged(-1)
            us = timed(lambda i: F.quantize_act(xs[i % nb], out=out))
            ref_q, ref_s = F.quantize_act(xs[0])
            res.append(("vec_auto", us))
            nvec = K * esz // 16
            for tpr, vpt in ((32, 8), (64, 4), (64, 8), (128, 4), (128, 8), (256, 2), (256, 4), (256, 6), (512, 3), (512, 4)):
                if tpr * vpt < nvec or tpr * vpt >= 2 * nvec + 256:
                    continue
                L.pq_debug_set_quant_config(tpr, vpt)
                res.append((f"vec_{tpr}x{vpt}", timed(lambda i: F.quantize_act(xs[i % nb], out=out))))
            L.pq_debug_set_quant_config(0, 0)
            L.pq_debug_set_quant_staged(1)
            us = timed(lambda i: F.quantize_act(xs[i % nb], out=out))
            q2, s2 = F.quantize_act(xs[0])
            same = bool(torch.equal(q2, ref_q) and torch.equal(s2, ref_s))
            res.append(("staged", us))
            L.pq_debug_set_quant_staged(0)
            res.append(("auto", timed(lambda i: F.quantize_act(xs[i % nb], out=out))))
            line = "  ".join(f"{n}={u:.2f}us({byt / u / 1e3 / HBM:.2f})" for n, u in res)
            print(f"{str(dt)[6:]} {M}x{K} {byt / 1e6:.1f}MB staged_bit_identical={same}: {line}", flush=True)
            del xs


if __name__ == "__main__":
    main()
