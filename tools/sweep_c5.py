"""BASELINE.json configs[4]: int8 qlinear GEMM sweep M = 1..8192 x N,K in {1024..16384}, next to cuBLASLt int8
(torch._int_mm, int32 output, no epilogue) and bf16 cuBLAS (torch.matmul) on the same shapes.
Writes CSV to stdout.  usage: sweep_c5.py [quick]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import protoquant_b200 as pq
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
Ms = [1, 16, 64, 128, 256, 512, 1024, 2048, 4096, 8192]
NK = [1024, 2048, 4096, 8192, 16384]
if quick:
    Ms, NK = [16, 512, 4096], [1024, 4096]
dev = "cuda"

def time_graph(fn, iters):
    fn(); torch.cuda.synchronize()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    return best * 1e3   # us

print("M,N,K,ours_us,ours_tops,int_mm_us,int_mm_tops,bf16_us,bf16_tflops,ours_vs_int_mm,weight_gbs")
for K in NK:
    for N in NK:
        b = torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev)
        bt = b.t()          # [K, N] column-major view: what _int_mm wants for a K-major weight
        wb = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
        sw = torch.rand(N, device=dev)
        for M in Ms:
            a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
            sx = torch.rand(M, device=dev)
            y = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
            ops = 2.0 * M * N * K
            iters = 5 if ops > 2e11 else 20
            t = time_graph(lambda: pq.qgemm(a, sx, b, sw, None, torch.bfloat16, out=y), iters)
            ti = float("nan")
            if M > 16:
                try:
                    ti = time_graph(lambda: torch._int_mm(a, bt), iters)
                except Exception:
                    ti = float("nan")
            xb = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
            tb = time_graph(lambda: torch.matmul(xb, wb.t()), iters)
            print(f"{M},{N},{K},{t:.2f},{ops/t/1e6:.0f},{ti:.2f},{ops/ti/1e6:.0f},{tb:.2f},{ops/tb/1e6:.0f},{ti/t:.3f},{N*K/t/1e3:.0f}", flush=True)
            del a, xb
        del b, wb
