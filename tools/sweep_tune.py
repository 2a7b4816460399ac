"""Tile-config x stream-K tuning sweep for the small/medium-M regime.  CSV: M,N,K,cfg,sk,us
usage: sweep_tune.py  (env PQ_MS, PQ_NKS override the grids)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import protoquant_b200 as pq
Ms = [int(v) for v in os.environ.get("PQ_MS", "128,256,512,1024,1536").split(",")]
NKs = [int(v) for v in os.environ.get("PQ_NKS", "1024,2048,4096,8192,16384").split(",")]
CFGS = [int(v) for v in os.environ.get("PQ_CFGS", "0,1,2,3,4").split(",")]
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
    return best * 1e3

print("M,N,K,cfg,sk,us")
for K in NKs:
    for N in NKs:
        b = torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev)
        sw = torch.rand(N, device=dev)
        for M in Ms:
            a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
            sx = torch.rand(M, device=dev)
            y = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
            for cfg in CFGS:
                for sk in (0, 1):
                    if sk == 1 and cfg in (1, 4) and False:
                        continue
                    pq.lib().pq_debug_set_gemm_config(cfg)
                    pq.lib().pq_debug_set_streamk(sk)
                    try:
                        t = time_graph(lambda: pq.qgemm(a, sx, b, sw, None, torch.bfloat16, out=y), 10)
                        print(f"{M},{N},{K},{cfg},{sk},{t:.2f}", flush=True)
                    except Exception as ex:
                        print(f"{M},{N},{K},{cfg},{sk},nan", flush=True)
                        torch.cuda.synchronize()
            pq.lib().pq_debug_set_gemm_config(-1)
            pq.lib().pq_debug_set_streamk(-1)
            t = time_graph(lambda: pq.qgemm(a, sx, b, sw, None, torch.bfloat16, out=y), 10)
            print(f"{M},{N},{K},-1,-1,{t:.2f}", flush=True)
            try:
                bt = b.t()
                t = time_graph(lambda: torch._int_mm(a, bt), 10)
                print(f"{M},{N},{K},int_mm,0,{t:.2f}", flush=True)
            except Exception:
                pass
