"""Round-2 tile-config sweep on the headline shapes (fused q/k/v and gate/up included) next to cuBLASLt int8.
usage: sweep_r2.py [hot|cold]   (cold: 8 distinct weight sets cycled so every launch streams its weights from DRAM)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import protoquant_b200 as pq
mode = sys.argv[1] if len(sys.argv) > 1 else "hot"
dev = "cuda"
SHAPES = [(2048, 4096, 4096), (2048, 12288, 4096), (2048, 11008, 4096), (2048, 22016, 4096), (2048, 4096, 11008),
          (4096, 3072, 768), (4096, 768, 3072), (4096, 2304, 768)]
CFGS = [-1, 0, 1, 8, 9, 10, 11, 12, 13]
if os.environ.get("PQ_CFGS"):
    CFGS = [int(v) for v in os.environ["PQ_CFGS"].split(",")]
if os.environ.get("PQ_SHAPES"):
    SHAPES = [tuple(int(v) for v in sh.split("x")) for sh in os.environ["PQ_SHAPES"].split(",")]
SKIP_LIBS = bool(os.environ.get("PQ_SKIP_LIBS"))

def time_graph(fn, iters):
    fn(0); torch.cuda.synchronize()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i)
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    return best * 1e3

print("mode,M,N,K,cfg,us,tops")
for M, N, K in SHAPES:
    nset = 1 if mode == "hot" else max(2, min(8, int(400e6 // (N * K))))
    bs = [torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev) for _ in range(nset)]
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
    sx = torch.rand(M, device=dev); sw = torch.rand(N, device=dev)
    y = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    ops = 2.0 * M * N * K
    iters = 8
    for cfg in CFGS:
        pq.lib().pq_debug_set_gemm_config(cfg)
        try:
            t = time_graph(lambda i: pq.qgemm(a, sx, bs[i % nset], sw, None, torch.bfloat16, out=y), iters)
            print(f"{mode},{M},{N},{K},{cfg},{t:.2f},{ops/t/1e6:.0f}", flush=True)
        except Exception as ex:
            print(f"{mode},{M},{N},{K},{cfg},error,{ex!r}"[:200], flush=True)
            torch.cuda.synchronize()
    pq.lib().pq_debug_set_gemm_config(-1)
    if SKIP_LIBS:
        del bs, a
        continue
    bts = [b.t() for b in bs]
    t = time_graph(lambda i: torch._int_mm(a, bts[i % nset]), iters)
    print(f"{mode},{M},{N},{K},int_mm,{t:.2f},{ops/t/1e6:.0f}", flush=True)
    xb = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    wbs = [torch.randn(N, K, device=dev, dtype=torch.bfloat16) for _ in range(nset)]
    t = time_graph(lambda i: torch.matmul(xb, wbs[i % nset].t()), iters)
    print(f"{mode},{M},{N},{K},bf16,{t:.2f},{ops/t/1e6:.0f}", flush=True)
    del bs, bts, wbs, a, xb
