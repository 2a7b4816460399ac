"""torchrun worker: column-sharded Llama-70B up projection (8192 -> 28672), fused all-gather variants side by side.
  TMA stores per destination (default) | LSU 256-byte stores per destination | multimem.st to the NVSwitch multicast address
usage: torchrun --nproc-per-node N tools/sharded_ab.py [M ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import protoquant_b200 as pq

rank = int(os.environ["RANK"]); local = int(os.environ.get("LOCAL_RANK", rank)); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
Ms = [int(v) for v in sys.argv[1:]] or [2048, 16]
K, N = 8192, 28672
g = torch.Generator(device=dev).manual_seed(7)
wq = torch.randint(-127, 128, (N, K), dtype=torch.int8, device=dev, generator=g)
sw = torch.rand(N, device=dev, generator=g) * 1e-3
full = pq.DynamicQuantLinear(K, N, bias=False, device=dev)
full.qweight_storage[:, :K].copy_(wq); full.weight_scale.copy_(sw)


def allmax(v):
    t = torch.tensor([v], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def time_it(fn, reps=20, graph=True):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    run, mode = fn, "eager"
    if graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn(); fn()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                fn(); fn()
            run, mode, reps = gr.replay, "graph", reps // 2
        except Exception as ex:
            if rank == 0:
                print("  (graph capture failed:", repr(ex)[:120], ")")
            torch.cuda.synchronize()
    for _ in range(2):
        run()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        run()
    b.record(); torch.cuda.synchronize()
    per = a.elapsed_time(b) / reps / (2 if mode == "graph" else 1)
    return allmax(per), mode


for M in Ms:
    x = torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16)
    ref = full(x)
    t_rep, _ = time_it(lambda: full(x))
    if rank == 0:
        print(f"M={M} world={world}: replicated {t_rep*1e3:.1f} us")
    for label, env_mc, tma, bn in (("lsu_staged_bn256", None, 0, 256), ("lsu_staged_bn224", None, 0, 224), ("lsu_staged_bn128", None, 0, 128),
                                   ("lsu_staged_model", None, 0, 0), ("multimem_st_model", "1", 0, 0)):
        if env_mc:
            os.environ["PQ_USE_MULTICAST"] = env_mc
        else:
            os.environ.pop("PQ_USE_MULTICAST", None)
        pq.lib().pq_debug_set_multi_tma(tma)
        pq.lib().pq_debug_set_multi_bn(bn)
        sh = pq.ShardedDynamicQuantLinear(wq, sw, None, fused=None)
        try:
            ok = all(torch.equal(sh(x), ref) for _ in range(3))
            for graph in (False, True):
                t, mode = time_it(lambda: sh(x), graph=graph)
                if rank == 0:
                    print(f"  {label:18s} {mode:6s}: {t*1e3:8.1f} us  {2*M*N*K/t/1e9:7.0f} TOPS  bit_identical={ok} fused={sh.fused}")
        except Exception as ex:
            print(f"  rank {rank} {label}: FAILED {ex!r}"[:300])
        del sh
    pq.lib().pq_debug_set_multi_tma(0)
    pq.lib().pq_debug_set_multi_bn(0)
    os.environ.pop("PQ_USE_MULTICAST", None)
    shn = pq.ShardedDynamicQuantLinear(wq, sw, None, fused=False)
    t, mode = time_it(lambda: shn(x), graph=False)
    if rank == 0:
        print(f"  nccl_allgather eager : {t*1e3:8.1f} us")
dist.barrier()
dist.destroy_process_group()
