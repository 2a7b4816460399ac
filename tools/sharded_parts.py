"""torchrun worker: where the time of a column-sharded forward goes (Llama-70B up projection shard).
usage: torchrun --nproc-per-node N tools/sharded_parts.py [M ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
import protoquant_b200 as pq
from protoquant_b200 import functional as F

rank = int(os.environ["RANK"]); local = int(os.environ.get("LOCAL_RANK", rank)); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
Ms = [int(v) for v in sys.argv[1:]] or [2048, 256, 16]
K, N = 8192, 28672
g = torch.Generator(device=dev).manual_seed(7)
wq = torch.randint(-127, 128, (N, K), dtype=torch.int8, device=dev, generator=g)
sw = torch.rand(N, device=dev, generator=g) * 1e-3


def allmax(v):
    t = torch.tensor([v], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def time_graph(fn, reps=10, inner=4):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(inner):
            fn()
    gr.replay()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        gr.replay()
    b.record(); torch.cuda.synchronize()
    return allmax(a.elapsed_time(b) / (reps * inner)) * 1e3


sh = pq.ShardedDynamicQuantLinear(wq, sw, None, fused=None)
for M in Ms:
    x = torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16)
    sh(x); sh(x)
    t, h = sh._symm[2][0]
    esz, ld, off = 2, sh.world * sh.per, sh.rank * sh.per * 2
    dests_all = [int(p) + off for p in h.buffer_ptrs]
    dest_local = [dests_all[rank]]
    xq_ws, sx_ws = sh._workspace(M, dev)
    y_local = torch.empty(M, sh.per, dtype=torch.bfloat16, device=dev)
    res = {}
    res["barrier_only"] = time_graph(lambda: h.barrier())
    res["quant_only"] = time_graph(lambda: F.quantize_act(x, out=(xq_ws[:M, :K], sx_ws[:M])))
    res["gemm_local_plain"] = time_graph(lambda: F.qgemm(xq_ws[:M, :K], sx_ws[:M], sh.qweight, sh.weight_scale, None, torch.bfloat16, out=y_local))
    res["quant+gemm_1dest(symm,ld=N)"] = time_graph(lambda: F.qlinear_multi_into(x, sh.qweight_storage, K, sh.weight_scale, None, dest_local, ld, torch.bfloat16, xq_ws, sx_ws))
    res["quant+gemm_all_dests"] = time_graph(lambda: F.qlinear_multi_into(x, sh.qweight_storage, K, sh.weight_scale, None, dests_all, ld, torch.bfloat16, xq_ws, sx_ws))
    def fwd():
        F.qlinear_multi_into(x, sh.qweight_storage, K, sh.weight_scale, None, dests_all, ld, torch.bfloat16, xq_ws, sx_ws)
        h.barrier()
    res["quant+gemm_all_dests+barrier"] = time_graph(fwd)
    res["module_forward"] = time_graph(lambda: sh(x), inner=4)
    if rank == 0:
        print(f"M={M} world={world} shard N={sh.per}: " + "  ".join(f"{k}={v:.1f}us" for k, v in res.items()), flush=True)
dist.barrier()
dist.destroy_process_group()
