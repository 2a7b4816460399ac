"""torchrun worker: what NVLink store bandwidth can SM-issued traffic reach on this box, for the traffic pattern of the
fused all-gather (every rank pushes its [2048 x N/world] bf16 column slice into all peers at once)?
  torchrun --nproc-per-node N tools/nvlink_probe.py        (build tools/libnvlink_probe.so first, see nvlink_probe.cu)"""
import ctypes
import os
import sys

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

HERE = os.path.dirname(os.path.abspath(__file__))
rank = int(os.environ["RANK"]); local = int(os.environ.get("LOCAL_RANK", rank)); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = ctypes.CDLL(os.path.join(HERE, "libnvlink_probe.so"))
lib.probe_launch.restype = ctypes.c_int
lib.probe_launch.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int,
                             ctypes.c_int, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]

M, N = 2048, 28672
per = N // world
out = symm_mem.empty((M, N), dtype=torch.bfloat16, device=dev)
h = symm_mem.rendezvous(out, dist.group.WORLD)
src = torch.full((M, N), float(rank + 1), dtype=torch.bfloat16, device=dev)      # local source, same layout
off = rank * per * 2
ld = N * 2
seg = per * 2
peers = [int(p) + off for i, p in enumerate(h.buffer_ptrs) if i != rank]
everyone = [int(p) + off for p in h.buffer_ptrs]
mc = int(getattr(h, "multicast_ptr", 0) or 0)
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def allmax(v):
    t = torch.tensor([v], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def run(mode, dests, ctas, threads, reps=20):
    arr = (ctypes.c_void_p * len(dests))(*[ctypes.c_void_p(d) for d in dests])

    def go():
        rc = lib.probe_launch(mode, ctypes.c_void_p(src.data_ptr() + off), arr, len(dests), M, seg, ld, ctas, threads, stream)
        assert rc == 0, rc
    for _ in range(3):
        go()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):     # back to back, no cross-rank barrier inside the loop: all ranks run the same loop at once
        go()
    b.record()
    torch.cuda.synchronize()
    return allmax(a.elapsed_time(b) / reps * 1e3)


def check():
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ok = all(bool((out[:, r * per:(r + 1) * per] == float(r + 1)).all()) for r in range(world) if r != rank)
    out.zero_()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    return ok


if rank == 0:
    print(f"world={world} slice {M} x {per} bf16 = {M * seg / 1e6:.1f} MB per rank; egress to peers {(world - 1) * M * seg / 1e6:.1f} MB; "
          f"ingress {(world - 1) * M * seg / 1e6:.1f} MB; multicast_ptr={'yes' if mc else 'no'}", flush=True)
t0 = None
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); dist.barrier(); a.record()
for _ in range(10):
    h.barrier()
b.record(); torch.cuda.synchronize()
if rank == 0:
    print(f"barrier alone: {a.elapsed_time(b) / 10 * 1e3:.1f} us", flush=True)
gb = (world - 1) * M * seg / 1e9
for mode, name, dests in ((0, "lsu_st_peers", peers), (0, "lsu_st_peers+self", everyone), (1, "multimem_st", [mc + off] if mc else None),
                          (2, "bulk_peers", peers)):
    if dests is None:
        continue
    for ctas, threads in ((148, 256), (148, 1024), (296, 512), (592, 256), (74, 1024), (32, 1024), (16, 1024), (8, 1024)):
        if mode == 2 and threads != 256 and not (ctas in (296, 592)):
            continue
        try:
            us = run(mode, dests, ctas, threads)
            ok = check()
        except Exception as ex:
            print(f"rank {rank} {name} {ctas}x{threads}: FAILED {ex!r}"[:200], flush=True)
            break
        if rank == 0:
            print(f"  {name:18s} ctas={ctas:4d} threads={threads:5d}: {us:8.1f} us  {gb / us * 1e6:7.1f} GB/s per direction per GPU  data_ok={ok}", flush=True)
# one-to-one: rank r -> rank r+1 only (pairwise peak for SM stores)
nxt = [int(h.buffer_ptrs[(rank + 1) % world]) + off]
for ctas, threads in ((148, 512), (296, 512)):
    us = run(0, nxt, ctas, threads)
    if rank == 0:
        print(f"  ring lsu_st 1 peer   ctas={ctas:4d} threads={threads:5d}: {us:8.1f} us  {M * seg / 1e9 / us * 1e6:7.1f} GB/s", flush=True)
dist.barrier()
dist.destroy_process_group()
