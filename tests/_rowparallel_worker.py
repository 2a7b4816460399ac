"""torchrun worker for test_row_parallel_module_nvlink_bit_identical: the K-split module (fused GEMM +
reduce-scatter + all-gather over symmetric memory, and the int32 all-reduce fallback) must be bit-identical to
the single-GPU module on every rank; prints timings of the Llama-70B down projection."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import protoquant_b200 as pq  # noqa: E402


def graph_of(fn):
    """Capture TWO calls of `fn` (one per half of the modules' double buffers) into a CUDA graph."""
    fn(); fn()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn(); fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    outs = []
    with torch.cuda.graph(g):
        outs.append(fn())
        outs.append(fn())
    return g, outs


def timed_graph(g, reps=10):
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / (2 * reps)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item() * 1e3


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    torch.manual_seed(0)
    ok = True
    for (N, K, M) in ((8192, 28672, 256), (4096, 11008, 2048), (1000, 512, 33), (8192, 28672, 16)):
        lin = torch.nn.Linear(K, N).to(torch.bfloat16).cuda()
        m = pq.DynamicQuantLinear.from_float(lin)
        x = torch.randn(M, K, dtype=torch.bfloat16, device="cuda")
        x[0, K - 1] = 25.0
        full = m(x)
        for fused in (False, True):
            for sharded_in in (False, True):
                for gather in (True, False):
                    rp = pq.RowParallelDynamicQuantLinear(m.qweight, m.weight_scale, m.bias, fused=fused,
                                                          input_is_sharded=sharded_in, gather_output=gather)
                    xin = x[:, rp.k_lo:rp.k_hi].contiguous() if sharded_in else x
                    want = full if gather else full[:, rp.n_lo:rp.n_hi]
                    for _ in range(3):                      # repeated forwards exercise the double buffering
                        y = rp(xin)
                        same = torch.equal(y, want)
                        ok = ok and same
                        if not same:
                            print(f"rank {rank} MISMATCH N={N} K={K} M={M} fused={fused} sharded_in={sharded_in} gather={gather}")
            if rank == 0:
                print(f"N={N} K={K} M={M} fused={fused} -> {rp.fused}")
    # the whole tensor-parallel gated MLP (column-parallel gate/up without a gather -> silu*up on the local slice ->
    # row-parallel down): bit-identical to the three modules chained on one GPU
    from protoquant_b200 import functional as F
    for (H, I, M) in ((4096, 11008, 512), (8192, 28672, 64), (512, 1000 * 16, 33)):
        gate = pq.DynamicQuantLinear.from_float(torch.nn.Linear(H, I, bias=False).to(torch.bfloat16).cuda())
        up = pq.DynamicQuantLinear.from_float(torch.nn.Linear(H, I, bias=False).to(torch.bfloat16).cuda())
        down = pq.DynamicQuantLinear.from_float(torch.nn.Linear(I, H, bias=True).to(torch.bfloat16).cuda())
        x = torch.randn(M, H, dtype=torch.bfloat16, device="cuda")
        want = down(F.act_mul(gate(x), up(x), "silu"))
        mlp = pq.ParallelGatedMLP(gate, up, down)
        for _ in range(3):
            same = torch.equal(mlp(x), want)
            ok = ok and same
            if not same:
                print(f"rank {rank} MLP MISMATCH H={H} I={I} M={M}")
        if rank == 0:
            print(f"MLP H={H} I={I} M={M} fused={mlp.down.fused}")
        # the forward is a fixed launch sequence (no NCCL, no host sync): capture it and replay it
        g, outs = graph_of(lambda: mlp(x))
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        same = all(torch.equal(o, want) for o in outs)
        ok = ok and same
        if rank == 0:
            print(f"MLP H={H} I={I} M={M} cuda_graph_replay bit_identical={same}")
        del g, outs
        if (H, I) == (8192, 28672):
            x2 = torch.randn(2048, H, dtype=torch.bfloat16, device="cuda")
            g2, _ = graph_of(lambda: mlp(x2))
            tg = timed_graph(g2)
            if rank == 0:
                print(f"TIMING llama70b_mlp 8192/28672 M=2048 world={dist.get_world_size()} tensor_parallel_cuda_graph: {tg:.1f} us")
            del g2
            for name, fn in (("one_gpu_chain", lambda: down(F.act_mul(gate(x2), up(x2), "silu"))), ("tensor_parallel", lambda: mlp(x2))):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                dist.barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(10):
                    fn()
                b.record()
                torch.cuda.synchronize()
                t = torch.tensor([a.elapsed_time(b) / 10], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                if rank == 0:
                    print(f"TIMING llama70b_mlp 8192/28672 M=2048 world={dist.get_world_size()} {name}: {t.item()*1e3:.1f} us")
    # Hugging Face's own LlamaMLP converted in place: same bits as this package's single-GPU chain, close to the float module
    try:
        from transformers.models.llama.modeling_llama import LlamaConfig, LlamaMLP
        import copy
        hf = LlamaMLP(LlamaConfig(hidden_size=1024, intermediate_size=2816, num_attention_heads=8, num_hidden_layers=1, vocab_size=64))
        hf = hf.to(torch.bfloat16).cuda().eval()
        xh = torch.randn(96, 1024, dtype=torch.bfloat16, device="cuda")
        with torch.no_grad():
            y_float = hf(xh)
            holder = torch.nn.ModuleDict({"mlp": copy.deepcopy(hf)})
            n_rep = pq.parallelize_gated_mlps(holder)
            y_tp = holder["mlp"](xh)
            g1, u1, d1 = (pq.DynamicQuantLinear.from_float(getattr(hf, n)) for n in ("gate_proj", "up_proj", "down_proj"))
            y_one = d1(F.act_mul(g1(xh), u1(xh), "silu"))
        same = n_rep == 1 and isinstance(holder["mlp"], pq.ParallelGatedMLP) and torch.equal(y_tp, y_one)
        close = (y_tp.float() - y_float.float()).abs().max().item() < 0.08 * y_float.float().abs().max().item() + 1e-3
        ok = ok and same and close
        if rank == 0:
            print(f"HF LlamaMLP -> ParallelGatedMLP: replaced={n_rep} bit_identical_to_one_gpu_chain={same} close_to_float={close}")
    except ImportError:
        if rank == 0:
            print("transformers not importable: HF drop-in check skipped")
    # timing: Llama-70B down projection 28672 -> 8192 at 2048 tokens
    N, K, M = 8192, 28672, 2048
    lin = torch.nn.Linear(K, N, bias=False).to(torch.bfloat16).cuda()
    m = pq.DynamicQuantLinear.from_float(lin)
    x = torch.randn(M, K, dtype=torch.bfloat16, device="cuda")
    mods = {"replicated": m,
            "row_parallel_int32_allreduce": pq.RowParallelDynamicQuantLinear(m.qweight, m.weight_scale, None, fused=False),
            "row_parallel_fused": pq.RowParallelDynamicQuantLinear(m.qweight, m.weight_scale, None, fused=True),
            "row_parallel_fused_sharded_in_scattered_out": pq.RowParallelDynamicQuantLinear(
                m.qweight, m.weight_scale, None, fused=True, input_is_sharded=True, gather_output=False)}
    for name, mod in mods.items():
        xin = x[:, mod.k_lo:mod.k_hi].contiguous() if getattr(mod, "input_is_sharded", False) else x
        for _ in range(3):
            mod(xin)
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            mod(xin)
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / 20], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"TIMING down_proj 28672->8192 M=2048 world={dist.get_world_size()} {name}: {t.item()*1e3:.1f} us")
        if name.startswith("row_parallel_fused"):
            g, outs = graph_of(lambda: mod(xin))
            tg = timed_graph(g)
            want = m(x) if mod.gather_output else m(x)[:, mod.n_lo:mod.n_hi]
            same = all(torch.equal(o, want) for o in outs)
            ok = ok and same
            if rank == 0:
                print(f"TIMING down_proj 28672->8192 M=2048 world={dist.get_world_size()} {name}_cuda_graph: {tg:.1f} us bit_identical={same}")
            del g, outs
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("ROWPARALLEL_OK" if t.item() == 1 else "ROWPARALLEL_MISMATCH")
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1 else 1)


if __name__ == "__main__":
    main()
