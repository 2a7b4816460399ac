"""torchrun worker for test_sharded_module_nccl_bit_identical: column-parallel result over NCCL
must be bit-identical to the single-GPU result on every rank."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import protoquant_b200 as pq  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    torch.manual_seed(0)
    ok = True
    for (N, K, M) in ((28672, 8192, 256), (8192, 28672, 64), (1000, 512, 33)):
        lin = torch.nn.Linear(K, N).to(torch.bfloat16).cuda()
        m = pq.DynamicQuantLinear.from_float(lin)
        x = torch.randn(M, K, dtype=torch.bfloat16, device="cuda")
        full = m(x)
        for fused in (False, True):
            sh = pq.ShardedDynamicQuantLinear(m.qweight, m.weight_scale, m.bias, fused=fused)
            for _ in range(3):                      # repeated forwards exercise the double buffering
                y = sh(x)
                same = torch.equal(y, full)
                ok = ok and same
                if not same and rank == 0:
                    print(f"MISMATCH N={N} K={K} M={M} fused={fused}")
            if rank == 0:
                print(f"N={N} K={K} M={M} fused={fused} -> {sh.fused}")
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED_OK" if t.item() == 1 else "SHARDED_MISMATCH")
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1 else 1)


if __name__ == "__main__":
    main()
