"""A few eager cuBLASLt int8 GEMMs (torch._int_mm), for `ncu` captures next to tools/one_gemm.py.  usage: one_intmm.py M N K"""
import sys
import torch
M, N, K = (int(v) for v in sys.argv[1:4])
a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device="cuda")
b = torch.randint(-128, 128, (N, K), dtype=torch.int8, device="cuda")
for _ in range(6):
    c = torch._int_mm(a, b.t())
torch.cuda.synchronize()
print(c.shape)
