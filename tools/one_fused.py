"""Launch the fused producer kernels a few times at streaming size (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from protoquant_b200 import functional as F
M, K = 131072, 4096
x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = torch.ones(K, dtype=torch.bfloat16, device="cuda")
q = F.alloc_q(M, K, "cuda"); s = torch.empty(M, dtype=torch.float32, device="cuda")
for _ in range(3):
    F.rmsnorm_quant(x, w, out=(q, s))
    F.quantize_act(x, out=(q, s))
torch.cuda.synchronize()
