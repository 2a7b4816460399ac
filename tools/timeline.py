"""Dump the in-kernel %globaltimer timeline of one GEMM launch.  usage: timeline.py M N K cfg sk"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import protoquant_b200 as pq
M, N, K, cfg, sk = (int(v) for v in sys.argv[1:6])
pq.lib().pq_debug_set_gemm_config(cfg); pq.lib().pq_debug_set_streamk(sk)
a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device="cuda")
b = torch.randint(-128, 128, (N, K), dtype=torch.int8, device="cuda")
sx = torch.rand(M, device="cuda"); sw = torch.rand(N, device="cuda")
y = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
for _ in range(5): pq.qgemm(a, sx, b, sw, None, torch.bfloat16, out=y)
torch.cuda.synchronize()
tl = torch.zeros(148 * 40, dtype=torch.int64, device="cuda")
pq.lib().pq_debug_set_timeline(tl.data_ptr())
pq.qgemm(a, sx, b, sw, None, torch.bfloat16, out=y)
torch.cuda.synchronize()
pq.lib().pq_debug_set_timeline(None)
t = tl.cpu().view(148, 40)
nz = t[:, 0] > 0
t0 = t[nz][:, 0].min().item()
names = ["start", "setup", "tma1", "tmaN", "mma1", "mmaN", "fin", "end"]
mhz = ((t[nz][:, 33] - t[nz][:, 32]).double() / (t[nz][:, 6] - t[nz][:, 1]).double().clamp(min=1) * 1e3).median().item()
print(f"M={M} N={N} K={K} cfg={cfg} sk={sk}: CTAs {int(nz.sum())}, span {(t[nz][:, 7].max().item() - t0) / 1e3:.1f} us, SM clock during the launch {mhz:.0f} MHz (clock64 / globaltimer)")
for cta in list(range(0, int(nz.sum()), max(1, int(nz.sum()) // 10)))[:12]:
    row = t[cta]
    base = " ".join(f"{n}={(row[i].item() - t0) / 1e3:6.1f}" if row[i] > 0 else f"{n}=   -  " for i, n in enumerate(names))
    segs = []
    for s_ in range(5):
        v = [row[8 + s_ * 4 + k].item() for k in range(4)]
        if any(v):
            segs.append("[" + " ".join(f"{(x - t0) / 1e3:5.1f}" if x else "  -  " for x in v) + f" r{row[28 + (s_ & 3)].item()}]")
    print(f"cta{cta:3d} {base}  segs(tfull,ticket,cwait,done): {' '.join(segs)}")
