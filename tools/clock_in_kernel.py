"""SM clock DURING a GEMM launch, measured in the kernel (clock64 / %globaltimer of the debug timeline), for a CUDA graph
of `iters` back-to-back launches (the stamps of the last launch survive).  usage: clock_in_kernel.py M N K cfg iters"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import protoquant_b200 as pq
M, N, K, cfg, iters = (int(v) for v in sys.argv[1:6])
pq.lib().pq_debug_set_gemm_config(cfg)
a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device="cuda")
b = torch.randint(-128, 128, (N, K), dtype=torch.int8, device="cuda")
sx = torch.rand(M, device="cuda"); sw = torch.rand(N, device="cuda")
y = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
tl = torch.zeros(148 * 40, dtype=torch.int64, device="cuda")
pq.lib().pq_debug_set_timeline(tl.data_ptr())
def run():
    for _ in range(iters):
        pq.qgemm(a, sx, b, sw, None, torch.bfloat16, out=y)
run(); torch.cuda.synchronize()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    run()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    run()
pq.lib().pq_debug_set_timeline(None)
g.replay(); torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    tl.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / iters)
t = tl.cpu().view(148, 40)
nz = t[:, 0] > 0
mhz = ((t[nz][:, 33] - t[nz][:, 32]).double() / (t[nz][:, 6] - t[nz][:, 1]).double().clamp(min=1) * 1e3)
span = (t[nz][:, 7].max().item() - t[nz][:, 0].min().item()) / 1e3
setup = (t[nz][:, 1] - t[nz][:, 0]).double().median().item() / 1e3
first_mma = (t[nz][:, 4] - t[nz][:, 0]).double().median().item() / 1e3
tail = (t[nz][:, 7].max().item() - t[nz][:, 5].double().median().item()) / 1e3
ops = 2.0 * M * N * K
print(f"M={M} N={N} K={K} cfg={cfg} iters={iters}: {best*1e3:.1f} us/launch {ops/best/1e9:.0f} TOPS | last launch: span {span:.1f} us, "
      f"setup {setup:.1f} us, first MMA at {first_mma:.1f} us, last-MMA-issue(median) to end {tail:.1f} us, "
      f"SM clock median {mhz.median().item():.0f} MHz (min {mhz.min().item():.0f}, max {mhz.max().item():.0f}) -> "
      f"tensor-peak at that clock {8192*2*148*mhz.median().item()/1e6:.0f} TOPS, utilisation {ops/best/1e9/(8192*2*148*mhz.median().item()/1e6):.3f}")
