"""Sustained GEMM loop with NVML clock/power sampling.  usage: clock_probe.py M N K cfg seconds"""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pynvml
import protoquant_b200 as pq
M, N, K, cfg = (int(v) for v in sys.argv[1:5]); secs = float(sys.argv[5])
pq.lib().pq_debug_set_gemm_config(cfg)
pq.lib().pq_debug_set_epilogue(int(os.environ.get("PQ_EPI", "0")))
a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device="cuda")
b = torch.randint(-128, 128, (N, K), dtype=torch.int8, device="cuda")
sx = torch.rand(M, device="cuda"); sw = torch.rand(N, device="cuda")
y = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples = []; stop = False
def samp():
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                        pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        time.sleep(0.02)
t = threading.Thread(target=samp); t.start()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): pq.qgemm(a, sx, b, sw, None, torch.bfloat16, out=y)
torch.cuda.synchronize()
with torch.cuda.graph(g):
    for _ in range(20): pq.qgemm(a, sx, b, sw, None, torch.bfloat16, out=y)
torch.cuda.synchronize()
t0 = time.time(); n = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < secs:
    g.replay(); n += 20
    if n % 200 == 0: torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
stop = True; t.join()
ms = e0.elapsed_time(e1) / n
half = samples[len(samples)//2:]
clk = sorted(x[0] for x in half); pw = sorted(x[1] for x in half); rs = 0
for x in half: rs |= x[2]
print(f"M={M} N={N} K={K} cfg={cfg} epi={os.environ.get('PQ_EPI', '0')}: {ms*1e3:.1f} us/launch {2*M*N*K/ms/1e9:.0f} TOPS sustained over {secs}s; SM clock median {clk[len(clk)//2]} MHz (min {clk[0]}), power median {pw[len(pw)//2]:.0f} W, reasons 0x{rs:x}, samples {len(half)}")
