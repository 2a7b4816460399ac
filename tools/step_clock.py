"""How the headline step's time depends on run length (burst vs sustained) with NVML clock/power samples."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pynvml
import protoquant_b200 as pq
from protoquant_b200 import functional as F
import bench
dev = torch.device("cuda", 0)
torch.manual_seed(0)
mods = {}
for name, k, n, src in bench.LINEARS:
    lin = torch.nn.Linear(k, n, bias=True).to(torch.bfloat16).to(dev)
    mods[name] = pq.DynamicQuantLinear.from_float(lin)
M = bench.M_TOKENS
acts = {a: torch.randn(M, k, device=dev).to(torch.bfloat16) for a, k in bench.ACTS.items()}
ws = {a: (F.alloc_q(M, k, dev), torch.empty(M, dtype=torch.float32, device=dev)) for a, k in bench.ACTS.items()}
outs = {name: torch.empty(M, n, dtype=torch.bfloat16, device=dev) for name, k, n, _ in bench.LINEARS}
def step():
    for name, k, n, src in bench.LINEARS:
        m = mods[name]
        F.qlinear_into(acts[src], m.qweight_storage, m.in_features, m.weight_scale, m.bias, outs[name], *ws[src])
step(); torch.cuda.synchronize()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    step()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step()
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples = []; stop = False
def samp():
    while not stop:
        samples.append((time.time(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                        pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        time.sleep(0.002)
t = threading.Thread(target=samp); t.start()
for reps in (1, 1, 5, 25, 100, 300, 1000, 5000):
    time.sleep(0.3)
    torch.cuda.synchronize()
    w0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    w1 = time.time()
    ms = e0.elapsed_time(e1) / reps
    sel = [x for x in samples if w0 <= x[0] <= w1]
    clk = sorted(x[1] for x in sel); pw = sorted(x[2] for x in sel); rs = 0
    for x in sel: rs |= x[3]
    info = f"clk med {clk[len(clk)//2]} min {clk[0]} MHz, power med {pw[len(pw)//2]:.0f} max {pw[-1]:.0f} W, reasons 0x{rs:x}, {len(sel)} samples" if sel else "no samples"
    print(f"{reps:5d} steps back to back: {ms*1e3:7.1f} us/step = {bench.OPS_PER_STEP/ms/1e9:6.0f} TOPS   ({info})")
stop = True; t.join()
