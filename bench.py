#!/usr/bin/env python
"""Benchmark of the dynamic-quantized int8 linear path (BASELINE.json `metric`).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N = 1 -- workload "llama7b_linears_2048tok" (BASELINE.json configs[2]): one "step" pushes one batch of 2048
  synthetic bf16 tokens through the seven linears of a Llama-7B block (q,k,v,o: 4096->4096; gate,up: 4096->11008;
  down: 11008->4096) the way `swap_linear` sets a block up: every DISTINCT activation is quantised once
  (x_attn for q/k/v, attn_out for o, x_mlp for gate/up, h_mlp for down) and linears that share an input run as one
  GEMM over their concatenated per-channel-quantised weights -- 4 act-quant launches + 4 tcgen05 GEMM launches with
  the fused dequant epilogue, outputs bit-identical to seven separate linears (tests/test_gpu_module.py).
    value   whole-job int8 TOPS = 2*M*sum(N*K) / step time, inputs resident in HBM, CUDA-graph replay.
    e2e     same metric through the public module API with HOST buffers: every step copies the four activation
            tensors from pinned host memory and brings the seven outputs back (copies inside the timed region).
    roofline  the tcgen05 GEMM (dominant kernel) against 2 x the measured cuBLAS bf16 peak, next to cuBLASLt int8
            (torch._int_mm) on the same shapes in the same run; act-quant against the measured HBM peak.
N > 1 -- workload "llama70b_up_proj_colsharded_2048tok" (BASELINE.json configs[3], north_star (4)): the Llama-70B
  up projection 8192 -> 28672 column-sharded over the N GPUs, 2048 tokens, all-gather of the output slices
  INCLUDED (fused into the GEMM epilogue: coalesced peer stores into every rank's symmetric output buffer over NVLink,
  one cross-rank barrier).  Total work is fixed -> "scaling": "strong".  value = 2*M*N*K / time (max over ranks);
  roofline.bound = "nvlink": bytes every rank must receive / time against the measured 770 GB/s peer bandwidth.
  The sharded output is compared bit for bit with the replicated layer on every rank; a mismatch fails the run.
--impl reference   times the oracle's CPU path (the reference checkout is absent, so the restatement stands in:
  oracle/protoquant_oracle.py) on all host threads, same workload and token count as N = 1.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_TOKENS = 2048
LINEARS = [  # name, K (in), N (out), which activation buffer feeds it
    ("q_proj", 4096, 4096, "x_attn"), ("k_proj", 4096, 4096, "x_attn"), ("v_proj", 4096, 4096, "x_attn"),
    ("o_proj", 4096, 4096, "attn_out"), ("gate_proj", 4096, 11008, "x_mlp"), ("up_proj", 4096, 11008, "x_mlp"),
    ("down_proj", 11008, 4096, "h_mlp"),
]
ACTS = {"x_attn": 4096, "attn_out": 4096, "x_mlp": 4096, "h_mlp": 11008}
# how swap_linear(fuse_shared_inputs=True) groups them: one act-quant + one GEMM per distinct activation
GROUPS = [("qkv_proj", ("q_proj", "k_proj", "v_proj"), "x_attn"), ("o_proj", ("o_proj",), "attn_out"),
          ("gate_up_proj", ("gate_proj", "up_proj"), "x_mlp"), ("down_proj", ("down_proj",), "h_mlp")]
NVLINK_PEER_GBS = 770.0   # measured peer-copy bandwidth per direction per GPU (B200_PROFILING.md)
NVLINK_A2A_GBS = 645.0    # measured here: SM-issued stores, 8 GPUs pushing to all peers at once (profiles/nvlink_probe_n8_r2.log)
S70_K, S70_N, S70_M = 8192, 28672, 2048
OPS_PER_STEP = sum(2 * M_TOKENS * n * k for _, k, n, _ in LINEARS)
NOMINAL_INT8_TOPS = 4500.0


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.stop_flag = threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def sample(self):
        nv = self.nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
            else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}
        for bit, n in names.items():
            if r & bit:
                self.reasons.add(n)

    def run(self):
        if self.nv is None:
            return
        while not self.stop_flag.is_set():
            try:
                self.sample()
            except Exception:
                break
            time.sleep(self.period)

    def result(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return rank, world, local


# ------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """CPU baseline arm: the oracle's torch-threaded restatement of the same workload as our arm at this --gpus
    (kind = "port": /root/reference holds no sources to compile or import).  N = 1: the 2048-token
    Llama-7B step; N > 1: the Llama-70B up projection at 2048 tokens (rank 0 alone runs it)."""
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import protoquant_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(0)
    if args.gpus > 1:
        workload, m_sample = "llama70b_up_proj_colsharded_2048tok", S70_M
        wq = torch.randint(-127, 128, (S70_N, S70_K), dtype=torch.int8, generator=g)
        layers = [(wq.t(), torch.rand(S70_N, generator=g) * 1e-3, None, "x")]
        acts = {"x": torch.randn(m_sample, S70_K, generator=g).to(torch.bfloat16)}
        ops = 2.0 * m_sample * S70_N * S70_K
        cfg = {"workload": workload, "tokens": m_sample, "layer": [S70_K, S70_N], "act_dtype": "bf16", "out_dtype": "bf16"}
        what = "the Llama-70B up projection (8192 -> 28672) on 2048 tokens"
        scaling = "strong"
    else:
        workload, m_sample = "llama7b_linears_2048tok", M_TOKENS
        layers = []
        for name, k, n, src in LINEARS:
            w = (torch.rand(n, k, generator=g) * 2 - 1) / k ** 0.5
            wq, sw = O.quantize_rowwise(w)
            layers.append((torch.from_numpy(wq).t(), torch.from_numpy(sw), torch.randn(n, generator=g), src))
        acts = {a: torch.randn(m_sample, k, generator=g).to(torch.bfloat16) for a, k in ACTS.items()}
        ops = float(OPS_PER_STEP)
        cfg = {"workload": workload, "tokens_per_gpu": m_sample, "act_dtype": "bf16", "out_dtype": "bf16",
               "linears": {l[0]: [l[1], l[2]] for l in LINEARS}}
        what = "the full 2048-token step, all 7 linears"
        scaling = "weak"

    def step():
        for wq_t, sw, b, src in layers:
            O.qlinear_torch_cpu(acts[src], wq_t, sw, b, torch.bfloat16)

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t_probe = time.perf_counter()
    step()
    t_probe = time.perf_counter() - t_probe
    steps = max(1, min(args.steps, 20, int(60.0 / max(t_probe, 1e-3))))    # keep the whole arm within about a minute
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    tops = ops / dt / 1e12
    sample = f"{what}, {steps} steps, torch CPU _int_mm on {cores} threads"
    line = {
        "impl": "reference", "metric": "int8_qlinear_tops", "value": tops, "unit": "TOPS", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "int8", "data": "synthetic",
        "config": cfg, "tokens_per_s": m_sample / dt,
        "cpu_baseline": {"value": tops, "unit": "TOPS", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": tops, "unit": "TOPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def cpu_baseline_leg(budget_s=12.0):
    """Same oracle path, timed beside the GPU numbers on rank 0 (bounded sample)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import protoquant_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m_sample = M_TOKENS
    g = torch.Generator().manual_seed(0)
    layers = []
    for name, k, n, src in LINEARS:
        wq = torch.randint(-127, 128, (n, k), dtype=torch.int8, generator=g)
        layers.append((wq.t(), torch.rand(n, generator=g) * 1e-3, torch.randn(n, generator=g), src))
    acts = {a: torch.randn(m_sample, k, generator=g).to(torch.bfloat16) for a, k in ACTS.items()}

    def step():
        for wq_t, sw, b, src in layers:
            O.qlinear_torch_cpu(acts[src], wq_t, sw, b, torch.bfloat16)

    step()
    n, t0 = 0, time.perf_counter()
    while True:
        step()
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 10:
            break
    dt = (time.perf_counter() - t0) / n
    tops = OPS_PER_STEP * m_sample / M_TOKENS / dt / 1e12
    return {"value": tops, "unit": "TOPS", "cores": cores, "kind": "port",
            "sample": f"the full {m_sample}-token step, all 7 linears, {n} steps, torch CPU _int_mm on {cores} threads",
            "tokens_per_s": m_sample / dt}


# ------------------------------------------------------------------------------ our arm
def capture(torch, dev, fn, no_graph=False):
    """Warm `fn` on a side stream and capture it into a CUDA graph; returns (callable, launch mode)."""
    fn()
    if no_graph:
        return fn, "eager"
    try:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            fn()
        return graph.replay, "cuda_graph_replay"
    except Exception as ex:  # capture unsupported: stay eager
        sys.stderr.write(f"bench: CUDA graph capture failed ({ex!r}); running eager\n")
        torch.cuda.synchronize()
        return fn, "eager"


def timed(torch, run, reps, warm=3):
    """CUDA-event time (ms) of `reps` calls of `run` on the current stream, after `warm` untimed calls."""
    for _ in range(warm):
        run()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        run()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import protoquant_b200 as pq
    from protoquant_b200 import functional as F

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; protoquant_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    ctx = {"torch": torch, "dist": dist, "pq": pq, "F": F, "rank": rank, "world": world, "local": local, "dev": dev,
           "peaks": peaks, "args": args}
    rc = run_sharded(ctx) if world > 1 else run_single(ctx)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return rc


class Llama7BBlockLinears:
    """The seven linears of a Llama-7B block, random-init, converted with the public swap helper."""

    def __init__(self, torch, pq, dev):
        from torch import nn

        class Block(nn.Module):
            def __init__(self):
                super().__init__()
                for name, k, n, _ in LINEARS:
                    setattr(self, name, nn.Linear(k, n, bias=True))

        blk = Block().to(torch.bfloat16).to(dev)
        self.block = pq.swap_linear(blk)      # fuse_shared_inputs=True: q/k/v and gate/up become one GEMM each
        b = self.block
        assert isinstance(b.q_proj, pq.SharedInputLinear) and isinstance(b.gate_proj, pq.SharedInputLinear)
        self.fused = {"qkv_proj": b.q_proj.fused, "o_proj": b.o_proj, "gate_up_proj": b.gate_proj.fused, "down_proj": b.down_proj}


def run_single(ctx):
    torch, pq, F, dev, peaks, args = ctx["torch"], ctx["pq"], ctx["F"], ctx["dev"], ctx["peaks"], ctx["args"]
    rank, world, local = ctx["rank"], ctx["world"], ctx["local"]
    torch.manual_seed(1234 + rank)

    # ---- load-time: random-init weights of the named architecture, quantised once (public swap helper) ----
    blk = Llama7BBlockLinears(torch, pq, dev)
    fused = blk.fused
    acts = {a: torch.randn(M_TOKENS, k, device=dev).to(torch.bfloat16) for a, k in ACTS.items()}
    xq_ws = {a: (F.alloc_q(M_TOKENS, k, dev), torch.empty(M_TOKENS, dtype=torch.float32, device=dev)) for a, k in ACTS.items()}
    outs = {g: torch.empty(M_TOKENS, fused[g].out_features, dtype=torch.bfloat16, device=dev) for g, _, _ in GROUPS}

    def step():
        # one pq_qlinear call (act-quant launch + GEMM launch) per distinct activation, on preallocated buffers so
        # that the step can be captured into a CUDA graph: exactly what the modules' forward does
        for g, members, src in GROUPS:
            m = fused[g]
            F.qlinear_into(acts[src], m.qweight_storage, m.in_features, m.weight_scale, m.bias, outs[g], *xq_ws[src])

    # round 1's step for comparison: one act-quant + one GEMM per linear (7 + 7 launches) on row slices of the same weights
    outs7 = {name: torch.empty(M_TOKENS, n, dtype=torch.bfloat16, device=dev) for name, k, n, _ in LINEARS}
    slices7 = {}
    for g, members, src in GROUPS:
        lo = 0
        for name in members:
            n = outs7[name].shape[1]
            m = fused[g]
            slices7[name] = (m.qweight_storage[lo:lo + n], m.in_features, m.weight_scale[lo:lo + n],
                             m.bias[lo:lo + n] if m.bias is not None else None, src)
            lo += n

    def step_unfused():
        for name, k, n, src in LINEARS:
            wq, kf, sw, b, _ = slices7[name]
            F.qlinear_into(acts[src], wq, kf, sw, b, outs7[name], *xq_ws[src])

    def barrier():
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches_per_step = pq.launch_count()
    step()
    launches_per_step = pq.launch_count() - launches_per_step
    run_step, launch_mode = capture(torch, dev, step, args.no_graph)
    for _ in range(max(args.warmup, 3)):
        run_step()
    barrier()

    # ---- timed region: device-resident inputs ----
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        run_step()
    e1.record()
    barrier()
    launches = launches_per_step * args.steps
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    if not sampler.samples and sampler.nv is not None:
        # region too short to be sampled: take samples under the same load right now
        for _ in range(200):
            run_step()
        try:
            sampler.sample()
        except Exception:
            pass
        torch.cuda.synchronize()
    ms_per_step = e0.elapsed_time(e1) / args.steps
    value = OPS_PER_STEP / (ms_per_step * 1e-3) / 1e12
    clk_now = sampler.result()

    # ---- roofline of the dominant kernel (the tcgen05 GEMM) ----
    # Average launch duration = CUDA-event time of a graph that replays ONLY the step's GEMM launches (same weights
    # cycling through 202 MB, pre-quantised activations), divided by the launch count; further graphs do the same
    # for cuBLASLt int8 (torch._int_mm: int32 output, no epilogue) on the same shapes and for the act-quant launches.
    # (Bracketing every launch with its own pair of events inflates a 30-130 us kernel by ~10 us of event latency.)
    for g, members, src in GROUPS:
        F.quantize_act(acts[src], out=xq_ws[src])

    def gemm_only():
        for g, members, src in GROUPS:
            m = fused[g]
            F.qgemm(xq_ws[src][0], xq_ws[src][1], m.qweight, m.weight_scale, m.bias, torch.bfloat16, out=outs[g])

    def quant_only():
        for g, members, src in GROUPS:
            F.quantize_act(acts[src], out=xq_ws[src])

    wts = {g: fused[g].qweight.t() for g, _, _ in GROUPS}      # [K, N] column-major views: what _int_mm wants
    acc32 = {g: torch.empty(M_TOKENS, fused[g].out_features, dtype=torch.int32, device=dev) for g, _, _ in GROUPS}

    def cublaslt_only():
        for g, members, src in GROUPS:
            torch._int_mm(xq_ws[src][0], wts[g], out=acc32[g])

    roof_steps = max(3, min(args.steps, 100))
    gemm_ms = timed(torch, capture(torch, dev, gemm_only, args.no_graph)[0], roof_steps)
    quant_ms = timed(torch, capture(torch, dev, quant_only, args.no_graph)[0], roof_steps)
    try:
        cublas_ms = timed(torch, capture(torch, dev, cublaslt_only, args.no_graph)[0], roof_steps)
    except Exception as ex:
        sys.stderr.write(f"bench: cuBLASLt int8 comparator unavailable ({ex!r})\n")
        cublas_ms = None
    del acc32
    unfused_ms = timed(torch, capture(torch, dev, step_unfused, args.no_graph)[0], roof_steps) / roof_steps
    n_gemm = len(GROUPS)
    quant_bytes = roof_steps * sum(M_TOKENS * (3 * k + 4) for k in ACTS.values())
    gemm_tops = OPS_PER_STEP * roof_steps / (gemm_ms * 1e-3) / 1e12
    cublas_tops = OPS_PER_STEP * roof_steps / (cublas_ms * 1e-3) / 1e12 if cublas_ms else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("qgemm_dram_bytes_per_launch")
        except Exception:
            traffic = None
    # Which measured peak applies: the burst figure for kernels timed alone at full clocks, the sustained one when
    # the timed region ran under the power cap (SM clock well below max or sw_power_cap seen by the sampler).
    capped = ("sw_power_cap" in clk_now["reasons"]) or bool(
        clk_now["sm_mhz"] and clk_now["sm_max_mhz"] and clk_now["sm_mhz"] < 0.93 * clk_now["sm_max_mhz"])
    peak_burst, peak_sust = 2.0 * peaks["bf16_tflops"], 2.0 * peaks["bf16_tflops_sustained"]
    peak_tops = peak_sust if capped else peak_burst
    act_gbs = quant_bytes / (quant_ms * 1e-3) / 1e9
    roofline = {
        "bound": "tensor", "kernel": "qgemm_kernel (tcgen05.mma.kind::i8, cta_group::2)", "achieved": gemm_tops, "peak": peak_tops,
        "unit": "TFLOP/s", "frac": gemm_tops / peak_tops, "traffic": traffic,
        "peak_regime": "sustained (power-capped run)" if capped else "burst",
        "frac_vs_burst": gemm_tops / peak_burst, "frac_vs_sustained": gemm_tops / peak_sust,
        "frac_of_nominal_int8": gemm_tops / NOMINAL_INT8_TOPS,
        "cublaslt_int8_tops": cublas_tops, "vs_cublaslt": (gemm_tops / cublas_tops) if cublas_tops else None,
        "cublaslt_note": "torch._int_mm (cuBLASLt int8, int32 output, no dequant epilogue) on the same 4 GEMM shapes and operands, "
                         "CUDA-graph replay, same run",
        "peak_note": f"2 x {peaks['source']} cuBLAS bf16 (int8 tensor rate = 2x bf16): burst {peaks['bf16_tflops']} TF/s, sustained "
                     f"{peaks['bf16_tflops_sustained']} TF/s; nominal dense int8 {NOMINAL_INT8_TOPS:.0f} TOPS; the GEMM-only graph is "
                     "timed right after the step loop, so it runs under the 1 kW power cap whenever the step loop did: `peak` "
                     "follows the clock sampler",
        "gemm_launches_per_step": n_gemm, "avg_launch_ms": gemm_ms / (roof_steps * n_gemm),
        "act_quant_step_gbs": act_gbs, "act_quant_step_frac": act_gbs / peaks["hbm_gbs"],
        "act_quant_step_avg_launch_ms": quant_ms / (roof_steps * n_gemm),
        "act_quant_step_note": "the step's four act-quant launches alone (2048 x 4096 / 11008 bf16, 25-68 MB each): latency bound",
    }
    # HBM-streaming point for the activation quantizer (traffic >> L2): SURVEY.md §8d
    Mbig, Kbig = 131072, 4096
    xb = torch.randn(Mbig, Kbig, device=dev).to(torch.bfloat16)
    qb, sb = F.alloc_q(Mbig, Kbig, dev), torch.empty(Mbig, dtype=torch.float32, device=dev)
    s_ms = timed(torch, lambda: F.quantize_act(xb, out=(qb, sb)), 10)
    gbs = 10 * Mbig * (3 * Kbig + 4) / (s_ms * 1e-3) / 1e9
    roofline["act_quant_stream_gbs"] = gbs
    roofline["act_quant_stream_frac"] = gbs / peaks["hbm_gbs"]
    roofline["act_quant_stream_shape"] = [Mbig, Kbig]
    # optional transposed output of the quantizer ([K, M]; two launches: row scales, then 128 x 128 tiles -- x is read twice)
    qt_big = F.alloc_q(Kbig, Mbig, dev)
    t_ms = timed(torch, lambda: F.quantize_act(xb, transpose=True, out=(qt_big, sb)), 5)
    roofline["act_quant_transposed_stream_frac"] = 5 * Mbig * (3 * Kbig + 4) / (t_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]
    del xb, qb, sb, qt_big
    qt = F.alloc_q(4096, M_TOKENS, dev)
    st_ = torch.empty(M_TOKENS, dtype=torch.float32, device=dev)
    t_run, _ = capture(torch, dev, lambda: [F.quantize_act(acts[a], transpose=True, out=(qt, st_)) for a in ("x_attn", "attn_out", "x_mlp")], args.no_graph)
    t_us = timed(torch, t_run, 20) / 60 * 1e3
    roofline["act_quant_transposed_2048x4096_us"] = t_us
    roofline["act_quant_transposed_2048x4096_frac"] = M_TOKENS * (3 * 4096 + 4) / (t_us * 1e-6) / 1e9 / peaks["hbm_gbs"]
    del qt, st_

    # ---- sustained regime: at least one second of back-to-back steps (the chip reaches its 1 kW power cap) ----
    sust_steps = int(min(20000, max(200, 1.2 / (ms_per_step * 1e-3))))
    s2 = ClockSampler(local)
    s2.start()
    sust_ms = timed(torch, run_step, sust_steps, warm=0) / sust_steps
    s2.stop_flag.set()
    s2.join(timeout=2)
    sustained = {"value": OPS_PER_STEP / (sust_ms * 1e-3) / 1e12, "unit": "TOPS", "ms_per_step": sust_ms, "steps": sust_steps,
                 "clocks": s2.result()}

    # ---- e2e: host buffers, copies inside the timed region, through the public module API ----
    host_in = {a: torch.randn(M_TOKENS, k).to(torch.bfloat16).pin_memory() for a, k in ACTS.items()}
    host_out = {g: torch.empty(M_TOKENS, fused[g].out_features, dtype=torch.bfloat16).pin_memory() for g, _, _ in GROUPS}
    h2d_bytes = sum(t.numel() * 2 for t in host_in.values())
    d2h_bytes = sum(t.numel() * 2 for t in host_out.values())
    copy_in, copy_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()

    # Software-pipelined over steps: device input buffers are double-buffered, so the H2D copies of
    # step s+1, the kernels of step s and the D2H copies of step s-1 run concurrently (three streams,
    # events only).  Every byte of every step still crosses PCIe inside the timed region.
    acts2 = [acts, {a: torch.empty_like(t) for a, t in acts.items()}]
    compute_done = [None, None]
    step_no = [0]

    def e2e_step():
        buf = step_no[0] & 1
        step_no[0] += 1
        ready = {}
        with torch.cuda.stream(copy_in):
            if compute_done[buf] is not None:
                copy_in.wait_event(compute_done[buf])      # kernels of step s-2 have consumed this buffer
            for a in ACTS:
                acts2[buf][a].copy_(host_in[a], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_in)
                ready[a] = ev
        for g, members, src in GROUPS:
            main.wait_event(ready[src])
            y = fused[g](acts2[buf][src])                  # public API: DynamicQuantLinear.forward (the nn.Linear replacement)
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(copy_out):
                copy_out.wait_event(done)
                host_out[g].copy_(y, non_blocking=True)
                y.record_stream(copy_out)
        ev = torch.cuda.Event()
        ev.record(main)
        compute_done[buf] = ev

    def e2e_drain():
        main.wait_stream(copy_out)
        main.wait_stream(copy_in)

    e2e_steps = max(3, min(args.steps, 60))
    for _ in range(3):
        e2e_step()
    e2e_drain()
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_drain()                                            # t1 is recorded after the last D2H copy has landed
    t1.record()
    barrier()
    e2e_ms_step = t0.elapsed_time(t1) / e2e_steps
    e2e = {"value": OPS_PER_STEP / (e2e_ms_step * 1e-3) / 1e12, "unit": "TOPS", "h2d_bytes_per_step": h2d_bytes,
           "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_ms_step, "steps": e2e_steps,
           "tokens_per_s": M_TOKENS / (e2e_ms_step * 1e-3),
           "pcie_gbs": (h2d_bytes + d2h_bytes) / (e2e_ms_step * 1e-3) / 1e9}
    del host_in, host_out, acts2

    # ---- side legs ----
    def leg(fn, *a):
        try:
            return fn(*a)
        except Exception as ex:  # never let a side measurement kill the headline line
            return {"error": repr(ex)[:200]}

    sharded = leg(sharded_leg, pq, torch, ctx["dist"], dev, rank, world)
    decode = leg(decode_leg, pq, F, torch, dev, peaks)
    fusedp = leg(fused_producer_leg, F, torch, dev, peaks)
    if "error" not in fusedp:
        fusedp["llama7b_block_fused_2048tok"] = leg(fused_block_leg, F, torch, dev, fused, acts)
    bert = leg(bert_leg, pq, F, torch, dev)
    cpu = cpu_baseline_leg()
    line = {
        "metric": "int8_qlinear_tops", "value": value, "unit": "TOPS", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int8", "data": "synthetic",
        "config": {"workload": "llama7b_linears_2048tok", "tokens_per_gpu": M_TOKENS, "act_dtype": "bf16",
                   "out_dtype": "bf16", "linears": {l[0]: [l[1], l[2]] for l in LINEARS},
                   "launches_per_step": launches_per_step,
                   "fusion": "swap_linear(fuse_shared_inputs=True): one act-quant + one GEMM per distinct activation "
                             "(qkv, o, gate_up, down); outputs bit-identical to seven separate linears",
                   "parallelism": "1 GPU", "launch": launch_mode,
                   "l2": "inputs larger than L2: each step streams 202 MB of int8 weights + 0.27 GB of activations/outputs (> 126 MB L2)"},
        "tokens_per_s": M_TOKENS / (ms_per_step * 1e-3),
        "frac_of_nominal_int8": value / NOMINAL_INT8_TOPS,
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clk_now, "sustained": sustained,
        "step_unfused_7x2_launches_ms": unfused_ms, "step_ms": ms_per_step,
        "gemm_tops": gemm_tops, "cublaslt_int8_tops": cublas_tops, "vs_cublaslt": roofline["vs_cublaslt"],
        "act_quant_stream_frac": roofline["act_quant_stream_frac"], "act_quant_step_frac": roofline["act_quant_step_frac"],
        "sharded_70b": sharded, "decode_16tok": decode, "fused_producers": fusedp, "bert_base_4096tok": bert,
    }
    print(json.dumps(line), flush=True)
    return 0


def run_sharded(ctx):
    """N > 1: the column-sharded Llama-70B up projection with its all-gather as the headline (strong scaling)."""
    torch, dist, pq, F, dev, peaks, args = ctx["torch"], ctx["dist"], ctx["pq"], ctx["F"], ctx["dev"], ctx["peaks"], ctx["args"]
    rank, world, local = ctx["rank"], ctx["world"], ctx["local"]
    K, N, M = S70_K, S70_N, S70_M
    g = torch.Generator(device=dev).manual_seed(7)            # same seed on every rank: replicated weights and input
    wq_full = torch.randint(-127, 128, (N, K), dtype=torch.int8, device=dev, generator=g)
    sw_full = torch.rand(N, device=dev, generator=g) * 1e-3
    x = torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16)
    full = pq.DynamicQuantLinear(K, N, bias=False, device=dev)
    full.qweight_storage[:, :K].copy_(wq_full)
    full.weight_scale.copy_(sw_full)
    shf = pq.ShardedDynamicQuantLinear(wq_full, sw_full, None, fused=None)

    def allmax(v):
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    # ---- parity gate: sharded + gathered output == replicated layer, bit for bit, on every rank ----
    y_ref = full(x)
    ok = True
    for _ in range(3):                                        # both halves of the double buffer, and a reuse
        ok = ok and bool(torch.equal(shf(x), y_ref))
    bad = allmax(0.0 if ok else 1.0)
    if bad:
        if rank == 0:
            sys.stderr.write("bench: column-sharded output differs from the replicated layer -- failing the run\n")
        return 3
    del y_ref
    barrier()

    # replicated layer on one GPU (the strong-scaling base), timed on every rank
    rep_run, _ = capture(torch, dev, lambda: full(x), args.no_graph)
    rep_ms = allmax(timed(torch, rep_run, 20) / 20)
    barrier()

    # Two forwards per graph: the module double-buffers its symmetric output (a rank may already be storing step
    # s+1 into a peer while that peer still reads step s), and a captured forward always uses the buffer it was
    # captured with -- so the graph holds one forward per buffer.
    def two_steps():
        shf(x)
        shf(x)

    for _ in range(max(args.warmup, 3)):
        two_steps()
    barrier()
    launches0 = pq.launch_count()
    two_steps()
    launches_per_step = (pq.launch_count() - launches0) // 2
    barrier()
    run2, launch_mode = capture(torch, dev, two_steps, args.no_graph)
    steps = max(2, args.steps + (args.steps & 1))             # even
    for _ in range(max((args.warmup + 1) // 2, 2)):
        run2()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps // 2):
        run2()
    e1.record()
    barrier()
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    if not sampler.samples and sampler.nv is not None:
        for _ in range(50):
            run2()
        try:
            sampler.sample()
        except Exception:
            pass
        torch.cuda.synchronize()
    ms_per_step = allmax(e0.elapsed_time(e1) / steps)
    ops = 2.0 * M * N * K
    value = ops / (ms_per_step * 1e-3) / 1e12

    # NCCL all-gather baseline (same shard GEMM, separate collective + layout copy)
    sh_nccl = pq.ShardedDynamicQuantLinear(wq_full, sw_full, None, fused=False)
    nccl_ms = allmax(timed(torch, lambda: sh_nccl(x), 20) / 20)
    nccl_ok = bool(torch.equal(sh_nccl(x), full(x)))
    del sh_nccl
    barrier()

    # ---- roofline: the slower of the shard GEMM at the measured tensor peak and the bytes every rank must receive ----
    bytes_in = (world - 1) / world * M * N * 2
    link_ms = bytes_in / (NVLINK_PEER_GBS * 1e9) * 1e3
    gemm_ideal_ms = ops / world / (2.0 * peaks["bf16_tflops"] * 1e12) * 1e3
    target_ms = max(link_ms, gemm_ideal_ms)
    in_gbs = bytes_in / (ms_per_step * 1e-3) / 1e9
    rank_tops = ops / world / (ms_per_step * 1e-3) / 1e12
    link_bound = link_ms >= gemm_ideal_ms
    # serial model of the fused kernel: act-quant at HBM peak, then the slower of (first tile pair + the bytes to receive at the
    # measured all-to-all store rate) and (whole waves of 256x256 tile pairs on 74 CTA pairs), then one cross-rank barrier
    quant_ms_model = M * (3 * K + 4) / (peaks["hbm_gbs"] * 1e9) * 1e3
    pair_ms = 2.0 * 256 * 256 * K / (2.0 * peaks["bf16_tflops"] * 1e12 / 74) * 1e3
    pair_tiles = ((M + 255) // 256) * ((N // world + 255) // 256)
    waves = (pair_tiles + 73) // 74
    serial_ms = quant_ms_model + max(pair_ms + bytes_in / (NVLINK_A2A_GBS * 1e9) * 1e3, waves * pair_ms) + 0.006
    roofline = {
        "bound": "nvlink" if link_bound else "tensor",
        "kernel": "qgemm_kernel<staged> (tcgen05 GEMM, epilogue stores every tile into all ranks' symmetric output buffers over NVLink: coalesced 256-byte LSU peer stores)",
        "achieved": in_gbs if link_bound else rank_tops, "peak": NVLINK_PEER_GBS if link_bound else 2.0 * peaks["bf16_tflops"],
        "unit": "GB/s" if link_bound else "TFLOP/s",
        "frac": (in_gbs / NVLINK_PEER_GBS) if link_bound else rank_tops / (2.0 * peaks["bf16_tflops"]), "traffic": None,
        "nvlink_in_gbs_per_rank": in_gbs, "nvlink_frac_of_770": in_gbs / NVLINK_PEER_GBS, "per_rank_tops": rank_tops,
        "bytes_received_per_rank": bytes_in, "link_floor_ms": link_ms, "shard_gemm_floor_ms": gemm_ideal_ms,
        "target_ms": target_ms, "frac_of_target": target_ms / ms_per_step,
        "nvlink_all_to_all_store_gbs_measured": NVLINK_A2A_GBS, "frac_of_measured_all_to_all": in_gbs / NVLINK_A2A_GBS,
        "serial_model_ms": serial_ms,
        "frac_of_serial_model": serial_ms / ms_per_step, "tile_pair_waves": waves,
        "serial_model_note": "act-quant at HBM peak + max(whole waves of 256x256 tile pairs on 74 CTA pairs; the first tile pair (nothing can "
                             "be sent before it) + the bytes every rank must receive at the measured all-to-all store rate (tools/nvlink_probe.py: 645 GB/s per direction per GPU "
                             "with all 8 GPUs pushing at once -- LSU stores, TMA bulk stores and multimem.st alike; the 770 GB/s figure "
                             "is a pairwise copy-engine number) + one cross-rank barrier (6 us)",
        "peak_note": "the bound is the slower of (a) the bytes every rank must receive at 770 GB/s = measured peer-copy bandwidth per "
                     "direction per GPU (B200_PROFILING.md; nominal 900) and (b) the shard GEMM 2MNK/" + str(world) + " at 2 x measured cuBLAS bf16 "
                     f"burst ({peaks['bf16_tflops']} TF/s); frac_of_target = that floor / measured time",
    }

    # ---- e2e: the batch lives ONCE in pinned host memory; every step brings it in and takes the result out ----
    # Every rank copies its 1/world block of the token rows host -> device (the H2D traffic is spread over all PCIe
    # links), an all-gather over NVLink (NCCL) replicates the activation, the column-sharded layer runs through the
    # public module API, and every rank copies ITS column slice of the gathered result device -> host.  Three streams,
    # software-pipelined over steps (H2D + gather of step s+1 | kernels of step s | D2H of step s-1); every byte of
    # every step crosses PCIe inside the timed region.
    lo, hi = shf.lo, shf.hi
    rows = M // world
    host_x = x.cpu().pin_memory()                                  # the whole batch, as a host process would hold it
    host_y = torch.empty(M, hi - lo, dtype=torch.bfloat16).pin_memory()
    x_part = [torch.empty(rows, K, dtype=torch.bfloat16, device=dev) for _ in range(2)]
    x_devs = [torch.empty_like(x) for _ in range(2)]
    copy_in, copy_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()
    compute_done, out_done = [None, None], [None, None]
    step_no = [0]

    def e2e_step():
        b = step_no[0] & 1
        step_no[0] += 1
        with torch.cuda.stream(copy_in):
            if compute_done[b] is not None:
                copy_in.wait_event(compute_done[b])               # step s-2 has consumed this activation buffer
            x_part[b].copy_(host_x[rank * rows:(rank + 1) * rows], non_blocking=True)
            dist.all_gather_into_tensor(x_devs[b], x_part[b])
            ready = torch.cuda.Event()
            ready.record(copy_in)
        main.wait_event(ready)
        if out_done[b] is not None:
            main.wait_event(out_done[b])                          # the module's double-buffered output: step s-2 was read out
        y = shf(x_devs[b])
        done = torch.cuda.Event()
        done.record(main)
        compute_done[b] = done
        with torch.cuda.stream(copy_out):
            copy_out.wait_event(done)
            host_y.copy_(y[:, lo:hi], non_blocking=True)
            od = torch.cuda.Event()
            od.record(copy_out)
            out_done[b] = od

    def e2e_drain():
        main.wait_stream(copy_out)
        main.wait_stream(copy_in)

    for _ in range(4):
        e2e_step()
    e2e_drain()
    barrier()
    e2e_steps = max(4, min(args.steps + (args.steps & 1), 40))
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_drain()
    t1.record()
    barrier()
    e2e_ms = allmax(t0.elapsed_time(t1) / e2e_steps)
    e2e_ok = bool(torch.equal(host_y, full(x)[:, lo:hi].cpu()))
    e2e = {"value": ops / (e2e_ms * 1e-3) / 1e12, "unit": "TOPS", "h2d_bytes_per_step": M * K * 2,
           "d2h_bytes_per_step": M * N * 2, "ms_per_step": e2e_ms, "steps": e2e_steps, "output_matches": e2e_ok,
           "note": "whole-job bytes: the batch is read from pinned host memory once per step (each rank copies its 1/world row block "
                   "and an NVLink all-gather replicates it), every rank writes its own column slice of the result back; H2D, "
                   "kernels and D2H of consecutive steps overlap on three streams"}
    del host_x, host_y, x_devs, x_part

    # ---- side legs: NCCL vs fused at M = 16 / 2048, row-parallel down projection, gated MLP; tokens-sharded 7B step ----
    try:
        sharded = sharded_leg(pq, torch, dist, dev, rank, world)
    except Exception as ex:
        sharded = {"error": repr(ex)[:300]}
    barrier()
    clocks = sampler.result()
    # parity gate, part 2: every multi-GPU side leg (NCCL vs fused at 16 / 2048 tokens, row-parallel down projection, gated MLP)
    # must have reproduced the single-GPU bits on every rank
    flags = [sharded.get("M16_bit_identical"), sharded.get("M2048_bit_identical"),
             (sharded.get("down_proj_row_parallel") or {}).get("bit_identical"),
             (sharded.get("gated_mlp_8192_28672") or {}).get("bit_identical")] if "error" not in sharded else []
    # (a side leg that raised is reported in the line as {"error": ...} and does not take the headline down; a leg that RAN
    # and produced different bits does)
    if allmax(1.0 if any(f is False for f in flags) else 0.0):
        if rank == 0:
            sys.stderr.write(f"bench: a multi-GPU leg differs from the single-GPU result ({flags}) -- failing the run\n")
        return 3
    if rank == 0:
        line = {
            "metric": "int8_qlinear_tops", "value": value, "unit": "TOPS", "n_gpus": world, "steps": steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": "llama70b_up_proj_colsharded_2048tok", "tokens": M, "layer": [K, N], "act_dtype": "bf16",
                       "out_dtype": "bf16", "launches_per_step": launches_per_step,
                       "parallelism": f"column-parallel x{world}: act-quant (replicated) + shard GEMM whose epilogue stores every tile "
                                      "into all ranks' output buffers over NVLink (fused all-gather) + 1 cross-rank barrier",
                       "launch": launch_mode,
                       "l2": "per-rank working set (weight shard + activation + 117 MB gathered output) is larger than L2"},
            "tokens_per_s": M / (ms_per_step * 1e-3),
            "roofline": roofline, "cpu_baseline": None, "e2e": e2e, "gpu_launches": int(launches_per_step * steps),
            "clocks": clocks,
            "bit_identical_to_replicated": True, "fused_path_active": bool(shf.fused), "fused_error": shf.fused_error,
            "replicated_1gpu_ms": rep_ms, "speedup_vs_replicated_1gpu": rep_ms / ms_per_step,
            "strong_scaling_efficiency": rep_ms / ms_per_step / world,
            "nccl_allgather_ms": nccl_ms, "nccl_bit_identical": nccl_ok,
            "sharded_70b": sharded,
        }
        print(json.dumps(line), flush=True)
    return 0


def decode_leg(pq, F, torch, dev, peaks):
    """BASELINE.json configs[0]: one nn.Linear 4096x4096 on 16 tokens (plus the Llama-7B MLP shapes at 16
    tokens).  HBM-bound weight streaming: achieved = int8 weight bytes / time, from a CUDA-graph replay of
    the module forward (act-quant + small-M tcgen05 GEMM) cycling 8 distinct weight sets (> L2)."""
    res = {}
    for name, K, N in (("linear_4096x4096", 4096, 4096), ("gate_4096x11008", 4096, 11008), ("down_11008x4096", 11008, 4096)):
        nset = 8
        mods = []
        for i in range(nset):
            m = pq.DynamicQuantLinear(K, N, bias=True, device=dev)
            m.qweight_storage.random_(-127, 128)
            m.weight_scale.uniform_(1e-4, 1e-3)
            mods.append(m)
        x = torch.randn(16, K, device=dev).to(torch.bfloat16)
        ws = (F.alloc_q(16, K, dev), torch.empty(16, dtype=torch.float32, device=dev))
        y = torch.empty(16, N, dtype=torch.bfloat16, device=dev)

        def run():
            for m in mods:
                xq, sx = F.quantize_act(x, out=ws)
                F.qgemm(xq, sx, m.qweight, m.weight_scale, m.bias, torch.bfloat16, out=y)
        run()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run()
        for _ in range(3):
            g.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / (20 * nset)
        gbs = (N * K + 16 * K * 3 + 16 * N * 2) / (us * 1e-6) / 1e9
        res[name] = {"us_per_forward": us, "tokens_per_s": 16 / (us * 1e-6), "achieved_gbs": gbs,
                     "frac_of_hbm_peak": gbs / peaks["hbm_gbs"]}
        del mods
    return res


def fused_producer_leg(F, torch, dev, peaks):
    """RMSNorm -> int8 and silu(gate)*up -> int8 written by one kernel each (HBM-bound;
    algorithmic bytes per row 3K+4 and 5K+4 for bf16), at a streaming size (>> L2) and at the step's size."""
    res = {}

    def timed(fn, iters):
        fn(0)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn(0)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(iters):
                fn(i)
        g.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / (3 * iters)

    for label, M, K in (("rmsnorm_quant_stream", 131072, 4096), ("rmsnorm_quant_2048tok", 2048, 4096)):
        nb = max(1, min(8, int(600e6 // (M * K * 2))))
        xs = [torch.randn(M, K, device=dev).to(torch.bfloat16) for _ in range(nb)]
        w = torch.ones(K, dtype=torch.bfloat16, device=dev)
        q, s = F.alloc_q(M, K, dev), torch.empty(M, dtype=torch.float32, device=dev)
        ms = timed(lambda i: F.rmsnorm_quant(xs[i % nb], w, out=(q, s)), 10 if M > 10000 else 50)
        gbs = M * (3 * K + 4) / (ms * 1e-3) / 1e9
        res[label] = {"shape": [M, K], "us": ms * 1e3, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peaks["hbm_gbs"]}
        del xs
    for label, M, K in (("silu_mul_quant_stream", 32768, 11008), ("silu_mul_quant_2048tok", 2048, 11008)):
        nb = max(1, min(8, int(600e6 // (M * K * 4))))
        gs = [torch.randn(M, 2 * K, device=dev).to(torch.bfloat16) for _ in range(nb)]
        q, s = F.alloc_q(M, K, dev), torch.empty(M, dtype=torch.float32, device=dev)
        ms = timed(lambda i: F.act_mul_quant(gs[i % nb][:, :K], gs[i % nb][:, K:], act="silu", out=(q, s)), 10 if M > 10000 else 50)
        gbs = M * (5 * K + 4) / (ms * 1e-3) / 1e9
        res[label] = {"shape": [M, K], "us": ms * 1e3, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peaks["hbm_gbs"]}
        del gs
    return res


def bert_leg(pq, F, torch, dev):
    """BASELINE.json configs[1]: the six linears of a BERT-base encoder layer (768 / 3072) at batch 32 x seq 128 =
    4096 tokens, each as act-quant + GEMM (12 launches), and the same layer with LayerNorm -> int8 and GELU -> int8
    fused into the producers (9 launches).  CUDA-graph replay."""
    M = 32 * 128
    shapes = [("q", 768, 768), ("k", 768, 768), ("v", 768, 768), ("o", 768, 768), ("ffn_up", 768, 3072), ("ffn_down", 3072, 768)]
    ops = sum(2 * M * k * n for _, k, n in shapes)
    mods = {}
    for name, k, n in shapes:
        m = pq.DynamicQuantLinear(k, n, bias=True, device=dev)
        m.qweight_storage.random_(-127, 128)
        m.weight_scale.uniform_(1e-4, 1e-3)
        mods[name] = m
    x = torch.randn(M, 768, device=dev).to(torch.bfloat16)
    ctx = torch.randn(M, 768, device=dev).to(torch.bfloat16)
    ln_w = torch.ones(768, dtype=torch.bfloat16, device=dev)
    ln_b = torch.zeros(768, dtype=torch.bfloat16, device=dev)
    outs = {name: torch.empty(M, n, dtype=torch.bfloat16, device=dev) for name, k, n in shapes}
    ws768 = (F.alloc_q(M, 768, dev), torch.empty(M, dtype=torch.float32, device=dev))
    ws3072 = (F.alloc_q(M, 3072, dev), torch.empty(M, dtype=torch.float32, device=dev))

    def lin(name, inp, ws):
        m = mods[name]
        F.qlinear_into(inp, m.qweight_storage, m.in_features, m.weight_scale, m.bias, outs[name], *ws)

    def gemm(name, ws):
        m = mods[name]
        F.qgemm(ws[0], ws[1], m.qweight, m.weight_scale, m.bias, torch.bfloat16, out=outs[name])

    def unfused():
        for n in ("q", "k", "v"):
            lin(n, x, ws768)
        lin("o", ctx, ws768)
        lin("ffn_up", outs["o"], ws768)
        lin("ffn_down", outs["ffn_up"], ws3072)       # (GELU left out: it is not on the path)

    def fused():
        F.quantize_act(x, out=ws768)
        for n in ("q", "k", "v"):
            gemm(n, ws768)
        F.quantize_act(ctx, out=ws768)
        gemm("o", ws768)
        F.layernorm_quant(outs["o"], ln_w, ln_b, out=ws768)        # attention-output LayerNorm -> int8
        gemm("ffn_up", ws768)
        F.act_mul_quant(outs["ffn_up"], None, act="gelu", out=ws3072)   # GELU -> int8
        gemm("ffn_down", ws3072)

    # what swap_linear(fuse_shared_inputs=True) builds for a BERT layer: query / key / value as ONE 768 -> 2304 GEMM
    qkv = pq.fuse_linears([mods["q"], mods["k"], mods["v"]])
    out_qkv = torch.empty(M, 2304, dtype=torch.bfloat16, device=dev)

    def shared_input():
        F.qlinear_into(x, qkv.qweight_storage, 768, qkv.weight_scale, qkv.bias, out_qkv, *ws768)
        lin("o", ctx, ws768)
        lin("ffn_up", outs["o"], ws768)
        lin("ffn_down", outs["ffn_up"], ws3072)

    def shared_input_fused_producers():
        F.qlinear_into(x, qkv.qweight_storage, 768, qkv.weight_scale, qkv.bias, out_qkv, *ws768)
        lin("o", ctx, ws768)
        F.layernorm_quant(outs["o"], ln_w, ln_b, out=ws768)
        gemm("ffn_up", ws768)
        F.act_mul_quant(outs["ffn_up"], None, act="gelu", out=ws3072)
        gemm("ffn_down", ws3072)

    tmp768 = torch.empty(M, 768, dtype=torch.bfloat16, device=dev)
    tmp3072 = torch.empty(M, 3072, dtype=torch.bfloat16, device=dev)

    def unfused_with_ops():
        # the same layer as `fused`, with torch's LayerNorm and GELU as separate kernels in front of the quantizers
        for n in ("q", "k", "v"):
            lin(n, x, ws768)
        lin("o", ctx, ws768)
        tmp768.copy_(torch.nn.functional.layer_norm(outs["o"], (768,), ln_w, ln_b, 1e-12))
        lin("ffn_up", tmp768, ws768)
        tmp3072.copy_(torch.nn.functional.gelu(outs["ffn_up"]))
        lin("ffn_down", tmp3072, ws3072)

    res = {"tokens": M, "linears": {n: [k, nn_] for n, k, nn_ in shapes},
           "note": "act_quant_per_linear = the six linears alone (12 launches); shared_input_fusion = the same work the way swap_linear sets a "
                   "BERT layer up (query/key/value as one 768 -> 2304 GEMM: 8 launches); with_torch_layernorm_gelu adds LayerNorm and GELU "
                   "as torch kernels in front of the quantizers; fused_producers computes them inside the quantizing kernels (9 launches); "
                   "the last leg combines both fusions (8 launches, LayerNorm and GELU included)"}
    for label, fn in (("act_quant_per_linear", unfused), ("shared_input_fusion", shared_input),
                      ("with_torch_layernorm_gelu", unfused_with_ops), ("fused_producers", fused),
                      ("shared_input_fusion_and_fused_producers", shared_input_fused_producers)):
        fn()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        for _ in range(3):
            g.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 50
        res[label] = {"ms_per_layer": ms, "tops": ops / (ms * 1e-3) / 1e12, "tokens_per_s": M / (ms * 1e-3)}
    return res


def fused_block_leg(F, torch, dev, fused, acts):
    """The same GEMMs wired the way a Llama block uses them, with the producer fusions (§8f-2): RMSNorm -> int8 feeds
    the fused q/k/v GEMM and the fused gate/up GEMM, silu(gate)*up -> int8 is written straight from the gate/up
    output: 4 quantising launches that also do the norm / activation work + 4 GEMMs = 8 launches."""
    M = M_TOKENS
    w_norm = torch.ones(4096, dtype=torch.bfloat16, device=dev)
    qa = (F.alloc_q(M, 4096, dev), torch.empty(M, dtype=torch.float32, device=dev))
    qo = (F.alloc_q(M, 4096, dev), torch.empty(M, dtype=torch.float32, device=dev))
    qm = (F.alloc_q(M, 4096, dev), torch.empty(M, dtype=torch.float32, device=dev))
    qh = (F.alloc_q(M, 11008, dev), torch.empty(M, dtype=torch.float32, device=dev))
    outs = {g: torch.empty(M, fused[g].out_features, dtype=torch.bfloat16, device=dev) for g in fused}

    def gemm(name, q):
        m = fused[name]
        F.qgemm(q[0], q[1], m.qweight, m.weight_scale, m.bias, torch.bfloat16, out=outs[name])

    def block():
        F.rmsnorm_quant(acts["x_attn"], w_norm, out=qa)
        gemm("qkv_proj", qa)
        F.quantize_act(acts["attn_out"], out=qo)          # attention itself is outside the path
        gemm("o_proj", qo)
        F.rmsnorm_quant(outs["o_proj"], w_norm, out=qm)   # post-attention norm reads the o_proj output
        gemm("gate_up_proj", qm)
        y_gu = outs["gate_up_proj"]
        F.act_mul_quant(y_gu[:, :11008], y_gu[:, 11008:], act="silu", out=qh)
        gemm("down_proj", qh)

    run, _ = capture(torch, dev, block)
    ms = timed(torch, run, 50) / 50
    return {"ms_per_block": ms, "tops": OPS_PER_STEP / (ms * 1e-3) / 1e12, "tokens_per_s": M / (ms * 1e-3), "launches": 8,
            "note": "rmsnorm_quant x2 + act_quant + silu_mul_quant + 4 GEMMs (qkv and gate/up fused), CUDA-graph replay x50"}


def sharded_leg(pq, torch, dist, dev, rank, world):
    """Llama-70B up-projection (8192 -> 28672): replicated vs column-sharded + all-gather, M in {16, 2048}."""
    K, N = 8192, 28672
    g = torch.Generator(device=dev).manual_seed(7)
    wq_full = torch.randint(-127, 128, (N, K), dtype=torch.int8, device=dev, generator=g)
    sw_full = torch.rand(N, device=dev, generator=g) * 1e-3
    res = {"layer": [K, N], "world": world}
    full = pq.DynamicQuantLinear(K, N, bias=False, device=dev)
    full.qweight_storage[:, :K].copy_(wq_full)
    full.weight_scale.copy_(sw_full)
    sh = pq.ShardedDynamicQuantLinear(wq_full, sw_full, None, fused=False) if world > 1 else None
    shf = pq.ShardedDynamicQuantLinear(wq_full, sw_full, None, fused=None) if world > 1 else None
    for M in (16, 2048):
        x = torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16)
        for label, mod in (("replicated", full), ("sharded_nccl_allgather", sh), ("sharded_fused_epilogue", shf)):
            if mod is None:
                continue
            for _ in range(3):
                mod(x)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                y = mod(x)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 20
            if world > 1:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = t.item()
            res[f"M{M}_{label}"] = {"ms": ms, "tops": 2 * M * N * K / (ms * 1e-3) / 1e12, "tokens_per_s": M / (ms * 1e-3)}
        if sh is not None:
            res[f"M{M}_bit_identical"] = bool(torch.equal(sh(x), full(x)) and torch.equal(shf(x), full(x)))
    if shf is not None:
        res["fused_path_active"] = bool(shf.fused)
    # row-parallel (K-split) down projection 28672 -> 8192: fused GEMM + reduce-scatter (+ all-gather) on int32 partials
    if world > 1:
        Kd, Nd = 28672, 8192
        wq_d = torch.randint(-127, 128, (Nd, Kd), dtype=torch.int8, device=dev, generator=g)
        sw_d = torch.rand(Nd, device=dev, generator=g) * 1e-3
        full_d = pq.DynamicQuantLinear(Kd, Nd, bias=False, device=dev)
        full_d.qweight_storage[:, :Kd].copy_(wq_d)
        full_d.weight_scale.copy_(sw_d)
        rp = pq.RowParallelDynamicQuantLinear(wq_d, sw_d, None, fused=None)
        rps = pq.RowParallelDynamicQuantLinear(wq_d, sw_d, None, fused=None, input_is_sharded=True, gather_output=False)
        x = torch.randn(2048, Kd, device=dev, generator=g).to(torch.bfloat16)
        down = {"layer": [Kd, Nd], "M": 2048}

        def graph_ms(fn):
            """CUDA-graph replay of TWO forwards (one per half of the modules' double buffers): ms per forward, max over ranks."""
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            run, mode = capture(torch, dev, lambda: (fn(), fn()))
            per = timed(torch, run, 10) / 20 if mode == "cuda_graph_replay" else timed(torch, fn, 20) / 20
            t = torch.tensor([per], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item(), mode

        for label, mod, xin in (("replicated", full_d, x), ("row_parallel_fused", rp, x),
                                ("row_parallel_fused_sharded_in_scattered_out", rps, x[:, rps.k_lo:rps.k_hi].contiguous())):
            ms, mode = graph_ms(lambda: mod(xin))
            down[label] = {"ms": ms, "tops": 2 * 2048 * Nd * Kd / (ms * 1e-3) / 1e12, "launch": mode}
        down["bit_identical"] = bool(torch.equal(rp(x), full_d(x)))
        down["fused_path_active"] = bool(rp.fused)
        down["note"] = ("one C call per forward (pq_rowparallel_forward): row-max -> quantise -> int32 GEMM scattering into the owners' "
                        "inboxes over NVLink -> signal-pad barrier -> reduce + dequant (+ all-gather store + barrier); no NCCL")
        res["down_proj_row_parallel"] = down
        # the whole Llama-70B gated MLP, Megatron layout (gate/up column-parallel, no gather; down row-parallel)
        F = pq.functional
        mk = lambda k, n: pq.DynamicQuantLinear(k, n, bias=False, device=dev)
        gate, up = mk(K, N), mk(K, N)
        for mm in (gate, up):
            mm.qweight_storage[:, :K].copy_(torch.randint(-127, 128, (N, K), dtype=torch.int8, device=dev, generator=g))
            mm.weight_scale.copy_(torch.rand(N, device=dev, generator=g) * 1e-3)
        mlp = pq.ParallelGatedMLP(gate, up, full_d)
        xm = torch.randn(2048, K, device=dev, generator=g).to(torch.bfloat16)
        one_gpu = lambda: full_d(F.act_mul(gate(xm), up(xm), "silu"))
        want = one_gpu()
        ms1, mode1 = graph_ms(one_gpu)
        msp, modep = graph_ms(lambda: mlp(xm))
        flops = 2 * 2048 * (2 * N * K + Nd * Kd)
        res["gated_mlp_8192_28672"] = {
            "M": 2048, "one_gpu_chain": {"ms": ms1, "tops": flops / (ms1 * 1e-3) / 1e12, "launch": mode1},
            "tensor_parallel": {"ms": msp, "tops": flops / (msp * 1e-3) / 1e12, "launch": modep},
            "bit_identical": bool(torch.equal(mlp(xm), want)), "fused_path_active": bool(mlp.down.fused),
            "note": "two C calls per forward: pq_qlinear (act-quant + fused gate/up GEMM on the local column slice) and "
                    "pq_rowparallel_forward with the gated input (silu(gate)*up max-exchanged and quantised on the fly)"}
        del gate, up, mlp
    return res


def main():
    # keep stdout to the one JSON line: NCCL prints its version banner (NCCL_DEBUG >= VERSION) to stdout
    # unless it is given a debug file
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
