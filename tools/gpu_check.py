"""Development harness: runs correctness + timing probes on a B200, one subprocess per probe
(with a timeout) so a trap or a hang in one kernel configuration cannot take the others down.
Results go to gpurun_out/gpu_check.json.  Not part of the product or of the test-suite.

usage:  python tools/gpu_check.py [probe ...]      (no args = all probes)
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def _time_ms(fn, iters=20, warmup=5):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[0], ts[len(ts) // 2]


def probe_quant(arg):
    import numpy as np
    import torch
    import protoquant_b200 as pq
    import protoquant_oracle as O
    out = {}
    torch.manual_seed(0)
    for dt in (torch.bfloat16, torch.float16, torch.float32):
        for (M, K) in ((1, 16), (3, 768), (17, 4096), (64, 11008), (8, 28672), (5, 100), (33, 4104), (2, 65536), (2, 70000)):
            x = torch.randn(M, K, dtype=torch.float32)
            x[0, 0] = 100.0
            if M > 1:
                x[1].zero_()
            x = x.to(dt)
            for mode in (0, 1, 2):
                spec = pq.QuantSpec(scale_mode=mode, eps=1e-5 if mode == 1 else 0.0)
                ospec = O.QuantSpec(scale_mode=mode, eps=1e-5 if mode == 1 else 0.0)
                q, s = pq.quantize_act(x.cuda(), spec=spec)
                qo, so = O.quantize_rowwise(x, ospec)
                nq = int((q.cpu().numpy() != qo).sum())
                ns = int((s.cpu().numpy().view(np.uint32) != so.view(np.uint32)).sum())
                out[f"{str(dt)[6:]}_{M}x{K}_m{mode}"] = [nq, ns]
            qt, st = pq.quantize_act(x.cuda(), transpose=True)
            qo, so = O.quantize_rowwise(x)
            out[f"{str(dt)[6:]}_{M}x{K}_T"] = [int((qt.cpu().numpy() != qo.T).sum()), int((st.cpu().numpy() != so).sum())]
    bad = {k: v for k, v in out.items() if v != [0, 0]}
    return {"cases": len(out), "bad": bad}


def probe_gemm(arg):
    import torch
    import protoquant_b200 as pq
    cfg = int(arg)
    pq.lib().pq_debug_set_gemm_config(cfg)
    res = {}
    torch.manual_seed(1)
    shapes = [(128, 256, 128), (128, 256, 512), (256, 256, 256), (256, 512, 4096), (300, 520, 1040),
              (1, 64, 16), (17, 40, 144), (2048, 4096, 4096), (513, 11008, 4096)]
    for (M, N, K) in shapes:
        a = torch.randint(-128, 128, (M, K), dtype=torch.int8)
        b = torch.randint(-128, 128, (N, K), dtype=torch.int8)
        ref = torch._int_mm(a, b.t()) if (M > 16 and N % 8 == 0 and K % 8 == 0) else (a.int() @ b.int().t())
        got = pq.qgemm_i32(a.cuda(), b.cuda())
        torch.cuda.synchronize()
        diff = (got.cpu() != ref)
        n = int(diff.sum())
        res[f"{M}x{N}x{K}"] = n
        if n:
            idx = diff.nonzero()[:5].tolist()
            res[f"{M}x{N}x{K}_first"] = [(i, j, int(got[i, j]), int(ref[i, j])) for i, j in idx]
            rows = diff.any(dim=1).nonzero().flatten()
            cols = diff.any(dim=0).nonzero().flatten()
            res[f"{M}x{N}x{K}_rows"] = [int(rows.min()), int(rows.max()), int(rows.numel())]
            res[f"{M}x{N}x{K}_cols"] = [int(cols.min()), int(cols.max()), int(cols.numel())]
    return res


def probe_epilogue(arg):
    import numpy as np
    import torch
    import protoquant_b200 as pq
    import protoquant_oracle as O
    res = {}
    torch.manual_seed(2)
    for cfg in (-1, 0, 1):
        pq.lib().pq_debug_set_gemm_config(cfg)
        for dt, name in ((torch.bfloat16, "bf16"), (torch.float16, "f16"), (torch.float32, "f32")):
            for (M, N, K, use_bias) in ((256, 512, 512, True), (100, 264, 272, False), (2048, 4096, 4096, True)):
                x = torch.randn(M, K).to(torch.bfloat16)
                w = (torch.rand(N, K) * 2 - 1) / K ** 0.5
                bias = torch.randn(N) if use_bias else None
                wq_o, sw_o = O.quantize_rowwise(w)
                y_o = O.qlinear(x, wq_o, sw_o, bias.numpy() if use_bias else None, out_dtype=name)
                wq, sw = pq.quantize_weight(w.cuda())
                y = pq.qlinear(x.cuda(), wq, sw, bias.cuda() if use_bias else None, out_dtype=dt)
                torch.cuda.synchronize()
                res[f"cfg{cfg}_{name}_{M}x{N}x{K}"] = int((y.cpu().view(torch.int16 if dt != torch.float32 else torch.int32)
                                                          != y_o.view(torch.int16 if dt != torch.float32 else torch.int32)).sum())
    return res


def probe_time_gemm(arg):
    import torch
    import protoquant_b200 as pq
    cfg = int(arg)
    pq.lib().pq_debug_set_gemm_config(cfg)
    res = {}
    for (M, N, K) in ((2048, 4096, 4096), (2048, 11008, 4096), (2048, 4096, 11008), (8192, 8192, 8192), (16, 4096, 4096), (4096, 3072, 768)):
        a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device="cuda")
        b = torch.randint(-128, 128, (N, K), dtype=torch.int8, device="cuda")
        sx = torch.rand(M, device="cuda")
        sw = torch.rand(N, device="cuda")
        y = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
        best, med = _time_ms(lambda: pq.qgemm(a, sx, b, sw, None, torch.bfloat16, out=y))
        res[f"{M}x{N}x{K}"] = {"best_ms": best, "med_ms": med, "tops_best": 2 * M * N * K / best / 1e9, "tops_med": 2 * M * N * K / med / 1e9}
    return res


def probe_time_ref(arg):
    import torch
    res = {}
    for (M, N, K) in ((2048, 4096, 4096), (2048, 11008, 4096), (2048, 4096, 11008), (8192, 8192, 8192)):
        a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device="cuda")
        b = torch.randint(-128, 128, (N, K), dtype=torch.int8, device="cuda").t()
        best, med = _time_ms(lambda: torch._int_mm(a, b))
        res[f"int_mm_{M}x{N}x{K}"] = {"best_ms": best, "tops_best": 2 * M * N * K / best / 1e9, "tops_med": 2 * M * N * K / med / 1e9}
        ab = torch.randn(M, K, dtype=torch.bfloat16, device="cuda")
        bb = torch.randn(N, K, dtype=torch.bfloat16, device="cuda").t()
        best, med = _time_ms(lambda: torch.matmul(ab, bb))
        res[f"bf16_{M}x{N}x{K}"] = {"best_ms": best, "tflops_best": 2 * M * N * K / best / 1e9, "tflops_med": 2 * M * N * K / med / 1e9}
    return res


def probe_time_quant(arg):
    import torch
    import protoquant_b200 as pq
    res = {}
    for dt, esz in ((torch.bfloat16, 2), (torch.float32, 4)):
        for (M, K) in ((2048, 4096), (2048, 11008), (131072, 4096), (65536, 11008), (32768, 28672), (262144, 768)):
            x = torch.randn(M, K, dtype=dt, device="cuda")
            q = pq.functional.alloc_q(M, K, "cuda")
            s = torch.empty(M, dtype=torch.float32, device="cuda")
            best, med = _time_ms(lambda: pq.quantize_act(x, out=(q, s)))
            byts = M * (K * esz + K + 4)
            res[f"{str(dt)[6:]}_{M}x{K}"] = {"best_ms": best, "gbs_best": byts / best / 1e6, "gbs_med": byts / med / 1e6}
            del x, q, s
    # reference point: torch copy of the same byte volume
    a = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    b = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    best, med = _time_ms(lambda: b.copy_(a))
    res["copy_1GiB"] = {"gbs_best": 2 * (1 << 30) / best / 1e6}
    return res


PROBES = {
    "quant": (probe_quant, [""]),
    "gemm": (probe_gemm, ["0", "2", "3", "1", "4"]),
    "epilogue": (probe_epilogue, [""]),
    "time_gemm": (probe_time_gemm, ["0", "1", "-1"]),
    "time_ref": (probe_time_ref, [""]),
    "time_quant": (probe_time_quant, [""]),
}

if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--one":
        name, arg = sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
        r = PROBES[name][0](arg)
        print("RESULT " + json.dumps(r))
        sys.exit(0)
    want = sys.argv[1:] or list(PROBES)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    allres = {}
    for name in want:
        for arg in PROBES[name][1]:
            key = f"{name}:{arg}"
            t0 = time.time()
            try:
                p = subprocess.run([sys.executable, __file__, "--one", name, arg], capture_output=True, text=True, timeout=240)
                lines = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
                if lines:
                    allres[key] = json.loads(lines[-1][7:])
                else:
                    allres[key] = {"error": f"rc={p.returncode}", "stdout": p.stdout[-1500:], "stderr": p.stderr[-2500:]}
            except subprocess.TimeoutExpired as e:
                allres[key] = {"error": "timeout", "stdout": (e.stdout or b"")[-1000:].decode(errors="replace") if isinstance(e.stdout, bytes) else str(e.stdout)[-1000:]}
            allres[key + ":secs"] = round(time.time() - t0, 1)
            print(key, json.dumps(allres[key])[:1500], flush=True)
            with open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w") as f:
                json.dump(allres, f, indent=1)
