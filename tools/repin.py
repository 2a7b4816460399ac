#!/usr/bin/env python
"""Re-pin the oracle against the REAL protoquant checkout, the day it is mounted.

  python tools/repin.py /path/to/protoquant [--write]

SURVEY.md §0: /root/reference holds only CODE_OF_CONDUCT.md, so parity is "unpinned": the oracle restates SPEC v0
and every arithmetic choice the reference could make differently is a `QuantSpec` knob.  This tool closes the gap
without touching a kernel:

1. imports the reference package from <reference_root> (nothing is copied) and finds its per-token / per-channel
   int8 quantisers -- by the names in CANDIDATES, or the ones given with --act-fn / --weight-fn
   ("module.sub:function" returning (int8 tensor, scale tensor) in either order, or an object with
   int_repr()/q_scale-like attributes);
2. runs them on the inputs of every committed golden file (tests/golden/torch_ao_*.npz, exact_*.npz) plus the
   seeded edge-case matrix of tests/test_gpu_quant.py::make_x, on the CPU;
3. for each knob set of the oracle (scale_mode x eps x qmin) reports whether codes AND scales match bit for bit;
   the first matching set is the pin.  If none matches it prints the first differing element (row, column, input
   value, reference code / scale, oracle code / scale per knob set) -- that is the new knob to add;
4. with --write: regenerates tests/golden/reference_*.npz from the reference's outputs (x_bits, q, s, the spec
   that matched) and writes protoquant_b200/_pinned_spec.json, which `protoquant_b200.functional.DEFAULT_SPEC`
   reads at import; prints the names found so that compat.py's alias table can be corrected by hand.

Exit status: 0 = pinned (a knob set matches everywhere), 1 = no knob set matches, 2 = reference not importable.
Run without a GPU; the CUDA kernels are then checked against the regenerated goldens by the ordinary
`pytest -m gpu` run (tests/test_gpu_quant.py picks up reference_*.npz automatically).
"""
import argparse
import glob
import importlib
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

CANDIDATES = {   # spellings seen in int8 dynamic-quant code bases; extend when the real names are known
    "act": ["quantize_per_token", "dynamically_quantize_per_token", "quant_per_token", "quantize_activation_per_token",
            "quantize_act", "per_token_quant", "dynamic_quant"],
    "weight": ["quantize_per_channel", "dynamically_quantize_per_channel", "quant_per_channel", "quantize_weight",
               "per_channel_quant"],
}


def find_callable(pkg_names, names):
    import pkgutil
    for pkg_name in pkg_names:
        try:
            pkg = importlib.import_module(pkg_name)
        except Exception as ex:   # noqa: BLE001
            print(f"  (cannot import {pkg_name}: {ex!r})")
            continue
        mods = [pkg]
        if hasattr(pkg, "__path__"):
            for m in pkgutil.walk_packages(pkg.__path__, pkg.__name__ + "."):
                try:
                    mods.append(importlib.import_module(m.name))
                except Exception:   # noqa: BLE001 - optional dependencies (triton, CUDA extensions) may be missing
                    pass
        for mod in mods:
            for n in names:
                fn = getattr(mod, n, None)
                if callable(fn):
                    return f"{mod.__name__}:{n}", fn
    return None, None


def load_spec(spec):
    mod, fn = spec.split(":")
    return spec, getattr(importlib.import_module(mod), fn)


def normalise(result):
    """(int8 ndarray [R, C], fp32 ndarray [R]) from whatever the reference returns."""
    import torch
    if hasattr(result, "int_repr"):
        q = result.int_repr()
        s = getattr(result, "scale", None)
        if s is None:
            s = result.q_per_channel_scales() if hasattr(result, "q_per_channel_scales") else result.q_scale()
        result = (q, s)
    parts = [r for r in result if isinstance(r, torch.Tensor)]
    q = next(p for p in parts if p.dtype == torch.int8)
    s = next(p for p in parts if p.is_floating_point())
    return q.cpu().numpy(), s.detach().to(torch.float32).flatten().cpu().numpy()


def inputs():
    import torch
    from conftest import load_golden_x
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "torch_ao_*.npz")) +
                       glob.glob(os.path.join(ROOT, "tests", "golden", "exact_*.npz"))):
        d = np.load(path)
        yield os.path.basename(path), load_golden_x(d)
    g = torch.Generator().manual_seed(0)
    for dt in (torch.float32, torch.bfloat16, torch.float16):
        x = torch.randn(64, 1031, generator=g) * torch.logspace(-3, 3, 64)[:, None]
        x[1].zero_()
        x[2] = torch.round(x[2]) / 2
        yield f"seeded_{str(dt)[6:]}_64x1031", x.to(dt)


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("reference_root")
    ap.add_argument("--package", default="protoquant", help="import name of the reference package")
    ap.add_argument("--act-fn", help="module:function of the per-token quantiser (skips the search)")
    ap.add_argument("--weight-fn", help="module:function of the per-channel quantiser")
    ap.add_argument("--write", action="store_true", help="regenerate goldens and write the pinned spec")
    args = ap.parse_args()

    if not os.path.isdir(args.reference_root) or not any(
            f.endswith(".py") for _, _, fs in os.walk(args.reference_root) for f in fs):
        print(f"repin: {args.reference_root} holds no Python sources -- the reference is still absent (SURVEY.md §0)")
        return 2
    sys.path.insert(0, args.reference_root)
    import protoquant_oracle as O
    found = {}
    for kind, opt in (("act", args.act_fn), ("weight", args.weight_fn)):
        name, fn = load_spec(opt) if opt else find_callable([args.package], CANDIDATES[kind])
        print(f"{kind} quantiser: {name or 'NOT FOUND (pass --' + kind + '-fn module:function)'}")
        if fn is not None:
            found[kind] = (name, fn)
    if "act" not in found:
        return 2

    knobs = [O.QuantSpec(scale_mode=m, eps=e, qmin=q) for m, e, q in
             itertools.product((O.DIV, O.RCP_MUL, O.INV_SCALE), (0.0, 1e-5, float(np.finfo(np.float32).eps)), (-128, -127))]
    alive = {k: True for k in knobs}
    first_diff = None
    outputs = []
    for label, x in inputs():
        q_ref, s_ref = normalise(found["act"][1](x))
        outputs.append((label, x, q_ref, s_ref))
        for k in knobs:
            if not alive[k]:
                continue
            q, s = O.quantize_rowwise(x, k)
            both_nan = np.isnan(s) & np.isnan(s_ref)
            same = np.array_equal(q, q_ref) and np.array_equal(np.where(both_nan, 0, s.view(np.uint32)),
                                                               np.where(both_nan, 0, s_ref.view(np.uint32)))
            if not same:
                alive[k] = False
                if first_diff is None or k == O.SPEC_V0:
                    bad = np.argwhere(q != q_ref)
                    r, c = (int(bad[0][0]), int(bad[0][1])) if len(bad) else (int(np.argwhere(s != s_ref)[0][0]), 0)
                    first_diff = (label, k, r, c, float(O.to_f32(x)[r, c]), int(q_ref[r, c]), float(s_ref[r]), int(q[r, c]), float(s[r]))
    winners = [k for k in knobs if alive[k]]
    if not winners:
        label, k, r, c, xv, qr, sr, qo, so = first_diff
        print(f"NO knob set reproduces the reference.  First difference ({label}, knobs {k}): row {r} col {c} x={xv!r}: "
              f"reference q={qr} s={sr!r}, oracle q={qo} s={so!r}.  Add the missing knob to oracle/protoquant_oracle.py, "
              f"quant_math.cuh and pq_quant_spec, then re-run.")
        return 1
    pin = winners[0]
    print(f"PINNED: {pin} reproduces the reference bit for bit on {len(outputs)} input sets "
          f"({len(winners)} knob set(s) are indistinguishable on these inputs)")
    if args.write:
        for label, x, q_ref, s_ref in outputs:
            import torch
            name = {torch.float32: "f32", torch.bfloat16: "bf16", torch.float16: "f16"}[x.dtype]
            bits = x.numpy().view(np.uint32) if name == "f32" else x.view(torch.int16).numpy().view(np.uint16)
            np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"reference_{label.replace('.npz', '')}.npz"),
                                x_bits=bits, q=q_ref, s=s_ref, shape=np.array(x.shape), dtype=name,
                                spec=np.array([pin.scale_mode, pin.eps, pin.qmin], dtype=np.float64))
        with open(os.path.join(ROOT, "protoquant_b200", "_pinned_spec.json"), "w") as f:
            json.dump({"scale_mode": pin.scale_mode, "eps": pin.eps, "qmin": pin.qmin,
                       "pinned_against": {k: v[0] for k, v in found.items()}}, f, indent=1)
        print("wrote tests/golden/reference_*.npz and protoquant_b200/_pinned_spec.json; now fix the alias table in "
              "protoquant_b200/compat.py by hand:", {k: v[0] for k, v in found.items()})
    return 0


if __name__ == "__main__":
    sys.exit(main())
