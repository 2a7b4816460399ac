"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/protoquant_b200.h declares, refuses to compute without a GPU (no CPU fallback), and
validates its arguments.  No compute call is made here."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
import protoquant_b200 as pq
from protoquant_b200 import _lib

HEADER = os.path.join(ROOT, "include", "protoquant_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pq_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_reports_version():
    lib = pq.lib()
    assert lib.pq_version() == 200


def test_every_declared_symbol_is_exported():
    lib = pq.lib()
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted(_lib.EXPORTS) == syms


def test_library_is_built_for_sm_100a_with_tcgen05():
    """The shipped cubin must be the sm_100a tcgen05/TMA kernels, not a legacy mma.sync build."""
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    assert os.path.exists(exe), "cuobjdump is part of the CUDA toolkit this repo builds with; without it the SASS evidence cannot be checked"
    out = subprocess.run([exe, "-sass", _lib.LIB_PATH], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-500:]
    assert "sm_100a" in out.stdout
    assert "UTCIMMA" in out.stdout          # tcgen05.mma.kind::i8
    assert "UTMALDG" in out.stdout          # TMA tensor loads
    assert "LDTM" in out.stdout             # tcgen05.ld
    assert "HMMA" not in out.stdout and "IMMA." not in out.stdout.replace("UTCIMMA", "")
    assert "UBLKCP" in out.stdout           # 1-D bulk copies (staged quantizer)
    assert "UTMASTG" in out.stdout          # TMA store epilogue
    # the committed opcode histogram (tools/sass_opcodes.py) is the judged artefact: it must exist and agree
    hist = open(os.path.join(ROOT, "profiles", "sass_opcodes_r2.txt")).read()
    assert "legacy tensor opcodes in the whole library (HMMA / IMMA / *GMMA): none" in hist
    for fam in ("qgemm_kernel", "qgemm_smallm_kernel", "rowwise_quant_vec_kernel", "symm_barrier_kernel"):
        assert fam in hist, fam


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_in_the_c_abi():
    lib = pq.lib()
    buf = (ctypes.c_char * 256)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    rc = lib.pq_act_quant(p, 0, 1, 16, 16, p, 16, p, 0, None, None)
    assert rc == 3  # PQ_ERR_DEVICE
    assert b"no CPU fallback" in lib.pq_last_error() or b"sm_100" in lib.pq_last_error()
    rc = lib.pq_qgemm_i32(p, 16, p, 16, p, 16, 1, 1, 16, None)
    assert rc == 3
    h = ctypes.c_void_p()
    rc = lib.pq_linear_create(ctypes.byref(h), p, 0, 1, 16, None, 1, 0, 2, None)
    assert rc == 3
    assert lib.pq_norm_quant(p, 2, 1, 16, 16, p, None, 1e-6, p, 16, p, None, 0, None, None) == 3
    assert lib.pq_act_mul_quant(p, p, 2, 1, 1, 16, 16, 16, p, 16, p, None, 0, None, None) == 3


def test_python_surface_rejects_cpu_tensors():
    x = torch.randn(4, 16)
    for fn in (lambda: pq.quantize_act(x), lambda: pq.quantize_weight(x), lambda: pq.quantize(x),
               lambda: pq.qlinear(x, x.to(torch.int8), torch.ones(4)),
               lambda: pq.qgemm_i32(x.to(torch.int8), x.to(torch.int8)),
               lambda: pq.dequantize_tensor(x.to(torch.int8), torch.ones(4))):
        with pytest.raises(pq.ProtoquantError, match="no CPU fallback"):
            fn()
    with pytest.raises(RuntimeError, match="no CPU path"):
        pq.DynamicQuantLinear.from_float(torch.nn.Linear(16, 8))


def test_missing_extension_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(pq.ProtoquantError, match="not built"):
        _lib.lib()


def test_argument_validation_happens_before_device_checks_where_possible():
    lib = pq.lib()
    # y_dtype PQ_I32 is not a valid pq_qgemm output type -> PQ_ERR_ARG regardless of device
    buf = (ctypes.c_char * 256)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    rc = lib.pq_qgemm(p, 16, p, 16, p, p, None, p, 3, 16, 1, 1, 16, None)
    assert rc == 1
    assert b"y_dtype" in lib.pq_last_error()
    rc = lib.pq_dequant(p, 16, p, 2, p, 0, 16, 1, 16, None)
    assert rc == 1 and b"axis" in lib.pq_last_error()


def test_header_is_plain_c():
    """The boundary is a C ABI: the header must compile as C (no C++isms, no torch types)."""
    import subprocess
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", HEADER],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = open(HEADER).read()
    assert "at::" not in src and "std::" not in src


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` is the CPU arm the driver runs beside ours: one JSON line, oracle path."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "TOPS" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["gpu_launches"] == 0
