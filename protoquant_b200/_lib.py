"""ctypes binding of libprotoquant_b200.so (the C ABI in include/protoquant_b200.h).

There is no fallback: if the shared library has not been built, or the process has no
sm_100 GPU, every entry point raises.  Build with ``python -c "import __graft_entry__ as g;
g.build()"`` or ``make -C protoquant_b200/csrc``.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# PQ_LIB_PATH: development override (A/B builds of the same ABI, e.g. tools/ab_build.sh); the product loads the in-tree library
LIB_PATH = os.environ.get("PQ_LIB_PATH") or os.path.join(_HERE, "libprotoquant_b200.so")

PQ_F32, PQ_F16, PQ_BF16, PQ_I32 = 0, 1, 2, 3
PQ_DIV, PQ_RCP_MUL, PQ_INV_SCALE = 0, 1, 2
PQ_ACT_IDENTITY, PQ_ACT_SILU, PQ_ACT_GELU, PQ_ACT_GELU_TANH = 0, 1, 2, 3
PQ_MULTI_MULTICAST = 1

# every symbol include/protoquant_b200.h declares (tests/test_abi.py checks the .so exports them)
EXPORTS = (
    "pq_version", "pq_last_error", "pq_launch_count", "pq_act_quant", "pq_weight_quant",
    "pq_qgemm", "pq_qgemm_multi", "pq_qlinear_multi", "pq_qgemm_i32", "pq_dequant", "pq_qlinear",
    "pq_linear_create", "pq_linear_forward_host", "pq_linear_destroy",
    "pq_norm_quant", "pq_act_mul_quant",
    "pq_row_absmax", "pq_act_quant_amax", "pq_qgemm_i32_scatter", "pq_reduce_dequant",
    "pq_symm_barrier", "pq_rowparallel_forward",
)


class PQQuantSpec(ctypes.Structure):
    _fields_ = [("scale_mode", ctypes.c_int32), ("eps", ctypes.c_float), ("qmin", ctypes.c_int32)]


class PQSymmGroup(ctypes.Structure):
    """pq_symm_group: per-rank peer-mapped addresses of the symmetric buffers of a row-parallel layer."""
    _fields_ = [("rank", ctypes.c_int32), ("world", ctypes.c_int32),
                ("inbox", ctypes.c_void_p * 8), ("out", ctypes.c_void_p * 8), ("amax", ctypes.c_void_p * 8),
                ("pads", ctypes.c_void_p * 8), ("cap", ctypes.c_int64)]


class ProtoquantError(RuntimeError):
    pass


_lib = None
_lock = threading.Lock()


def _declare(lib):
    c = ctypes
    vp, i64, i32 = c.c_void_p, c.c_int64, c.c_int
    specp = c.POINTER(PQQuantSpec)
    lib.pq_version.restype = i32
    lib.pq_version.argtypes = []
    lib.pq_last_error.restype = c.c_char_p
    lib.pq_last_error.argtypes = []
    lib.pq_launch_count.restype = c.c_uint64
    lib.pq_launch_count.argtypes = []
    lib.pq_act_quant.restype = i32
    lib.pq_act_quant.argtypes = [vp, i32, i64, i64, i64, vp, i64, vp, i32, specp, vp]
    lib.pq_weight_quant.restype = i32
    lib.pq_weight_quant.argtypes = [vp, i32, i64, i64, i64, vp, i64, vp, specp, vp]
    lib.pq_qgemm.restype = i32
    lib.pq_qgemm.argtypes = [vp, i64, vp, i64, vp, vp, vp, vp, i32, i64, i64, i64, i64, vp]
    lib.pq_qgemm_multi.restype = i32
    lib.pq_qgemm_multi.argtypes = [vp, i64, vp, i64, vp, vp, vp, c.POINTER(vp), i32, i32, i64, i64, i64, i64, vp]
    lib.pq_qlinear_multi.restype = i32
    lib.pq_qlinear_multi.argtypes = [vp, i32, i64, vp, i64, vp, vp, c.POINTER(vp), i32, i32, i64, vp, vp, i64, i64, i64, specp, i32, vp]
    lib.pq_qgemm_i32.restype = i32
    lib.pq_qgemm_i32.argtypes = [vp, i64, vp, i64, vp, i64, i64, i64, i64, vp]
    lib.pq_dequant.restype = i32
    lib.pq_dequant.argtypes = [vp, i64, vp, i32, vp, i32, i64, i64, i64, vp]
    lib.pq_qlinear.restype = i32
    lib.pq_qlinear.argtypes = [vp, i32, i64, vp, i64, vp, vp, vp, i32, i64, vp, vp, i64, i64, i64, specp, vp]
    lib.pq_linear_create.restype = i32
    lib.pq_linear_create.argtypes = [c.POINTER(vp), vp, i32, i64, i64, vp, i64, i32, i32, specp]
    lib.pq_linear_forward_host.restype = i32
    lib.pq_linear_forward_host.argtypes = [vp, vp, vp, i64]
    lib.pq_linear_destroy.restype = None
    lib.pq_linear_destroy.argtypes = [vp]
    lib.pq_norm_quant.restype = i32
    lib.pq_norm_quant.argtypes = [vp, i32, i64, i64, i64, vp, vp, c.c_float, vp, i64, vp, vp, i64, specp, vp]
    lib.pq_row_absmax.restype = i32
    lib.pq_row_absmax.argtypes = [vp, i32, i64, i64, i64, vp, vp]
    lib.pq_act_quant_amax.restype = i32
    lib.pq_act_quant_amax.argtypes = [vp, i32, i64, i64, i64, vp, vp, i64, vp, specp, vp]
    lib.pq_qgemm_i32_scatter.restype = i32
    lib.pq_qgemm_i32_scatter.argtypes = [vp, i64, vp, i64, c.POINTER(vp), i32, i64, i64, i64, i64, i64, vp]
    lib.pq_reduce_dequant.restype = i32
    lib.pq_reduce_dequant.argtypes = [c.POINTER(vp), i32, i64, vp, vp, vp, c.POINTER(vp), i32, i32, i64, i64, i64, vp]
    lib.pq_act_mul_quant.restype = i32
    lib.pq_act_mul_quant.argtypes = [vp, vp, i32, i32, i64, i64, i64, i64, vp, i64, vp, vp, i64, specp, vp]
    sgp = c.POINTER(PQSymmGroup)
    lib.pq_symm_barrier.restype = i32
    lib.pq_symm_barrier.argtypes = [sgp, i32, vp]
    lib.pq_rowparallel_forward.restype = i32
    lib.pq_rowparallel_forward.argtypes = [vp, vp, i32, i32, i64, i64, i32, i64, i64, vp, i64, vp, vp, sgp, i32, vp, i32, i64,
                                           vp, vp, vp, i64, i64, i64, i64, specp, vp]
    if hasattr(lib, "pq_debug_set_quant_config"):
        lib.pq_debug_set_quant_config.restype = None
        lib.pq_debug_set_quant_config.argtypes = [i32, i32]
    if hasattr(lib, "pq_debug_set_timeline"):
        lib.pq_debug_set_timeline.restype = None
        lib.pq_debug_set_timeline.argtypes = [vp]
    for dbg in ("pq_debug_set_gemm_config", "pq_debug_set_streamk", "pq_debug_set_pdl", "pq_debug_set_staged", "pq_debug_set_prefetch", "pq_debug_set_fused_decode", "pq_debug_set_tma_store", "pq_debug_set_epilogue", "pq_debug_set_narrow_tiles", "pq_debug_set_weight_prefetch", "pq_debug_set_multi_tma", "pq_debug_set_quant_staged", "pq_debug_set_smallm_splits", "pq_debug_set_multi_bn", "pq_debug_set_tile_rotation"):
        if hasattr(lib, dbg):
            getattr(lib, dbg).restype = None
            getattr(lib, dbg).argtypes = [i32]


def lib():
    """The loaded shared library; raises ProtoquantError if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise ProtoquantError(
                        f"{LIB_PATH} not found: the CUDA extension is not built and protoquant_b200 "
                        "has no CPU fallback. Run __graft_entry__.build() (or make -C protoquant_b200/csrc).")
                handle = ctypes.CDLL(LIB_PATH)
                _declare(handle)
                _lib = handle
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().pq_last_error()
        raise ProtoquantError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(lib().pq_launch_count())
