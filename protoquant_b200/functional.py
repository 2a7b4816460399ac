"""Functional API of the dynamic-quantized linear path.

Every function takes CUDA tensors, enqueues on the current CUDA stream and never
synchronises.  CPU tensors raise: there is no CPU path in this package.
"""
from __future__ import annotations

import ctypes
import dataclasses
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import PQ_BF16, PQ_F16, PQ_F32, PQQuantSpec

_DT = {torch.float32: PQ_F32, torch.float16: PQ_F16, torch.bfloat16: PQ_BF16}


@dataclasses.dataclass(frozen=True)
class QuantSpec:
    """Scale-computation knobs.  Default == SPEC v0 (true division, no eps)."""
    scale_mode: int = _lib.PQ_DIV
    eps: float = 0.0
    qmin: int = -128

    def c(self) -> PQQuantSpec:
        return PQQuantSpec(self.scale_mode, self.eps, self.qmin)


def _default_spec() -> QuantSpec:
    """SPEC v0, unless tools/repin.py has pinned the knobs against the real reference (protoquant_b200/_pinned_spec.json)."""
    import json
    import os
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_pinned_spec.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return QuantSpec(int(d["scale_mode"]), float(d["eps"]), int(d["qmin"]))
    return QuantSpec()


DEFAULT_SPEC = _default_spec()


def _stream(device=None) -> ctypes.c_void_p:
    """The current stream of `device` (default: the current device) as the void* the C ABI takes."""
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(t: torch.Tensor, name: str):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.ProtoquantError(f"{name} must be a CUDA tensor: protoquant_b200 has no CPU fallback")


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def _on(t: torch.Tensor, *others):
    """Device guard for one C-ABI call: the library launches on the CURRENT device, so make the device that owns
    `t` current for the duration of the call (a no-op when it already is), and insist that every other tensor
    argument lives on the same device -- single-process model parallelism (HF device_map, pipeline stages) calls
    modules whose weights are on cuda:1 while cuda:0 is current."""
    dev = t.device
    for o in others:
        if o is not None and o.device != dev:
            raise _lib.ProtoquantError(f"all tensor arguments must be on one device: got {dev} and {o.device}")
    if dev.index == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(dev)


def _pad16(k: int) -> int:
    return (k + 15) // 16 * 16


def _specp(spec: Optional[QuantSpec]):
    return ctypes.byref((spec or DEFAULT_SPEC).c())


def _rows2d(x: torch.Tensor) -> torch.Tensor:
    if x.dim() != 2:
        raise ValueError(f"expected a 2-D tensor, got shape {tuple(x.shape)}")
    if x.stride(1) != 1:
        x = x.contiguous()
    return x


def alloc_q(rows: int, cols: int, device) -> torch.Tensor:
    """int8 [rows, cols] whose row stride is padded to a multiple of 16 bytes (TMA requirement)."""
    buf = torch.empty((rows, _pad16(cols)), dtype=torch.int8, device=device)
    return buf[:, :cols]


def quantize_act(x: torch.Tensor, transpose: bool = False, spec: Optional[QuantSpec] = None,
                 out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-token int8 quantisation of x [M,K] -> (xq int8 [M,K] (or [K,M] if transpose), s_x fp32 [M])."""
    _require_cuda(x, "x")
    x = _rows2d(x)
    if x.dtype not in _DT:
        raise TypeError(f"unsupported activation dtype {x.dtype}")
    M, K = x.shape
    if out is None:
        xq = alloc_q(K, M, x.device) if transpose else alloc_q(M, K, x.device)
        s = torch.empty((M,), dtype=torch.float32, device=x.device)
    else:
        xq, s = out
    if M == 0:
        return xq, s
    with _on(x, xq, s):
        rc = _lib.lib().pq_act_quant(x.data_ptr(), _DT[x.dtype], M, K, x.stride(0), xq.data_ptr(), xq.stride(0),
                                     s.data_ptr(), int(transpose), _specp(spec), _stream(x.device))
    _lib.check(rc, "pq_act_quant")
    return xq, s


def quantize_weight(w: torch.Tensor, spec: Optional[QuantSpec] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-output-channel int8 quantisation of W [N,K] -> (Wq int8 [N,K], s_w fp32 [N])."""
    _require_cuda(w, "w")
    w = _rows2d(w)
    if w.dtype not in _DT:
        raise TypeError(f"unsupported weight dtype {w.dtype}")
    N, K = w.shape
    wq = alloc_q(N, K, w.device)
    s = torch.empty((N,), dtype=torch.float32, device=w.device)
    if N:
        with _on(w):
            rc = _lib.lib().pq_weight_quant(w.data_ptr(), _DT[w.dtype], N, K, w.stride(0), wq.data_ptr(), wq.stride(0),
                                            s.data_ptr(), _specp(spec), _stream(w.device))
        _lib.check(rc, "pq_weight_quant")
    return wq, s


def _gemm_operand(t: torch.Tensor, name: str) -> torch.Tensor:
    _require_cuda(t, name)
    if t.dtype != torch.int8 or t.dim() != 2:
        raise TypeError(f"{name} must be a 2-D int8 tensor")
    if t.stride(1) != 1 or t.stride(0) % 16 or t.data_ptr() % 16:
        p = alloc_q(t.shape[0], t.shape[1], t.device)
        p.copy_(t)
        t = p
    return t


def qgemm_i32(xq: torch.Tensor, wq: torch.Tensor) -> torch.Tensor:
    """Exact int32 accumulators acc[M,N] = xq[M,K] . wq[N,K]^T (parity hook)."""
    xq = _gemm_operand(xq, "xq")
    wq = _gemm_operand(wq, "wq")
    M, K = xq.shape
    N = wq.shape[0]
    if wq.shape[1] != K:
        raise ValueError("K mismatch")
    acc = torch.empty((M, N), dtype=torch.int32, device=xq.device)
    if M and N:
        with _on(xq, wq):
            rc = _lib.lib().pq_qgemm_i32(xq.data_ptr(), xq.stride(0), wq.data_ptr(), wq.stride(0), acc.data_ptr(), N,
                                         M, N, K, _stream(xq.device))
        _lib.check(rc, "pq_qgemm_i32")
    return acc


def qgemm(xq: torch.Tensor, s_x: torch.Tensor, wq: torch.Tensor, s_w: torch.Tensor,
          bias: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.bfloat16,
          out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y[M,N] = cast((float(xq.wq^T) * s_x[m]) * s_w[n] + bias[n])."""
    xq = _gemm_operand(xq, "xq")
    wq = _gemm_operand(wq, "wq")
    M, K = xq.shape
    N = wq.shape[0]
    if wq.shape[1] != K:
        raise ValueError("K mismatch")
    if out_dtype not in _DT:
        raise TypeError(f"unsupported output dtype {out_dtype}")
    for t, n, ln in ((s_x, "s_x", M), (s_w, "s_w", N)):
        _require_cuda(t, n)
        if t.dtype != torch.float32 or t.numel() != ln or not t.is_contiguous():
            raise TypeError(f"{n} must be a contiguous fp32 tensor of {ln} elements")
    if bias is not None:
        _require_cuda(bias, "bias")
        if bias.dtype != torch.float32 or not bias.is_contiguous():
            bias = bias.to(torch.float32).contiguous()
        if bias.numel() != N:
            raise ValueError("bias size mismatch")
    y = out if out is not None else torch.empty((M, N), dtype=out_dtype, device=xq.device)
    if M and N:
        with _on(xq, wq, s_x, s_w, bias, y):
            rc = _lib.lib().pq_qgemm(xq.data_ptr(), xq.stride(0), wq.data_ptr(), wq.stride(0), s_x.data_ptr(),
                                     s_w.data_ptr(), bias.data_ptr() if bias is not None else None,
                                     y.data_ptr(), _DT[y.dtype], y.stride(0), M, N, K, _stream(xq.device))
        _lib.check(rc, "pq_qgemm")
    return y


def qgemm_multi(xq: torch.Tensor, s_x: torch.Tensor, wq: torch.Tensor, s_w: torch.Tensor,
                bias: Optional[torch.Tensor], dest_ptrs, ldy: int, out_dtype: torch.dtype) -> None:
    """Fused GEMM + all-gather store: the [M,N] result is written to every raw device address in
    `dest_ptrs` (row stride `ldy` elements).  The addresses are peer-mapped (NVLink) or NVSwitch
    multicast pointers into the ranks' full output buffers, already offset to this rank's first
    column; the caller issues the cross-rank barrier afterwards."""
    xq = _gemm_operand(xq, "xq")
    wq = _gemm_operand(wq, "wq")
    M, K = xq.shape
    N = wq.shape[0]
    if not (1 <= len(dest_ptrs) <= 8):
        raise ValueError("1..8 destinations")
    if bias is not None and (bias.dtype != torch.float32 or not bias.is_contiguous()):
        bias = bias.to(torch.float32).contiguous()
    arr = (ctypes.c_void_p * len(dest_ptrs))(*[ctypes.c_void_p(int(p)) for p in dest_ptrs])
    if M and N:
        with _on(xq, wq, s_x, s_w, bias):
            rc = _lib.lib().pq_qgemm_multi(xq.data_ptr(), xq.stride(0), wq.data_ptr(), wq.stride(0), s_x.data_ptr(),
                                           s_w.data_ptr(), bias.data_ptr() if bias is not None else None,
                                           arr, len(dest_ptrs), _DT[out_dtype], ldy, M, N, K, _stream(xq.device))
        _lib.check(rc, "pq_qgemm_multi")


def qlinear(x: torch.Tensor, wq: torch.Tensor, s_w: torch.Tensor, bias: Optional[torch.Tensor] = None,
            out_dtype: Optional[torch.dtype] = None, spec: Optional[QuantSpec] = None) -> torch.Tensor:
    """Dynamic-quant linear forward: x[...,K] -> y[...,N] with cached (wq, s_w)."""
    _require_cuda(x, "x")
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    xq, s_x = quantize_act(x2, spec=spec)
    y = qgemm(xq, s_x, wq, s_w, bias, out_dtype or x.dtype)
    return y.reshape(*lead, wq.shape[0])


def qlinear_into(x2: torch.Tensor, wq_storage: torch.Tensor, in_features: int, s_w: torch.Tensor,
                 bias: Optional[torch.Tensor], y: torch.Tensor, xq_ws: torch.Tensor, sx_ws: torch.Tensor,
                 spec: Optional[QuantSpec] = None) -> torch.Tensor:
    """One `pq_qlinear` call (act-quant launch + GEMM launch, or the fused decode kernel) on caller-owned
    buffers: x2 [M,K] row-major, wq_storage [N, ld>=K] int8 with a 16-byte row stride, y [M,N],
    xq_ws [M, ld16(K)] int8 and sx_ws [M] fp32 scratch.  No allocation, no checks beyond the C ABI's."""
    M, N, K = x2.shape[0], wq_storage.shape[0], in_features
    if M:
        with _on(x2, wq_storage, s_w, bias, y, xq_ws, sx_ws):
            rc = _lib.lib().pq_qlinear(x2.data_ptr(), _DT[x2.dtype], x2.stride(0), wq_storage.data_ptr(), wq_storage.stride(0),
                                       s_w.data_ptr(), bias.data_ptr() if bias is not None else None,
                                       y.data_ptr(), _DT[y.dtype], y.stride(0), xq_ws.data_ptr(), sx_ws.data_ptr(), M, N, K,
                                       _specp(spec), _stream(x2.device))
        _lib.check(rc, "pq_qlinear")
    return y


def qlinear_multi_into(x2: torch.Tensor, wq_storage: torch.Tensor, in_features: int, s_w: torch.Tensor,
                       bias: Optional[torch.Tensor], dest_ptrs, ldy: int, out_dtype: torch.dtype,
                       xq_ws: torch.Tensor, sx_ws: torch.Tensor, spec: Optional[QuantSpec] = None,
                       multicast: bool = False) -> None:
    """One `pq_qlinear_multi` call: act-quant into the caller's workspace + GEMM whose epilogue stores the [M, N]
    result into every raw device address of `dest_ptrs` (row stride `ldy` elements): the local output buffer and
    the peers' buffers over NVLink (coalesced 256-byte peer stores from the CTA-staged tile), or, with `multicast`, one NVSwitch multicast address
    (multimem.st).  The caller issues the cross-rank barrier afterwards."""
    _require_cuda(x2, "x")
    M, N, K = x2.shape[0], wq_storage.shape[0], in_features
    if x2.dim() != 2 or x2.stride(1) != 1 or x2.dtype not in _DT:
        raise TypeError("x must be a 2-D tensor of a supported dtype with unit column stride")
    if not (1 <= len(dest_ptrs) <= 8) or (multicast and len(dest_ptrs) != 1):
        raise ValueError("1..8 destinations (exactly 1 with multicast)")
    for t, n in ((s_w, "s_w"), (bias, "bias")):
        if t is not None and (t.dtype != torch.float32 or not t.is_contiguous()):
            raise TypeError(f"{n} must be a contiguous fp32 tensor")
    if xq_ws.shape[0] < M or xq_ws.shape[1] < _pad16(K) or xq_ws.stride(0) % 16 or sx_ws.numel() < M:
        raise ValueError("act-quant workspace too small")
    arr = (ctypes.c_void_p * len(dest_ptrs))(*[ctypes.c_void_p(int(p)) for p in dest_ptrs])
    if M and N:
        with _on(x2, wq_storage, s_w, bias, xq_ws, sx_ws):
            rc = _lib.lib().pq_qlinear_multi(x2.data_ptr(), _DT[x2.dtype], x2.stride(0), wq_storage.data_ptr(),
                                             wq_storage.stride(0), s_w.data_ptr(),
                                             bias.data_ptr() if bias is not None else None, arr, len(dest_ptrs),
                                             _DT[out_dtype], ldy, xq_ws.data_ptr(), sx_ws.data_ptr(), M, N, K,
                                             _specp(spec), _lib.PQ_MULTI_MULTICAST if multicast else 0,
                                             _stream(x2.device))
        _lib.check(rc, "pq_qlinear_multi")


def row_absmax(x: torch.Tensor) -> torch.Tensor:
    """amax[m] = max_k |x[m,k]| in fp32 (the per-token statistic a K-sharded quantizer all-reduces)."""
    _require_cuda(x, "x")
    x2 = _rows2d(x) if x.dim() == 2 and x.stride(1) == 1 else x.reshape(-1, x.shape[-1]).contiguous()
    M, K = x2.shape
    amax = torch.empty((M,), dtype=torch.float32, device=x.device)
    if M:
        with _on(x2):
            rc = _lib.lib().pq_row_absmax(x2.data_ptr(), _DT[x2.dtype], M, K, x2.stride(0), amax.data_ptr(), _stream(x2.device))
        _lib.check(rc, "pq_row_absmax")
    return amax


def quantize_act_with_amax(x: torch.Tensor, amax: torch.Tensor, spec: Optional[QuantSpec] = None):
    """Quantise x [M,Ks] (a K-slice of the activation) with the given per-row maxima of the WHOLE row."""
    _require_cuda(x, "x")
    if x.dim() != 2 or x.stride(1) != 1:
        x = x.reshape(-1, x.shape[-1]).contiguous()
    M, K = x.shape
    if amax.dtype != torch.float32 or amax.numel() != M or not amax.is_contiguous():
        raise TypeError("amax must be a contiguous fp32 tensor with one entry per row")
    xq = alloc_q(M, K, x.device)
    s_x = torch.empty((M,), dtype=torch.float32, device=x.device)
    if M:
        with _on(x, amax):
            rc = _lib.lib().pq_act_quant_amax(x.data_ptr(), _DT[x.dtype], M, K, x.stride(0), amax.data_ptr(),
                                              xq.data_ptr(), xq.stride(0), s_x.data_ptr(), _specp(spec), _stream(x.device))
        _lib.check(rc, "pq_act_quant_amax")
    return xq, s_x


def qgemm_i32_scatter(xq: torch.Tensor, wq: torch.Tensor, dest_ptrs, ld_dest: int, cols_per_dest: int) -> None:
    """int32 GEMM whose output columns [d*cols_per_dest, (d+1)*cols_per_dest) are stored to the raw device
    address dest_ptrs[d] as an [M, cols_per_dest] int32 matrix (row stride ld_dest): the GEMM half of the fused
    GEMM + reduce-scatter (dest_ptrs are peer-mapped inboxes)."""
    xq = _gemm_operand(xq, "xq")
    wq = _gemm_operand(wq, "wq")
    M, K = xq.shape
    N = wq.shape[0]
    arr = (ctypes.c_void_p * len(dest_ptrs))(*[ctypes.c_void_p(int(p)) for p in dest_ptrs])
    if M and N:
        with _on(xq, wq):
            rc = _lib.lib().pq_qgemm_i32_scatter(xq.data_ptr(), xq.stride(0), wq.data_ptr(), wq.stride(0), arr,
                                                 len(dest_ptrs), ld_dest, cols_per_dest, M, N, K, _stream(xq.device))
        _lib.check(rc, "pq_qgemm_i32_scatter")


def reduce_dequant(part_ptrs, ld_part: int, s_x: torch.Tensor, s_w: torch.Tensor, bias: Optional[torch.Tensor],
                   dest_ptrs, ldy: int, out_dtype: torch.dtype, M: int, N: int) -> None:
    """y = cast(((float(sum of the int32 parts) * s_x[m]) * s_w[n]) + bias[n]) written to every dest_ptrs[d]."""
    pa = (ctypes.c_void_p * len(part_ptrs))(*[ctypes.c_void_p(int(p)) for p in part_ptrs])
    da = (ctypes.c_void_p * len(dest_ptrs))(*[ctypes.c_void_p(int(p)) for p in dest_ptrs])
    if M and N:
        with _on(s_x, s_w, bias):
            rc = _lib.lib().pq_reduce_dequant(pa, len(part_ptrs), ld_part, s_x.data_ptr(), s_w.data_ptr(),
                                              bias.data_ptr() if bias is not None else None, da, len(dest_ptrs),
                                              _DT[out_dtype], ldy, M, N, _stream(s_x.device))
        _lib.check(rc, "pq_reduce_dequant")


def dequant_accumulators(acc: torch.Tensor, s_x: torch.Tensor, s_w: torch.Tensor, bias: Optional[torch.Tensor] = None,
                         out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """The fused epilogue as a stand-alone op on int32 accumulators acc [M,N] (e.g. after an exact all-reduce)."""
    _require_cuda(acc, "acc")
    if acc.dtype != torch.int32 or acc.dim() != 2 or acc.stride(1) != 1:
        raise TypeError("acc must be a 2-D int32 tensor with unit column stride")
    M, N = acc.shape
    y = torch.empty((M, N), dtype=out_dtype, device=acc.device)
    reduce_dequant([acc.data_ptr()], acc.stride(0), s_x, s_w, bias, [y.data_ptr()], N, out_dtype, M, N)
    return y


_ACTS = {"identity": _lib.PQ_ACT_IDENTITY, "silu": _lib.PQ_ACT_SILU, "gelu": _lib.PQ_ACT_GELU,
         "gelu_tanh": _lib.PQ_ACT_GELU_TANH}


def _fused_out(M: int, K: int, like: torch.Tensor, return_float: bool, out):
    if out is not None:
        xq, s_x = out
    else:
        xq = alloc_q(M, K, like.device)
        s_x = torch.empty((M,), dtype=torch.float32, device=like.device)
    y = torch.empty((M, K), dtype=like.dtype, device=like.device) if return_float else None
    return xq, s_x, y


def norm_quant(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, eps: float = 1e-6,
               spec: Optional[QuantSpec] = None, return_normed: bool = False,
               out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
    """Fused RMSNorm (bias=None, Llama) / LayerNorm (bias given, BERT) + per-token int8 quantisation
    (SURVEY.md §8f-2): x[..., K] -> (xq int8 [M,K], s_x fp32 [M]) ready for `qgemm`, plus the normalised
    tensor itself (dtype of x) when `return_normed`.  (xq, s_x) is bit-identical to
    `quantize_act(normed)`; one read of x instead of a norm kernel + a quantizer kernel."""
    _require_cuda(x, "x")
    if x.dtype not in _DT:
        raise TypeError(f"unsupported dtype {x.dtype}")
    K = x.shape[-1]
    x2 = _rows2d(x.reshape(-1, K))
    M = x2.shape[0]
    w = weight.to(device=x.device, dtype=x.dtype).contiguous()
    b = bias.to(device=x.device, dtype=x.dtype).contiguous() if bias is not None else None
    if w.numel() != K or (b is not None and b.numel() != K):
        raise ValueError("norm weight / bias must have K elements")
    xq, s_x, y = _fused_out(M, K, x2, return_normed, out)
    if M:
        with _on(x2, w, b, xq, s_x):
            rc = _lib.lib().pq_norm_quant(x2.data_ptr(), _DT[x2.dtype], M, K, x2.stride(0), w.data_ptr(),
                                          b.data_ptr() if b is not None else None, float(eps),
                                          xq.data_ptr(), xq.stride(0), s_x.data_ptr(),
                                          y.data_ptr() if y is not None else None, K, _specp(spec), _stream(x2.device))
        _lib.check(rc, "pq_norm_quant")
    return (xq, s_x, y.reshape(x.shape)) if return_normed else (xq, s_x)


def rmsnorm_quant(x, weight, eps: float = 1e-6, **kw):
    return norm_quant(x, weight, None, eps, **kw)


def layernorm_quant(x, weight, bias, eps: float = 1e-5, **kw):
    return norm_quant(x, weight, bias, eps, **kw)


def act_mul_quant(gate: torch.Tensor, up: Optional[torch.Tensor] = None, act: str = "silu",
                  spec: Optional[QuantSpec] = None, return_float: bool = False,
                  out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
    """Fused act(gate) [* up] + per-token int8 quantisation: the Llama MLP's `silu(gate) * up` (or BERT's
    `gelu(x)`) written straight as the int8 operand of the down projection.  gate / up may be column
    slices of one [M, 2K] GEMM output (any row stride that keeps rows 16-byte aligned)."""
    _require_cuda(gate, "gate")
    if gate.dtype not in _DT:
        raise TypeError(f"unsupported dtype {gate.dtype}")
    if act not in _ACTS:
        raise ValueError(f"act must be one of {sorted(_ACTS)}")
    K = gate.shape[-1]
    g2 = gate.reshape(-1, K) if gate.dim() != 2 else gate
    if g2.stride(1) != 1:
        g2 = g2.contiguous()
    u2 = None
    if up is not None:
        _require_cuda(up, "up")
        if up.shape != gate.shape or up.dtype != gate.dtype:
            raise ValueError("up must match gate in shape and dtype")
        u2 = up.reshape(-1, K) if up.dim() != 2 else up
        if u2.stride(1) != 1:
            u2 = u2.contiguous()
    M = g2.shape[0]
    hq, s_h, h = _fused_out(M, K, g2, return_float, out)
    if M:
        with _on(g2, u2, hq, s_h):
            rc = _lib.lib().pq_act_mul_quant(g2.data_ptr(), u2.data_ptr() if u2 is not None else None, _DT[g2.dtype],
                                             _ACTS[act], M, K, g2.stride(0), u2.stride(0) if u2 is not None else 0,
                                             hq.data_ptr(), hq.stride(0), s_h.data_ptr(),
                                             h.data_ptr() if h is not None else None, K, _specp(spec), _stream(g2.device))
        _lib.check(rc, "pq_act_mul_quant")
    return (hq, s_h, h.reshape(gate.shape)) if return_float else (hq, s_h)


def act_mul(gate: torch.Tensor, up: Optional[torch.Tensor] = None, act: str = "silu") -> torch.Tensor:
    """T(T(act(gate)) * up) as a tensor of the input dtype: exactly the tensor `act_mul_quant` quantises (used where
    the quantisation has to wait for a cross-rank row maximum)."""
    return act_mul_quant(gate, up, act, return_float=True)[2]


def dequantize(q: torch.Tensor, s: torch.Tensor, axis: int = 0, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """out[r,c] = q[r,c] * s[r] (axis=0) or q[r,c] * s[c] (axis=1)."""
    _require_cuda(q, "q")
    _require_cuda(s, "s")
    if q.dim() != 2 or q.dtype != torch.int8:
        raise TypeError("q must be a 2-D int8 tensor")
    if q.stride(1) != 1:
        q = q.contiguous()
    rows, cols = q.shape
    if s.dtype != torch.float32 or not s.is_contiguous() or s.numel() != (rows if axis == 0 else cols):
        raise TypeError("s must be contiguous fp32 with one entry per slice along `axis`")
    out = torch.empty((rows, cols), dtype=out_dtype, device=q.device)
    if rows and cols:
        with _on(q, s):
            rc = _lib.lib().pq_dequant(q.data_ptr(), q.stride(0), s.data_ptr(), axis, out.data_ptr(), _DT[out_dtype],
                                       out.stride(0), rows, cols, _stream(q.device))
        _lib.check(rc, "pq_dequant")
    return out
