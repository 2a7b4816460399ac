"""CPU oracle for protoquant's dynamic-quantized linear path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product
(``protoquant_b200``) never does: it has no CPU path at all.

PARITY UNPINNED.  The reference checkout is absent in this environment
(``/root/reference`` holds only CODE_OF_CONDUCT.md — SURVEY.md §0), so no
reference file:line, test or golden vector exists to pin this restatement
against protoquant itself.  What it follows instead:

* BASELINE.json ``north_star``: per-token absmax -> scale -> round-to-nearest-even
  int8 for activations; per-output-channel int8 + scale for weights; exact
  int8 x int8 -> int32; epilogue ``row scale x column scale, bias, cast``.
* SURVEY.md §8c "SPEC v0" (frozen default, every item a knob):
  upcast to fp32 · amax = max|x| over the last dim · s = amax/127 (fp32 true
  division) · amax == 0 -> s = 1 · q = rne(x / s) · clamp [-128,127] · scales
  stored fp32 · epilogue ((float(acc)*s_x)*s_w)+bias in fp32, one RNE cast ·
  NaN / inf / denormal-scale rows: the same formula evaluated literally in IEEE arithmetic
  with NaN -> 0 at the int8 conversion (see quantize_rowwise).
* The nearest *verifiable* definition in this container, a different project:
  ``torch.ao.quantization.fx._decomposed`` ``choose_qparams_per_token`` (:778-810)
  and ``quantize_per_token`` (:930-965).  ``QuantSpec.torch_ao()`` selects the knob
  values that reproduce it bit-exactly; ``tests/golden/make_golden.py`` generated the
  committed golden vectors from those torch ops, and ``tests/test_oracle.py`` pins
  this oracle against them.

* Exact-rational golden vectors (``tests/golden/make_golden_exact.py`` -> ``exact_*.npz``,
  ``epilogue_exact_24x40.npz``): every binary32 operation of SPEC v0 done in exact rational
  arithmetic + integer round-to-nearest-even -- a third, independent producer that pins the
  DEFAULT spec (and the other knob sets), the non-finite / denormal policy, the dequant
  epilogue and dequantize.  It pins this module to the written spec, still not to protoquant.
* ``tools/repin.py <reference_root>`` re-pins the knobs against the real reference the day it
  is mounted.

All arithmetic is explicit numpy float32 (no fused multiply-add, no fast-math).
"""
from __future__ import annotations

import dataclasses
from typing import Optional, Tuple

import numpy as np

try:  # torch is used only for bf16 <-> fp32 conversion and a fast exact int32 matmul
    import torch
except Exception:  # pragma: no cover
    torch = None

DIV, RCP_MUL, INV_SCALE = 0, 1, 2


@dataclasses.dataclass(frozen=True)
class QuantSpec:
    """Knobs of SURVEY.md §8c.  Default == SPEC v0."""

    scale_mode: int = DIV   # DIV: x / s ; RCP_MUL: x * (1/s) ; INV_SCALE: x * (127/amax)
    eps: float = 0.0        # 0 = none; else amax is clamped to >= eps before /127
    qmin: int = -128        # -128 or -127
    qmax: int = 127

    @staticmethod
    def torch_ao() -> "QuantSpec":
        # torch/ao/quantization/fx/_decomposed.py:808 clamp(min=1e-5)/127, :959 x*(1/scale)
        return QuantSpec(scale_mode=RCP_MUL, eps=1e-5, qmin=-128)


SPEC_V0 = QuantSpec()


def to_f32(x) -> np.ndarray:
    """Exact upcast of fp32 / fp16 / bf16 input (numpy array or torch tensor) to float32."""
    if torch is not None and isinstance(x, torch.Tensor):
        return x.detach().to("cpu").to(torch.float32).numpy()
    x = np.asarray(x)
    if x.dtype == np.float32:
        return x
    if x.dtype == np.float16:
        return x.astype(np.float32)
    raise TypeError(f"unsupported input dtype {x.dtype}")


def rowwise_scale(x32: np.ndarray, spec: QuantSpec = SPEC_V0, amax: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
    """(amax_eff, scale) per row of a [R, C] float32 matrix.  `amax` overrides the row maxima (a K-shard of a
    row-parallel layer quantises its slice with the maximum of the whole row)."""
    if amax is not None:
        amax = np.asarray(amax, dtype=np.float32)
    else:
        amax = np.max(np.abs(x32), axis=-1).astype(np.float32) if x32.shape[-1] else np.zeros(x32.shape[:-1], np.float32)
    if spec.eps > 0:
        amax = np.maximum(amax, np.float32(spec.eps))
    with np.errstate(divide="ignore", invalid="ignore"):
        s = (amax / np.float32(127.0)).astype(np.float32)
    s = np.where(amax == 0, np.float32(1.0), s).astype(np.float32)
    return amax, s


def quantize_rowwise(x, spec: QuantSpec = SPEC_V0, amax: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Per-row symmetric int8 quantisation of a 2-D matrix.  Returns (q int8 [R,C], s fp32 [R]).

    Used both for activations (rows = tokens) and for weights
    W[N,K] (rows = output channels, row a2)."""
    x32 = to_f32(x)
    assert x32.ndim == 2
    amax, s = rowwise_scale(x32, spec, amax)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        if spec.scale_mode == DIV:
            r = (x32 / s[:, None]).astype(np.float32)
        elif spec.scale_mode == RCP_MUL:
            inv = (np.float32(1.0) / s).astype(np.float32)
            r = (x32 * inv[:, None]).astype(np.float32)
        elif spec.scale_mode == INV_SCALE:
            inv = np.where(amax == 0, np.float32(1.0), np.float32(127.0) / np.where(amax == 0, np.float32(1), amax)).astype(np.float32)
            r = (x32 * inv[:, None]).astype(np.float32)
        else:
            raise ValueError(spec.scale_mode)
    q = np.rint(r)  # round half to even
    # Non-finite policy (SPEC v0 left it "undefined, documented"; this is the documentation): the formula is
    # evaluated literally in IEEE arithmetic -- amax and s propagate NaN / inf (np.max, np.maximum and the division
    # do), x/inf = 0, inf/inf = NaN, x/0 = +-inf -- and the float -> int8 conversion maps NaN to 0 (what the
    # conversion gives on x86 and on the GPU) and saturates everything else through the clamp.  So a row holding
    # NaN or +-inf gets scale NaN / inf and all-zero codes, and an fp32 row whose scale underflows to 0 (amax <
    # 127 * 2^-150) gets +-qmax / qmin for its non-zero elements.
    q = np.where(np.isnan(q), np.float32(0), q)
    q = np.clip(q, spec.qmin, spec.qmax)
    return q.astype(np.int8), s


def quantize_act(x, spec: QuantSpec = SPEC_V0, transpose: bool = False):
    q, s = quantize_rowwise(x, spec)
    return (np.ascontiguousarray(q.T) if transpose else q), s


def quantize_weight(w, spec: QuantSpec = SPEC_V0):
    return quantize_rowwise(w, spec)


def int_mm(xq: np.ndarray, wq: np.ndarray) -> np.ndarray:
    """Exact acc[m,n] = sum_k xq[m,k] * wq[n,k] in int32 (row a3)."""
    xq = np.ascontiguousarray(xq, dtype=np.int8)
    wq = np.ascontiguousarray(wq, dtype=np.int8)
    M, K = xq.shape
    N, K2 = wq.shape
    assert K == K2
    if torch is not None and M > 16 and K % 8 == 0 and N % 8 == 0:
        # torch._int_mm on CPU is exact int32 accumulation (torch/_meta_registrations.py:3762)
        return torch._int_mm(torch.from_numpy(xq), torch.from_numpy(wq).t()).numpy()
    return xq.astype(np.int32) @ wq.astype(np.int32).T


def dequant_epilogue(acc: np.ndarray, s_x: np.ndarray, s_w: np.ndarray,
                     bias: Optional[np.ndarray] = None) -> np.ndarray:
    """fp32 result of ((float(acc) * s_x[m]) * s_w[n]) + bias[n]  (row a4, before the cast)."""
    y = acc.astype(np.float32)
    y = (y * s_x.astype(np.float32)[:, None]).astype(np.float32)
    y = (y * s_w.astype(np.float32)[None, :]).astype(np.float32)
    if bias is not None:
        y = (y + bias.astype(np.float32)[None, :]).astype(np.float32)
    return y


def cast_out(y32: np.ndarray, dtype: str):
    """Single RNE cast of the fp32 epilogue result.  Returns a torch tensor for bf16/fp16."""
    if dtype in ("f32", "float32"):
        return torch.from_numpy(y32) if torch is not None else y32
    t = torch.from_numpy(np.ascontiguousarray(y32))
    if dtype in ("bf16", "bfloat16"):
        return t.to(torch.bfloat16)
    if dtype in ("f16", "float16"):
        return t.to(torch.float16)
    raise ValueError(dtype)


def dequantize(q: np.ndarray, s: np.ndarray, axis: int = 0) -> np.ndarray:
    """QTensor.dequantize(): q * s broadcast along `axis` (0: per-row, 1: per-column), fp32 (row a5)."""
    q32 = q.astype(np.float32)
    s = s.astype(np.float32)
    return (q32 * (s[:, None] if axis == 0 else s[None, :])).astype(np.float32)


def qlinear(x, wq: np.ndarray, s_w: np.ndarray, bias: Optional[np.ndarray] = None,
            spec: QuantSpec = SPEC_V0, out_dtype: str = "bf16"):
    """The whole forward of row a6: act-quant -> int32 GEMM -> dequant epilogue -> cast."""
    x32 = to_f32(x)
    lead = x32.shape[:-1]
    x2 = x32.reshape(-1, x32.shape[-1])
    xq, s_x = quantize_rowwise(x2, spec)
    acc = int_mm(xq, wq)
    y = cast_out(dequant_epilogue(acc, s_x, s_w, bias), out_dtype)
    return y.reshape(*lead, wq.shape[0])


# ---- producer ops in front of the path ------------------------------
# fp64 references of the tensors the fused kernels emit: the floating-point half of their parity
# (tolerance stated in tests/test_gpu_fused.py); the integer half is quantize_rowwise() of the
# emitted tensor, bit-exact.
def rmsnorm_ref(x, gamma, eps: float) -> np.ndarray:
    """Llama RMSNorm, exact arithmetic: gamma * x * rsqrt(mean(x^2) + eps)."""
    x64 = to_f32(x).astype(np.float64)
    g64 = to_f32(gamma).astype(np.float64)
    r = 1.0 / np.sqrt(np.mean(x64 * x64, axis=-1, keepdims=True) + eps)
    return x64 * r * g64


def layernorm_ref(x, gamma, beta, eps: float) -> np.ndarray:
    x64 = to_f32(x).astype(np.float64)
    mu = np.mean(x64, axis=-1, keepdims=True)
    var = np.mean((x64 - mu) ** 2, axis=-1, keepdims=True)
    return (x64 - mu) / np.sqrt(var + eps) * to_f32(gamma).astype(np.float64) + to_f32(beta).astype(np.float64)


def act_ref(x64: np.ndarray, act: str) -> np.ndarray:
    import math
    if act == "identity":
        return x64
    if act == "silu":
        return x64 / (1.0 + np.exp(-x64))
    if act == "gelu":
        erf = np.vectorize(math.erf)
        return 0.5 * x64 * (1.0 + erf(x64 / math.sqrt(2.0)))
    if act == "gelu_tanh":
        return 0.5 * x64 * (1.0 + np.tanh(math.sqrt(2.0 / math.pi) * (x64 + 0.044715 * x64 ** 3)))
    raise ValueError(act)


def act_mul_ref(gate, up=None, act: str = "silu") -> np.ndarray:
    a = act_ref(to_f32(gate).astype(np.float64), act)
    return a if up is None else a * to_f32(up).astype(np.float64)


# ---- torch-threaded variant used only as the timed CPU baseline (bench.py) ----------
def qlinear_torch_cpu(x_t, wq_t_kn, s_w_t, bias_t, out_dtype):
    """Same math as `qlinear` written with torch CPU ops so it uses every host thread.
    x_t [M,K] float tensor; wq_t_kn = Wq.t() ([K,N] int8 view); returns [M,N] out_dtype."""
    xf = x_t.to(torch.float32)
    amax = xf.abs().amax(dim=-1, keepdim=True)
    s = amax / 127.0
    s = torch.where(amax == 0, torch.ones_like(s), s)
    xq = torch.round(xf / s).clamp_(-128, 127).to(torch.int8)
    acc = torch._int_mm(xq, wq_t_kn)
    y = acc.to(torch.float32) * s
    y = y * s_w_t[None, :]
    if bias_t is not None:
        y = y + bias_t[None, :]
    return y.to(out_dtype)
