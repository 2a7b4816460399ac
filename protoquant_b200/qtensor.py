"""QTensor: int8 payload + fp32 scales + quantisation axis.

The reference's real class is unavailable (REFERENCE ABSENT); this keeps the
"QTensor-style quantize/dequantize" surface BASELINE.json names.  Names the reference may
use are collected in compat.py.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import functional as F


class QTensor:
    """A row-wise (axis=-1 reduced) symmetric int8 quantised 2-D tensor.

    ``data``  int8  [R, C]  (row stride padded to 16 bytes so it can feed the GEMM directly)
    ``scale`` fp32  [R]     one scale per row: per token for activations, per output channel
                            for weights W[N,K]
    ``axis``  the reduced axis of the original tensor: -1 (the last one; one scale per row of ``data``) or, for 2-D
              tensors, 0 (one scale per column of ``data``)
    """

    __slots__ = ("data", "scale", "axis", "orig_dtype", "orig_shape")

    def __init__(self, data: torch.Tensor, scale: torch.Tensor, axis: int = -1,
                 orig_dtype: torch.dtype = torch.float32, orig_shape=None):
        self.data = data
        self.scale = scale
        self.axis = axis
        self.orig_dtype = orig_dtype
        self.orig_shape = tuple(orig_shape) if orig_shape is not None else tuple(data.shape)

    @property
    def shape(self):
        return self.orig_shape

    @property
    def device(self):
        return self.data.device

    def dequantize(self, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        # axis == -1: one scale per row of `data`; axis == 0 (2-D tensors): one scale per column
        out = F.dequantize(self.data, self.scale, axis=0 if self.axis == -1 else 1, out_dtype=dtype or self.orig_dtype)
        return out.reshape(self.orig_shape)

    def int_repr(self) -> torch.Tensor:
        return self.data

    # ---- serialisation ------------------------------------------------------------------
    # The reference's own QTensor format cannot be read (REFERENCE ABSENT); this one is versioned so that a
    # converter can be added when it can.  Only tensors, ints, strings and tuples are stored, so the file loads
    # with torch.load(weights_only=True).
    FORMAT = "protoquant_b200.QTensor"
    FORMAT_VERSION = 1

    def state(self) -> dict:
        """Plain-dict form: the int8 payload WITHOUT the row padding ([R, C] contiguous), the fp32 scales and the
        metadata needed to rebuild the tensor."""
        return {"format": self.FORMAT, "format_version": self.FORMAT_VERSION,
                "data": self.data.contiguous(), "scale": self.scale.contiguous(), "axis": int(self.axis),
                "orig_dtype": str(self.orig_dtype).replace("torch.", ""), "orig_shape": tuple(self.orig_shape)}

    @classmethod
    def from_state(cls, state: dict, device=None) -> "QTensor":
        """Inverse of `state()`.  The payload is re-laid out with a 16-byte row stride (the GEMM / TMA requirement),
        on `device` if given (else wherever the stored tensors live).  Raises on a foreign or newer format."""
        if state.get("format", cls.FORMAT) != cls.FORMAT:
            raise ValueError(f"not a {cls.FORMAT} state: format={state.get('format')!r}")
        ver = int(state.get("format_version", 0))        # 0: round-1 state() (torch.dtype object, no version field)
        if ver > cls.FORMAT_VERSION:
            raise ValueError(f"QTensor state has format_version {ver}; this build reads <= {cls.FORMAT_VERSION}")
        data, scale = state["data"], state["scale"]
        if data.dtype != torch.int8 or data.dim() != 2:
            raise TypeError("QTensor state: `data` must be a 2-D int8 tensor")
        axis = int(state["axis"])
        if axis not in (-1, 0):
            raise ValueError(f"QTensor state: unsupported axis {axis}")
        n_scale = data.shape[0] if axis == -1 else data.shape[1]
        if scale.dtype != torch.float32 or scale.numel() != n_scale:
            raise TypeError(f"QTensor state: `scale` must be fp32 with {n_scale} entries")
        od = state["orig_dtype"]
        orig_dtype = od if isinstance(od, torch.dtype) else getattr(torch, str(od))
        dev = torch.device(device) if device is not None else data.device
        q = F.alloc_q(data.shape[0], data.shape[1], dev)
        q.copy_(data)
        return cls(q, scale.to(dev).contiguous(), axis, orig_dtype, tuple(state["orig_shape"]))

    def save(self, f) -> None:
        torch.save(self.state(), f)

    @classmethod
    def load(cls, f, device=None) -> "QTensor":
        return cls.from_state(torch.load(f, map_location="cpu" if device is not None else None, weights_only=True), device)

    def to(self, device) -> "QTensor":
        return self.from_state(self.state(), device)

    def __repr__(self):
        return (f"QTensor(shape={self.orig_shape}, dtype=int8, scale=fp32[{self.scale.numel()}], "
                f"axis={self.axis}, orig_dtype={self.orig_dtype}, device={self.data.device})")


def quantize(t: torch.Tensor, axis: int = -1, spec: Optional[F.QuantSpec] = None) -> QTensor:
    """Symmetric int8 quantisation with one scale per slice along the last axis.

    ``t`` [..., C] is flattened to [R, C]; each row gets ``scale = absmax/127`` and
    ``q = rne(t / scale)``.  Only ``axis=-1`` (reduce over the last dim) is supported,
    which covers both per-token activations and per-output-channel weights W[N,K]."""
    if t.dim() == 2 and axis in (0, -2):
        # reduce over the rows: one scale per column (e.g. a weight stored [K, N]).  The row-wise kernel runs on
        # the transposed copy; `data` is returned in the original orientation, its scale vector has N entries.
        q, s = F.quantize_act(t.t().contiguous(), spec=spec)
        return QTensor(q.t().contiguous(), s, axis=0, orig_dtype=t.dtype, orig_shape=t.shape)
    if axis not in (-1, t.dim() - 1):
        raise NotImplementedError("QTensor quantisation reduces over the last axis (any rank) or axis 0 of a 2-D tensor")
    t2 = t.reshape(-1, t.shape[-1])
    q, s = F.quantize_act(t2, spec=spec)
    return QTensor(q, s, axis=-1, orig_dtype=t.dtype, orig_shape=t.shape)


def dequantize(qt: QTensor, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    return qt.dequantize(dtype)
