"""Drop-in replacement for nn.Linear on the dynamic int8 path.

forward:  x[...,K] --per-token int8--> xq,s_x --tcgen05 int8 GEMM + fused dequant--> y[...,N]
Weights are quantised once, per output channel, when the module is built.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
from torch import nn

from . import functional as F
from .qtensor import QTensor

_NVTX = os.environ.get("PQ_NVTX", "0") not in ("", "0")


class DynamicQuantLinear(nn.Module):
    """int8 weight (per-output-channel scale) + dynamic per-token int8 activations."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, device=None,
                 out_dtype: Optional[torch.dtype] = None, spec: Optional[F.QuantSpec] = None):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.out_dtype = out_dtype
        self.spec = spec
        kp = (in_features + 15) // 16 * 16
        self.register_buffer("qweight_storage", torch.zeros((out_features, kp), dtype=torch.int8, device=device))
        self.register_buffer("weight_scale", torch.ones((out_features,), dtype=torch.float32, device=device))
        if bias:
            self.register_buffer("bias", torch.zeros((out_features,), dtype=torch.float32, device=device))
        else:
            self.bias = None

    @property
    def qweight(self) -> torch.Tensor:
        """int8 [N, K] view whose row stride is a multiple of 16 bytes."""
        return self.qweight_storage[:, : self.in_features]

    @classmethod
    def from_float(cls, linear: nn.Linear, out_dtype: Optional[torch.dtype] = None,
                   spec: Optional[F.QuantSpec] = None) -> "DynamicQuantLinear":
        w = linear.weight.detach()
        if not w.is_cuda:
            raise RuntimeError("DynamicQuantLinear.from_float needs the nn.Linear on a CUDA device "
                               "(weights are quantised by the GPU kernel; there is no CPU path)")
        m = cls(linear.in_features, linear.out_features, linear.bias is not None, device=w.device,
                out_dtype=out_dtype, spec=spec)
        wq, s = F.quantize_weight(w, spec=spec)
        m.qweight_storage[:, : linear.in_features].copy_(wq)
        m.weight_scale.copy_(s)
        if linear.bias is not None:
            m.bias.copy_(linear.bias.detach().to(torch.float32))
        return m

    def forward_quantized(self, xq: torch.Tensor, s_x: torch.Tensor, out_dtype: Optional[torch.dtype] = None,
                          out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """GEMM + dequant epilogue only, for an input that is already per-token int8 (a QTensor's payload,
        or the output of `rmsnorm_quant` / `act_mul_quant`): xq int8 [M,K], s_x fp32 [M] -> y [M,N]."""
        return F.qgemm(xq, s_x, self.qweight, self.weight_scale, self.bias,
                       out_dtype or self.out_dtype or torch.bfloat16, out=out)

    def forward(self, x) -> torch.Tensor:
        if _NVTX:      # PQ_NVTX=1: one range per linear so that nsys / ncu timelines show the two launches together
            torch.cuda.nvtx.range_push(f"pq.DynamicQuantLinear[{self.in_features}->{self.out_features}]")
            try:
                return self._forward(x)
            finally:
                torch.cuda.nvtx.range_pop()
        return self._forward(x)

    def _forward(self, x) -> torch.Tensor:
        # A QTensor (or an (int8, scale) pair) is an activation that is already quantised per token.
        if isinstance(x, QTensor):
            y = self.forward_quantized(x.data, x.scale, self.out_dtype or x.orig_dtype)
            return y.reshape(*x.orig_shape[:-1], self.out_features)
        if isinstance(x, tuple) and len(x) == 2:
            return self.forward_quantized(x[0], x[1])
        # Lean path: one C-ABI call (pq_qlinear = act-quant launch + GEMM launch), three allocations.
        K, N = self.in_features, self.out_features
        if (not x.is_cuda) or x.dtype not in F._DT or x.shape[-1] != K:
            return F.qlinear(x, self.qweight, self.weight_scale, self.bias, self.out_dtype or x.dtype, self.spec)
        x2 = x.reshape(-1, K)
        if x2.stride(-1) != 1:
            x2 = x2.contiguous()
        M = x2.shape[0]
        out_dtype = self.out_dtype or x.dtype
        if out_dtype not in F._DT:
            raise TypeError(f"unsupported output dtype {out_dtype}")
        wq = self.qweight_storage
        y = torch.empty((M, N), dtype=out_dtype, device=x.device)
        if M:
            xq = torch.empty((M, wq.shape[1]), dtype=torch.int8, device=x.device)
            sx = torch.empty((M,), dtype=torch.float32, device=x.device)
            F.qlinear_into(x2, wq, K, self.weight_scale, self.bias, y, xq, sx, self.spec)
        return y.reshape(*x.shape[:-1], N)

    def dequantized_weight(self, dtype: torch.dtype = torch.float32) -> torch.Tensor:
        return F.dequantize(self.qweight, self.weight_scale, axis=0, out_dtype=dtype)

    def extra_repr(self) -> str:
        return f"in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None}, int8"


def fuse_linears(mods) -> DynamicQuantLinear:
    """One DynamicQuantLinear computing several linears that read the SAME input (q/k/v, gate/up): the int8
    weights, per-channel scales and biases are concatenated along the output dimension, so the activation is
    quantised once and one GEMM runs instead of len(mods).  Per-output-channel scales make this exact: the fused
    output is the concatenation of the separate outputs, bit for bit (split it with `y.split(sizes, -1)`)."""
    mods = list(mods)
    if not mods:
        raise ValueError("fuse_linears needs at least one module")
    K = mods[0].in_features
    if any(m.in_features != K for m in mods):
        raise ValueError("fused linears must share in_features")
    if any((m.bias is None) != (mods[0].bias is None) for m in mods):
        raise ValueError("fused linears must all have a bias or all have none")
    if any(m.spec != mods[0].spec or m.out_dtype != mods[0].out_dtype for m in mods):
        raise ValueError("fused linears must share the quantisation spec and output dtype")
    dev = mods[0].qweight_storage.device
    fused = DynamicQuantLinear(K, sum(m.out_features for m in mods), mods[0].bias is not None, device=dev,
                               out_dtype=mods[0].out_dtype, spec=mods[0].spec)
    fused.qweight_storage.copy_(torch.cat([m.qweight_storage for m in mods], dim=0))
    fused.weight_scale.copy_(torch.cat([m.weight_scale for m in mods]))
    if fused.bias is not None:
        fused.bias.copy_(torch.cat([m.bias for m in mods]))
    return fused


def swap_linear(model: nn.Module, min_features: int = 0, out_dtype: Optional[torch.dtype] = None,
                spec: Optional[F.QuantSpec] = None, skip=()) -> nn.Module:
    """Replace every nn.Linear (in/out features >= min_features, name not in `skip`) in place."""
    for name, child in list(model.named_children()):
        if isinstance(child, nn.Linear) and name not in skip and \
                min(child.in_features, child.out_features) >= min_features:
            setattr(model, name, DynamicQuantLinear.from_float(child, out_dtype=out_dtype, spec=spec))
        else:
            swap_linear(child, min_features, out_dtype, spec, skip)
    return model
