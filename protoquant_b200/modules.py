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

    # ---- serialisation ------------------------------------------------------------------
    # state_dict(): qweight_storage int8 [N, ld16(K)] (rows padded to 16 bytes), weight_scale fp32 [N], bias fp32 [N].
    # nn.Module records `_version` in the state_dict metadata; version 1 = this layout.  A checkpoint may also carry
    # the payload UNPADDED under `qweight` ([N, K], what QTensor.state()["data"] holds) -- it is re-padded on load.
    _version = 1

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        ver = local_metadata.get("version", 1)
        if ver is not None and ver > self._version:
            error_msgs.append(f"{prefix}: checkpoint written by DynamicQuantLinear format {ver}, this build reads <= {self._version}")
            return
        key, alt = prefix + "qweight_storage", prefix + "qweight"
        src = state_dict.get(key, state_dict.get(alt))
        if src is not None and (alt in state_dict or tuple(src.shape) != tuple(self.qweight_storage.shape)):
            # unpadded (or differently padded) payload: copy the [N, K] block into this module's padded storage
            if src.dim() != 2 or src.shape[0] != self.out_features or src.shape[1] < self.in_features or src.dtype != torch.int8:
                error_msgs.append(f"{prefix}qweight: expected int8 [{self.out_features}, >={self.in_features}], got "
                                  f"{src.dtype} {tuple(src.shape)}")
                return
            state_dict = dict(state_dict)
            state_dict.pop(alt, None)
            fixed = torch.zeros_like(self.qweight_storage, device=src.device)
            fixed[:, : self.in_features].copy_(src[:, : self.in_features])
            state_dict[key] = fixed
        for name in self._FP32_BUFFERS:      # scales / bias saved in another float dtype are widened, never narrowed
            k = prefix + name
            if k in state_dict and state_dict[k].dtype != torch.float32:
                state_dict = dict(state_dict)
                state_dict[k] = state_dict[k].to(torch.float32)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    def weight_qtensor(self) -> QTensor:
        """The packed weight as a QTensor (int8 [N, K] payload + per-output-channel scales)."""
        return QTensor(self.qweight, self.weight_scale, axis=-1, orig_dtype=torch.float32,
                       orig_shape=(self.out_features, self.in_features))

    @classmethod
    def from_qtensor(cls, qt: QTensor, bias: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None,
                     spec: Optional[F.QuantSpec] = None) -> "DynamicQuantLinear":
        """Build the module from an already-quantised weight (e.g. QTensor.load(path, device))."""
        if qt.axis != -1 or qt.data.dim() != 2:
            raise ValueError("the weight QTensor must be [out_features, in_features] with one scale per row")
        n, k = qt.data.shape
        m = cls(k, n, bias is not None, device=qt.data.device, out_dtype=out_dtype, spec=spec)
        m.qweight_storage[:, :k].copy_(qt.data)
        m.weight_scale.copy_(qt.scale)
        if bias is not None:
            m.bias.copy_(bias.to(torch.float32))
        return m

    # buffers whose dtype is part of the kernel ABI: module.half() / .to(torch.bfloat16) / model.to(dtype) must not
    # cast them (the C entry points read weight_scale and bias as fp32 through raw pointers)
    _FP32_BUFFERS = ("weight_scale", "bias")

    def _apply(self, fn, recurse=True):
        keep = {name: self._buffers.get(name) for name in self._FP32_BUFFERS}
        super()._apply(fn, recurse)
        for name, old in keep.items():
            buf = self._buffers.get(name)
            if old is not None and buf is not None and buf.dtype != torch.float32:
                # a dtype cast: keep the ORIGINAL fp32 values (a round trip through fp16 would change them), follow the device
                self._buffers[name] = old.to(device=buf.device)
        return self

    def _check_abi_buffers(self, device):
        """The lean path hands raw pointers to the C ABI: they must be fp32, contiguous and on the input's device."""
        for name in self._FP32_BUFFERS:
            buf = getattr(self, name)
            if buf is None:
                continue
            if buf.dtype != torch.float32 or not buf.is_contiguous() or buf.device != device:
                raise TypeError(f"DynamicQuantLinear.{name} must be a contiguous fp32 tensor on {device} "
                                f"(got {buf.dtype} on {buf.device})")
        w = self.qweight_storage
        if w.dtype != torch.int8 or w.device != device or w.stride(1) != 1 or w.stride(0) % 16 or w.data_ptr() % 16:
            raise TypeError("DynamicQuantLinear.qweight_storage must be int8 on the input's device with 16-byte aligned rows")

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
        self._check_abi_buffers(x.device)
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


class _SharedInputGroup:
    """Run-time state of several linears that read the SAME activation (q/k/v, gate/up): their weights live in one
    fused DynamicQuantLinear, so the activation is quantised once and one GEMM produces every member's output.

    The first member called with an input x runs the fused forward and the result is kept until every member has
    taken its slice (or another input arrives).  "Same input" = same storage address, tensor version, shape,
    strides, dtype and CUDA stream; the group holds a reference to x meanwhile, so the address cannot be recycled.
    A member that arrives with a different input computes only its own slice (one act-quant + one GEMM on its rows
    of the fused weight); after `MAX_DIVERGENT` such calls the group stops fusing (e.g. cross-attention, where
    q and k/v read different tensors)."""

    MAX_DIVERGENT = 4

    def __init__(self, fused: DynamicQuantLinear, sizes):
        self.fused = fused
        self.bounds = []
        lo = 0
        for n in sizes:
            self.bounds.append((lo, lo + n))
            lo += n
        self.all_mask = (1 << len(sizes)) - 1
        self.divergent = 0
        self._key = None
        self._x = None
        self._y = None
        self._served = 0

    @staticmethod
    def _key_of(x: torch.Tensor):
        return (x.data_ptr(), x._version, tuple(x.shape), tuple(x.stride()), x.dtype, x.device,
                torch.cuda.current_stream(x.device).cuda_stream if x.is_cuda else 0)

    def _slice_forward(self, x, idx):
        lo, hi = self.bounds[idx]
        f = self.fused
        bias = f.bias[lo:hi] if f.bias is not None else None
        return F.qlinear(x, f.qweight[lo:hi], f.weight_scale[lo:hi], bias, f.out_dtype or x.dtype, f.spec)

    def output_for(self, x: torch.Tensor, idx: int) -> torch.Tensor:
        if not isinstance(x, torch.Tensor) or self.divergent >= self.MAX_DIVERGENT:
            return self._slice_forward(x, idx)
        key = self._key_of(x)
        bit = 1 << idx
        if self._key == key and not (self._served & bit):
            y = self._y
        elif self._key is not None and self._served and self._served != self.all_mask and self._key != key:
            # the group is in the middle of serving another input: this member reads a different tensor
            self.divergent += 1
            return self._slice_forward(x, idx)
        else:
            y = self.fused(x)
            self._key, self._x, self._y, self._served = key, x, y, 0
        self._served |= bit
        lo, hi = self.bounds[idx]
        out = y[..., lo:hi]
        if self._served == self.all_mask:
            self._key = self._x = self._y = None
            self._served = 0
        return out


class SharedInputLinear(nn.Module):
    """One member (e.g. `k_proj`) of a group of linears fused by `swap_linear(fuse_shared_inputs=True)`.  The fused
    parameters are registered once, on the parent module (`<parent>._pq_fused_<first member>`); this module only
    knows its group and its position, and returns its column slice of the fused output (a view: rows keep the fused
    row stride).  Bit-identical to the separate DynamicQuantLinear: per-output-channel scales make the column
    blocks of the fused GEMM independent."""

    def __init__(self, group: _SharedInputGroup, index: int, in_features: int, out_features: int):
        super().__init__()
        object.__setattr__(self, "_group", group)      # not a registered submodule: the parent owns the parameters
        self.index = index
        self.in_features = in_features
        self.out_features = out_features

    @property
    def fused(self) -> DynamicQuantLinear:
        return self._group.fused

    def forward(self, x) -> torch.Tensor:
        return self._group.output_for(x, self.index)

    def extra_repr(self) -> str:
        lo, hi = self._group.bounds[self.index]
        return f"in_features={self.in_features}, out_features={self.out_features}, rows [{lo}, {hi}) of the fused int8 weight"


# sibling linears that read the same activation in the common transformer blocks
SHARED_INPUT_PATTERNS = (
    ("q_proj", "k_proj", "v_proj"), ("gate_proj", "up_proj"),          # Llama / Mistral / Qwen (HF naming)
    ("query", "key", "value"),                                         # BERT self-attention
    ("wq", "wk", "wv"), ("w1", "w3"),                                  # Meta Llama naming
    ("q_lin", "k_lin", "v_lin"),                                       # DistilBERT
)


def _fuse_group(parent: nn.Module, names, out_dtype, spec) -> bool:
    lins = [getattr(parent, n, None) for n in names]
    if not all(isinstance(l, nn.Linear) for l in lins):
        return False
    k = lins[0].in_features
    if any(l.in_features != k or (l.bias is None) != (lins[0].bias is None) or l.weight.device != lins[0].weight.device
           or l.weight.dtype != lins[0].weight.dtype for l in lins):
        return False
    fused = fuse_linears([DynamicQuantLinear.from_float(l, out_dtype=out_dtype, spec=spec) for l in lins])
    group = _SharedInputGroup(fused, [l.out_features for l in lins])
    setattr(parent, "_pq_fused_" + names[0], fused)
    for i, (n, l) in enumerate(zip(names, lins)):
        setattr(parent, n, SharedInputLinear(group, i, k, l.out_features))
    return True


def swap_linear(model: nn.Module, min_features: int = 0, out_dtype: Optional[torch.dtype] = None,
                spec: Optional[F.QuantSpec] = None, skip=(), fuse_shared_inputs: bool = True,
                patterns=SHARED_INPUT_PATTERNS) -> nn.Module:
    """Replace every nn.Linear (in/out features >= min_features, name not in `skip`) in place.

    With `fuse_shared_inputs` (default) sibling linears that read the same activation -- `patterns`, e.g.
    q_proj/k_proj/v_proj and gate_proj/up_proj -- are fused: the activation is quantised ONCE and one GEMM computes
    all of them (SharedInputLinear).  The outputs are bit-identical to separate modules; a Llama block then runs 4
    quantising launches and 4 GEMMs instead of 7 + 7."""
    if fuse_shared_inputs:
        for names in patterns:
            if any(n in skip for n in names):
                continue
            lins = [getattr(model, n, None) for n in names]
            if all(isinstance(l, nn.Linear) and min(l.in_features, l.out_features) >= min_features for l in lins):
                _fuse_group(model, names, out_dtype, spec)
    for name, child in list(model.named_children()):
        if isinstance(child, (DynamicQuantLinear, SharedInputLinear)):
            continue
        if isinstance(child, nn.Linear) and name not in skip and \
                min(child.in_features, child.out_features) >= min_features:
            setattr(model, name, DynamicQuantLinear.from_float(child, out_dtype=out_dtype, spec=spec))
        else:
            swap_linear(child, min_features, out_dtype, spec, skip, fuse_shared_inputs, patterns)
    return model
