"""Column-parallel dynamic-quant linear (BASELINE.json north_star (4)).

Rank r of a G-rank group holds rows [r*N/G, (r+1)*N/G) of (Wq, s_w, bias).  The activation
is replicated; every rank quantises it (cheap, HBM-bound), runs the int8 GEMM on its weight
slice and the output slices are all-gathered along N over NCCL/NVLink.  Column sharding does
not change any accumulation order, so the gathered result is bit-identical to the 1-GPU one.

The gather is the path's only exchange step.  Two implementations:

* fused (default on CUDA when torch's symmetric memory can be set up): every rank owns a
  symmetric [tokens, N] output buffer; the GEMM epilogue stores each finished tile straight
  into ALL ranks' buffers -- one coalesced NVLink peer store per rank (or, opt-in with
  PQ_USE_MULTICAST=1, a single store to the NVSwitch multicast address) -- so the transfer overlaps the MMAs tile
  by tile and no separate collective or layout-fixing copy runs.  One symmetric-memory barrier
  follows the kernel.  Output buffers are double-buffered: the tensor returned by forward() is a view
  that stays valid until the next-but-one forward() of the same module (`copy_output=True` returns a
  private copy instead).
* NCCL all-gather of the [tokens, N/G] slices (baseline; also the gloo path used by CPU tests).

`min_out_features` implements the north star's "used only for layers big enough to benefit":
smaller layers stay replicated.
"""
from __future__ import annotations

import logging
import os
from typing import Callable, Optional

import torch
import torch.distributed as dist
from torch import nn

from . import functional as F

log = logging.getLogger("protoquant_b200.sharded")


def _agree(ok: bool, device, group) -> bool:
    """All ranks of `group` learn whether EVERY rank succeeded (MIN all-reduce of a flag): a rank-local failure of
    the symmetric-memory setup must flip every rank to the fallback together, or the job hangs with some ranks in
    a symmetric-memory barrier and others in an NCCL collective."""
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(flag.item())


def shard_bounds(n: int, world: int, rank: int, align: int = 8):
    """[lo, hi) of rank's slice of n output channels; slices are `align`-aligned except the last."""
    per = (n + world - 1) // world
    per = (per + align - 1) // align * align
    lo = min(rank * per, n)
    hi = min(lo + per, n)
    return lo, hi


def _full_from_state_dict(state_dict, prefix: str, device):
    """(qweight [N,K] int8, weight_scale [N] fp32, bias [N] fp32 | None) of ONE unsharded DynamicQuantLinear read out
    of a full-model checkpoint (`prefix` = the module's name + '.'), moved to `device`.  Accepts the padded
    `qweight_storage` layout and the unpadded `qweight` one (modules.DynamicQuantLinear._load_from_state_dict)."""
    w = state_dict.get(prefix + "qweight_storage", state_dict.get(prefix + "qweight"))
    sw = state_dict.get(prefix + "weight_scale")
    if w is None or sw is None:
        raise KeyError(f"checkpoint has no '{prefix}qweight_storage' / '{prefix}weight_scale'")
    if w.dtype != torch.int8 or w.dim() != 2 or sw.numel() != w.shape[0]:
        raise TypeError(f"'{prefix}qweight_storage' must be int8 [N, K] with one scale per row")
    b = state_dict.get(prefix + "bias")
    return (w.to(device), sw.to(device=device, dtype=torch.float32),
            b.to(device=device, dtype=torch.float32) if b is not None else None)


class ShardedDynamicQuantLinear(nn.Module):
    @classmethod
    def from_full_state_dict(cls, state_dict, prefix: str = "", in_features: Optional[int] = None, device=None, **kw):
        """Every rank loads the SAME full (unsharded) checkpoint of a DynamicQuantLinear -- e.g.
        torch.load(path, map_location="cpu") -- and keeps only its column slice on `device`.
        `in_features` trims the 16-byte row padding of `qweight_storage` (default: the stored width)."""
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        w, sw, b = _full_from_state_dict(state_dict, prefix, dev)
        return cls(w[:, : (in_features or w.shape[1])], sw, b, **kw)

    def __init__(self, qweight_full: torch.Tensor, weight_scale_full: torch.Tensor,
                 bias_full: Optional[torch.Tensor], group=None, out_dtype: Optional[torch.dtype] = None,
                 spec: Optional[F.QuantSpec] = None,
                 local_forward: Optional[Callable] = None, fused: Optional[bool] = None,
                 gather_output: bool = True, align: int = 8, copy_output: bool = False):
        """qweight_full [N,K] int8, weight_scale_full [N] fp32, bias_full [N] fp32|None: the
        UNSHARDED quantised weight (every rank passes the same tensors; each keeps its slice).
        `local_forward(x2d, wq, s_w, bias, out_dtype)` defaults to the CUDA path; tests on a
        CPU-only box inject the oracle there to exercise the shard/gather logic over gloo."""
        super().__init__()
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.out_features, self.in_features = qweight_full.shape
        self.out_dtype = out_dtype
        self.spec = spec
        self._local_forward = local_forward
        # fused epilogue all-gather needs CUDA + symmetric memory; None = try it, fall back to NCCL
        self.fused = fused
        self.fused_error = None   # why the fused path was turned off (repr of the setup exception), if it was
        self.copy_output = copy_output
        self._symm = None      # (capacity_rows, dtype) -> [(tensor, handle), (tensor, handle)]
        self._ws = None        # act-quant workspace (xq, s_x) reused across calls
        self._flip = 0
        # gather_output=False keeps the [tokens, hi-lo] slice local (it feeds a row-parallel layer, §8f-3)
        self.gather_output = gather_output
        per = shard_bounds(self.out_features, self.world, 0, align)[1]
        self.per = per
        lo, hi = shard_bounds(self.out_features, self.world, self.rank, align)
        self.lo, self.hi = lo, hi
        dev = qweight_full.device
        kp = (self.in_features + 15) // 16 * 16
        # every rank allocates the same padded slice height so the all-gather is regular
        wq = torch.zeros((per, kp), dtype=torch.int8, device=dev)
        sw = torch.ones((per,), dtype=torch.float32, device=dev)
        wq[: hi - lo, : self.in_features].copy_(qweight_full[lo:hi])
        sw[: hi - lo].copy_(weight_scale_full[lo:hi])
        self.register_buffer("qweight_storage", wq)
        self.register_buffer("weight_scale", sw)
        if bias_full is not None:
            b = torch.zeros((per,), dtype=torch.float32, device=dev)
            b[: hi - lo].copy_(bias_full[lo:hi].to(torch.float32))
            self.register_buffer("bias", b)
        else:
            self.bias = None

    @property
    def qweight(self):
        return self.qweight_storage[:, : self.in_features]

    def local(self, x2: torch.Tensor, out_dtype) -> torch.Tensor:
        if self._local_forward is not None:
            return self._local_forward(x2, self.qweight, self.weight_scale, self.bias, out_dtype)
        return F.qlinear(x2, self.qweight, self.weight_scale, self.bias, out_dtype, self.spec)

    # ---- fused path -------------------------------------------------------------------
    def _symm_buffers(self, rows: int, dtype, device):
        import torch.distributed._symmetric_memory as symm_mem
        key = (dtype,)
        if self._symm is not None and self._symm[0] == key and self._symm[1] >= rows:
            return self._symm[2]
        cap = max(rows, 16)
        group = self.group if self.group is not None else dist.group.WORLD
        bufs = []
        for _ in range(2):
            t = symm_mem.empty((cap, self.world * self.per), dtype=dtype, device=device)
            h = symm_mem.rendezvous(t, group)
            bufs.append((t, h))
        self._symm = (key, cap, bufs)
        return bufs

    def _enable_fused(self, rows: int, dtype, device) -> bool:
        """Set up (or grow) the symmetric output buffers.  Only THIS step may fail softly -- symmetric memory can
        be unavailable -- and the ranks agree on the outcome, so they all take the same path; the reason is kept
        in `fused_error` and logged once.  Errors of the kernels themselves are never swallowed."""
        if self._symm is not None and self._symm[0] == (dtype,) and self._symm[1] >= rows:
            return True
        err = None
        try:
            self._symm_buffers(rows, dtype, device)
        except Exception as ex:        # noqa: BLE001 - any setup failure means "no symmetric memory here"
            err = ex
        group = self.group if self.group is not None else dist.group.WORLD
        if _agree(err is None, device, group):
            return True
        self._symm = None
        self.fused_error = repr(err) if err is not None else "symmetric-memory setup failed on another rank"
        if self.fused is True:
            raise RuntimeError(f"fused epilogue all-gather was requested (fused=True) but is unavailable: {self.fused_error}")
        log.warning("ShardedDynamicQuantLinear: fused epilogue all-gather unavailable (%s); using the NCCL all-gather",
                    self.fused_error)
        self.fused = False
        return False

    def _workspace(self, rows: int, device):
        if self._ws is None or self._ws[0].shape[0] < rows or self._ws[0].device != device:
            kp = self.qweight_storage.shape[1]
            self._ws = (torch.empty((max(rows, 16), kp), dtype=torch.int8, device=device),
                        torch.empty((max(rows, 16),), dtype=torch.float32, device=device))
        return self._ws

    def _forward_fused(self, x2: torch.Tensor, out_dtype) -> torch.Tensor:
        M = x2.shape[0]
        bufs = self._symm[2]
        t, h = bufs[self._flip]
        self._flip ^= 1
        esz = t.element_size()
        ld = self.world * self.per
        off = self.rank * self.per * esz
        mc = int(getattr(h, "multicast_ptr", 0) or 0) if getattr(h, "has_multicast_support", False) else 0
        # PQ_USE_MULTICAST=1: one multimem.st per 16 bytes to the NVSwitch multicast address (the switch replicates
        # it into every rank); default: one TMA bulk store per destination (local buffer + each peer over NVLink).
        multicast = bool(mc and os.environ.get("PQ_USE_MULTICAST"))
        dests = [mc + off] if multicast else [int(p) + off for p in h.buffer_ptrs]
        if x2.stride(-1) != 1:
            x2 = x2.contiguous()
        xq_ws, sx_ws = self._workspace(M, x2.device)
        F.qlinear_multi_into(x2, self.qweight_storage, self.in_features, self.weight_scale, self.bias, dests, ld,
                             out_dtype, xq_ws, sx_ws, self.spec, multicast=multicast)
        h.barrier()
        y = t[:M, : self.out_features]
        return y.clone() if self.copy_output else y

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [..., K] (replicated) -> y [..., N] (gathered).  On the fused path the result is a VIEW into a
        double-buffered symmetric-memory buffer: it is overwritten by the second forward() after this one
        (construct with copy_output=True to get a private tensor)."""
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        out_dtype = self.out_dtype or x.dtype
        if not self.gather_output:
            y_local = self.local(x2, out_dtype)
            return y_local[:, : self.hi - self.lo].reshape(*lead, self.hi - self.lo)
        if self.world > 1 and self._local_forward is None and x2.is_cuda and self.fused is not False:
            if self._enable_fused(x2.shape[0], out_dtype, x2.device):
                self.fused = True
                return self._forward_fused(x2, out_dtype).reshape(*lead, self.out_features)
        y_local = self.local(x2, out_dtype).contiguous()          # [M, per]
        if self.world == 1:
            return y_local[:, : self.out_features].reshape(*lead, self.out_features)
        M = y_local.shape[0]
        gathered = torch.empty((self.world * M, self.per), dtype=y_local.dtype, device=y_local.device)
        dist.all_gather_into_tensor(gathered, y_local, group=self.group)
        y = gathered.view(self.world, M, self.per).permute(1, 0, 2).reshape(M, self.world * self.per)[:, : self.out_features]
        return y.reshape(*lead, self.out_features)


class _CudaShardOps:
    """The device ops of the row-parallel module (tests on a CPU-only box inject oracle versions)."""
    row_absmax = staticmethod(F.row_absmax)
    quantize_with_amax = staticmethod(F.quantize_act_with_amax)
    int_mm = staticmethod(F.qgemm_i32)
    epilogue = staticmethod(F.dequant_accumulators)


class RowParallelDynamicQuantLinear(nn.Module):
    """Row-parallel (K-split) dynamic-quant linear: rank r holds columns [k_lo, k_hi) of Wq, so
    a column-parallel up-projection can feed it without a gather in between (Megatron MLP layout).

    * The per-token scale is the maximum over the WHOLE row: with `input_is_sharded` each rank takes the maximum
      of its K-slice and the M floats are all-reduced with MAX; a replicated input needs no communication.
      Every rank then quantises its slice with the global maximum, i.e. produces exactly the codes and scales of
      the unsharded quantizer.
    * The K-slices' int32 partial sums are exact and associative, so the reduce step is done on int32 and the
      dequant epilogue runs once on the sum: the output is bit-identical to the 1-GPU module for any world size.
    * Fused exchange (CUDA + symmetric memory), ONE C call per forward (`pq_rowparallel_forward`): the row maxima
      of a K-sharded input are exchanged through symmetric memory, the GEMM epilogue stores each output-column
      block straight into the inbox of the rank that owns it (NVLink peer stores); after a signal-pad barrier the
      owner sums its `world` inboxes, applies scales and bias and -- with `gather_output` -- writes the finished
      slice into every rank's output buffer: GEMM + reduce-scatter + all-gather as a fixed sequence of launches
      with in-stream cross-rank barriers, no NCCL call and no host synchronisation, CUDA-graph capturable.
    * Fallback (gloo, or no symmetric memory): int32 all-reduce(SUM) of the partial sums + local epilogue.
    """

    @classmethod
    def from_full_state_dict(cls, state_dict, prefix: str = "", in_features: Optional[int] = None, device=None, **kw):
        """Every rank loads the same full checkpoint of a DynamicQuantLinear and keeps its K-slice (see
        ShardedDynamicQuantLinear.from_full_state_dict)."""
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        w, sw, b = _full_from_state_dict(state_dict, prefix, dev)
        return cls(w[:, : (in_features or w.shape[1])], sw, b, **kw)

    def __init__(self, qweight_full: torch.Tensor, weight_scale_full: torch.Tensor,
                 bias_full: Optional[torch.Tensor], group=None, out_dtype: Optional[torch.dtype] = None,
                 spec: Optional[F.QuantSpec] = None, input_is_sharded: bool = False, gather_output: bool = True,
                 ops=None, fused: Optional[bool] = None):
        super().__init__()
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.out_features, self.in_features = qweight_full.shape
        self.out_dtype = out_dtype
        self.spec = spec
        self.input_is_sharded = input_is_sharded
        self.gather_output = gather_output
        self.ops = ops or _CudaShardOps
        self.fused = fused
        self.fused_error = None
        self._symm = None
        self._ws = None
        self._flip = 0
        self.k_lo, self.k_hi = shard_bounds(self.in_features, self.world, self.rank, align=16)
        if self.k_hi <= self.k_lo:
            raise ValueError(f"in_features={self.in_features} is too small to split over {self.world} ranks")
        self.per_n = shard_bounds(self.out_features, self.world, 0, align=8)[1]
        self.n_lo, self.n_hi = shard_bounds(self.out_features, self.world, self.rank, align=8)
        dev = qweight_full.device
        ks = self.k_hi - self.k_lo
        wq = torch.zeros((self.out_features, (ks + 15) // 16 * 16), dtype=torch.int8, device=dev)
        wq[:, :ks].copy_(qweight_full[:, self.k_lo:self.k_hi])
        self.register_buffer("qweight_storage", wq)
        self.register_buffer("weight_scale", weight_scale_full.to(torch.float32).clone())
        if bias_full is not None:
            self.register_buffer("bias", bias_full.to(torch.float32).clone())
        else:
            self.bias = None

    @property
    def qweight(self):
        return self.qweight_storage[:, : self.k_hi - self.k_lo]

    def _quantize(self, x2: torch.Tensor):
        amax = self.ops.row_absmax(x2)
        if self.input_is_sharded and self.world > 1:
            dist.all_reduce(amax, op=dist.ReduceOp.MAX, group=self.group)
        xs = x2 if self.input_is_sharded else x2[:, self.k_lo:self.k_hi]
        return self.ops.quantize_with_amax(xs, amax, self.spec)

    def _symm_buffers(self, rows: int, dtype, device):
        """Two sets (double buffering) of symmetric buffers + one signal pad, and the pq_symm_group structs the C
        entry point takes: every rank's peer-mapped addresses of inbox / out / amax / pads."""
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        key = (dtype,)
        if self._symm is not None and self._symm[0] == key and self._symm[1] >= rows:
            return self._symm[1], self._symm[2]
        cap = max(rows, 16)
        group = self.group if self.group is not None else dist.group.WORLD
        pads = symm_mem.empty((64,), dtype=torch.int32, device=device)
        pads.zero_()
        hp = symm_mem.rendezvous(pads, group)
        sets = []
        for _ in range(2):
            inbox = symm_mem.empty((self.world, cap, self.per_n), dtype=torch.int32, device=device)
            hi = symm_mem.rendezvous(inbox, group)
            out = symm_mem.empty((cap, self.world * self.per_n), dtype=dtype, device=device)
            ho = symm_mem.rendezvous(out, group)
            amax = symm_mem.empty((self.world, cap), dtype=torch.float32, device=device)
            ha = symm_mem.rendezvous(amax, group)
            sg = _lib.PQSymmGroup()
            sg.rank, sg.world, sg.cap = self.rank, self.world, cap
            for r in range(self.world):
                sg.inbox[r], sg.out[r], sg.amax[r], sg.pads[r] = (int(hi.buffer_ptrs[r]), int(ho.buffer_ptrs[r]),
                                                                  int(ha.buffer_ptrs[r]), int(hp.buffer_ptrs[r]))
            sets.append({"inbox": inbox, "out": out, "amax": amax, "handles": (hi, ho, ha), "sg": sg, "sgp": ctypes.byref(sg)})
        torch.cuda.synchronize(device)
        dist.barrier(group)            # every rank's pad is zero before anybody signals
        self._symm = (key, cap, sets, (pads, hp))
        return cap, sets

    def _enable_fused(self, rows: int, dtype, device) -> bool:
        """Symmetric-memory setup: the only step that may fail softly; all ranks agree on the outcome."""
        if self._symm is not None and self._symm[0] == (dtype,) and self._symm[1] >= rows:
            return True
        err = None
        try:
            self._symm_buffers(rows, dtype, device)
        except Exception as ex:        # noqa: BLE001
            err = ex
        group = self.group if self.group is not None else dist.group.WORLD
        if _agree(err is None, device, group):
            return True
        self._symm = None
        self.fused_error = repr(err) if err is not None else "symmetric-memory setup failed on another rank"
        if self.fused is True:
            raise RuntimeError(f"fused reduce-scatter was requested (fused=True) but is unavailable: {self.fused_error}")
        log.warning("RowParallelDynamicQuantLinear: fused reduce-scatter unavailable (%s); using the int32 all-reduce",
                    self.fused_error)
        self.fused = False
        return False

    def _workspace(self, rows: int, device):
        if self._ws is None or self._ws[0].shape[0] < rows or self._ws[0].device != device:
            cap = max(rows, 16)
            self._ws = (torch.empty((cap, self.qweight_storage.shape[1]), dtype=torch.int8, device=device),
                        torch.empty((cap,), dtype=torch.float32, device=device),
                        torch.empty((cap,), dtype=torch.float32, device=device))
        return self._ws

    def _forward_fused(self, x2: torch.Tensor, up2: Optional[torch.Tensor], act: str, M: int, out_dtype) -> torch.Tensor:
        """ONE C call (pq_rowparallel_forward): row maxima (exchanged through symmetric memory when the input is
        K-sharded) -> quantise -> int32 GEMM scattering into the owners' inboxes -> barrier -> reduce + dequant
        (+ all-gather store + barrier).  No NCCL, no host sync: a fixed launch sequence, CUDA-graph capturable."""
        from . import _lib
        sets = self._symm[2]
        st = sets[self._flip]
        self._flip ^= 1
        xq_ws, sx_ws, amax_ws = self._workspace(M, x2.device)
        ks = self.k_hi - self.k_lo
        if not self.input_is_sharded:
            # a replicated input is handed over as its K-slice view: every rank reads 1/world of the activation for the
            # row maxima and exchanges them (tokens x 4 bytes per peer) instead of reading all of it
            x2 = x2[:, self.k_lo:self.k_hi]
        n_mine = self.n_hi - self.n_lo
        y_local = None
        if not self.gather_output:
            y_local = torch.empty((M, n_mine), dtype=out_dtype, device=x2.device)
        ldy = self.world * self.per_n if self.gather_output else max(n_mine, 1)
        with F._on(x2, up2, self.qweight_storage, self.weight_scale, self.bias):
            rc = _lib.lib().pq_rowparallel_forward(
                x2.data_ptr(), up2.data_ptr() if up2 is not None else None, F._DT[x2.dtype], F._ACTS[act],
                x2.stride(0), up2.stride(0) if up2 is not None else 0,
                1, x2.shape[1], 0,
                self.qweight_storage.data_ptr(), self.qweight_storage.stride(0), self.weight_scale.data_ptr(),
                self.bias.data_ptr() if self.bias is not None else None,
                st["sgp"], 1 if self.gather_output else 0,
                y_local.data_ptr() if (y_local is not None and n_mine) else (xq_ws.data_ptr() if not self.gather_output else None),
                F._DT[out_dtype], ldy, xq_ws.data_ptr(), sx_ws.data_ptr(), amax_ws.data_ptr(),
                M, self.out_features, ks, self.per_n, F._specp(self.spec), F._stream(x2.device))
        _lib.check(rc, "pq_rowparallel_forward")
        if not self.gather_output:
            return y_local
        return st["out"][:M, : self.out_features]

    def _check_fused_input(self, x2, up2):
        want = (self.k_hi - self.k_lo) if self.input_is_sharded else self.in_features
        if x2.shape[1] != want:
            raise ValueError(f"expected {want} input columns, got {x2.shape[1]}")
        for t in (x2, up2):
            if t is not None and (t.stride(1) != 1 or t.dtype not in F._DT):
                raise TypeError("inputs must have unit column stride and a supported dtype")
        if up2 is not None and (up2.shape != x2.shape or up2.dtype != x2.dtype or not self.input_is_sharded):
            raise ValueError("a gated input needs input_is_sharded=True and gate / up of equal shape and dtype")

    def forward(self, x: torch.Tensor, up: Optional[torch.Tensor] = None, act: str = "silu") -> torch.Tensor:
        """y = linear(x), or -- `up` given, K-sharded input only -- y = linear(act(x) * up) with the activation product
        computed and quantised on the fly (the Llama MLP's down projection fed by column-parallel gate / up slices).
        With `gather_output` on the fused path the result is a VIEW into a double-buffered symmetric-memory buffer,
        overwritten by the second forward() after this one; clone it to keep it longer."""
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        up2 = up.reshape(-1, up.shape[-1]) if up is not None else None
        out_dtype = self.out_dtype or x.dtype
        M = x2.shape[0]
        n_out = self.out_features if (self.gather_output or self.world == 1) else self.n_hi - self.n_lo
        if M and self.world > 1 and self.ops is _CudaShardOps and x2.is_cuda and self.fused is not False:
            if self._enable_fused(M, out_dtype, x2.device):
                self.fused = True
                if x2.stride(-1) != 1:
                    x2 = x2.contiguous()
                if up2 is not None and up2.stride(-1) != 1:
                    up2 = up2.contiguous()
                self._check_fused_input(x2, up2)
                return self._forward_fused(x2, up2, act, M, out_dtype).reshape(*lead, n_out)
        if up2 is not None:
            x2 = getattr(self.ops, "act_mul", F.act_mul)(x2, up2, act)
        xq, s_x = self._quantize(x2)
        acc = self.ops.int_mm(xq, self.qweight)              # exact int32 partial sums of this K-slice
        if self.world > 1:
            dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=self.group)
        if self.gather_output or self.world == 1:
            y = self.ops.epilogue(acc, s_x, self.weight_scale, self.bias, out_dtype)
        else:
            y = self.ops.epilogue(acc[:, self.n_lo:self.n_hi].contiguous(), s_x, self.weight_scale[self.n_lo:self.n_hi],
                                  self.bias[self.n_lo:self.n_hi] if self.bias is not None else None, out_dtype)
        return y.reshape(*lead, n_out)


class ParallelGatedMLP(nn.Module):
    """Tensor-parallel gated MLP (Llama: down(silu(gate(x)) * up(x))) on the dynamic int8 path, Megatron layout:
    gate / up column-parallel with NO gather, the activation product computed on the local column slice, down
    row-parallel with a K-sharded input.  Exchanges per forward: the row maxima of the hidden activation (`tokens`
    floats per rank, through symmetric memory -- the per-token scale needs the maximum over all shards) and the fused
    int32 reduce-scatter + all-gather of the down projection.  Because the hidden slices, the global row maxima and the
    int32 partial sums are all exact, the output is bit-identical to the same three DynamicQuantLinear modules
    chained on one GPU (with `act_mul_quant`'s definition of the activation product)."""

    def __init__(self, gate, up, down, group=None, act: str = "silu", out_dtype: Optional[torch.dtype] = None,
                 fused: Optional[bool] = None):
        """gate, up, down: unsharded DynamicQuantLinear modules (every rank passes the same ones)."""
        super().__init__()
        if gate.out_features != up.out_features or down.in_features != gate.out_features:
            raise ValueError("gate / up / down shapes do not form a gated MLP")
        self.act = act
        self.out_dtype = out_dtype
        kw = dict(group=group, out_dtype=out_dtype)
        self.gate = ShardedDynamicQuantLinear(gate.qweight, gate.weight_scale, gate.bias, spec=gate.spec,
                                              gather_output=False, align=16, **kw)
        self.up = ShardedDynamicQuantLinear(up.qweight, up.weight_scale, up.bias, spec=up.spec,
                                            gather_output=False, align=16, **kw)
        self.down = RowParallelDynamicQuantLinear(down.qweight, down.weight_scale, down.bias, spec=down.spec,
                                                  input_is_sharded=True, gather_output=True, fused=fused, **kw)
        if (self.gate.lo, self.gate.hi) != (self.down.k_lo, self.down.k_hi):
            raise ValueError("column shards of gate/up and K shards of down do not line up")

    def _gate_up(self, device):
        """gate and up shards concatenated into ONE local DynamicQuantLinear ([2 * per, K] int8): one act-quant + one
        GEMM per forward (per-output-channel scales make that exact).  Built on first use; not a registered submodule,
        so `state_dict()` keeps the two shards only."""
        g, u = self.gate, self.up
        # the copy is rebuilt when the shards change under it (load_state_dict, in-place edits): tensor version counters
        key = (device,) + tuple(t._version for m in (g, u) for t in (m.qweight_storage, m.weight_scale, m.bias) if t is not None)
        gu = self.__dict__.get("_gu")
        if gu is not None and self.__dict__.get("_gu_key") == key:
            return gu
        if (g.bias is None) != (u.bias is None) or g.spec != u.spec:
            return None
        from .modules import DynamicQuantLinear
        m = DynamicQuantLinear(g.in_features, 2 * g.per, g.bias is not None, device=device, out_dtype=self.out_dtype, spec=g.spec)
        m.qweight_storage.copy_(torch.cat([g.qweight_storage, u.qweight_storage], dim=0))
        m.weight_scale.copy_(torch.cat([g.weight_scale, u.weight_scale]))
        if m.bias is not None:
            m.bias.copy_(torch.cat([g.bias, u.bias]))
        object.__setattr__(self, "_gu", m)
        object.__setattr__(self, "_gu_key", key)
        return m

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Two C calls per forward: pq_qlinear (act-quant + the fused gate/up GEMM on the local column slice) and
        pq_rowparallel_forward with the gated input (silu(gate) * up computed, max-exchanged and quantised on the
        fly, int32 GEMM + reduce-scatter + all-gather): 9 launches, no NCCL, no host sync -- CUDA-graph capturable."""
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        dt = self.out_dtype or x.dtype
        n, per = self.gate.hi - self.gate.lo, self.gate.per
        gu_mod = self._gate_up(x2.device) if x2.is_cuda else None
        if gu_mod is not None:
            gu = gu_mod(x2)                                                # [M, 2 * per]
            g, u = gu[:, :n], gu[:, per:per + n]
        else:
            xq, s_x = F.quantize_act(x2, spec=self.gate.spec)             # one quantisation shared by gate and up
            g = F.qgemm(xq, s_x, self.gate.qweight, self.gate.weight_scale, self.gate.bias, dt)[:, :n]
            u = F.qgemm(xq, s_x, self.up.qweight, self.up.weight_scale, self.up.bias, dt)[:, :n]
        y = self.down(g, u, self.act)                                      # h = act(g) * u never leaves the kernels
        return y.reshape(*lead, self.down.out_features)


_HF_ACTS = {"SiLU": "silu", "SiLUActivation": "silu", "GELU": "gelu", "GELUActivation": "gelu", "NewGELUActivation": "gelu_tanh",
            "PytorchGELUTanh": "gelu_tanh", "GELUTanh": "gelu_tanh", "Identity": "identity"}


def parallelize_gated_mlps(model: nn.Module, group=None, names=("gate_proj", "up_proj", "down_proj"), act_attr: str = "act_fn",
                           out_dtype: Optional[torch.dtype] = None, spec: Optional[F.QuantSpec] = None,
                           fused: Optional[bool] = None) -> int:
    """Replace, in place, every sub-module of `model` that is a gated MLP -- children `names` = (gate, up, down) that are
    nn.Linear (or DynamicQuantLinear) and an activation module `act_attr` (Hugging Face's LlamaMLP / MistralMLP / Qwen2MLP
    layout) -- by a tensor-parallel `ParallelGatedMLP` over `group`.  Every rank calls it on the SAME full model (CUDA);
    each keeps its shards only.  Returns the number of modules replaced.  `forward(x)` keeps its signature.  The result
    equals the single-GPU chain `down(act_mul(gate(x), up(x)))` of this package bit for bit; against the float module it
    differs by the int8 quantisation (and by one bf16 ulp on a few elements: SiLU is evaluated in fp32 here)."""
    from .modules import DynamicQuantLinear
    replaced = 0
    for name, child in list(model.named_children()):
        parts = [getattr(child, n, None) for n in names]
        if all(isinstance(p, (nn.Linear, DynamicQuantLinear)) for p in parts):
            act_mod = getattr(child, act_attr, None)
            act = _HF_ACTS.get(type(act_mod).__name__) if act_mod is not None else "identity"
            if isinstance(act_mod, nn.GELU) and getattr(act_mod, "approximate", "none") == "tanh":
                act = "gelu_tanh"
            if act is None:
                raise ValueError(f"{name}: unsupported activation {type(act_mod).__name__} (silu / gelu / gelu_tanh / identity)")
            q = [p if isinstance(p, DynamicQuantLinear) else DynamicQuantLinear.from_float(p, out_dtype=out_dtype, spec=spec) for p in parts]
            setattr(model, name, ParallelGatedMLP(q[0], q[1], q[2], group=group, act=act, out_dtype=out_dtype, fused=fused))
            replaced += 1
        else:
            replaced += parallelize_gated_mlps(child, group, names, act_attr, out_dtype, spec, fused)
    return replaced


class TokenAdaptiveLinear(nn.Module):
    """Keeps the replicated DynamicQuantLinear next to its column-sharded twin and picks per call by token count:
    calls with fewer than `min_tokens` tokens run on the replicated copy.  Round 1 needed this (M = 16 on 8 GPUs:
    replicated 52 us vs sharded 64-71 us); since round 2 the sharded forward is one C call whose decode kernel writes
    all destinations itself and wins at every token count measured (profiles/README_r2.md), so `maybe_shard` no
    longer uses it by default -- it remains for fabrics where the cross-rank barrier is expensive.  The input is
    replicated, so every rank takes the same branch.  Costs (1 + 1/world) x the layer's int8 weights."""

    def __init__(self, replicated: nn.Module, sharded: nn.Module, min_tokens: int):
        super().__init__()
        self.replicated = replicated
        self.sharded = sharded
        self.min_tokens = int(min_tokens)
        self.in_features, self.out_features = sharded.in_features, sharded.out_features

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        tokens = x.numel() // max(1, x.shape[-1])
        return self.replicated(x) if tokens < self.min_tokens else self.sharded(x)


def maybe_shard(linear_q, group=None, min_out_features: int = 16384, min_tokens: int = 0, **kw):
    """Shard a DynamicQuantLinear across `group` only where that is measured to pay ("used only for layers big
    enough to benefit"): layers with fewer than `min_out_features` outputs stay replicated, and -- `min_tokens` > 0 --
    calls with fewer tokens than that run on the replicated copy (TokenAdaptiveLinear)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1 or linear_q.out_features < min_out_features:
        return linear_q
    sharded = ShardedDynamicQuantLinear(linear_q.qweight, linear_q.weight_scale, linear_q.bias, group=group,
                                        out_dtype=linear_q.out_dtype, spec=linear_q.spec, **kw)
    if min_tokens > 0:
        return TokenAdaptiveLinear(linear_q, sharded, min_tokens)
    return sharded
