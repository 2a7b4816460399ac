"""protoquant_b200 — B200-native (sm_100a) dynamic-quantized int8 linear path.

Python surface over the C ABI in include/protoquant_b200.h.
No Triton, no backend dispatch, no CPU fallback.
"""
from ._lib import ProtoquantError, launch_count, lib
from .functional import (DEFAULT_SPEC, QuantSpec, act_mul_quant, dequantize as dequantize_tensor, layernorm_quant,
                         norm_quant, qgemm, qgemm_i32, qlinear, quantize_act, quantize_weight, rmsnorm_quant)
from .modules import DynamicQuantLinear, SharedInputLinear, fuse_linears, swap_linear
from .qtensor import QTensor, dequantize, quantize
from .sharded import (ParallelGatedMLP, RowParallelDynamicQuantLinear, ShardedDynamicQuantLinear, TokenAdaptiveLinear,
                      maybe_shard, parallelize_gated_mlps, shard_bounds)

__version__ = "0.2.0"
__all__ = [
    "ProtoquantError", "launch_count", "lib", "QuantSpec", "DEFAULT_SPEC",
    "quantize_act", "quantize_weight", "qgemm", "qgemm_i32", "qlinear", "dequantize_tensor",
    "norm_quant", "rmsnorm_quant", "layernorm_quant", "act_mul_quant",
    "QTensor", "quantize", "dequantize", "DynamicQuantLinear", "SharedInputLinear", "swap_linear", "fuse_linears",
    "ShardedDynamicQuantLinear", "RowParallelDynamicQuantLinear", "ParallelGatedMLP", "TokenAdaptiveLinear", "maybe_shard", "parallelize_gated_mlps", "shard_bounds",
]
