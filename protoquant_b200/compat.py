"""Alias table for the reference's public names.

REFERENCE ABSENT: the real protoquant checkout was not mounted, so its exact
function/class names could not be read.  Every reference-facing name lives in this one file
so that re-pointing them when the source is visible is a one-file change.  The aliases below
are PROVISIONAL guesses at common spellings, not citations.
"""
from .functional import dequantize as dequantize_tensor
from .functional import qgemm, qgemm_i32, qlinear, quantize_act, quantize_weight
from .modules import DynamicQuantLinear, swap_linear
from .qtensor import QTensor, dequantize, quantize

# provisional spellings
QLinear = DynamicQuantLinear
qlinear_from_linear = DynamicQuantLinear.from_float
quantize_per_token = quantize_act
quantize_per_channel = quantize_weight
int8_mm = qgemm_i32
int8_mm_dequant = qgemm
replace_linear = swap_linear

__all__ = [
    "QTensor", "quantize", "dequantize", "dequantize_tensor", "quantize_act", "quantize_weight",
    "qgemm", "qgemm_i32", "qlinear", "DynamicQuantLinear", "swap_linear",
    "QLinear", "qlinear_from_linear", "quantize_per_token", "quantize_per_channel",
    "int8_mm", "int8_mm_dequant", "replace_linear",
]
