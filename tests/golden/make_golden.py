"""Generates tests/golden/*.npz.  Run once in the build container:  python tests/golden/make_golden.py

Two families of vectors:

* ``torch_ao_*``  — produced by torch.ao.quantization.fx._decomposed
  (choose_qparams_per_token :778-810, quantize_per_token :930-965, torch 2.11.0).  This is the
  nearest verifiable per-token symmetric int8 definition available in the container; it is
  a DIFFERENT PROJECT from protoquant (whose checkout is absent).  The oracle
  reproduces it with QuantSpec.torch_ao() (x*(1/s), eps=1e-5).
* ``int_mm_*``    — exact int32 products computed with a plain int64 numpy matmul.

The SPEC v0 default (true division, no eps) has no external producer; its known-answer
tests are hand-derived in tests/test_oracle.py.
"""
import os

import numpy as np
import torch
import torch.ao.quantization.fx._decomposed as D

HERE = os.path.dirname(os.path.abspath(__file__))


def bits(t: torch.Tensor) -> np.ndarray:
    if t.dtype == torch.float32:
        return t.numpy().view(np.uint32)
    return t.view(torch.int16).numpy().view(np.uint16)


def make_inputs(seed, M, K, dtype):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(M, K, generator=g)
    x[0, K // 2] = 100.0            # outlier row
    if M > 1:
        x[1].zero_()                # all-zero row
    if M > 2:
        x[2] *= 1e-7                # below the 1e-5 eps clamp
    if M > 3:
        x[3] = torch.round(x[3] * 4) / 4   # many exact ties after scaling
        x[3, 0] = 127.0
    return x.to(dtype)


def main():
    for name, dtype in (("bf16", torch.bfloat16), ("f16", torch.float16), ("f32", torch.float32)):
        for (M, K) in ((6, 64), (5, 768), (4, 1000)):
            x = make_inputs(1234 + M + K, M, K, dtype)
            # torch.ao runs bf16 inputs in NATIVE bf16 arithmetic (only fp16 is upcast, :796-799);
            # SPEC v0 fixes compute_dtype = fp32, so bf16 vectors are produced from the exact
            # fp32 upcast of the bf16 input (x_bits still stores the bf16 payload).
            xin = x.float() if dtype == torch.bfloat16 else x
            s, zp = D.choose_qparams_per_token(xin, torch.int8)
            q = D.quantize_per_token(xin, s, zp, -128, 127, torch.int8)
            np.savez_compressed(os.path.join(HERE, f"torch_ao_{name}_{M}x{K}.npz"),
                                x_bits=bits(x), q=q.numpy(), s=s.to(torch.float32).flatten().numpy(),
                                shape=np.array([M, K]), dtype=name)
    g = np.random.default_rng(7)
    for (M, N, K) in ((5, 24, 48), (33, 40, 256)):
        a = g.integers(-128, 128, (M, K), dtype=np.int8)
        b = g.integers(-128, 128, (N, K), dtype=np.int8)
        acc = (a.astype(np.int64) @ b.astype(np.int64).T).astype(np.int32)
        np.savez_compressed(os.path.join(HERE, f"int_mm_{M}x{N}x{K}.npz"), a=a, b=b, acc=acc)
    # worst case magnitudes: every product is (-128)*(-128), K = 28672 (the largest K in BASELINE.json)
    a = np.full((2, 28672), -128, np.int8)
    b = np.full((8, 28672), -128, np.int8)
    acc = (a.astype(np.int64) @ b.astype(np.int64).T)
    assert acc.max() < 2 ** 31
    np.savez_compressed(os.path.join(HERE, "int_mm_extreme_2x8x28672.npz"), a=a, b=b, acc=acc.astype(np.int32))


if __name__ == "__main__":
    main()
