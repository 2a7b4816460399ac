import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu through gpurun")


def load_golden_x(d):
    """Rebuild the exact input tensor of a torch_ao_*.npz golden file."""
    import numpy as np
    import torch
    name = str(d["dtype"])
    M, K = (int(v) for v in d["shape"])
    if name == "f32":
        return torch.from_numpy(d["x_bits"].view(np.float32).copy()).reshape(M, K)
    t = torch.from_numpy(d["x_bits"].view(np.int16).copy()).reshape(M, K)
    return t.view(torch.bfloat16 if name == "bf16" else torch.float16)


@pytest.fixture(scope="session")
def oracle():
    import protoquant_oracle
    return protoquant_oracle


EXACT_SPECS = (  # label in exact_*.npz -> (scale_mode, eps, qmin)
    ("div", 0, 0.0, -128), ("div_qmin127", 0, 0.0, -127), ("rcp_mul_eps1e5", 1, 1e-5, -128), ("inv_scale", 2, 0.0, -128))


def same_scales(a, b):
    """Bit equality of two fp32 scale vectors, except that any NaN equals any NaN (payload and sign of a NaN are
    not part of the contract: x86, numpy and the GPU produce different quiet NaNs)."""
    import numpy as np
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    both = np.isnan(a) & np.isnan(b)
    return bool(np.array_equal(np.where(both, 0, a.view(np.uint32)), np.where(both, 0, b.view(np.uint32))))
