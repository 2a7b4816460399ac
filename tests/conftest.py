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
