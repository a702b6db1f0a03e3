"""Shared test plumbing.  `-m gpu` tests need a B200; everything else runs on CPU."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = os.environ.get("DYNAMORPH_REFERENCE", "/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


class Golden:
    """A fixture file written by oracle/gen_golden.py (reference-module outputs)."""

    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"))

    def __getitem__(self, k):
        return self.z[k]

    def t(self, k):
        a = self.z[k]
        if a.dtype == np.float16:
            a = a.astype(np.float32)
        a = np.array(a, copy=True, order='C')            # keeps 0-dim scalars 0-dim
        return torch.from_numpy(a)

    def group(self, prefix):
        p = prefix.rstrip("/") + "/"
        return {k[len(p):]: self.t(k) for k in self.z.files if k.startswith(p)}

    def state(self, prefix="state"):
        st = self.group(prefix)
        # np.savez orders keys as written; restore reference state_dict order by name sort fallback
        return st

    def has(self, k):
        return k in self.z.files


@pytest.fixture(scope="session", params=["vqvae_default", "z16_masked", "z32_default", "vqvae_heavy"])
def golden_case(request):
    return Golden(request.param)


@pytest.fixture(scope="session")
def golden_default():
    return Golden("vqvae_default")


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))
