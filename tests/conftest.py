import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) where they cannot run: no CUDA device, or no built library."""
    reason = None
    try:
        import torch
        if not torch.cuda.is_available():
            reason = "no CUDA device"
    except Exception as e:  # pragma: no cover
        reason = f"torch unavailable: {e}"
    lib = os.path.join(ROOT, "tante_b200", "lib", "libtante_b200.so")
    if reason is None and not os.path.exists(lib):
        reason = "tante_b200/lib/libtante_b200.so is not built (python -m tante_b200.build)"
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(bytes(z["meta_json"]).decode())
    return z, meta


def golden_cfg(meta):
    from oracle import tante_oracle as O
    return O.OracleConfig(**meta["cfg"])


def rel_l2(a, b):
    a, b = np.asarray(a), np.asarray(b)
    dt = np.complex128 if (np.iscomplexobj(a) or np.iscomplexobj(b)) else np.float64
    a = a.astype(dt)
    b = b.astype(dt)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
