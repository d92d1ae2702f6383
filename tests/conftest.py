import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "graph-transformer_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a box without an sm_100 device; the driver's GPU tier runs them on a B200."""
    try:
        import u2gnn_b200
        have = u2gnn_b200.LIB.cdll.u2gnn_device_check() == 0
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no sm_100 (B200) device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def split_case(case):
    params = {k[len("param."):]: v for k, v in case.items() if k.startswith("param.")}
    grads = {k[len("grad."):]: v for k, v in case.items() if k.startswith("grad.")}
    after = {k[len("after."):]: v for k, v in case.items() if k.startswith("after.")}
    return params, grads, after


@pytest.fixture
def golden():
    return load_golden


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))
