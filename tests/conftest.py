import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "attributing-image-generative-models-using-latent-fingerprints-sg2_b200")
for p in (ROOT, PKG, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # fp32 parity is checked against an fp32 oracle: pin convolutions to fp32 exactly as one would for the reference
    # (its cuDNN convs default to TF32); the tf32 tests select the tensor-core path explicitly
    import torch
    torch.backends.cudnn.allow_tf32 = False


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        # a deadlocked kernel must not hold the GPU box until the job's own limit: the "thread" method ends the process even
        # when the main thread is blocked inside a CUDA call
        try:
            import pytest_timeout  # noqa: F401
            for item in items:
                if "gpu" in item.keywords and item.get_closest_marker("timeout") is None:
                    item.add_marker(pytest.mark.timeout(900, method="thread"))
        except ImportError:
            pass
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    d = os.path.join(ROOT, "tests", "golden")
    out = {}
    for f in ("ops.npz", "modconv.npz", "generator.npz", "attribution.npz"):
        with np.load(os.path.join(d, f)) as z:
            out.update({k: z[k] for k in z.files})
    return out
