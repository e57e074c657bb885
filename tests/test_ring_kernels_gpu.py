"""The bulk-copy ring kernels of the synthesis path (csrc/synth_kernels.cu: fir_ring_nhwc_kernel, act_bwd_ring_kernel) against
the register-streaming kernels they replace on large maps, at 512 px where the 128- and 64-channel variants run (the 32-channel
ones are exercised by the 1024 px tests).  The FIR ring forms its sums in the same order as the streaming kernel, so the image
must be bit-identical; the activation-backward ring groups the per-pixel style reductions into 2048-pixel segments instead
of 256-pixel ones, which changes the latent gradient at rounding level only."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def run_probe(tmp_path, name, **env):
    out = str(tmp_path / (name + ".npz"))
    e = dict(os.environ)
    e.update({k: str(v) for k, v in env.items()})
    subprocess.run([sys.executable, os.path.join(HERE, "_ring_probe.py"), out], check=True, env=e, timeout=600)
    return np.load(out)


def test_ring_kernels_match_streaming_kernels(tmp_path):
    ring = run_probe(tmp_path, "ring", LFP_FIR_RING=1, LFP_ACTBWD_RING=1)
    plain = run_probe(tmp_path, "plain", LFP_FIR_RING=0, LFP_ACTBWD_RING=0)
    fir_only = run_probe(tmp_path, "fir_only", LFP_FIR_RING=1, LFP_ACTBWD_RING=0)
    assert np.isfinite(ring["img"]).all() and np.isfinite(ring["d_lat"]).all()
    # forward: only the FIR differs, and it is bit-identical
    np.testing.assert_array_equal(ring["img"], plain["img"])
    # backward with the FIR ring alone: still bit-identical
    np.testing.assert_array_equal(fir_only["d_lat"], plain["d_lat"])
    # activation-backward ring: same per-thread sums, different partial grouping
    num = np.linalg.norm((ring["d_lat"] - plain["d_lat"]).ravel())
    den = np.linalg.norm(plain["d_lat"].ravel())
    assert num <= 1e-5 * den, (num, den)
