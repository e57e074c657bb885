"""The drop-in boundary, executed: the UNMODIFIED reference ``model.py`` (shipped copy under baseline/_ref/src, see
tools/ship_reference.py) running on this repo's native ops through the two bindings INTEGRATION.md describes, and on its
own JIT-built CUDA ops as the GPU-vs-GPU cross-check.  Each runs in a fresh interpreter (tests/dropin_runner.py) and
is compared with the CPU oracle at 256 px: image max-abs <= 1e-3 * max(1, max|ref|) (north_star's fp32 bar), latent
gradient <= 5e-3 relative (free-running: leaky-ReLU kink flips included, see test_synthesis_gpu.py)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HAVE_REF = os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "src", "model.py"))


def run(mode, size, timeout=900):
    out = subprocess.run([sys.executable, os.path.join(HERE, "dropin_runner.py"), mode, str(size)], capture_output=True,
                         text=True, timeout=timeout)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])


@pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref/src not shipped (run tools/ship_reference.py in the build container)")
@pytest.mark.parametrize("mode", ["optionA", "optionB"])
def test_reference_model_runs_on_the_native_ops(mode):
    r = run(mode, 256)
    print(r)
    assert r["libs"] == ["liblfp_sg2.so"], r["libs"]     # the native library, and none of the reference's extensions
    assert r["img_err"] <= 1e-3 * max(1.0, r["img_scale"]), r
    assert r["grad_rel"] <= 5e-3, r


@pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref/src not shipped")
def test_reference_on_its_own_cuda_ops_matches_the_oracle():
    """Pins the oracle against the reference's CUDA path (JIT-built upfirdn2d / fused_bias_act kernels + cuDNN fp32)."""
    r = run("refgpu", 256, timeout=1500)
    print(r)
    assert "liblfp_sg2.so" not in r["libs"]
    assert r["img_err"] <= 1e-3 * max(1.0, r["img_scale"]), r
    assert r["grad_rel"] <= 5e-3, r
