"""CPU oracle for the latent-fingerprint attribution hot path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement (torch-CPU fp32/fp64
and numpy) of the reference's algorithm for the StyleGAN2 synthesis forward /
backward pass and the fingerprint embed.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker or the timed CPU
baseline - never as part of the product path.  The product path
(``attributing-image-generative-models-using-latent-fingerprints-sg2_b200/``)
fails loudly when its CUDA library is missing; it never routes through here.

Pinning: the reference holds no tests or golden vectors, so
the oracle is pinned against outputs of the reference itself, imported in the
build container from /root/reference/src by ``tests/golden/make_golden.py``; the
vectors live in ``tests/golden/*.npz`` and ``tests/test_oracle_golden.py``
checks the oracle against every one of them.
"""
from .sg2_oracle import *  # noqa: F401,F403
