"""Native mapping network and PCA set-up (lfp_mapping_*, lfp_pca_covariance) against the oracle's mapping
(pinned to the reference by tests/golden/generator.npz) and against sklearn.decomposition.PCA, the library the reference
calls (src/PCA.py:62-108)."""
import numpy as np
import pytest
import torch

import fixtures as fx
import oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_mapping_matches_oracle_and_module():
    from lfp_native.mapping import MappingPlan
    from model import Generator
    params = fx.make_params(32, 11)
    z = fx.seeded((257, 512), 5)
    plan = MappingPlan(512, 8, 0.01, device=DEV)
    plan.load(params)
    w = plan.forward(z.to(DEV)).cpu()
    ref = oracle.mapping(params, z)
    np.testing.assert_allclose(w.numpy(), ref.numpy(), rtol=1e-4, atol=1e-5 * float(ref.abs().max()))
    g = Generator(32, 512, 8)
    g.load_state_dict(params, strict=False)
    g = g.eval().to(DEV)
    with torch.no_grad():
        wm = g.style(z.to(DEV)).cpu()
    np.testing.assert_allclose(w.numpy(), wm.numpy(), rtol=1e-3, atol=1e-4 * float(ref.abs().max()))


def test_perform_pca_matches_sklearn():
    from sklearn.decomposition import PCA
    from generator import perform_pca
    from lfp_native.mapping import MappingPlan
    from model import Generator
    params = fx.make_params(32, 11)
    g = Generator(32, 512, 8)
    g.load_state_dict(params, strict=False)
    g = g.eval().to(DEV)
    n = 4000
    pc, sigma, mean = perform_pca(g, n_samples=n, seed=7)
    # the same latents through the same mapping, then the reference's library call
    gen = torch.Generator(device=DEV).manual_seed(7)
    z = torch.randn(n, 512, device=DEV, generator=gen)
    plan = MappingPlan(512, 8, 0.01, device=DEV)
    plan.load(params)
    w = plan.forward(z).cpu().numpy()
    sk = PCA().fit(w)                                       # src/PCA.py:72-73
    np.testing.assert_allclose(mean.cpu().numpy()[:, 0], sk.mean_, rtol=1e-4, atol=1e-5)
    sig_ref = np.sqrt(sk.explained_variance_)               # src/PCA.py sigma = sqrt(explained_variance_)
    np.testing.assert_allclose(sigma.cpu().numpy()[:400, 0], sig_ref[:400], rtol=2e-3)
    # principal axes agree up to sign where the spectrum is not degenerate
    dots = np.abs(np.sum(pc.cpu().numpy()[:20] * sk.components_[:20], axis=1))
    gaps = np.abs(np.diff(sk.explained_variance_[:21])) / sk.explained_variance_[:20]
    assert np.all(dots[gaps > 1e-2] > 0.999), (dots, gaps)
    # orthonormal basis
    eye = pc.cpu().double() @ pc.cpu().double().t()
    assert float((eye - torch.eye(512, dtype=torch.float64)).abs().max()) < 1e-4
