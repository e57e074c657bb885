"""Drive the UNMODIFIED reference (a copy of /root/reference/src shipped under baseline/_ref/src) through its own
public API: ``model.Generator``, ``GetGen.get_new_latent`` / ``generate_image`` and ``main.optimization``.

TEST / BENCH INFRASTRUCTURE ONLY (like everything under oracle/): imported by tests/, by ``bench.py --impl
reference`` and ``bench.py``'s cpu_baseline leg, and by tools/ref_gpu_bench.py.  Nothing under the product package
imports it.

The reference is a script drop (no setup.py), so "install" = copy: ``tools/ship_reference.py`` (called by
``__graft_entry__.build()`` when /root/reference is mounted) copies ``src/`` to the git-ignored ``baseline/_ref/src``
and pre-builds the reference's two JIT extensions for sm_100a into ``baseline/_ref/ext``; ``gpurun`` ships both to the
GPU box, where /root/reference does not exist.

Three things are stubbed, none of which is on the path being timed or compared:
  * ``custom_lpips`` (needs skimage / pip lpips / downloaded VGG16 weights, all absent offline) -> the reference's
    own MSE alternative (src/utils.py:46-47), stated wherever a number from this harness is reported;
  * ``params.opt`` is built from a synthetic argv (src/params.py:35 parses sys.argv at import);
  * on a box without a GPU (or with ``real_ops=False``) ``torch.utils.cpp_extension.load`` returns None - the
    reference never calls the extension modules on CPU tensors (src/op/upfirdn2d.py:150-153, src/op/fused_act.py:111).
"""
from __future__ import annotations

import os
import sys
import types
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = os.path.join(ROOT, "baseline", "_ref")
REF_SRC = os.path.join(REF_ROOT, "src")
REF_EXT = os.path.join(REF_ROOT, "ext")

_REF = None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "model.py")) and os.path.isdir(os.path.join(REF_SRC, "op"))


def import_reference(device: str = "cpu", real_ops: bool = False, img_size: int = 1024, key_len: int = 64,
                     shift: int = 448, sigma: float = 1.0) -> dict:
    """Import the shipped reference once per process.  ``real_ops=True`` lets the reference JIT-load its own CUDA
    extensions (needs a GPU box; the build directory is baseline/_ref/ext)."""
    global _REF
    if _REF is not None:
        if _REF["real_ops"] != real_ops:
            raise RuntimeError("the reference was already imported with a different real_ops setting in this process")
        _REF["opt"].device = device
        _REF["opt"].img_size, _REF["opt"].key_len, _REF["opt"].shift, _REF["opt"].sigma = img_size, key_len, shift, sigma
        return _REF
    if not available():
        raise RuntimeError(f"no reference under {REF_SRC}: run tools/ship_reference.py in the build container")
    for name in ("model", "op", "utils", "generator", "main", "params", "PCA"):
        if name in sys.modules and not getattr(sys.modules[name], "__file__", "").startswith(REF_SRC):
            raise RuntimeError(f"module '{name}' is already imported from {sys.modules[name].__file__}; the reference "
                               "harness must run in a process that has not imported this repo's package")
    import torch.utils.cpp_extension as ce
    if real_ops:
        os.makedirs(REF_EXT, exist_ok=True)
        os.environ.setdefault("TORCH_EXTENSIONS_DIR", REF_EXT)
        os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    else:
        ce.load = lambda *a, **k: None
    lp = types.ModuleType("custom_lpips")

    class PerceptualLoss:  # stand-in: the reference's own MSE option (src/utils.py:46-47)
        def __init__(self, *a, **k):
            pass

        def __call__(self, a, b):
            return F.mse_loss(a, b)

    lp.PerceptualLoss = PerceptualLoss
    sys.modules["custom_lpips"] = lp
    sys.path.insert(0, REF_SRC)
    argv = sys.argv
    sys.argv = ["main.py", "--model", "sg2", "--img_size", str(img_size), "--steps", "1", "--n", "1",
                "--key_len", str(key_len), "--shift", str(shift), "--sigma", str(sigma)]
    try:
        import params
    finally:
        sys.argv = argv
    params.opt.device = device
    import model
    import op
    from op.upfirdn2d import upfirdn2d_native
    import utils as rutils
    import generator as rgen
    import main as rmain
    _REF = dict(model=model, op=op, native=upfirdn2d_native, utils=rutils, gen=rgen, main=rmain, opt=params.opt,
                real_ops=real_ops)
    return _REF


def reference_generator(ref: dict, size: int, params: dict, cm: int = 2, device: str = "cpu"):
    g = ref["model"].Generator(size, 512, 8, channel_multiplier=cm)
    missing = g.load_state_dict(params, strict=False)
    assert not missing.unexpected_keys, missing
    assert all(("kernel" in k or k.startswith("noises.")) for k in missing.missing_keys), missing
    return g.eval().to(device)


class ReferenceLoop:
    """``main.optimization`` (src/main.py:45-89) of the shipped reference on one target image, with the module
    globals the reference's ``__main__`` block would have set (src/main.py:93-112) filled in from the same seeded
    fixtures bench.py and the tests use.  Parameters keep ``requires_grad=True`` as in the reference."""

    def __init__(self, size: int, params: dict, noise, pc, sigma_512, mean, key_len: int = 64, shift: int = 448,
                 sigma: float = 1.0, device: str = "cpu", real_ops: bool = False, alpha_seed: int = 5, key_seed: int = 6):
        import fixtures as fx
        self.ref = ref = import_reference(device, real_ops, size, key_len, shift, sigma)
        self.device, self.key_len, self.n_main = device, key_len, pc.shape[0] - key_len
        g = reference_generator(ref, size, params, device=device)
        for p in g.parameters():
            p.requires_grad_(True)
        sp = fx.split_basis(pc, sigma_512, key_len, shift, sigma)
        self.sp = sp = {k: v.to(device) for k, v in sp.items()}
        G = ref["gen"].GetGen
        fake = types.SimpleNamespace(sd_moved=1, key_len=key_len, batch_size=1, device=device, model="sg2",
                                     style_mixing=False, g_ema=g, latent_mean=mean.to(device), style_space_dim=pc.shape[0],
                                     num_main_pc=self.n_main)
        fake.get_new_latent = types.MethodType(G.get_new_latent, fake)
        fake.generate_image = types.MethodType(G.generate_image, fake)
        self.gen = fake
        self.noise = [n.to(device) for n in noise]
        alpha = sp["sigma_main"] * fx.seeded((self.n_main, 1), alpha_seed).to(device)
        torch.manual_seed(key_seed)
        with torch.no_grad():
            img, w0_t, wx_t, key = G.generate_with_alpha(fake, alpha, sp["u_cap"].t(), sp["sigma_key"], sp["v_cap"],
                                                         self.noise)
        self.target, self.key_true = img, key
        fake.key = key
        rm = ref["main"]
        rm.sigma_448 = sp["sigma_main"]
        rm.generator = fake
        rm.u_cap, rm.v_cap, rm.sigma_64 = sp["u_cap"], sp["v_cap"], sp["sigma_key"]
        rm.noise = self.noise
        rm.sigmoid = torch.nn.Sigmoid()
        rm.max_alpha, rm.min_alpha = sp["max_alpha"], sp["min_alpha"]
        rm.target_w0 = w0_t
        rm.tqdm = lambda it, *a, **k: it     # progress bar only (src/main.py:57)

    def run(self, lhs: np.ndarray, steps: int):
        """``lhs`` [n, n_main] in (0,1): the Latin-hypercube sample ``samlping.random(n)`` would return
        (scipy 1.18 rejects the reference's ``centered=True``).  Returns (losses, alphas, keys, acc)."""
        rm, opt = self.ref["main"], self.ref["opt"]
        opt.steps, opt.n = int(steps), int(lhs.shape[0])
        rm.samlping = types.SimpleNamespace(random=lambda n: lhs)
        rm.loss, rm.a, rm.k = [], [], []
        _, _, acc = rm.optimization(self.target)
        return list(rm.loss), [t.detach() for t in rm.a], [t.detach() for t in rm.k], float(acc)


def lhs_sample(n: int, d: int, seed: int) -> np.ndarray:
    """Centered Latin hypercube, ``(perm + 0.5) / n`` per dimension (what ``LatinHypercube(centered=True)`` gave
    the reference on scipy 1.7, src/main.py:103)."""
    return np.stack([(np.random.RandomState(seed + j).permutation(n) + 0.5) / n for j in range(d)], 1).astype(np.float32)
