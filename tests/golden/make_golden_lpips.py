"""Generate tests/golden/lpips.npz by running the reference's own LPIPS network (src/custom_lpips/networks_basic.py PNetLin,
pnet_type='vgg', version 0.1, eval mode) on CPU with seeded RANDOM VGG16 weights (pnet_rand=True: the ImageNet weights cannot be
downloaded in this container) and the reference's shipped linear heads (src/custom_lpips/weights/v0.1/vgg.pth).

    python tests/golden/make_golden_lpips.py        (build container only: needs /root/reference)

Modules the reference imports but never uses on this path (skimage, IPython, pip `lpips`) are stubbed; `lpips.normalize_tensor`
is bound to the reference's identical local copy (custom_lpips/__init__.py:42-44).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.dirname(HERE)]
import fixtures as fx  # noqa: E402
from oracle import lpips_oracle as lo  # noqa: E402

REF = "/root/reference/src/custom_lpips"

CASES = [("b2_64", 2, 64, 7), ("b1_96x64", 1, (96, 64), 8)]


def import_reference_pnetlin():
    def normalize_tensor(in_feat, eps=1e-10):      # custom_lpips/__init__.py:42-44, verbatim semantics
        norm_factor = torch.sqrt(torch.sum(in_feat ** 2, dim=1, keepdim=True))
        return in_feat / (norm_factor + eps)

    for name in ("skimage", "skimage.color", "IPython", "pdb_stub"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["skimage"].color = sys.modules["skimage.color"]
    sys.modules["IPython"].embed = lambda *a, **k: None
    lp = types.ModuleType("lpips")
    lp.normalize_tensor = normalize_tensor
    sys.modules["lpips"] = lp
    pkg = types.ModuleType("custom_lpips")
    pkg.__path__ = [REF]
    sys.modules["custom_lpips"] = pkg
    for sub in ("pretrained_networks", "networks_basic"):
        spec = importlib.util.spec_from_file_location(f"custom_lpips.{sub}", os.path.join(REF, sub + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"custom_lpips.{sub}"] = mod
        spec.loader.exec_module(mod)
    return sys.modules["custom_lpips.networks_basic"]


def main():
    nb = import_reference_pnetlin()
    net = nb.PNetLin(pnet_type="vgg", pnet_rand=True, pnet_tune=False, use_dropout=True, spatial=False, version="0.1", lpips=True)
    heads = torch.load(os.path.join(REF, "weights", "v0.1", "vgg.pth"), map_location="cpu")
    params = lo.make_vgg_params(seed=5, lin_weights=heads)
    missing = net.load_state_dict(params, strict=False)
    assert not missing.unexpected_keys, missing
    assert all(k.startswith("scaling_layer") for k in missing.missing_keys), missing
    net.eval()
    out = {}
    for name, B, size, seed in CASES:
        h, w = (size, size) if isinstance(size, int) else size
        a = fx.seeded((B, 3, h, w), seed, scale=0.5)
        b = fx.seeded((B, 3, h, w), seed + 100, scale=0.5).requires_grad_(True)
        val = net.forward(a, b)
        (g,) = torch.autograd.grad(val.sum(), b)
        out[f"lpips/{name}/val"] = val.detach().numpy()
        out[f"lpips/{name}/grad_in1"] = g.numpy()
    for k in range(5):
        out[f"lpips/head{k}"] = heads[f"lin{k}.model.1.weight"].numpy()
    path = os.path.join(HERE, "lpips.npz")
    np.savez_compressed(path, **out)
    print("lpips.npz", len(out), "arrays", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
