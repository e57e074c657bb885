"""Generate tests/golden/attacks.npz by running the reference's own attack modules (src/attack_methods) on CPU at their
test-time settings.  Build container only (needs /root/reference).  `params` is stubbed (attack_initializer.py imports
`opt` from it); the modules under test are imported unmodified."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import fixtures as fx  # noqa: E402

SIZE, B = 64, 2


def main():
    params = types.ModuleType("params")
    params.opt = types.SimpleNamespace(noise_sigma=0.1, blur_sigma=0.5, jpeg_quality=50, img_size=SIZE, device="cpu")
    sys.modules["params"] = params
    sys.path.insert(0, "/root/reference/src")
    from attack_methods.Gaussian_blur import Gaussian_blur
    from attack_methods.Jpeg_compression import Jpeg
    img = torch.tanh(fx.seeded((B, 3, SIZE, SIZE), 77))
    out = {"img": img.numpy()}
    for sigma in (0.5, 2.0):
        out[f"blur_{sigma}"] = Gaussian_blur(sigma=[sigma], is_train=False)(img).numpy()
    for q in (50, 90, 20):
        out[f"jpeg_{q}"] = Jpeg(False, q, SIZE)(img).detach().numpy()
    path = os.path.join(HERE, "attacks.npz")
    np.savez_compressed(path, **out)
    print("attacks.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
