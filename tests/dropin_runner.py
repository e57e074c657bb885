"""Subprocess body of tests/test_reference_dropin_gpu.py: run the UNMODIFIED reference ``model.py`` (baseline/_ref/src)
on top of this repo's native ops and compare with the CPU oracle.

    python tests/dropin_runner.py optionA|optionB|refgpu SIZE

optionA  INTEGRATION.md Option A: this repo's ``op`` package is importable as ``op`` before the reference's model.py.
optionB  INTEGRATION.md Option B: the reference's own ``op/*.py`` with the two ``load(...)`` statements replaced by the
         ctypes stub printed in INTEGRATION.md (extracted from the document, so the document is what is tested).
refgpu   the reference on its own JIT-built CUDA ops (baseline/_ref/ext) - the GPU-vs-GPU cross-check of the oracle.
A separate process is required because the reference's module names (model, op) are the drop-in's names too.
Prints one JSON line.
"""
import importlib.util
import json
import os
import re
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = os.path.join(ROOT, "attributing-image-generative-models-using-latent-fingerprints-sg2_b200")
REF_SRC = os.path.join(ROOT, "baseline", "_ref", "src")


def load_reference_model():
    spec = importlib.util.spec_from_file_location("ref_model", os.path.join(REF_SRC, "model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    mode, size = sys.argv[1], int(sys.argv[2])
    import torch
    torch.backends.cudnn.allow_tf32 = False      # fp32 oracle: pin the reference's cuDNN convs to fp32
    sys.path[:0] = [ROOT, HERE]
    import fixtures as fx
    import oracle
    loaded = []
    if mode == "optionA":
        sys.path.insert(0, PKG)
        import op  # noqa: F401  this repo's drop-in package
        assert op.__file__.startswith(PKG)
    elif mode == "optionB":
        tmp = tempfile.mkdtemp(prefix="lfp_optB_")
        shutil.copytree(os.path.join(REF_SRC, "op"), os.path.join(tmp, "op"))
        doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
        stub = re.search(r"```python\n# src/op/_lfp.py.*?\n(.*?)```", doc, re.S).group(0)
        stub = stub[len("```python\n"):-3].replace("/path/to/liblfp_sg2.so", os.path.join(PKG, "lfp_native", "liblfp_sg2.so"))
        open(os.path.join(tmp, "op", "_lfp.py"), "w").write(stub)
        for fname, sym in (("upfirdn2d.py", "upfirdn2d_op"), ("fused_act.py", "fused")):
            p = os.path.join(tmp, "op", fname)
            src = open(p).read()
            new, n = re.subn(r"\n" + sym + r" = load\(.*?\n\)\n", f"\nfrom ._lfp import {sym}\n", src, flags=re.S)
            assert n == 1, (fname, n)
            open(p, "w").write(new)
        sys.path.insert(0, tmp)
        import op  # noqa: F401  the reference's op/*.py over the ctypes stub
        assert op.__file__.startswith(tmp)
    elif mode == "refgpu":
        os.environ.setdefault("TORCH_EXTENSIONS_DIR", os.path.join(ROOT, "baseline", "_ref", "ext"))
        os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
        sys.path.insert(0, REF_SRC)
        import op  # noqa: F401  JIT-loads the reference's own extensions
        assert op.__file__.startswith(REF_SRC)
    else:
        raise SystemExit(f"unknown mode {mode}")
    M = load_reference_model()
    seed, B = 40 + size, 2
    params = fx.make_params(size, seed)
    g = M.Generator(size, 512, 8)
    missing = g.load_state_dict(params, strict=False)
    assert not missing.unexpected_keys
    g = g.eval().cuda()
    noise = fx.make_noise(size, seed + 1)
    w = fx.seeded((B, 512), seed + 2)
    wr = w.clone().requires_grad_(True)
    ref = oracle.generator_forward(params, [wr], size, input_is_latent=True, noise=noise)
    ct = fx.seeded(tuple(ref.shape), seed + 3)
    (gref,) = torch.autograd.grad((ref * ct).sum(), wr)
    wg = w.cuda().requires_grad_(True)
    img, _ = g([wg], input_is_latent=True, noise=[n.cuda() for n in noise])
    (gw,) = torch.autograd.grad((img * ct.cuda()).sum(), wg)
    torch.cuda.synchronize()
    with open("/proc/self/maps") as f:
        for line in f:
            if line.rstrip().endswith(".so") and ("liblfp_sg2" in line or "/ext/" in line):
                loaded.append(os.path.basename(line.split()[-1]))
    print(json.dumps({"mode": mode, "size": size,
                      "img_err": float((img.detach().cpu() - ref.detach()).abs().max()),
                      "img_scale": float(ref.detach().abs().max()),
                      "grad_rel": float((gw.cpu() - gref).norm() / gref.norm()),
                      "libs": sorted(set(loaded))}))


if __name__ == "__main__":
    main()
