#!/usr/bin/env python
"""Copy the reference's live sources to the git-ignored baseline/_ref/ so that they travel to the GPU box.

    python tools/ship_reference.py [--prebuild]

The reference is a script drop without setup.py / pyproject.toml, so `pip install --target baseline/_ref
/root/reference` has nothing to install; a verbatim copy of `src/` (minus the authors' dead-code folders and images)
is the install.  Nothing here is committed: baseline/_ref/ is in .gitignore (not in .gpurunignore).

--prebuild also JIT-builds the reference's own two CUDA extensions (src/op/upfirdn2d.py:11-17, src/op/fused_act.py:11-17)
for sm_100a into baseline/_ref/ext with the recipe of SURVEY.md section 9, so that the GPU box does not spend ~100 s on
it (it rebuilds by itself if ninja decides the shipped objects are stale).
"""
import argparse
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/src"
DST = os.path.join(ROOT, "baseline", "_ref")
SKIP_DIRS = {"unused_code_backup", "back_up_code", "__pycache__", "image", "weights"}


def ship() -> bool:
    if not os.path.isdir(SRC):
        print(f"ship_reference: {SRC} is not mounted; keeping whatever is under {DST}")
        return os.path.isdir(os.path.join(DST, "src"))
    dst_src = os.path.join(DST, "src")
    if os.path.isdir(dst_src):
        shutil.rmtree(dst_src)
    shutil.copytree(SRC, dst_src, ignore=lambda d, names: [n for n in names if n in SKIP_DIRS or n.endswith((".png", ".jpg", ".pth", ".pt"))])
    for lic in ("LICENSE", "LICENSE-NVIDIA", "LICENSE-Rosinality", "LICENSE-LPIPS", "LICENSE-FID"):
        p = os.path.join("/root/reference", lic)
        if os.path.isfile(p):
            shutil.copy(p, os.path.join(DST, lic))
    n = sum(len(f) for _, _, f in os.walk(dst_src))
    print(f"ship_reference: {n} files -> {dst_src}")
    return True


def prebuild() -> None:
    ext = os.path.join(DST, "ext")
    os.makedirs(ext, exist_ok=True)
    env = dict(os.environ, TORCH_EXTENSIONS_DIR=ext, TORCH_CUDA_ARCH_LIST="10.0a", MAX_JOBS="8")
    code = ("import sys; sys.path.insert(0, %r); import op.upfirdn2d, op.fused_act; print('reference extensions built')"
            % os.path.join(DST, "src"))
    subprocess.run([sys.executable, "-c", code], check=True, env=env, cwd=os.path.join(DST, "src"))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--prebuild", action="store_true")
    a = ap.parse_args()
    ok = ship()
    if ok and a.prebuild:
        prebuild()
