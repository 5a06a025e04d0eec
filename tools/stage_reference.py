#!/usr/bin/env python
"""Stage the UNMODIFIED reference into baseline/_ref/ so that `bench.py --impl reference` can run it on the GPU
box (where /root/reference does not exist).

The reference has no setup.py / pyproject.toml, so `pip install --target baseline/_ref /root/reference` has
nothing to build (recorded in DESIGN.md); its "installation" is its source tree on sys.path, which is how its own
scripts import it (`sys.path.append(parent)`, tokenizer/hyperbolic_merge.py:34).  Only the packages on the merge-loop
path are staged: embedding/ and tokenizer/ (pure Python).  baseline/_ref is git-ignored and NOT gpurun-ignored:
it travels with the snapshot and never enters the history.

    python tools/stage_reference.py [--src /root/reference] [--dst baseline/_ref]
"""
import argparse
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PACKAGES = ("embedding", "tokenizer")


def stage(src: str = "/root/reference", dst: str = os.path.join(ROOT, "baseline", "_ref")) -> bool:
    if not os.path.isdir(src):
        return False
    os.makedirs(dst, exist_ok=True)
    for pkg in PACKAGES:
        out = os.path.join(dst, pkg)
        if os.path.isdir(out):
            shutil.rmtree(out)
        shutil.copytree(os.path.join(src, pkg), out, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(dst, "STAGED_FROM"), "w") as f:
        f.write(src + "\n")
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ap.add_argument("--dst", default=os.path.join(ROOT, "baseline", "_ref"))
    a = ap.parse_args()
    ok = stage(a.src, a.dst)
    print("staged" if ok else f"{a.src} not found: nothing staged", a.dst)
    sys.exit(0 if ok else 1)
