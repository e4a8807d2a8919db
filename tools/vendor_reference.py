#!/usr/bin/env python
"""Vendors the UNMODIFIED reference (`/root/reference/src`, pure Python) into `baseline/_ref/src`.

`baseline/_ref/` is git-ignored (never part of the history) but NOT gpurun-ignored, so it travels to the GPU
box with the snapshot: `bench.py --impl reference` (host cores) and the `gpu_eager_baseline` leg (the same code
through PyTorch eager on the B200) import the reference from there.  The reference has no setup.py /
pyproject.toml, so `pip install --target baseline/_ref /root/reference` has nothing to install; its `src`
tree is an implicit namespace package that only needs its parent on sys.path (reference Dockerfile:28).

Nothing is edited: files are byte-for-byte copies, and a MANIFEST with their sha256 is written beside them so
the bench line can state which tree it timed.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("DDPM_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def vendor(force: bool = False) -> str:
    src_tree = os.path.join(SRC, "src")
    if not os.path.isdir(src_tree):
        if os.path.isdir(os.path.join(DST, "src")):
            return DST                       # GPU box: the prebuilt copy travelled with the snapshot
        raise FileNotFoundError(f"{src_tree} not found and no vendored copy under {DST}")
    man_path = os.path.join(DST, "MANIFEST.json")
    files = {}
    for d, _, fs in os.walk(src_tree):
        for f in fs:
            if f.endswith(".py"):
                p = os.path.join(d, f)
                files[os.path.relpath(p, SRC)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    if not force and os.path.exists(man_path):
        try:
            if json.load(open(man_path)).get("files") == files:
                return DST
        except Exception:
            pass
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    for rel in files:
        out = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), out)
    json.dump({"source": SRC, "files": files}, open(man_path, "w"), indent=1, sort_keys=True)
    return DST


if __name__ == "__main__":
    print(vendor(force="--force" in sys.argv))
