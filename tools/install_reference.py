#!/usr/bin/env python
"""Installs the UNMODIFIED reference package into ``baseline/_ref`` (git-ignored, ships to the GPU box with gpurun).

``pip install --target baseline/_ref /root/reference/DeFT`` builds an empty wheel (the reference's pyproject has a flat
layout without package discovery), so the install is what pip would have done: the ``deft`` package directory copied
byte for byte to ``baseline/_ref/DeFT/deft``, plus a manifest of sha256 sums (``baseline/_ref/MANIFEST.json``) that
``tools/ref_triton_probe.py`` re-checks on the box.  Nothing under ``baseline/_ref`` is product source and nothing
under ``deft_b200/`` imports it; it exists so that the reference's own Triton operators can be timed and compared
on the same B200 as ours.

    python tools/install_reference.py            # in the build container (needs /root/reference)
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/DeFT/deft"
DST = os.path.join(ROOT, "baseline", "_ref", "DeFT", "deft")


def install() -> str:
    if not os.path.isdir(SRC):
        raise SystemExit(f"{SRC} is absent: the reference can only be installed in the build container")
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    manifest = {}
    for base, _, files in os.walk(DST):
        for f in sorted(files):
            path = os.path.join(base, f)
            manifest[os.path.relpath(path, DST)] = hashlib.sha256(open(path, "rb").read()).hexdigest()
    with open(os.path.join(ROOT, "baseline", "_ref", "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=1, sort_keys=True)
    return DST


if __name__ == "__main__":
    print(install(), file=sys.stderr)
