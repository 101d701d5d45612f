"""Recipe: put an UNMODIFIED copy of the reference's SNAG_MMEA tree under baseline/_ref/ so that it travels to the GPU
box (baseline/_ref/ is git-ignored — the reference's sources never enter this repository's history — but not
gpurun-ignored). The reference is a pure-Python research checkout without packaging (no setup.py / pyproject), so
`pip install --target baseline/_ref /root/reference` has nothing to install; the recipe is a byte-for-byte copy plus
a manifest of sha256 digests that tests/test_patch.py re-checks.

    python baseline/install_ref.py            # no-op when /root/reference is absent (e.g. on the GPU box)
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/SNAG_MMEA"
DST = os.path.join(HERE, "_ref", "SNAG_MMEA")
MANIFEST = os.path.join(HERE, "_ref", "MANIFEST.json")


def _digests(root: str) -> dict:
    out = {}
    for base, _, files in os.walk(root):
        if "__pycache__" in base:
            continue
        for f in sorted(files):
            if f.endswith(".pyc"):
                continue
            p = os.path.join(base, f)
            with open(p, "rb") as fh:
                out[os.path.relpath(p, root)] = hashlib.sha256(fh.read()).hexdigest()
    return out


def install(force: bool = False) -> str | None:
    """Copy the reference tree; returns the installed path, or None when there is no reference checkout here."""
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(DST) else None
    want = _digests(SRC)
    if not force and os.path.isfile(MANIFEST):
        with open(MANIFEST) as fh:
            if json.load(fh).get("files") == want and _digests(DST) == want:
                return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(MANIFEST, "w") as fh:
        json.dump({"source": SRC, "files": want}, fh, indent=1, sort_keys=True)
    return DST


def verify() -> bool:
    """True when baseline/_ref holds exactly the files the manifest lists (nothing edited, added or removed)."""
    if not (os.path.isdir(DST) and os.path.isfile(MANIFEST)):
        return False
    with open(MANIFEST) as fh:
        return json.load(fh).get("files") == _digests(DST)


if __name__ == "__main__":
    path = install(force="--force" in sys.argv)
    print(path if path else "no reference checkout at " + SRC)
