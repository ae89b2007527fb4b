"""TEST INFRASTRUCTURE ONLY -- installs the UNMODIFIED reference package into ``oracle/_ref/``.

    python oracle/make_ref.py

The reference (open2c/cooler v0.10.4) is pure Python, so "installing" it is what
``pip install --target oracle/_ref --no-deps /root/reference`` would do: the package directory
``src/cooler`` is copied as it is (pip itself cannot run here: the hatchling build backend is not in
the offline wheelhouse).  ``oracle/_ref/`` is git-ignored -- reference sources never enter the
history -- but it is NOT gpurun-ignored, so the copy travels to the GPU box, where
``/root/reference`` does not exist, and ``bench.py --impl reference`` / the ``cpu_baseline`` leg can
time the reference's own code (``kind: "reference"``) through ``oracle/refshim.py``.
Nothing under ``cooler_b200/`` imports it.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/cooler"
DST = os.path.join(HERE, "_ref", "cooler")


def tree_digest(root):
    h = hashlib.sha256()
    for d, _, files in sorted(os.walk(root)):
        for f in sorted(files):
            if f.endswith(".py"):
                p = os.path.join(d, f)
                h.update(os.path.relpath(p, root).encode())
                with open(p, "rb") as fh:
                    h.update(fh.read())
    return h.hexdigest()


def install(force=False):
    """Copy the reference package; returns the destination or None when the reference is absent."""
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(DST) else None
    if os.path.isdir(DST) and not force and tree_digest(DST) == tree_digest(SRC):
        return DST
    shutil.rmtree(os.path.dirname(DST), ignore_errors=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(os.path.dirname(DST), "INSTALLED_FROM"), "w") as f:
        f.write(f"{SRC}\nsha256(py files) {tree_digest(DST)}\n")
    return DST


if __name__ == "__main__":
    out = install(force="--force" in sys.argv)
    print(out or "reference sources not present; nothing installed")
