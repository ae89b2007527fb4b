"""Build ``libcooler_b200.so`` in-tree with nvcc for sm_100a.

    python -m cooler_b200.build [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC_DIR = os.path.join(_HERE, "csrc")
SOURCES = ["prims.cu", "table.cu", "ice.cu", "coarsen.cu", "fetch.cu", "h5decode.cpp"]
HEADERS = ["common.cuh", "walk.cuh", "far.cuh", os.path.join("..", "..", "include", "cooler_b200.h")]
OUT = os.path.join(_HERE, "libcooler_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "-lz", "-lpthread",
]


def source_hash(extra=()):
    """sha256 over the sources, headers and compiler flags the library is built from."""
    import hashlib
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        with open(os.path.join(SRC_DIR, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    h.update(" ".join(NVCC_FLAGS + list(extra)).encode())
    return h.hexdigest()


def _stale(out=OUT, extra=()):
    """The library is rebuilt unless the hash recorded next to it matches the current sources
    (modification times are not trusted: a stale-but-newer .so must never be used silently)."""
    tag = out + ".srchash"
    if not (os.path.exists(out) and os.path.exists(tag)):
        return True
    with open(tag) as f:
        return f.read().strip() != source_hash(extra)


def build(force=False, verbose=False, far_row_shift=None):
    """Build the library.  ``far_row_shift`` (12..14) builds a variant with another far-field
    row-block size next to it (``libcooler_b200_rs<k>.so``; select it with ``CB200_LIB``)."""
    out, extra = OUT, []
    if far_row_shift is not None:
        out = os.path.join(_HERE, f"libcooler_b200_rs{int(far_row_shift)}.so")
        extra = [f"-DCB200_FAR_ROW_SHIFT_VALUE={int(far_row_shift)}"]
    if not force and not _stale(out, extra):
        return out
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-o", out] + [os.path.join(SRC_DIR, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    with open(out + ".srchash", "w") as f:
        f.write(source_hash(extra) + "\n")
    return out


if __name__ == "__main__":
    rs = int(sys.argv[sys.argv.index("--far-row-shift") + 1]) if "--far-row-shift" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, far_row_shift=rs))
