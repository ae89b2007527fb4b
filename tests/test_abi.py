"""No-GPU checks of the drop-in boundary: the C-ABI library loads and exports
every symbol include/cooler_b200.h declares; the product never imports oracle/."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "cooler_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cb200_[a-z0-9_]+)\s*\(", text)))


def test_build_and_load():
    import __graft_entry__ as g
    out = g.build()
    assert os.path.exists(out)


def test_library_exports_every_declared_symbol():
    from cooler_b200 import _lib
    lib = _lib.load()
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), s
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(syms)
    assert lib.cb200_abi_version() == 3
    assert lib.cb200_warp_tile() == 2048
    assert lib.cb200_scan_workspace_bytes(10_000_000) > 0
    assert lib.cb200_sort_workspace_bytes(10_000_000) > 0


def test_struct_layout_matches_header(tmp_path):
    """ctypes mirrors of cb200_copy / cb200_ice agree with what gcc makes of the header."""
    import ctypes as C
    import subprocess

    from cooler_b200 import _lib
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "cooler_b200.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(cb200_copy), sizeof(cb200_ice),'
        ' offsetof(cb200_ice, n_bins), offsetof(cb200_ice, bias), offsetof(cb200_ice, tol),'
        ' offsetof(cb200_ice, max_iters), offsetof(cb200_ice, far_copies), offsetof(cb200_ice, far_warps),'
        ' sizeof(cb200_far_copy), offsetof(cb200_far_copy, rb_unit), sizeof(cb200_far_unit), sizeof(cb200_sub),'
        ' offsetof(cb200_sub, val), offsetof(cb200_ice, float_counts));'
        ' return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    want = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    got = [C.sizeof(_lib.Copy), C.sizeof(_lib.Ice), _lib.Ice.n_bins.offset, _lib.Ice.bias.offset,
           _lib.Ice.tol.offset, _lib.Ice.max_iters.offset, _lib.Ice.far_copies.offset, _lib.Ice.far_warps.offset,
           C.sizeof(_lib.FarCopy), _lib.FarCopy.rb_unit.offset, C.sizeof(_lib.FarUnit), C.sizeof(_lib.Sub),
           _lib.Sub.val.offset, _lib.Ice.float_counts.offset]
    assert got == want


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "cooler_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
                assert "ice_numpy" not in src and "coarsen_numpy" not in src, fn


def test_no_cpu_fallback():
    import numpy as np
    import torch

    from cooler_b200 import PixelTable, _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises((_lib.NativeLibraryError, RuntimeError, AssertionError)):
        PixelTable.from_arrays(np.array([0, 1, 2]), np.array([0, 1]), np.array([1, 1], dtype=np.int32),
                               np.array([0, 2]), device="cpu")
