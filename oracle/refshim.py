"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Runs the UNMODIFIED reference (open2c/cooler, /root/reference/src) without
h5py / libhdf5 by injecting stub modules for its missing dependencies and an
in-memory dict-of-arrays that stands in for an open HDF5 group (the same trick
the reference's own tests use: /root/reference/tests/conftest.py:8-106).

/root/reference only exists in the build container.  ``oracle/make_ref.py`` installs an
unmodified copy of the package into the git-ignored ``oracle/_ref/`` (it travels to the GPU box
with the built libraries), which is preferred when present.  This module is used
(a) by ``tests/golden/make_golden.py`` to freeze golden vectors,
(b) by CPU tests that are skipped when the reference is absent, and
(c) by ``bench.py``'s ``--impl reference`` / ``cpu_baseline`` legs to time the reference's own
    ``balance_cooler`` on the box's host cores (``kind: "reference"``).
The product package never imports it.
"""
from __future__ import annotations

import json
import logging
import os
import sys
import types

import numpy as np

_INSTALLED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_SRC = _INSTALLED if os.path.isdir(os.path.join(_INSTALLED, "cooler")) else "/root/reference/src"
_REGISTRY: dict = {}


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "cooler"))


def _install_stubs() -> None:
    if "h5py" in sys.modules and getattr(sys.modules["h5py"], "_cb200_stub", False):
        return
    h5py = types.ModuleType("h5py")
    h5py._cb200_stub = True

    class Dataset:  # isinstance checks only
        pass

    class AttributeManager(dict):
        pass

    class Group:
        # util.closing_hdf5 subclasses h5py.Group and calls super().__init__(grp.id)
        def __init__(self, id):
            self.id = id

        def __getitem__(self, k):
            return self.id[k]

        def __contains__(self, k):
            return k in self.id

        @property
        def attrs(self):
            return self.id.attrs

        @property
        def file(self):
            return self.id.file

        def keys(self):
            return self.id.keys()

    def File(filename, mode="r", **kw):
        return _REGISTRY[filename]

    h5py.Dataset, h5py.Group, h5py.File = Dataset, Group, File
    h5py.AttributeManager = AttributeManager
    h5py.is_hdf5 = lambda p: p in _REGISTRY
    h5py.check_dtype = lambda **kw: None
    h5py.special_dtype = lambda **kw: None
    sys.modules["h5py"] = h5py

    sj = types.ModuleType("simplejson")
    sj.__dict__.update({k: v for k, v in json.__dict__.items() if not k.startswith("__")})
    sys.modules["simplejson"] = sj

    cz = types.ModuleType("cytoolz")

    def compose(*fs):
        def f(*args, **kwargs):
            x = fs[-1](*args, **kwargs)
            for g in reversed(fs[:-1]):
                x = g(x)
            return x

        return f

    cz.compose = compose
    sys.modules["cytoolz"] = cz

    at = types.ModuleType("asciitree")
    at.BoxStyle = object
    at.LeftAligned = object
    att = types.ModuleType("asciitree.traversal")
    att.Traversal = object
    sys.modules["asciitree"] = at
    sys.modules["asciitree.traversal"] = att


def import_reference():
    """Return the unmodified reference package ``cooler`` (shimmed)."""
    if not available():
        raise RuntimeError("reference sources not present (GPU box?)")
    _install_stubs()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import cooler  # noqa: PLC0415

    import cooler._balance  # noqa: F401,PLC0415
    import cooler._reduce  # noqa: F401,PLC0415

    return cooler


class MemGroup(dict):
    """dict-of-dicts-of-arrays that quacks like an open h5py root group."""

    def __init__(self, name, tables, attrs):
        super().__init__(tables)
        self.attrs = attrs
        self.name = "/"
        self.filename = name
        self.mode = "r"
        self.file = self
        self.id = self
        _REGISTRY[name] = self

    def __getitem__(self, path):
        cur = self
        if set(path) == {"/"}:
            return self
        for item in path.strip("/").split("/"):
            cur = dict.__getitem__(cur, item)
        return cur

    def close(self):
        pass

    def __enter__(self):                  # `with h5py.File(path, mode) as h5:` (cli/balance.py:181, 195)
        return self

    def __exit__(self, *exc):
        return False


def register(name, chromnames, chromsizes, bins_chrom, bins_start, bins_end,
             bin1, bin2, count, binsize, symmetric_upper=True, extra_columns=None):
    """Register an in-memory cooler under file name ``name``.

    ``binsize`` None means variable-size bins.
    """
    bins_chrom = np.asarray(bins_chrom, dtype=np.int32)
    n_bins = len(bins_chrom)
    n_chroms = len(chromnames)
    chrom_offset = np.searchsorted(bins_chrom, np.arange(n_chroms + 1), side="left").astype(np.int64)
    bin1 = np.asarray(bin1, np.int64)
    bin1_offset = np.searchsorted(bin1, np.arange(n_bins + 1), side="left").astype(np.int64)
    tables = {
        "chroms": {"name": np.array(chromnames, dtype="S"),
                   "length": np.asarray(chromsizes, dtype=np.int32)},
        "bins": {"chrom": bins_chrom,
                 "start": np.asarray(bins_start, np.int32),
                 "end": np.asarray(bins_end, np.int32)},
        "pixels": {"bin1_id": bin1, "bin2_id": np.asarray(bin2, np.int64),
                   "count": np.asarray(count), **{k: np.asarray(v) for k, v in (extra_columns or {}).items()}},
        "indexes": {"chrom_offset": chrom_offset, "bin1_offset": bin1_offset},
    }
    attrs = {
        "bin-type": "fixed" if binsize is not None else "variable",
        "storage-mode": "symmetric-upper" if symmetric_upper else "square",
        "nchroms": n_chroms, "nbins": n_bins, "nnz": len(bin1), "metadata": "{}",
    }
    if binsize is not None:
        attrs["bin-size"] = int(binsize)
    else:
        attrs["bin-size"] = "null"  # reference: api.py binsize -> None when 'null'
    return MemGroup(name, tables, attrs)


class _CountHandler(logging.Handler):
    def __init__(self):
        super().__init__(level=logging.INFO)
        self.records = []

    def emit(self, record):
        self.records.append(record.getMessage())


def ref_balance(name, **kwargs):
    """Run the reference's balance_cooler on registered cooler ``name``.

    Returns (bias, stats, passes) where ``passes`` is the list of iteration
    counts: [n] genome-wide / trans, [n_chr1, n_chr2, ...] for cis_only
    (counted from the reference's own "variance is" log lines, _balance.py:109).
    """
    import warnings

    cooler = import_reference()
    clr = cooler.Cooler(name)
    lg = logging.getLogger("cooler")
    h = _CountHandler()
    old = lg.level
    lg.addHandler(h)
    lg.setLevel(logging.INFO)
    try:
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            bias, stats = cooler._balance.balance_cooler(clr, **kwargs)
        warned = any(issubclass(x.category, cooler._balance.ConvergenceWarning) for x in w)
    finally:
        lg.removeHandler(h)
        lg.setLevel(old)
    passes, cur = [], None
    cis = kwargs.get("cis_only", False)
    for m in h.records:
        if m.startswith("variance is"):
            if cur is None:
                cur = 0
            cur += 1
        elif cis:  # a chromosome-name line starts a new segment
            if cur is not None:
                passes.append(cur)
            cur = 0
    if cur is not None:
        passes.append(cur)
    return bias, stats, passes, warned


def ref_coarsen(name, factor, chunksize=10_000_000, columns=None, agg=None):
    """Concatenated chunk stream of the reference's CoolerCoarsener (_reduce.py:501-642)."""
    cooler = import_reference()
    columns = ["count"] if columns is None else list(columns)
    it = cooler._reduce.CoolerCoarsener(name, int(factor), chunksize,
                                       columns=columns, agg=agg, batchsize=1)
    chunks = list(it)
    out = {k: np.concatenate([c[k] for c in chunks]) if chunks else np.array([])
           for k in ["bin1_id", "bin2_id", *columns]}
    nb = it.new_bins
    new_bins = {"chrom": np.asarray(nb["chrom"].cat.codes if hasattr(nb["chrom"], "cat") else nb["chrom"]),
                "start": nb["start"].values, "end": nb["end"].values}
    return out, new_bins


def ref_merge(names, mergebuf=20_000_000, columns=None, agg=None):
    """Concatenated chunk stream of the reference's CoolerMerger (_reduce.py:132-208)."""
    cooler = import_reference()
    clrs = [cooler.Cooler(n) for n in names]
    it = cooler._reduce.CoolerMerger(clrs, mergebuf=mergebuf, columns=columns, agg=agg)
    chunks = list(it)
    return {k: np.concatenate([c[k] for c in chunks]) for k in chunks[0]}
