"""In-memory cooler store.

``MemCooler`` exposes the part of ``cooler.Cooler`` (reference
src/cooler/api.py:32-409) that the balancing / coarsening hot path touches, on
top of plain numpy arrays laid out like the cooler schema (docs/schema_v3.rst):
groups ``chroms``, ``bins``, ``pixels``, ``indexes`` plus the info attributes.
It is the input/output object of the array-level API and of the tests (h5py /
libhdf5 are not available in the build image); a real ``cooler.Cooler`` works
with the same façade functions because only this protocol is used.
"""
from __future__ import annotations

from contextlib import contextmanager

import numpy as np


# ---------------------------------------------------------------------------
# URI resolution: "mem://name" (or any registered name) -> in-memory store, anything
# else -> the real cooler package (h5py).  A registered value may be a MemCooler or a
# dict {resolution: MemCooler} standing for an .mcool ("name::resolutions/<res>").
# ---------------------------------------------------------------------------
REGISTRY: dict = {}


def parse_cooler_uri(s):
    """'path.cool::/group' -> (path, '/group')   (util.parse_cooler_uri, util.py:36-52)."""
    parts = s.split("::")
    if len(parts) == 1:
        return parts[0], "/"
    if len(parts) == 2:
        g = parts[1]
        return parts[0], g if g.startswith("/") else "/" + g
    raise ValueError("Invalid Cooler URI string")


def register(name, obj):
    REGISTRY[parse_cooler_uri(name)[0]] = obj
    return obj


def is_memory_uri(uri):
    return parse_cooler_uri(uri)[0] in REGISTRY


def open_cooler(uri):
    """Resolve a cooler URI to a Cooler-like object."""
    if hasattr(uri, "_load_dset"):
        return uri
    path, group = parse_cooler_uri(uri)
    if path in REGISTRY:
        obj = REGISTRY[path]
        if isinstance(obj, dict):
            parts = [p for p in group.split("/") if p]
            if len(parts) == 2 and parts[0] == "resolutions" and int(parts[1]) in obj:
                return obj[int(parts[1])]
            raise KeyError(f"No cooler found at: {path}::{group}")      # api.py:101-109
        if group != "/":
            raise KeyError(f"No cooler found at: {path}::{group}")
        return obj
    try:
        import cooler
    except ImportError:
        # no h5py / libhdf5 here: read the file with the built-in HDF5 subset reader (read-only)
        return H5Cooler(path, group)
    return cooler.Cooler(uri)


class MemDataset:
    """A numpy array with ``attrs`` (stand-in for h5py.Dataset)."""

    def __init__(self, data):
        self.data = np.asarray(data)
        self.attrs = {}

    def __getitem__(self, key):
        return self.data[key]

    def __len__(self):
        return len(self.data)

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def shape(self):
        return self.data.shape

    def __array__(self, dtype=None, copy=None):
        return self.data if dtype is None else self.data.astype(dtype)


class MemGroup(dict):
    """dict of MemDataset / MemGroup with an ``attrs`` dict (stand-in for h5py.Group)."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.attrs = {}

    def __getitem__(self, path):
        cur = self
        for item in str(path).strip("/").split("/"):
            if item:
                cur = dict.__getitem__(cur, item)
        return cur

    def __contains__(self, path):
        try:
            self[path]
            return True
        except KeyError:
            return False

    def create_dataset(self, name, data=None, **_h5opts):
        ds = MemDataset(np.array(data))
        dict.__setitem__(self, name, ds)
        return ds

    def create_group(self, name):
        g = MemGroup()
        dict.__setitem__(self, name, g)
        return g


class _Selector:
    """``clr.bins()[:]`` / ``clr.chroms()["name"][:]`` style access."""

    def __init__(self, group, fields=None, decode=None):
        self._g, self._fields, self._decode = group, fields, decode or {}

    def _frame(self, sl):
        import pandas as pd
        cols = self._fields or list(self._g.keys())
        data = {}
        for c in cols:
            v = self._g[c][sl]
            if c in self._decode:
                v = self._decode[c](v)
            data[c] = v
        return pd.DataFrame(data)

    def __getitem__(self, key):
        if isinstance(key, str):
            return _Selector(self._g, [key], self._decode)._column()
        if isinstance(key, list):
            return _Selector(self._g, key, self._decode)
        return self._frame(key)

    def _column(self):
        sel = self

        class _Col:
            def __getitem__(self, sl):
                return sel._frame(sl)[sel._fields[0]]
        return _Col()

    @property
    def dtypes(self):
        import pandas as pd
        return pd.Series({k: self._g[k].dtype for k in (self._fields or self._g.keys())})

    def __len__(self):
        k = next(iter(self._g.keys()))
        return len(self._g[k])


class MemCooler:
    """Duck-typed ``cooler.Cooler`` over an in-memory ``MemGroup``."""

    def __init__(self, root: MemGroup, filename="<memory>"):
        self.root = root
        self.filename = filename
        self.uri = filename

    # ---- construction ----------------------------------------------------
    @classmethod
    def from_arrays(cls, chromnames, chromsizes, bins_chrom, bins_start, bins_end,
                    bin1_id, bin2_id, count, *, binsize, symmetric_upper=True, filename="<memory>",
                    extra_pixel_columns=None):
        bins_chrom = np.asarray(bins_chrom, dtype=np.int32)
        n_bins = len(bins_chrom)
        n_chroms = len(chromnames)
        bin1_id = np.asarray(bin1_id, dtype=np.int64)
        root = MemGroup()
        g = root.create_group("chroms")
        g.create_dataset("name", np.asarray(chromnames, dtype="S"))
        g.create_dataset("length", np.asarray(chromsizes, dtype=np.int32))
        g = root.create_group("bins")
        g.create_dataset("chrom", bins_chrom)
        g.create_dataset("start", np.asarray(bins_start, dtype=np.int32))
        g.create_dataset("end", np.asarray(bins_end, dtype=np.int32))
        g = root.create_group("pixels")
        g.create_dataset("bin1_id", bin1_id)
        g.create_dataset("bin2_id", np.asarray(bin2_id, dtype=np.int64))
        g.create_dataset("count", np.asarray(count))
        for k, v in (extra_pixel_columns or {}).items():
            g.create_dataset(k, np.asarray(v))
        g = root.create_group("indexes")
        g.create_dataset("chrom_offset",
                         np.searchsorted(bins_chrom, np.arange(n_chroms + 1)).astype(np.int64))
        g.create_dataset("bin1_offset",
                         np.searchsorted(bin1_id, np.arange(n_bins + 1)).astype(np.int64))
        cnt = np.asarray(count)
        root.attrs.update({
            "format": "HDF5::Cooler", "format-version": 3,
            "bin-type": "fixed" if binsize is not None else "variable",
            "bin-size": int(binsize) if binsize is not None else None,
            "storage-mode": "symmetric-upper" if symmetric_upper else "square",
            "nchroms": n_chroms, "nbins": n_bins, "nnz": len(bin1_id),
            "sum": cnt.sum().item() if len(cnt) else 0,
        })
        return cls(root, filename=filename)

    @classmethod
    def from_chromsizes(cls, chromnames, chromsizes, binsize, bin1_id, bin2_id, count, **kw):
        """Uniform ``binsize`` bins over the chromosomes (util.binnify semantics)."""
        c, s, e = [], [], []
        for i, L in enumerate(chromsizes):
            st = np.arange(0, int(L), int(binsize), dtype=np.int64)
            c.append(np.full(len(st), i, np.int32))
            s.append(st)
            e.append(np.minimum(st + int(binsize), int(L)))
        return cls.from_arrays(chromnames, chromsizes, np.concatenate(c), np.concatenate(s),
                               np.concatenate(e), bin1_id, bin2_id, count, binsize=binsize, **kw)

    # ---- the cooler.Cooler protocol used by the hot path -----------------
    @property
    def info(self):
        return dict(self.root.attrs)

    @property
    def binsize(self):
        return self.root.attrs.get("bin-size")

    @property
    def storage_mode(self):
        return self.root.attrs.get("storage-mode", "symmetric-upper")

    @property
    def chromnames(self):
        return [n.decode() if isinstance(n, bytes) else str(n) for n in self.root["chroms/name"][:]]

    @property
    def chromsizes(self):
        import pandas as pd
        return pd.Series(self.root["chroms/length"][:].astype(np.int64), index=self.chromnames,
                         name="length")

    def _load_dset(self, path):
        return self.root[path][:]

    @contextmanager
    def open(self, mode="r", **kwargs):
        yield self.root

    def chroms(self, **kwargs):
        return _Selector(self.root["chroms"],
                         decode={"name": lambda v: np.array([x.decode() if isinstance(x, bytes) else x for x in v],
                                                            dtype=object)})

    def bins(self, **kwargs):
        return _Selector(self.root["bins"])

    def pixels(self, join=False, **kwargs):
        if join:
            raise NotImplementedError("MemCooler.pixels(join=True)")
        return _Selector(self.root["pixels"])

    @property
    def shape(self):
        n = self.root.attrs["nbins"]
        return (n, n)

    def matrix(self, **kwargs):
        """``Cooler.matrix(...)`` (api.py:326-409) on the GPU: a selector with ``.fetch(region[, region2])``
        and ``[i0:i1, j0:j1]`` (see ``cooler_b200.api.matrix``).  The uploaded pixel table is kept for
        the next call."""
        from .api import matrix
        sel = matrix(self, table=getattr(self, "_b200_table", None), **kwargs)
        if sel.values is None:            # a table built for another field carries no counts: do not keep it
            self._b200_table = sel.table
        return sel


class H5Cooler(MemCooler):
    """A ``.cool`` file (or one resolution of an ``.mcool``) read with ``cooler_b200.h5lite``.

    Same read protocol as ``cooler.Cooler`` for the hot path (api.py:32-409): columns are
    decoded (gzip + shuffle) on access, so ``balance_cooler`` / ``CoolerCoarsener`` read the
    pixel table exactly once.  ``open('r+')`` (``balance_cooler(..., store=True)``, the
    ``cooler balance`` CLI) appends through ``cooler_b200.h5write``.
    """

    def __init__(self, path, group="/"):
        from . import h5lite
        self._file = h5lite.File(path)
        try:
            root = self._file[group] if group not in ("", "/") else self._file
        except KeyError:
            raise KeyError(f"No cooler found at: {path}::{group}") from None     # api.py:101-109
        if not isinstance(root, h5lite.Group) or any(k not in root for k in ("chroms", "bins", "pixels", "indexes")):
            raise KeyError(f"No cooler found at: {path}::{group}")
        super().__init__(root, filename=path)
        self.uri = path if group in ("", "/") else f"{path}::{group}"
        self.group = group

    @property
    def info(self):
        d = dict(self.root.attrs)
        for k in ("storage-mode",):                  # some old files keep it as a scalar dataset
            if k not in d and k in self.root:
                d[k] = self.root[k][()]
        if "nbins" not in d:
            d["nbins"] = len(self.root["bins/chrom"])
        if "nnz" not in d:
            d["nnz"] = len(self.root["pixels/bin1_id"])
        return d

    @property
    def binsize(self):
        v = self.root.attrs.get("bin-size")
        return None if v in (None, "null", "None") else int(v)

    @property
    def storage_mode(self):
        return self.info.get("storage-mode", "symmetric-upper")

    @property
    def shape(self):
        n = self.info["nbins"]
        return (n, n)

    @staticmethod
    def _ordered(group, first):
        keys = group.keys()
        return [k for k in first if k in keys] + [k for k in keys if k not in first]

    def bins(self, **kwargs):
        return _Selector(self.root["bins"], self._ordered(self.root["bins"], ["chrom", "start", "end"]))

    def pixels(self, join=False, **kwargs):
        if join:
            raise NotImplementedError("H5Cooler.pixels(join=True)")
        return _Selector(self.root["pixels"], self._ordered(self.root["pixels"], ["bin1_id", "bin2_id", "count"]))

    def _reload(self):
        from . import h5lite
        self._file = h5lite.File(self.filename)
        self.root = self._file[self.group] if self.group not in ("", "/") else self._file

    @contextmanager
    def open(self, mode="r", **kwargs):
        """``'r'``: the h5lite group.  ``'r+'`` / ``'a'``: the same group of a writable
        ``h5write.File`` (append-only writer: create_dataset / del / attrs.update); the reader
        is refreshed when the block ends, like ``Cooler.open`` re-opening the file (api.py:139-150)."""
        if mode == "r":
            yield self.root
            return
        from . import h5write
        f = h5write.File(self.filename, "r+")
        try:
            yield f[self.group] if self.group not in ("", "/") else f
        finally:
            f.close()
            self._reload()


def list_coolers(path):
    """Group paths of every cooler in a file (fileops.list_coolers, fileops.py:113-150)."""
    from . import h5lite
    f = h5lite.File(path)
    out = []

    def visit(g, name):
        is_clr = all(k in g for k in ("chroms", "bins", "pixels", "indexes"))
        if is_clr:
            out.append(name or "/")
        for k in g.keys():
            if is_clr and k in ("chroms", "bins", "pixels", "indexes"):
                continue
            o = g[k]
            if isinstance(o, h5lite.Group):
                visit(o, f"{name}/{k}")
    visit(f, "")
    return out
