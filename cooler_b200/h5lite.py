"""Minimal read-only HDF5 reader for ``.cool`` / ``.mcool`` files.

h5py / libhdf5 are not part of this image, and the hot path only needs to read a cooler's
columns ONCE into (pinned) host memory.  This module parses the subset of the HDF5 file
format that h5py writes for coolers (reference: src/cooler/create/_create.py:205-323 --
1-D chunked datasets with shuffle + gzip, enum / fixed-string / integer / float types,
scalar and vlen-string attributes):

* superblock versions 0-3; object headers v1 and v2 (with continuation blocks);
* old-style groups (symbol table: B-tree v1 + local heap) and new-style groups with compact
  link messages (dense link storage in fractal heaps is not supported);
* dataset layouts compact / contiguous / chunked with a B-tree v1 chunk index (layout
  message v3; v4 "single chunk" and "implicit" indexes as written by newer libhdf5 for
  small datasets), filters deflate, shuffle, fletcher32 (checksum skipped);
* datatypes fixed-point, floating-point, fixed-length string, enum, vlen string (global heap).

The object model mirrors the small part of h5py the callers use: ``File[...]`` -> ``Group`` /
``Dataset``; ``Dataset[...]`` (any numpy index; reads the whole column), ``.astype(dt)[:]``,
``.dtype``, ``.shape``, ``.attrs``; ``Group.keys()``, ``in``, ``.attrs``.

This is host-side I/O plumbing, not the compute path; nothing here touches the GPU.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

__all__ = ["File", "Group", "Dataset", "is_hdf5", "H5FormatError"]

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(ValueError):
    pass


def is_hdf5(path):
    try:
        with open(path, "rb") as f:
            return f.read(8) == _SIG
    except OSError:
        return False


class _Reader:
    """Random access to the file bytes (whole file in memory: .cool metadata is scattered)."""

    def __init__(self, path):
        self.path = path
        self.buf = np.memmap(path, dtype=np.uint8, mode="r")
        self.mv = memoryview(self.buf)
        self.O = 8      # size of offsets
        self.L = 8      # size of lengths
        self.base = 0

    def bytes(self, off, n):
        if off < 0 or off + n > len(self.mv):
            raise H5FormatError(f"read past end of file ({off}+{n})")
        return bytes(self.mv[off:off + n])

    def u(self, off, n):
        return int.from_bytes(self.mv[off:off + n], "little")

    def off_(self, off):
        v = self.u(off, self.O)
        return None if v == (1 << (8 * self.O)) - 1 else v + self.base

    def len_(self, off):
        return self.u(off, self.L)


# ---------------------------------------------------------------------------------- datatypes
class _Type:
    """Parsed datatype message."""

    def __init__(self, np_dtype=None, kind="plain", size=0, base=None, enum=None, vlen_str=False,
                 consumed=0):
        self.np_dtype, self.kind, self.size, self.base = np_dtype, kind, size, base
        self.enum, self.vlen_str, self.consumed = enum, vlen_str, consumed


def _parse_type(b, p=0):
    cv = b[p]
    cls, ver = cv & 0x0F, cv >> 4
    bits = b[p + 1] | (b[p + 2] << 8) | (b[p + 3] << 16)
    size = struct.unpack_from("<I", b, p + 4)[0]
    q = p + 8
    if cls == 0:                                        # fixed-point
        order = ">" if bits & 1 else "<"
        signed = bool(bits & 8)
        q += 4
        if size not in (1, 2, 4, 8):
            raise H5FormatError(f"unsupported integer size {size}")
        return _Type(np.dtype(f"{order}{'i' if signed else 'u'}{size}"), size=size, consumed=q - p)
    if cls == 1:                                        # floating point
        order = ">" if bits & 1 else "<"
        q += 12
        if size not in (2, 4, 8):
            raise H5FormatError(f"unsupported float size {size}")
        return _Type(np.dtype(f"{order}f{size}"), size=size, consumed=q - p)
    if cls == 3:                                        # fixed-length string
        return _Type(np.dtype(f"S{size}"), kind="string", size=size, consumed=q - p)
    if cls == 8:                                        # enum
        n = bits & 0xFFFF
        base = _parse_type(b, q)
        q += base.consumed
        names = []
        for _ in range(n):
            e = b.index(b"\x00", q)
            names.append(b[q:e].decode("utf-8", "replace"))
            ln = e - q + 1
            if ver < 3:
                ln = (ln + 7) & ~7
            q += ln
        vals = np.frombuffer(b, dtype=base.np_dtype, count=n, offset=q)
        q += n * base.size
        return _Type(base.np_dtype, kind="enum", size=size, base=base,
                     enum=dict(zip(names, (int(v) for v in vals))), consumed=q - p)
    if cls == 9:                                        # variable length
        is_str = (bits & 0x0F) == 1
        base = _parse_type(b, q)
        q += base.consumed
        return _Type(None, kind="vlen", size=size, base=base, vlen_str=is_str, consumed=q - p)
    if cls == 6:                                        # compound: only skipped over / not readable
        raise H5FormatError("compound datatypes are not supported")
    raise H5FormatError(f"unsupported datatype class {cls}")


def _parse_space(b, L):
    ver, rank, flags = b[0], b[1], b[2]
    if ver == 1:
        p = 8
    elif ver == 2:
        p = 4
        if b[3] == 2:                                   # null dataspace
            return None
    else:
        raise H5FormatError(f"unsupported dataspace version {ver}")
    return tuple(int.from_bytes(b[p + i * L:p + (i + 1) * L], "little") for i in range(rank))


# ---------------------------------------------------------------------------------- objects
class _Obj:
    """Parsed object header: list of (type, flags, bytes)."""

    def __init__(self, rd: _Reader, addr):
        self.rd, self.addr = rd, addr
        self.msgs = []
        self.moffs = []                                 # file offset of every message body
        self.version = 2 if rd.bytes(addr, 4) == b"OHDR" else 1
        if rd.bytes(addr, 4) == b"OHDR":
            self._parse_v2(addr)
        else:
            self._parse_v1(addr)

    def _parse_v1(self, addr):
        rd = self.rd
        if rd.u(addr, 1) != 1:
            raise H5FormatError(f"bad object header at {addr}")
        nmsg = rd.u(addr + 2, 2)
        size = rd.u(addr + 8, 4)
        blocks = [(addr + 16, size)]
        while blocks and len(self.msgs) < nmsg:
            p, n = blocks.pop(0)
            end = p + n
            while p + 8 <= end and len(self.msgs) < nmsg:
                mtype, msize, mflags = rd.u(p, 2), rd.u(p + 2, 2), rd.u(p + 4, 1)
                body = rd.bytes(p + 8, msize)
                self.moffs.append(p + 8)
                p += 8 + msize
                if mtype == 0x10:
                    blocks.append((int.from_bytes(body[:rd.O], "little") + rd.base,
                                   int.from_bytes(body[rd.O:rd.O + rd.L], "little")))
                self.msgs.append((mtype, mflags, body))

    def _parse_v2(self, addr):
        rd = self.rd
        flags = rd.u(addr + 5, 1)
        p = addr + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        szb = 1 << (flags & 3)
        size0 = rd.u(p, szb)
        p += szb
        track = bool(flags & 0x04)
        blocks = [(p, size0)]                           # (start of messages, bytes of messages incl. gap)
        while blocks:
            p, n = blocks.pop(0)
            end = p + n
            while p + 4 <= end:
                mtype, msize, mflags = rd.u(p, 1), rd.u(p + 1, 2), rd.u(p + 3, 1)
                p += 4 + (2 if track else 0)
                if p + msize > end:
                    break
                body = rd.bytes(p, msize)
                self.moffs.append(p)
                p += msize
                if mtype == 0x10:
                    caddr = int.from_bytes(body[:rd.O], "little") + rd.base
                    clen = int.from_bytes(body[rd.O:rd.O + rd.L], "little")
                    if rd.bytes(caddr, 4) != b"OCHK":
                        raise H5FormatError("bad object header continuation")
                    blocks.append((caddr + 4, clen - 8))    # minus signature and checksum
                self.msgs.append((mtype, mflags, body))

    def find(self, mtype):
        return [m for m in self.msgs if m[0] == mtype]


class _Attrs(dict):
    incomplete = False      # True when the object also has densely stored attributes (not read)


def _read_vlen_strings(rd: _Reader, raw, n):
    out = []
    step = 4 + rd.O + 4
    for i in range(n):
        p = i * step
        ln = int.from_bytes(raw[p:p + 4], "little")
        gaddr = int.from_bytes(raw[p + 4:p + 4 + rd.O], "little")
        idx = int.from_bytes(raw[p + 4 + rd.O:p + 8 + rd.O], "little")
        if ln == 0 or gaddr in (0, (1 << (8 * rd.O)) - 1):
            out.append("")
            continue
        out.append(_global_heap_object(rd, gaddr + rd.base, idx)[:ln].decode("utf-8", "replace"))
    return out


def _global_heap_object(rd: _Reader, addr, index):
    if rd.bytes(addr, 4) != b"GCOL":
        raise H5FormatError("bad global heap collection")
    size = rd.len_(addr + 8)
    p = addr + 8 + rd.L
    end = addr + size
    while p + 8 + rd.L <= end:
        idx = rd.u(p, 2)
        osize = rd.len_(p + 8)
        if idx == 0:
            break
        if idx == index:
            return rd.bytes(p + 8 + rd.L, osize)
        p += 8 + rd.L + ((osize + 7) & ~7)
    raise H5FormatError("global heap object not found")


def _decode_values(rd, typ: _Type, shape, raw):
    n = int(np.prod(shape)) if shape else 1
    if typ.kind == "vlen":
        if not typ.vlen_str:
            raise H5FormatError("variable-length sequences are not supported")
        vals = _read_vlen_strings(rd, raw, n)
        return vals[0] if shape == () else np.array(vals, dtype=object).reshape(shape)
    arr = np.frombuffer(raw, dtype=typ.np_dtype, count=n)
    if typ.kind == "enum" and set(typ.enum or ()) == {"FALSE", "TRUE"}:      # h5py's encoding of numpy bool
        arr = arr != typ.enum["FALSE"]
        return bool(arr[0]) if shape == () else arr.reshape(shape).copy()
    if typ.kind == "string":
        if shape == ():
            return arr[0].decode("utf-8", "replace")
        return arr.reshape(shape).copy()
    if shape == ():
        return arr[0].item() if arr.dtype.kind in "iuf" else arr[0]
    return arr.reshape(shape).astype(arr.dtype.newbyteorder("="), copy=True)


def _parse_attribute(rd, body):
    ver = body[0]
    nsz, tsz, ssz = struct.unpack_from("<HHH", body, 2)
    if ver == 1:
        p = 8
        pad = lambda k: (k + 7) & ~7                     # noqa: E731
    elif ver == 2:
        p = 8
        pad = lambda k: k                                # noqa: E731
    elif ver == 3:
        p = 9
        pad = lambda k: k                                # noqa: E731
    else:
        raise H5FormatError(f"unsupported attribute version {ver}")
    name = body[p:p + nsz].split(b"\x00")[0].decode("utf-8", "replace")
    p += pad(nsz)
    typ = _parse_type(body, p)
    p += pad(tsz)
    shape = _parse_space(body[p:p + ssz], rd.L)
    p += pad(ssz)
    if shape is None:
        return name, None
    return name, _decode_values(rd, typ, shape, body[p:])


def _attrs_of(obj: _Obj):
    out = _Attrs()
    for _, _, body in obj.find(0x0C):
        try:
            k, v = _parse_attribute(obj.rd, body)
        except H5FormatError:
            continue
        out[k] = v
    for _, _, b in obj.find(0x15):                      # attribute info: dense storage (fractal heap)?
        p = 2 + (2 if b[1] & 1 else 0)
        if int.from_bytes(b[p:p + obj.rd.O], "little") != (1 << (8 * obj.rd.O)) - 1:
            out.incomplete = True                       # dense storage (fractal heap): not read
    return out


def _narrowable(storage, want):
    """A little-endian integer column may be read as a narrower integer of the same signedness class
    (values are checked to fit: the dropped byte planes must be all zero)."""
    storage, want = np.dtype(storage), np.dtype(want)
    if storage == want:
        return True
    return (storage.kind in "iu" and want.kind in "iu" and storage.byteorder in "<=|"
            and want.itemsize < storage.itemsize)


class _ChunkTask:
    """Decode one stored chunk into its slice of a destination array.  Runs on a worker thread: inflate
    releases the GIL, and so do the numpy passes that assemble the byte planes of a shuffled chunk."""

    __slots__ = ("ds", "ch", "dt", "c", "i0", "i1", "dest")

    def __init__(self, ds, ch, dt, c, i0, i1, dest):
        self.ds, self.ch, self.dt, self.c, self.i0, self.i1, self.dest = ds, ch, dt, c, i0, i1, dest

    def __call__(self):
        ds, dt, dest = self.ds, self.dt, self.dest
        _offs, addr, csize, mask = self.ch
        rd = ds._rd
        if addr < 0 or addr + csize > len(rd.mv):
            raise H5FormatError(f"read past end of file ({addr}+{csize})")
        data = rd.mv[addr:addr + csize]
        es = dt.itemsize
        planes = None
        for i in range(len(ds._filters) - 1, -1, -1):
            if mask & (1 << i):
                continue
            fid, cd = ds._filters[i]
            if fid == 1:
                # the inflated size is known (HDF5 stores whole chunks): one allocation instead of the
                # doubling re-allocations of the default 16 KB buffer, which also scale badly over threads
                data = zlib.decompress(data, bufsize=self.c * es + 4)
            elif fid == 2:
                if i != 0:
                    data = ds._unfilter_shuffle(data, cd, es)
                else:
                    planes = cd[0] if cd else es            # outermost on write = last to undo: keep the planes
            elif fid == 3:
                data = data[:len(data) - 4]
            else:
                raise H5FormatError(f"unsupported HDF5 filter id {fid}")
        i0, i1 = self.i0, self.i1
        if planes is not None and planes == es and es > 1 and dt.kind in "iuf" and dt.byteorder in "<=|":
            n = len(data) // es
            p = np.frombuffer(data, dtype=np.uint8, count=n * es).reshape(es, n)
            ws = dest.dtype.itemsize
            if ws < es and p[ws:, i0:i1].any():
                raise OverflowError(f"values of {ds.name} do not fit {dest.dtype}")
            u = np.dtype(f"<u{ws}")
            dv = dest.view(u)
            np.copyto(dv, p[0, i0:i1], casting="unsafe")
            for k in range(1, ws):
                t = p[k, i0:i1].astype(u)
                t <<= 8 * k
                dv |= t
            return
        if planes is not None:
            data = ds._unfilter_shuffle(data, (planes,), es)
        arr = np.frombuffer(data, dtype=dt, count=min(self.c, len(data) // es))[i0:i1]
        if dest.dtype != arr.dtype and dest.dtype.itemsize < arr.dtype.itemsize and len(arr):
            lo_, hi_ = arr.min(), arr.max()
            info = np.iinfo(dest.dtype)
            if lo_ < info.min or hi_ > info.max:
                raise OverflowError(f"values of {ds.name} do not fit {dest.dtype}")
        dest[:] = arr


_POOL = None


def decode_pool():
    """The process-wide pool of decode threads (one per core, at most 32)."""
    global _POOL
    if _POOL is None:
        import os
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1), thread_name_prefix="h5lite")
    return _POOL


_CHUNK_DT = np.dtype([("src", "<u8"), ("src_bytes", "<i8"), ("dst_row", "<i8"), ("row0", "<i4"), ("row1", "<i4"),
                      ("filter_mask", "<u4")], align=True)          # cb200_chunk (include/cooler_b200.h)


def _native_decode(ds, lo, hi, dest):
    """Rows [lo, hi) of ``ds`` into ``dest`` through ``cb200_decode_chunks`` (inflate + un-shuffle (+ narrowing)
    of all overlapping chunks by native host threads inside one call: no interpreter lock is held while the
    file is decoded).  Returns False when the dataset's filter pipeline or types are outside what the native
    decoder covers -- the caller then uses the Python chunk tasks."""
    try:
        from . import _lib
        lib = _lib.load()
    except Exception:
        return False
    dt = ds._np_storage_dtype()
    es, ws = dt.itemsize, dest.dtype.itemsize
    ids = [fid for fid, _ in ds._filters]
    order = {2: 0, 1: 1, 3: 2}                               # only [shuffle][deflate][fletcher32], in that order
    if any(f not in order for f in ids) or [order[f] for f in ids] != sorted({order[f] for f in ids}):
        return False
    for fid, cd in ds._filters:
        if fid == 2 and cd and cd[0] != es:
            return False
    if dt.kind not in "iuf" or (es > 1 and dt.byteorder not in "<=|") or not dest.flags.c_contiguous:
        return False
    if ws != es and not (_narrowable(dt, dest.dtype) and dest.dtype.kind in "iu"):
        return False
    if ws == es and dest.dtype.kind != dt.kind:
        return False
    flags = (1 if 2 in ids else 0) | (2 if 1 in ids else 0) | (4 if 3 in ids else 0)
    # HDF5 numbers mask bits by pipeline position; the native side re-derives positions from `flags`
    a, addr, csize, mask, c = ds._chunk_arrays()
    n_rows = ds.shape[0]
    sel = np.flatnonzero((addr >= 0) & (a < hi) & (a + c > lo))
    s0 = np.maximum(a[sel], lo)
    s1 = np.minimum(np.minimum(a[sel] + c, hi), n_rows)
    ok = s1 > s0
    sel, s0, s1 = sel[ok], s0[ok], s1[ok]
    if int((s1 - s0).sum()) < hi - lo:
        dest[:] = 0                                          # chunks that were never written read as 0
    if len(sel) == 0:
        return True
    rd = ds._rd
    if int((addr[sel] + csize[sel]).max()) > len(rd.mv):
        raise H5FormatError("read past end of file")
    arr = np.zeros(len(sel), dtype=_CHUNK_DT)
    arr["src"] = (rd.buf.ctypes.data + addr[sel]).astype(np.uint64)
    arr["src_bytes"], arr["dst_row"] = csize[sel], s0 - lo
    arr["row0"], arr["row1"], arr["filter_mask"] = s0 - a[sel], s1 - a[sel], mask[sel]
    rc = lib.cb200_decode_chunks(arr.ctypes.data, len(sel), int(c), es, ws, flags, dest.ctypes.data, 0)
    if rc == -3:
        raise OverflowError(f"values of {ds.name} do not fit {dest.dtype}")
    if rc != 0:
        raise H5FormatError(f"{ds.name}: " + lib.cb200_last_error().decode("utf-8", "replace"))
    return True


def read_rows_into(jobs):
    """``jobs``: [(dataset, lo, hi, dest)] -- rows [lo, hi) of 1-D chunked datasets into numpy arrays
    (``dest`` may be pinned staging memory), straight into the destinations: no zero-filled staging array,
    no whole-column dtype cast.  Columns written the way h5py / cooler write them (shuffle + gzip) are
    decoded by native host threads (``cb200_decode_chunks``); anything else by a pool of Python threads
    (zlib and the plane assembly release the GIL)."""
    tasks = []
    for ds, lo, hi, dest in jobs:
        if hi <= lo or _native_decode(ds, lo, hi, dest):
            continue
        t, covered = ds.chunk_tasks(lo, hi, dest)
        if covered < hi - lo:
            dest[:] = 0                                   # chunks that were never written read as the fill value 0
        tasks += t
    if len(tasks) > 4:
        for f in [decode_pool().submit(t) for t in tasks]:
            f.result()
    else:
        for t in tasks:
            t()


class Dataset:
    def __init__(self, rd: _Reader, obj: _Obj, name):
        self._rd, self._obj, self.name = rd, obj, name
        self._type = _parse_type(obj.find(0x03)[0][2])
        self.shape = _parse_space(obj.find(0x01)[0][2], rd.L) or ()
        self._layout = obj.find(0x08)[0][2]
        self._filters = self._parse_filters()
        self._attrs = None
        self._as = None
        self._cidx = None
        self._cidx_np = None

    # -- h5py-like surface ---------------------------------------------------------------
    @property
    def dtype(self):
        if self._type.kind == "vlen":
            return np.dtype(object)
        return self._type.np_dtype.newbyteorder("=") if self._type.np_dtype.kind in "iuf" else self._type.np_dtype

    @property
    def enum(self):
        """name -> value for enum-typed datasets (bins/chrom), else None."""
        return self._type.enum

    @property
    def attrs(self):
        if self._attrs is None:
            self._attrs = _attrs_of(self._obj)
        return self._attrs

    @property
    def id(self):          # h5py datasets have .id; read_cooler_arrays uses it to detect them
        return self._obj.addr

    def __len__(self):
        return self.shape[0] if self.shape else 0

    def astype(self, dtype):
        view = Dataset.__new__(Dataset)
        view.__dict__.update(self.__dict__)
        view._as = np.dtype(dtype)
        return view

    def __getitem__(self, key):
        out = None
        if isinstance(key, slice) and len(self.shape) == 1 and self._layout[1] == 2:
            lo, hi, step = key.indices(self.shape[0])
            if step == 1:
                out = self._read_chunked_1d(lo, max(hi, lo))
                key = slice(None)
        if out is None:
            out = self.read()
        if self._as is not None:
            out = out.astype(self._as, copy=False)
        if key is Ellipsis or (isinstance(key, slice) and key == slice(None)):
            return out
        if isinstance(key, tuple) and key == ():
            return out
        return out[key]

    def __array__(self, dtype=None, copy=None):
        a = self.read()
        return a if dtype is None else a.astype(dtype)

    # -- decoding ------------------------------------------------------------------------
    def _parse_filters(self):
        msgs = self._obj.find(0x0B)
        if not msgs:
            return []
        b = msgs[0][2]
        ver, nf = b[0], b[1]
        p = 8 if ver == 1 else 2
        out = []
        for _ in range(nf):
            fid = struct.unpack_from("<H", b, p)[0]
            if ver == 1 or fid >= 256:
                nlen = struct.unpack_from("<H", b, p + 2)[0]
                p += 4
            else:
                nlen = 0
                p += 2
            _flags, ncd = struct.unpack_from("<HH", b, p)
            p += 4
            p += ((nlen + 7) & ~7) if ver == 1 else nlen
            cd = struct.unpack_from(f"<{ncd}I", b, p)
            p += 4 * ncd
            if ver == 1 and ncd % 2:
                p += 4
            out.append((fid, cd))
        return out

    @staticmethod
    def _unfilter_shuffle(data, cd, elsize):
        es = cd[0] if cd else elsize
        n = len(data) // es
        if es > 1 and n > 0:
            a = np.frombuffer(data, dtype=np.uint8, count=n * es).reshape(es, n).T
            data = np.ascontiguousarray(a).tobytes() + bytes(data[n * es:])
        return data

    def _unfilter(self, data, mask, elsize):
        for i in range(len(self._filters) - 1, -1, -1):
            if mask & (1 << i):
                continue
            fid, cd = self._filters[i]
            if fid == 1:
                data = zlib.decompress(data)
            elif fid == 2:
                es = cd[0] if cd else elsize
                n = len(data) // es
                if es > 1 and n > 0:
                    a = np.frombuffer(data, dtype=np.uint8, count=n * es).reshape(es, n).T
                    data = np.ascontiguousarray(a).tobytes() + data[n * es:]
            elif fid == 3:
                data = data[:-4]
            else:
                raise H5FormatError(f"unsupported HDF5 filter id {fid}")
        return data

    def _np_storage_dtype(self):
        if self._type.kind == "vlen":
            raise H5FormatError("variable-length datasets are not supported")
        return self._type.np_dtype

    def _chunk_index(self):
        """[(offset tuple, address, stored size, filter mask)] of a chunked dataset (parsed once)."""
        if getattr(self, "_cidx", None) is None:
            self._cidx = self._parse_chunk_index()
        return self._cidx

    def _parse_chunk_index(self):
        rd, b = self._rd, self._layout
        ver = b[0]
        if ver == 3:
            ndim = b[2]
            baddr = int.from_bytes(b[3:3 + rd.O], "little")
            dims = struct.unpack_from(f"<{ndim}I", b, 3 + rd.O)
            chunks = []
            if baddr != (1 << (8 * rd.O)) - 1:
                self._walk_chunk_btree(baddr + rd.base, ndim, chunks)
            return dims[:-1], chunks
        if ver == 4:
            flags, ndim, enc = b[2], b[3], b[4]
            dims = tuple(int.from_bytes(b[5 + i * enc:5 + (i + 1) * enc], "little") for i in range(ndim))
            p = 5 + ndim * enc
            itype = b[p]
            p += 1
            cdims = dims[:-1]
            if itype == 1:                              # single chunk
                if flags & 2:
                    size = int.from_bytes(b[p:p + rd.L], "little")
                    mask = struct.unpack_from("<I", b, p + rd.L)[0]
                    p += rd.L + 4
                else:
                    size, mask = int(np.prod(cdims)) * self._type.size, 0
                addr = int.from_bytes(b[p:p + rd.O], "little")
                return cdims, ([((0,) * len(cdims), addr + rd.base, size, mask)]
                               if addr != (1 << (8 * rd.O)) - 1 else [])
            if itype == 2:                              # implicit: contiguous array of unfiltered chunks
                addr = int.from_bytes(b[p:p + rd.O], "little")
                if addr == (1 << (8 * rd.O)) - 1:
                    return cdims, []
                csize = int(np.prod(cdims)) * self._type.size
                grid = [-(-s // c) for s, c in zip(self.shape, cdims)]
                out, k = [], 0
                for idx in np.ndindex(*grid):
                    out.append((tuple(i * c for i, c in zip(idx, cdims)), addr + rd.base + k * csize, csize, 0))
                    k += 1
                return cdims, out
            raise H5FormatError(f"chunk index type {itype} (HDF5 1.10+ latest-format files) is not supported; "
                                "re-write the file with libver='earliest' (the h5py default)")
        raise H5FormatError(f"unsupported data layout version {ver}")

    def _walk_chunk_btree(self, addr, ndim, out):
        rd = self._rd
        if rd.bytes(addr, 4) != b"TREE":
            raise H5FormatError("bad chunk B-tree node")
        level, nent = rd.u(addr + 5, 1), rd.u(addr + 6, 2)
        p = addr + 8 + 2 * rd.O
        ksz = 8 + 8 * ndim
        for _ in range(nent):
            csize, mask = rd.u(p, 4), rd.u(p + 4, 4)
            offs = tuple(rd.u(p + 8 + 8 * i, 8) for i in range(ndim - 1))
            child = rd.off_(p + ksz)
            p += ksz + rd.O
            if level == 0:
                out.append((offs, child, csize, mask))
            else:
                self._walk_chunk_btree(child, ndim, out)

    def _read_chunked_1d(self, lo, hi):
        """Rows [lo, hi) of a 1-D chunked dataset (only the chunks that overlap are decoded)."""
        dt = self._np_storage_dtype()
        want = dt
        if self._as is not None and np.dtype(self._as) != dt and _narrowable(dt, self._as):
            want = np.dtype(self._as)                    # assembled from the low byte planes, no wide staging copy
        out = np.empty(hi - lo, dtype=want)
        read_rows_into([(self, lo, hi, out)])
        return self._finish(out)

    def _chunk_arrays(self):
        """The chunk index as numpy columns: (first row, address or -1, stored bytes, filter mask, rows per chunk)."""
        if getattr(self, "_cidx_np", None) is None:
            cdims, chunks = self._chunk_index()
            self._cidx_np = (np.array([ch[0][0] for ch in chunks], dtype=np.int64),
                             np.array([-1 if ch[1] is None else ch[1] for ch in chunks], dtype=np.int64),
                             np.array([ch[2] for ch in chunks], dtype=np.int64),
                             np.array([ch[3] for ch in chunks], dtype=np.uint32), int(cdims[0]))
        return self._cidx_np

    def is_chunked_1d(self):
        return len(self.shape) == 1 and self._layout[1] == 2

    def chunk_rows(self):
        return int(self._chunk_index()[0][0])

    def chunk_tasks(self, lo, hi, dest):
        """Decode tasks (callables) that fill ``dest[0:hi-lo]`` with rows [lo, hi) of this 1-D chunked
        dataset, one per overlapping chunk, plus the number of rows they cover (rows of chunks that were
        never written are not covered: the caller zero-fills).  ``dest`` is a numpy array of the
        dataset's dtype, or of a narrower integer type of the same kind (see ``_narrowable``: only the
        low byte planes of a shuffled chunk are then assembled)."""
        dt = self._np_storage_dtype()
        cdims, chunks = self._chunk_index()
        c = cdims[0]
        n_rows = self.shape[0]
        tasks, covered = [], 0
        for ch in chunks:
            a = ch[0][0]
            if ch[1] is None or a >= hi or a + c <= lo:
                continue
            s0, s1 = max(a, lo), min(a + c, hi, n_rows)
            if s1 <= s0:
                continue
            covered += s1 - s0
            tasks.append(_ChunkTask(self, ch, dt, c, s0 - a, s1 - a, dest[s0 - lo:s1 - lo]))
        return tasks, covered

    def _finish(self, arr):
        if arr.dtype.kind in "iuf" and arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("="))
        return arr

    def read(self):
        """The whole dataset as a numpy array."""
        rd, b = self._rd, self._layout
        ver, cls = b[0], b[1]
        shape = self.shape
        n = int(np.prod(shape)) if shape else 1
        if self._type.kind == "vlen":
            if cls == 1:
                addr = int.from_bytes(b[2:2 + rd.O], "little")
                raw = rd.bytes(addr + rd.base, n * (8 + rd.O))
            elif cls == 0:
                sz = struct.unpack_from("<H", b, 2)[0]
                raw = b[4:4 + sz]
            else:
                raise H5FormatError("chunked variable-length datasets are not supported")
            return _decode_values(rd, self._type, shape, raw)
        dt = self._np_storage_dtype()
        if ver not in (3, 4):
            raise H5FormatError(f"unsupported data layout version {ver}")
        if cls == 0:                                    # compact
            sz = struct.unpack_from("<H", b, 2)[0]
            arr = np.frombuffer(b[4:4 + sz], dtype=dt, count=n).reshape(shape).copy()
            return self._finish(arr)
        if cls == 1:                                    # contiguous
            addr = int.from_bytes(b[2:2 + rd.O], "little")
            if addr == (1 << (8 * rd.O)) - 1 or n == 0:
                return self._finish(np.zeros(shape, dtype=dt))
            arr = np.frombuffer(rd.mv, dtype=dt, count=n, offset=addr + rd.base).reshape(shape).copy()
            return self._finish(arr)
        if cls == 2:
            if len(shape) == 1:
                return self._read_chunked_1d(0, shape[0])
            cdims, chunks = self._chunk_index()
            out = np.zeros(shape, dtype=dt)
            for offs, addr, csize, mask in chunks:
                raw = self._unfilter(rd.bytes(addr, csize), mask, dt.itemsize)
                blk = np.frombuffer(raw, dtype=dt, count=int(np.prod(cdims))).reshape(cdims)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
                out[sl] = blk[tuple(slice(0, s.stop - s.start) for s in sl)]
            return self._finish(out)
        raise H5FormatError(f"unsupported data layout class {cls}")


class Group:
    def __init__(self, rd: _Reader, obj: _Obj, name="/"):
        self._rd, self._obj, self.name = rd, obj, name
        self._links = None
        self._attrs = None

    @property
    def attrs(self):
        if self._attrs is None:
            self._attrs = _attrs_of(self._obj)
        return self._attrs

    def _load_links(self):
        if self._links is not None:
            return self._links
        rd, links = self._rd, {}
        for _, _, body in self._obj.find(0x11):         # symbol table message
            btree = int.from_bytes(body[:rd.O], "little") + rd.base
            heap = int.from_bytes(body[rd.O:2 * rd.O], "little") + rd.base
            if rd.bytes(heap, 4) != b"HEAP":
                raise H5FormatError("bad local heap")
            hdata = rd.off_(heap + 8 + 2 * rd.L)
            self._walk_group_btree(btree, hdata, links)
        for _, _, body in self._obj.find(0x06):         # link message (new-style compact groups)
            name, addr = self._parse_link(body)
            if addr is not None:
                links[name] = addr
        for _, _, body in self._obj.find(0x02):         # link info: dense storage?
            flags = body[1]
            p = 2 + (8 if flags & 1 else 0)
            fheap = int.from_bytes(body[p:p + rd.O], "little")
            if fheap != (1 << (8 * rd.O)) - 1 and not links:
                raise H5FormatError("groups with dense link storage (fractal heap) are not supported")
        self._links = links
        return links

    def _parse_link(self, body):
        rd = self._rd
        flags = body[1]
        p = 2
        ltype = 0
        if flags & 0x08:
            ltype = body[p]
            p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        lsz = 1 << (flags & 3)
        nlen = int.from_bytes(body[p:p + lsz], "little")
        p += lsz
        name = body[p:p + nlen].decode("utf-8", "replace")
        p += nlen
        if ltype != 0:
            return name, None                           # soft / external links are ignored
        return name, int.from_bytes(body[p:p + rd.O], "little") + rd.base

    def _walk_group_btree(self, addr, hdata, links):
        rd = self._rd
        sig = rd.bytes(addr, 4)
        if sig == b"SNOD":
            nsym = rd.u(addr + 6, 2)
            p = addr + 8
            esz = 2 * rd.O + 24
            for _ in range(nsym):
                noff = rd.u(p, rd.O)
                oaddr = rd.off_(p + rd.O)
                end = hdata + noff
                while rd.mv[end] != 0:
                    end += 1
                links[rd.bytes(hdata + noff, end - hdata - noff).decode("utf-8", "replace")] = oaddr
                p += esz
            return
        if sig != b"TREE":
            raise H5FormatError("bad group B-tree node")
        nent = rd.u(addr + 6, 2)
        p = addr + 8 + 2 * rd.O + rd.L                  # skip key 0
        for _ in range(nent):
            child = rd.off_(p)
            p += rd.O + rd.L
            self._walk_group_btree(child, hdata, links)

    # -- h5py-like surface ---------------------------------------------------------------
    def keys(self):
        return list(self._load_links().keys())

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self._load_links())

    def __contains__(self, path):
        try:
            self[path]
            return True
        except KeyError:
            return False

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def __getitem__(self, path):
        cur = self
        parts = [p for p in str(path).split("/") if p]
        if str(path).startswith("/") and self.name != "/":
            cur = self._root
        for i, part in enumerate(parts):
            if not isinstance(cur, Group):
                raise KeyError(path)
            links = cur._load_links()
            if part not in links:
                raise KeyError(f"Unable to open object (component not found: '{part}')")
            obj = _Obj(self._rd, links[part])
            name = (cur.name.rstrip("/") + "/" + part)
            if obj.find(0x08) and obj.find(0x03):
                nxt = Dataset(self._rd, obj, name)
            else:
                nxt = Group(self._rd, obj, name)
                nxt._root = getattr(self, "_root", self)
            cur = nxt
        return cur


class File(Group):
    """Read-only HDF5 file.  ``File(path)[...]`` like ``h5py.File(path, 'r')``."""

    def __init__(self, path, mode="r"):
        if mode not in ("r",):
            raise H5FormatError("cooler_b200.h5lite is read-only (h5py / libhdf5 are not installed)")
        rd = _Reader(path)
        self.filename = path
        sb = 0
        n = len(rd.mv)
        while sb < n and rd.bytes(sb, 8) != _SIG:       # the superblock may sit at 0, 512, 1024, ...
            sb = 512 if sb == 0 else sb * 2
        if sb >= n:
            raise H5FormatError(f"{path} is not an HDF5 file")
        ver = rd.u(sb + 8, 1)
        if ver in (0, 1):
            rd.O, rd.L = rd.u(sb + 13, 1), rd.u(sb + 14, 1)
            p = sb + 24 + (4 if ver == 1 else 0)
            rd.base = rd.u(p, rd.O)
            self._sb = {"off": sb, "ver": ver, "leaf_k": rd.u(sb + 16, 2), "internal_k": rd.u(sb + 18, 2),
                        "chunk_k": rd.u(sb + 24, 2) if ver == 1 else 32, "eof_off": p + 2 * rd.O,
                        "eof": rd.u(p + 2 * rd.O, rd.O), "root_entry_off": p + 4 * rd.O}
            p += 4 * rd.O                               # base, free-space, eof, driver info
            root_addr = rd.off_(p + rd.O)               # symbol table entry: name offset, header address
        elif ver in (2, 3):
            rd.O, rd.L = rd.u(sb + 9, 1), rd.u(sb + 10, 1)
            p = sb + 12
            rd.base = rd.u(p, rd.O)
            self._sb = {"off": sb, "ver": ver}
            root_addr = rd.off_(p + 3 * rd.O)
        else:
            raise H5FormatError(f"unsupported HDF5 superblock version {ver}")
        if sb and rd.base == 0:
            rd.base = sb
        super().__init__(rd, _Obj(rd, root_addr), "/")
        self._root = self

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False
