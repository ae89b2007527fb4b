"""Minimal append-only HDF5 writer for ``.cool`` / ``.mcool`` files.

h5py / libhdf5 are not part of this image, so the step on the far side of the hot path --
storing ``bins/weight`` and writing coarsened pixel tables (reference:
src/cooler/create/_create.py:44-323, src/cooler/_balance.py:469-475) -- is done here, in the
same subset of the file format that h5py (``libver='earliest'``, its default) produces for
coolers and that ``cooler_b200.h5lite`` reads:

* superblock version 0; object headers version 1; old-style groups (symbol table message ->
  B-tree v1 + local heap + symbol table nodes);
* 1-D datasets, chunked with a B-tree v1 chunk index, shuffle + deflate (layout message v3,
  filter pipeline v1, fill value v2); empty datasets are contiguous with no storage;
* datatypes: fixed-point, IEEE float, fixed-length string, enum, variable-length UTF-8 string
  (attributes only; values in a global heap collection);
* attributes (message v1): scalars and 1-D arrays of bool (h5py's FALSE/TRUE int8 enum), int,
  float, str.

Every structure mirrors, field for field, what libhdf5 wrote into the reference's own test
files (``tests/golden/cool/*.cool``; see tests/test_h5write.py for the structural checks).

**Append-only.**  Nothing that exists in a file is moved or overwritten except (a) the 16-byte
body of a group's symbol-table message, (b) the superblock's end-of-file address and root
entry.  New datasets, object headers (an attribute update re-writes the object's header at the
end of the file and re-points its link) and group indexes are appended, so a file stays valid
until the final patch at ``close()``.  Space of replaced structures is not reclaimed
(``h5repack`` does that).  Files whose groups use the newer link-message / fractal-heap form or
version-2 object headers (``libver='latest'``) are opened read-only by h5lite and rejected
here with a clear error.

This is host-side I/O plumbing, not the compute path; nothing here touches the GPU.
"""
from __future__ import annotations

import os
import struct
import zlib

import numpy as np

from . import h5lite
from .h5lite import H5FormatError

__all__ = ["File", "Group", "Dataset", "StreamDataset"]

_UNDEF = 0xFFFFFFFFFFFFFFFF
_LEAF_K, _INTERNAL_K, _CHUNK_K = 16, 16, 32      # new files; existing files keep their own
_DEFAULT_CHUNK = 65536                           # elements per chunk of a streamed column


def _pad8(b):
    return b + b"\x00" * (-len(b) % 8)


# ---------------------------------------------------------------------------------- datatypes
def _enc_int(size, signed):
    return struct.pack("<BBBBIHH", 0x10, 0x08 if signed else 0x00, 0, 0, size, 0, 8 * size)


def _enc_float(size):
    if size == 8:
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 63, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
    if size == 4:
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 31, 0, 4, 0, 32, 23, 8, 0, 23, 127)
    if size == 2:
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 15, 0, 2, 0, 16, 10, 5, 0, 10, 15)
    raise H5FormatError(f"unsupported float size {size}")


def _enc_enum(base_dtype, mapping):
    base_dtype = np.dtype(base_dtype)
    names = list(mapping.keys())
    n = len(names)
    if n > 0xFFFF:
        raise H5FormatError("too many enum members")
    out = struct.pack("<BBBBI", 0x18, n & 0xFF, n >> 8, 0, base_dtype.itemsize)
    out += _enc_int(base_dtype.itemsize, base_dtype.kind == "i")
    for nm in names:
        out += _pad8(str(nm).encode("utf-8") + b"\x00")
    out += np.asarray([mapping[k] for k in names], dtype=base_dtype.newbyteorder("<")).tobytes()
    return out


_VLEN_STR = struct.pack("<BBBBI", 0x19, 0x01, 0x01, 0x00, 16) + struct.pack("<BBBBIHH", 0x10, 0, 0, 0, 1, 0, 8)
_BOOL_ENUM = {"FALSE": 0, "TRUE": 1}


def _enc_dtype(dt, enum=None):
    """numpy dtype (+ optional enum mapping) -> datatype message bytes (unpadded)."""
    dt = np.dtype(dt)
    if enum is not None:
        return _enc_enum(dt, enum)
    if dt.kind == "b":
        return _enc_enum(np.int8, _BOOL_ENUM)
    if dt.kind in "iu":
        return _enc_int(dt.itemsize, dt.kind == "i")
    if dt.kind == "f":
        return _enc_float(dt.itemsize)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, max(1, dt.itemsize))     # null-padded, ASCII
    raise H5FormatError(f"unsupported dtype {dt} for HDF5 output")


def _storage_array(a):
    """Array in its on-disk form (little-endian; bool as int8; S0 -> S1)."""
    a = np.ascontiguousarray(a)
    if a.dtype.kind == "b":
        return a.astype(np.int8)
    if a.dtype.kind == "S" and a.dtype.itemsize == 0:
        return a.astype("S1")
    if a.dtype.kind in "iuf" and a.dtype.byteorder == ">":
        return a.astype(a.dtype.newbyteorder("<"))
    return a


def _enc_space(shape):
    if shape == ():
        return struct.pack("<BBB5x", 1, 0, 0)
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(s)) for s in shape)


# ---------------------------------------------------------------------------------- raw output
class _Out:
    def __init__(self, fh, eof):
        self.fh, self.eof = fh, eof

    def alloc(self, n):
        self.eof += -self.eof % 8
        a = self.eof
        self.eof += n
        return a

    def put(self, addr, data):
        self.fh.seek(addr)
        self.fh.write(data)

    def append(self, data):
        a = self.alloc(len(data))
        self.put(a, data)
        return a


def _split(n, cap):
    """Consecutive, balanced groups [a, b) of at most ``cap`` items covering range(n)."""
    if n == 0:
        return [(0, 0)]
    k = -(-n // cap)
    edges = [(i * n) // k for i in range(k + 1)]
    return list(zip(edges[:-1], edges[1:]))


def _build_btree(out: _Out, ntype, K, keys, children):
    """B-tree v1 over ``children`` (n addresses) with ``keys`` (n + 1 byte strings); returns the root."""
    cap = 2 * K
    ksz = len(keys[0])
    node_size = 24 + (cap + 1) * ksz + cap * 8
    level = 0
    while True:
        n = len(children)
        groups = _split(n, cap)
        addrs = [out.alloc(node_size) for _ in groups]
        for gi, (a, b) in enumerate(groups):
            left = addrs[gi - 1] if gi > 0 else _UNDEF
            right = addrs[gi + 1] if gi + 1 < len(groups) else _UNDEF
            buf = bytearray(b"TREE" + struct.pack("<BBHQQ", ntype, level, b - a, left, right))
            for i in range(a, b):
                buf += keys[i] + struct.pack("<Q", children[i])
            buf += keys[b]
            buf += b"\x00" * (node_size - len(buf))
            out.put(addrs[gi], bytes(buf))
        if len(groups) == 1:
            return addrs[0]
        keys = [keys[a] for a, _ in groups] + [keys[n]]
        children = addrs
        level += 1


def _build_group_index(out: _Out, links, leaf_k, internal_k):
    """Local heap + symbol table nodes + B-tree for ``links`` (name -> header address)."""
    names = sorted(links, key=lambda s: s.encode("utf-8"))
    heap = bytearray(8)                                   # offset 0: the empty name (left-most key)
    offs = {}
    for nm in names:
        offs[nm] = len(heap)
        heap += _pad8(nm.encode("utf-8") + b"\x00")
    free_off = len(heap)
    heap += struct.pack("<QQ", 1, 32) + b"\x00" * 16      # one free block (next = none, 32 bytes)
    data_addr = out.append(bytes(heap))
    heap_addr = out.append(b"HEAP" + b"\x00" * 4 + struct.pack("<QQQ", len(heap), free_off, data_addr))
    cap = 2 * leaf_k
    snods, keys = [], [struct.pack("<Q", 0)]
    if names:
        for a, b in _split(len(names), cap):
            buf = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, b - a))
            for nm in names[a:b]:
                buf += struct.pack("<QQII16x", offs[nm], links[nm], 0, 0)
            buf += b"\x00" * (8 + cap * 40 - len(buf))
            snods.append(out.append(bytes(buf)))
            keys.append(struct.pack("<Q", offs[names[b - 1]]))
    btree = _build_btree(out, 0, internal_k, keys, snods)
    return btree, heap_addr


# ---------------------------------------------------------------------------------- attributes
def _global_heap(out: _Out, blobs):
    """Write one global heap collection holding ``blobs``; returns (address, [index...])."""
    body = bytearray()
    for i, b in enumerate(blobs, 1):
        body += struct.pack("<HHIQ", i, 0, 0, len(b)) + _pad8(b)
    size = max(4096, 16 + len(body) + 16)
    size += -size % 8
    free = size - 16 - len(body)
    body += struct.pack("<HHIQ", 0, 0, 0, free) + b"\x00" * (free - 16)
    addr = out.append(b"GCOL" + struct.pack("<B3xQ", 1, size) + bytes(body))
    return addr, list(range(1, len(blobs) + 1))


def _normalize_attr(v):
    """Python / numpy value -> ('str', [str...], shape) or ('num', ndarray, shape)."""
    if v is None:
        v = "null"                                        # the reference stores missing values as "null"
    if isinstance(v, bytes):
        v = v.decode("utf-8", "replace")
    if isinstance(v, str):
        return "str", [v], ()
    a = np.asarray(v)
    if a.dtype.kind in "US" or a.dtype == object:
        flat = [x.decode("utf-8", "replace") if isinstance(x, bytes) else str(x) for x in a.ravel().tolist()]
        return "str", flat, tuple(a.shape)
    if a.dtype.kind not in "biuf":
        raise H5FormatError(f"unsupported attribute value type {a.dtype}")
    if a.ndim > 1:
        raise H5FormatError("only scalar and 1-D attributes are supported")
    if a.dtype.kind == "i" and a.dtype.itemsize < 8 and not isinstance(v, np.generic) and not isinstance(v, np.ndarray):
        a = a.astype(np.int64)                            # python ints are stored as int64, like h5py
    return "num", a, tuple(a.shape)


def _enc_attrs(out: _Out, attrs):
    """dict -> list of attribute messages (type 0x0C, flags, body).  vlen strings go to one heap."""
    norm = {k: _normalize_attr(v) for k, v in attrs.items()}
    blobs, where = [], {}
    for k, (kind, val, _shape) in norm.items():
        if kind == "str":
            for j, s in enumerate(val):
                where[(k, j)] = len(blobs)
                blobs.append(s.encode("utf-8"))
    gaddr, idx = _global_heap(out, blobs) if blobs else (0, [])
    msgs = []
    for k, (kind, val, shape) in norm.items():
        name = k.encode("utf-8") + b"\x00"
        if kind == "str":
            typ = _VLEN_STR
            data = b"".join(struct.pack("<IQI", len(blobs[where[(k, j)]]), gaddr, idx[where[(k, j)]])
                            for j in range(len(val)))
        else:
            typ = _enc_dtype(val.dtype)
            data = _storage_array(val).tobytes()
        space = _enc_space(shape)
        body = struct.pack("<BBHHH", 1, 0, len(name), len(typ), len(space))
        body += _pad8(name) + _pad8(typ) + _pad8(space) + data
        body = _pad8(body)
        if len(body) > 0xFFF8:
            raise H5FormatError(f"attribute '{k}' is too large for an object-header message")
        msgs.append((0x0C, 0x04, body))
    return msgs


def _attr_name(body):
    ver = body[0]
    nsz = struct.unpack_from("<H", body, 2)[0]
    p = 9 if ver == 3 else 8
    return body[p:p + nsz].split(b"\x00")[0].decode("utf-8", "replace")


def _write_ohdr(out: _Out, msgs):
    """Version-1 object header with ``msgs`` [(type, flags, body)]; returns (address, body offsets)."""
    size = sum(8 + len(b) for _, _, b in msgs)
    if any(len(b) % 8 or len(b) > 0xFFF8 for _, _, b in msgs):
        raise H5FormatError("object-header message of invalid size")
    buf = bytearray(struct.pack("<BBHII4x", 1, 0, len(msgs), 1, size))
    offs = []
    for t, f, b in msgs:
        buf += struct.pack("<HHB3x", t, len(b), f)
        offs.append(len(buf))
        buf += b
    addr = out.append(bytes(buf))
    return addr, [addr + o for o in offs]


# ---------------------------------------------------------------------------------- chunk output
def _shuffle(raw, es):
    if es <= 1:
        return raw
    a = np.frombuffer(raw, dtype=np.uint8).reshape(-1, es)
    return np.ascontiguousarray(a.T).tobytes()


class _ChunkSink:
    """Compresses and appends the chunks of one 1-D dataset; builds its chunk B-tree at the end."""

    def __init__(self, fobj, dtype, chunk_len, level, shuffle):
        self.f, self.dt = fobj, np.dtype(dtype)
        self.chunk_len, self.level, self.shuffle = int(chunk_len), level, shuffle
        self.entries = []                                 # (element offset, address, stored bytes)
        self.n = 0
        self._tail = np.zeros(0, dtype=self.dt)

    def _encode(self, arr):
        if len(arr) < self.chunk_len:                     # the last chunk is stored at full size
            arr = np.concatenate([arr, np.zeros(self.chunk_len - len(arr), dtype=self.dt)])
        raw = arr.tobytes()
        if self.shuffle:
            raw = _shuffle(raw, self.dt.itemsize)
        if self.level is not None:
            raw = zlib.compress(raw, self.level)
        return raw

    def _emit(self, blocks, first_off):
        if len(blocks) > 4 and self.level is not None:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
                enc = list(ex.map(self._encode, blocks))
        else:
            enc = [self._encode(b) for b in blocks]
        out = self.f._out
        off = first_off
        for raw in enc:
            self.entries.append((off, out.append(raw), len(raw)))
            off += self.chunk_len

    def append(self, arr):
        arr = _storage_array(np.asarray(arr).astype(self.dt, copy=False))
        if arr.ndim != 1:
            raise H5FormatError("only 1-D datasets are supported")
        if len(self._tail):
            arr = np.concatenate([self._tail, arr])
        c = self.chunk_len
        nfull = len(arr) // c
        base = self.n - len(self._tail)
        batch = 256
        for s in range(0, nfull, batch):
            e = min(nfull, s + batch)
            self._emit([arr[i * c:(i + 1) * c] for i in range(s, e)], base + s * c)
        self._tail = arr[nfull * c:].copy()
        self.n = base + len(arr)

    def finish(self):
        """Flush the partial chunk, write the B-tree; returns its address (or UNDEF)."""
        if len(self._tail):
            self._emit([self._tail], self.n - len(self._tail))
            self._tail = self._tail[:0]
        if not self.entries:
            return _UNDEF
        es = self.dt.itemsize
        keys = [struct.pack("<IIQQ", nb, 0, off, 0) for off, _, nb in self.entries]
        keys.append(struct.pack("<IIQQ", 0, 0, self.entries[-1][0], es))
        return _build_btree(self.f._out, 1, self.f._chunk_k, keys, [a for _, a, _ in self.entries])


def _dataset_msgs(dtype, enum, n, chunk_len, btree, level, shuffle):
    dt = np.dtype(dtype)
    es = max(1, dt.itemsize) if dt.kind != "b" else 1
    msgs = [(0x01, 0, struct.pack("<BBB5xQQ", 1, 1, 1, n, n)),
            (0x03, 1, _pad8(_enc_dtype(dt, enum)))]
    if n == 0 or btree == _UNDEF:
        msgs.append((0x05, 1, struct.pack("<BBBBI", 2, 2, 2, 1, 0)))
        msgs.append((0x08, 0, _pad8(struct.pack("<BBQQ", 3, 1, _UNDEF, 0))))
        return msgs
    msgs.append((0x05, 1, struct.pack("<BBBBI", 2, 3, 0, 1, 0)))
    filt = []
    if shuffle:
        filt.append(struct.pack("<HHHH", 2, 8, 1, 1) + b"shuffle\x00" + struct.pack("<II", es, 0))
    if level is not None:
        filt.append(struct.pack("<HHHH", 1, 8, 1, 1) + b"deflate\x00" + struct.pack("<II", level, 0))
    if filt:
        msgs.append((0x0B, 1, struct.pack("<BB6x", 1, len(filt)) + b"".join(filt)))
    msgs.append((0x08, 0, _pad8(struct.pack("<BBBQII", 3, 2, 2, btree, chunk_len, es))))
    return msgs


# ---------------------------------------------------------------------------------- object model
class _Attrs:
    """``obj.attrs``: reads come from the current header, ``update`` / ``[]=`` re-write it."""

    def __init__(self, fobj, parent, name):
        self._f, self._parent, self._name = fobj, parent, name

    def _current(self):
        out = {}
        for t, _f, body in self._f._messages(self._f._addr_of(self._parent, self._name)):
            if t == 0x0C:
                try:
                    k, v = h5lite._parse_attribute(self._f._reader_for_values(), body)
                except Exception:                         # noqa: BLE001 - unreadable attribute: keep the name
                    k, v = _attr_name(body), None
                out[k] = v
        return out

    def update(self, mapping=(), **kw):
        d = dict(mapping)
        d.update(kw)
        if d:
            self._f._set_attrs(self._parent, self._name, d)

    def __setitem__(self, k, v):
        self.update({k: v})

    def __getitem__(self, k):
        return self._current()[k]

    def get(self, k, default=None):
        return self._current().get(k, default)

    def __contains__(self, k):
        return k in self._current()

    def keys(self):
        return self._current().keys()

    def items(self):
        return self._current().items()

    def __iter__(self):
        return iter(self._current())


class Dataset:
    """A dataset reachable from a writable file: attributes can be updated, old data read."""

    def __init__(self, fobj, parent, name):
        self._f, self._parent, self._name = fobj, parent, name

    @property
    def attrs(self):
        return _Attrs(self._f, self._parent, self._name)

    def _old(self):
        addr = self._f._addr_of(self._parent, self._name)
        if self._f._rd_file is None:
            raise H5FormatError("datasets written in this session can be read after close()")
        rd = self._f._rd_file._rd
        if addr < self._f._old_eof:
            return h5lite.Dataset(rd, h5lite._Obj(rd, addr), self._name)
        ds = h5lite.Dataset(rd, _SessionObj(rd, addr, self._f._session[addr]), self._name)
        layout = ds._layout                               # header re-written here, data possibly old
        data_addr = int.from_bytes(layout[3:11] if layout[1] == 2 else layout[2:10], "little")
        if data_addr != _UNDEF and data_addr >= self._f._old_eof:
            raise H5FormatError("datasets written in this session can be read after close()")
        return ds

    def __getitem__(self, key):
        return self._old()[key]

    @property
    def shape(self):
        return self._old().shape

    @property
    def dtype(self):
        return self._old().dtype


class _SessionObj:
    """Object header written in this session, in the shape h5lite.Dataset expects."""

    def __init__(self, rd, addr, msgs):
        self.rd, self.addr, self.msgs, self.version = rd, addr, list(msgs), 1

    def find(self, mtype):
        return [m for m in self.msgs if m[0] == mtype]


class StreamDataset:
    """1-D dataset written by ``append()`` calls; linked into its group by ``close()``."""

    def __init__(self, fobj, group, name, dtype, enum, chunk_len, level, shuffle):
        self._f, self._group, self._name = fobj, group, name
        self._dtype, self._enum = np.dtype(dtype), enum
        self._sink = _ChunkSink(fobj, _storage_array(np.zeros(0, dtype)).dtype, chunk_len, level, shuffle)
        self._level, self._shuffle = level, shuffle
        self._attrs = {}
        self.closed = False

    def append(self, arr):
        self._sink.append(arr)

    def __len__(self):
        return self._sink.n

    def close(self, attrs=None):
        if self.closed:
            return
        self.closed = True
        btree = self._sink.finish()
        msgs = _dataset_msgs(self._dtype, self._enum, self._sink.n, self._sink.chunk_len, btree,
                             self._level, self._shuffle)
        if attrs:
            msgs += _enc_attrs(self._f._out, attrs)
        addr, _ = _write_ohdr(self._f._out, msgs)
        self._f._session[addr] = msgs
        self._group._state.links[self._name] = addr
        self._f._touch(self._group._state)
        self._f._streams.remove(self)


class _GroupState:
    def __init__(self, path, hdr_addr, stab_off, links, parent, name):
        self.path, self.hdr_addr, self.stab_off = path, hdr_addr, stab_off
        self.links, self.parent, self.name = links, parent, name
        self.dirty = False


class Group:
    def __init__(self, fobj, state):
        self._f, self._state = fobj, state

    @property
    def name(self):
        return self._state.path

    @property
    def attrs(self):
        return _Attrs(self._f, self._state.parent, self._state.name)

    def keys(self):
        return list(self._state.links.keys())

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self._state.links)

    def _walk(self, path, create=False):
        cur = self._f._root_state if str(path).startswith("/") else self._state
        parts = [p for p in str(path).split("/") if p]
        return cur, parts

    def __contains__(self, path):
        try:
            self[path]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        cur, parts = self._walk(path)
        for i, part in enumerate(parts):
            if part not in cur.links:
                raise KeyError(f"Unable to open object (component not found: '{part}')")
            child = self._f._group_state(cur, part)
            if child is None:
                if i != len(parts) - 1:
                    raise KeyError(path)
                return Dataset(self._f, cur, part)
            cur = child
        return Group(self._f, cur)

    def __delitem__(self, path):
        cur, parts = self._walk(path)
        for part in parts[:-1]:
            if part not in cur.links:
                raise KeyError(path)
            cur = self._f._group_state(cur, part)
            if cur is None:
                raise KeyError(path)
        if parts[-1] not in cur.links:
            raise KeyError(f"Couldn't delete link (name doesn't exist: '{parts[-1]}')")
        self._f._forget(cur, parts[-1])
        del cur.links[parts[-1]]
        self._f._touch(cur)

    def create_group(self, path):
        cur, parts = self._walk(path)
        for i, part in enumerate(parts):
            if part in cur.links:
                if i == len(parts) - 1:
                    raise ValueError(f"Unable to create group (name already exists: '{part}')")
                nxt = self._f._group_state(cur, part)
                if nxt is None:
                    raise ValueError(f"'{part}' is a dataset")
                cur = nxt
            else:
                cur = self._f._new_group(cur, part)
        return Group(self._f, cur)

    def require_group(self, path):
        try:
            g = self[path]
        except KeyError:
            return self.create_group(path)
        if not isinstance(g, Group):
            raise TypeError(f"'{path}' is not a group")
        return g

    def copy_dataset(self, src, name):
        """Raw copy of an ``h5lite.Dataset`` (chunks are not re-encoded); see ``_copy_dataset_raw``."""
        return _copy_dataset_raw(self, name, src)

    def _target(self, name):
        parts = [p for p in str(name).split("/") if p]
        grp = self if len(parts) == 1 else self.require_group("/".join(parts[:-1]))
        if parts[-1] in grp._state.links:
            raise ValueError(f"Unable to create dataset (name already exists: '{parts[-1]}')")
        return grp, parts[-1]

    def create_dataset(self, name, shape=None, dtype=None, data=None, *, enum=None, compression="gzip",
                       compression_opts=None, shuffle=True, chunks=None, attrs=None, **_ignored):
        """h5py-like; 1-D only.  ``data=None`` returns a ``StreamDataset`` (``append`` / ``close``)."""
        grp, leaf = self._target(name)
        level = None
        if compression in ("gzip", True):
            level = 6 if compression_opts is None else int(compression_opts)
        elif compression not in (None, False):
            raise H5FormatError(f"unsupported compression '{compression}'")
        if data is None:
            if dtype is None:
                raise TypeError("create_dataset needs data= or dtype=")
            clen = int(chunks[0]) if isinstance(chunks, (tuple, list)) else _DEFAULT_CHUNK
            ds = StreamDataset(self._f, grp, leaf, dtype, enum, clen, level, shuffle)
            self._f._streams.append(ds)
            return ds
        arr = np.asarray(data)
        if dtype is not None:
            arr = arr.astype(dtype, copy=False)
        if arr.dtype.kind == "U":
            arr = np.char.encode(arr, "utf-8")
        if arr.ndim != 1:
            raise H5FormatError("only 1-D datasets are supported")
        if shape is not None and tuple(np.atleast_1d(shape)) != arr.shape:
            raise ValueError("shape does not match data")
        es = max(1, arr.dtype.itemsize)
        clen = int(chunks[0]) if isinstance(chunks, (tuple, list)) else max(1, min(len(arr), (1 << 19) // es))
        ds = StreamDataset(self._f, grp, leaf, arr.dtype, enum, clen, level, shuffle)
        self._f._streams.append(ds)
        ds.append(arr)
        ds.close(attrs)
        return Dataset(self._f, grp._state, leaf)


def _copy_dataset_raw(dst_group, name, src):
    """Copy the h5lite dataset ``src`` into ``dst_group`` without re-encoding its chunks.

    The header messages (dataspace, datatype, fill value, filter pipeline) are taken verbatim and the
    stored chunk bytes are copied as they are, like ``h5py.Group.copy``; attributes are re-encoded
    from their values (vlen strings live in the source file's global heap).  Falls back to a
    decode / encode copy for layouts other than version-3 chunked / contiguous.
    """
    f = dst_group._f
    grp, leaf = dst_group._target(name)
    rd, lay = src._rd, src._layout
    attrs = {k: v for k, v in dict(src.attrs).items() if v is not None}
    keep = [m for m in src._obj.msgs if m[0] in (0x01, 0x03, 0x05, 0x0B)]
    if src._obj.version != 1 or lay[0] != 3 or lay[1] not in (1, 2) or len(src.shape) != 1 or \
            any(len(b) % 8 for _, _, b in keep):
        return grp.create_dataset(leaf, data=src[:], enum=src.enum, attrs=attrs or None)
    if lay[1] == 2:
        cdims, chunks = src._chunk_index()
        es = struct.unpack_from("<I", lay, 3 + rd.O + 4)[0]
        entries = []
        for offs, addr, nbytes, mask in sorted(chunks, key=lambda c: c[0]):
            entries.append((offs[0], f._out.append(rd.bytes(addr, nbytes)), nbytes, mask))
        if entries:
            keys = [struct.pack("<IIQQ", nb, mk, off, 0) for off, _, nb, mk in entries]
            keys.append(struct.pack("<IIQQ", 0, 0, entries[-1][0], es))
            btree = _build_btree(f._out, 1, f._chunk_k, keys, [a for _, a, _, _ in entries])
        else:
            btree = _UNDEF
        layout = _pad8(struct.pack("<BBBQII", 3, 2, 2, btree, cdims[0], es))
    else:
        addr = int.from_bytes(lay[2:2 + rd.O], "little")
        size = int.from_bytes(lay[2 + rd.O:2 + rd.O + rd.L], "little")
        new = _UNDEF if addr == _UNDEF or size == 0 else f._out.append(rd.bytes(addr + rd.base, size))
        layout = _pad8(struct.pack("<BBQQ", 3, 1, new, size))
    msgs = [(t, fl, b) for t, fl, b in keep] + [(0x08, 0, layout)]
    if attrs:
        msgs += _enc_attrs(f._out, attrs)
    addr, _ = _write_ohdr(f._out, msgs)
    f._session[addr] = msgs
    grp._state.links[leaf] = addr
    f._touch(grp._state)
    return Dataset(f, grp._state, leaf)


_NOTICE_GIVEN = False


def _experimental_notice():
    """One log line per process: this writer's files have never been opened by libhdf5 / h5py (neither is
    in the build image); they are checked by tests/_h5check.py and read back by h5lite only."""
    global _NOTICE_GIVEN
    if not _NOTICE_GIVEN:
        _NOTICE_GIVEN = True
        import logging
        logging.getLogger("cooler").warning(
            "h5py is not installed: writing with cooler_b200's built-in HDF5 writer (EXPERIMENTAL -- verify "
            "the output with h5py / `cooler info` before relying on it)")


class File(Group):
    """Writable HDF5 file: ``File(path, 'w' | 'r+' | 'a')``; use as a context manager.  EXPERIMENTAL:
    used only when ``cooler`` + h5py are not importable (see ``_experimental_notice``)."""

    def __init__(self, path, mode="r+"):
        if mode not in ("w", "r+", "a"):
            raise ValueError("mode must be 'w', 'r+' or 'a'")
        _experimental_notice()
        exists = os.path.exists(path) and os.path.getsize(path) > 0
        if mode == "a":
            mode = "r+" if exists else "w"
        if mode == "r+" and not exists:
            raise FileNotFoundError(path)
        self.filename = path
        self._session = {}            # header address -> messages written in this session
        self._states = {}             # header address -> _GroupState
        self._streams = []
        self._closed = False
        if mode == "w":
            self._rd_file = None
            self._old_eof = 0
            self._fh = open(path, "w+b")
            self._sb_off, self._leaf_k, self._internal_k, self._chunk_k = 0, _LEAF_K, _INTERNAL_K, _CHUNK_K
            self._root_entry_off, self._eof_off = 56, 40
            self._out = _Out(self._fh, 96)
            self._fh.write(b"\x00" * 96)
            msgs = [(0x11, 0, struct.pack("<QQ", _UNDEF, _UNDEF))]
            addr, offs = _write_ohdr(self._out, msgs)
            self._session[addr] = msgs
            self._root_state = _GroupState("/", addr, offs[0], {}, None, None)
            self._root_state.dirty = True
        else:
            rf = h5lite.File(path)
            sb = rf._sb
            if sb["ver"] not in (0, 1) or rf._rd.O != 8 or rf._rd.L != 8 or rf._rd.base != 0:
                raise H5FormatError(f"{path}: only files with a version 0/1 superblock and 8-byte offsets "
                                    "(h5py's default libver='earliest') can be modified by cooler_b200.h5write")
            self._rd_file = rf
            self._sb_off, self._leaf_k, self._internal_k = sb["off"], sb["leaf_k"], sb["internal_k"]
            self._chunk_k = sb["chunk_k"]
            self._root_entry_off, self._eof_off = sb["root_entry_off"], sb["eof_off"]
            self._old_eof = max(sb["eof"], os.path.getsize(path))
            self._fh = open(path, "r+b")
            self._out = _Out(self._fh, self._old_eof)
            self._root_state = self._load_state("/", rf._obj.addr, None, None)
        self._states[self._root_state.hdr_addr] = self._root_state
        super().__init__(self, self._root_state)

    # -- state -----------------------------------------------------------------------------
    def _messages(self, addr):
        if addr in self._session:
            return self._session[addr]
        obj = h5lite._Obj(self._rd_file._rd, addr)
        if obj.version != 1:
            raise H5FormatError("version-2 object headers (libver='latest') cannot be modified")
        return [m for m in obj.msgs if m[0] not in (0x00, 0x10)]

    def _reader_for_values(self):
        """A reader able to resolve global-heap strings of OLD attributes (new ones are in-session)."""
        if self._rd_file is not None:
            return self._rd_file._rd
        return _NullReader()

    def _load_state(self, path, addr, parent, name):
        rd = self._rd_file._rd
        obj = h5lite._Obj(rd, addr)
        if obj.version != 1:
            raise H5FormatError("version-2 object headers (libver='latest') cannot be modified")
        stab = [i for i, m in enumerate(obj.msgs) if m[0] == 0x11]
        if not stab:
            if obj.find(0x02) or obj.find(0x06):
                raise H5FormatError(f"group '{path}' uses link messages (libver='latest'); it cannot be modified")
            return None
        g = h5lite.Group(rd, obj, path)
        return _GroupState(path, addr, obj.moffs[stab[0]], dict(g._load_links()), parent, name)

    def _group_state(self, parent, name):
        """State of child group ``name`` of ``parent`` (None when the child is a dataset)."""
        addr = parent.links[name]
        if addr in self._states:
            return self._states[addr]
        if addr in self._session:
            return None                                   # a dataset written in this session
        obj = h5lite._Obj(self._rd_file._rd, addr)
        if obj.find(0x08) and obj.find(0x03):
            return None
        st = self._load_state(parent.path.rstrip("/") + "/" + name, addr, parent, name)
        if st is not None:
            self._states[addr] = st
        return st

    def _new_group(self, parent, name):
        msgs = [(0x11, 0, struct.pack("<QQ", _UNDEF, _UNDEF))]
        addr, offs = _write_ohdr(self._out, msgs)
        self._session[addr] = msgs
        st = _GroupState(parent.path.rstrip("/") + "/" + name, addr, offs[0], {}, parent, name)
        self._states[addr] = st
        parent.links[name] = addr
        self._touch(st)
        return st

    def _forget(self, parent, name):
        self._states.pop(parent.links.get(name), None)

    def _touch(self, st):
        while st is not None:
            st.dirty = True
            st = st.parent

    def _addr_of(self, parent, name):
        return self._root_state.hdr_addr if parent is None else parent.links[name]

    def _set_attrs(self, parent, name, new):
        """Re-write the header of ``parent/name`` (root when parent is None) with ``new`` attributes."""
        old_addr = self._addr_of(parent, name)
        msgs = [m for m in self._messages(old_addr) if not (m[0] == 0x0C and _attr_name(m[2]) in new)]
        if any(m[0] == 0x15 for m in msgs):
            raise H5FormatError("objects with densely stored attributes cannot be modified")
        msgs = msgs + _enc_attrs(self._out, new)
        addr, offs = _write_ohdr(self._out, msgs)
        self._session[addr] = msgs
        st = self._states.pop(old_addr, None)
        if st is not None:                                # a group: its stab message moved with the header
            k = [i for i, m in enumerate(msgs) if m[0] == 0x11][0]
            st.hdr_addr, st.stab_off = addr, offs[k]
            self._states[addr] = st
            self._touch(st)
        if parent is not None:
            parent.links[name] = addr
            self._touch(parent)

    # -- finish ----------------------------------------------------------------------------
    def flush(self):
        for ds in list(self._streams):
            ds.close()
        root = self._root_state
        if self._rd_file is not None and self._out.eof == self._old_eof and \
                not any(st.dirty for st in self._states.values()):
            return                                        # opened 'r+' but nothing was changed
        for st in list(self._states.values()):
            if st.dirty:
                btree, heap = _build_group_index(self._out, st.links, self._leaf_k, self._internal_k)
                self._out.put(st.stab_off, struct.pack("<QQ", btree, heap))
                # keep the in-session copy of the header in sync (used when it is re-written later)
                if st.hdr_addr in self._session:
                    self._session[st.hdr_addr] = [
                        (t, f, struct.pack("<QQ", btree, heap)) if t == 0x11 else (t, f, b)
                        for t, f, b in self._session[st.hdr_addr]]
                st.dirty = False
                if st is root:
                    self._root_cache = (btree, heap)
        out = self._out
        out.eof += -out.eof % 8
        self._fh.seek(0, os.SEEK_END)
        if self._fh.tell() < out.eof:
            self._fh.write(b"\x00" * (out.eof - self._fh.tell()))
        if self._rd_file is None:
            sb = h5lite._SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self._leaf_k, self._internal_k, 0)
            sb += struct.pack("<QQQQ", 0, _UNDEF, out.eof, _UNDEF)
            out.put(0, sb)
        else:
            out.put(self._eof_off, struct.pack("<Q", out.eof))
        if hasattr(self, "_root_cache"):
            bt, hp = self._root_cache
        else:                                             # root index untouched: keep the cached addresses
            stab = [m for m in h5lite._Obj(self._rd_file._rd, root.hdr_addr).msgs if m[0] == 0x11][0][2]
            bt, hp = struct.unpack("<QQ", stab[:16])
        out.put(self._root_entry_off, struct.pack("<QQIIQQ", 0, root.hdr_addr, 1, 0, bt, hp))
        self._fh.flush()

    def close(self):
        if self._closed:
            return
        self.flush()
        self._closed = True
        try:
            os.fsync(self._fh.fileno())
        except OSError:
            pass
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class _NullReader:
    """Placeholder reader for brand-new files (attribute read-back of in-session values)."""
    O = L = 8
    base = 0

    def bytes(self, off, n):
        raise H5FormatError("attribute values written in this session can be read after close()")

    u = len_ = bytes
