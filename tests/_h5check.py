"""Strict structural checker for the HDF5 subset cooler files use (test helper).

Independent of ``cooler_b200.h5lite`` (which is lenient by design): walks the raw bytes and
asserts the invariants libhdf5 relies on when it opens a file written with
``libver='earliest'`` -- superblock v0, v1 object headers, symbol-table groups (B-tree v1,
local heap, SNOD), chunk B-trees, global heap collections.  It is calibrated on the
reference's own ``.cool`` files (written by libhdf5): tests run it on those first, then on
files produced by ``cooler_b200.h5write``.
"""
import struct

UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b"\x89HDF\r\n\x1a\n"


class Check:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.b = f.read()
        self.n = len(self.b)
        self.objects = 0
        self.chunks = 0
        self.gcols = set()

    def u(self, off, size):
        assert 0 <= off and off + size <= self.n, f"read past EOF at {off}+{size}"
        return int.from_bytes(self.b[off:off + size], "little")

    def run(self):
        b = self.b
        assert b[:8] == SIG
        assert b[8] == 0, "superblock version 0 expected"
        assert b[13] == 8 and b[14] == 8
        self.leaf_k, self.internal_k = self.u(16, 2), self.u(18, 2)
        assert self.leaf_k > 0 and self.internal_k > 0
        base, free, eof, drv = (self.u(24 + 8 * i, 8) for i in range(4))
        assert base == 0 and free == UNDEF and drv == UNDEF
        assert eof == self.n, f"EOF address {eof} != file size {self.n}"
        name_off, hdr, cache = self.u(56, 8), self.u(64, 8), self.u(72, 4)
        assert name_off == 0
        root_stab = self.group(hdr, "/")
        if cache == 1:
            assert (self.u(80, 8), self.u(88, 8)) == root_stab, "root entry cache disagrees with the header"
        else:
            assert cache == 0
        return self

    # -- object headers ---------------------------------------------------------------------
    def messages(self, addr):
        assert self.u(addr, 1) == 1, f"object header version at {addr}"
        assert self.u(addr + 1, 1) == 0
        nmsg, refs, size = self.u(addr + 2, 2), self.u(addr + 4, 4), self.u(addr + 8, 4)
        assert refs >= 1
        blocks, msgs = [(addr + 16, size)], []
        while blocks:
            p, sz = blocks.pop(0)
            end = p + sz
            assert end <= self.n
            while p < end:
                assert p + 8 <= end, "truncated message header"
                t, ms, fl = self.u(p, 2), self.u(p + 2, 2), self.u(p + 4, 1)
                assert ms % 8 == 0, f"message size {ms} not a multiple of 8"
                assert p + 8 + ms <= end, "message overruns its block"
                body = self.b[p + 8:p + 8 + ms]
                if t == 0x10:
                    blocks.append((self.u(p + 8, 8), self.u(p + 16, 8)))
                msgs.append((t, fl, body))
                p += 8 + ms
        assert len(msgs) == nmsg, f"header at {addr}: {len(msgs)} messages, header says {nmsg}"
        self.objects += 1
        return msgs

    def attributes(self, msgs):
        for t, _fl, body in msgs:
            if t != 0x0C:
                continue
            assert body[0] == 1
            nsz, tsz, ssz = struct.unpack_from("<HHH", body, 2)
            p = 8
            name = body[p:p + nsz]
            assert name.endswith(b"\x00") and b"\x00" not in name[:-1]
            p += (nsz + 7) & ~7
            typ = body[p:p + tsz]
            p += (tsz + 7) & ~7
            space = body[p:p + ssz]
            p += (ssz + 7) & ~7
            assert space[0] == 1
            rank = space[1]
            count = 1
            for i in range(rank):
                count *= int.from_bytes(space[8 + 8 * i:16 + 8 * i], "little")
            cls = typ[0] & 0x0F
            tsize = struct.unpack_from("<I", typ, 4)[0]
            assert len(body) - p >= count * tsize, "attribute data shorter than its dataspace"
            if cls == 9:                                  # vlen string: every element points into a GCOL
                for i in range(count):
                    ln, gaddr, idx = struct.unpack_from("<IQI", body, p + 16 * i)
                    if ln:
                        assert self.global_heap_object(gaddr, idx) >= ln

    def global_heap_object(self, addr, index):
        assert self.b[addr:addr + 4] == b"GCOL" and self.b[addr + 4] == 1
        size = self.u(addr + 8, 8)
        assert size >= 4096 and addr + size <= self.n
        p, end, found = addr + 16, addr + size, None
        while p + 16 <= end:
            idx, osz = self.u(p, 2), self.u(p + 8, 8)
            if idx == 0:
                assert p + osz == end, "global heap free space must run to the end of the collection"
                break
            if idx == index:
                found = osz
            p += 16 + ((osz + 7) & ~7)
        assert found is not None, "global heap object not found"
        self.gcols.add(addr)
        return found

    # -- groups -----------------------------------------------------------------------------
    def group(self, addr, path):
        msgs = self.messages(addr)
        self.attributes(msgs)
        stab = [m for m in msgs if m[0] == 0x11]
        assert len(stab) == 1, f"group {path} without a symbol table message"
        btree, heap = struct.unpack("<QQ", stab[0][2][:16])
        assert self.b[heap:heap + 4] == b"HEAP" and self.b[heap + 4] == 0
        dsize, free, daddr = self.u(heap + 8, 8), self.u(heap + 16, 8), self.u(heap + 24, 8)
        assert daddr + dsize <= self.n
        seen = set()
        while free != 1:                                  # H5HL_FREE_NULL
            assert free % 8 == 0 and free + 16 <= dsize and free not in seen, "bad local-heap free list"
            seen.add(free)
            nxt, fsz = self.u(daddr + free, 8), self.u(daddr + free + 8, 8)
            assert fsz >= 16 and free + fsz <= dsize
            free = nxt

        def name_at(off):
            assert off < dsize
            e = self.b.index(b"\x00", daddr + off)
            assert e < daddr + dsize
            return self.b[daddr + off:e]

        links = []
        lo, hi = self.group_node(btree, name_at, links, is_root=True)
        names = [n for n, _ in links]
        assert names == sorted(names) and len(set(names)) == len(names), f"links of {path} not sorted/unique"
        assert lo == b"", "left-most group key must be the empty string"
        for name, child in links:
            cpath = path.rstrip("/") + "/" + name.decode()
            cm = self.messages(child)
            self.objects -= 1
            if any(m[0] == 0x11 for m in cm):
                self.group(child, cpath)
            else:
                self.dataset(child, cpath)
        return btree, heap

    def _siblings(self, tree, level, addr):
        """Nodes of one level form a doubly linked list in key order, across parents."""
        prev = tree.get(level)
        left, right = self.u(addr + 8, 8), self.u(addr + 16, 8)
        if prev is None:
            assert left == UNDEF, "left-most node has a left sibling"
        else:
            assert left == prev[0] and prev[1] == addr, "sibling pointers"
        tree[level] = (addr, right)

    def _siblings_done(self, tree):
        assert all(right == UNDEF for _, right in tree.values()), "right-most node has a right sibling"

    def group_node(self, addr, name_at, links, is_root=False, tree=None):
        assert self.b[addr:addr + 4] == b"TREE", f"group B-tree node at {addr}"
        ntype, level, nent = self.b[addr + 4], self.b[addr + 5], self.u(addr + 6, 2)
        assert ntype == 0
        cap = 2 * self.internal_k
        assert nent <= cap and (nent > 0 or is_root)
        assert addr + 24 + (cap + 1) * 8 + cap * 8 <= self.n, "group B-tree node is not allocated at full size"
        tree = {} if tree is None else tree
        self._siblings(tree, level, addr)
        p = addr + 24
        keys, kids = [], []
        for i in range(nent):
            keys.append(name_at(self.u(p, 8)))
            kids.append(self.u(p + 8, 8))
            p += 16
        keys.append(name_at(self.u(p, 8)))
        assert keys == sorted(keys), "group B-tree keys out of order"
        for i, kid in enumerate(kids):
            if level == 0:
                assert self.b[kid:kid + 4] == b"SNOD" and self.b[kid + 4] == 1
                nsym = self.u(kid + 6, 2)
                assert 0 < nsym <= 2 * self.leaf_k
                assert kid + 8 + 2 * self.leaf_k * 40 <= self.n, "SNOD is not allocated at full size"
                names = []
                for j in range(nsym):
                    e = kid + 8 + 40 * j
                    nm = name_at(self.u(e, 8))
                    cache = self.u(e + 16, 4)
                    assert cache in (0, 1, 2)
                    names.append(nm)
                    links.append((nm, self.u(e + 8, 8)))
                assert names == sorted(names)
                assert keys[i] < names[0] or (keys[i] == b"" and i == 0 and names[0] >= b""), "left key bound"
                assert names[-1] == keys[i + 1], "right key must be the largest name of the node"
            else:
                lo, hi = self.group_node(kid, name_at, links, False, tree)
                assert lo == keys[i] and hi == keys[i + 1], "internal keys must repeat the children's bounds"
        if is_root:
            self._siblings_done(tree)
        return keys[0], keys[-1]

    # -- datasets ---------------------------------------------------------------------------
    def dataset(self, addr, path):
        msgs = self.messages(addr)
        self.attributes(msgs)
        by = {}
        for t, fl, body in msgs:
            by.setdefault(t, []).append(body)
        for t in (0x01, 0x03, 0x08):
            assert len(by.get(t, [])) == 1, f"{path}: message {t:#x}"
        space = by[0x01][0]
        assert space[0] == 1 and space[1] == 1, f"{path}: 1-D simple dataspace expected"
        n = int.from_bytes(space[8:16], "little")
        typ = by[0x03][0]
        tsize = struct.unpack_from("<I", typ, 4)[0]
        lay = by[0x08][0]
        assert lay[0] == 3
        if lay[1] == 1:                                   # contiguous
            a, sz = struct.unpack_from("<QQ", lay, 2)
            assert 0x0B not in by, "filters need chunked storage"
            if a == UNDEF:
                assert sz == 0 or n == 0 or True
            else:
                assert sz == n * tsize and a + sz <= self.n
            return
        assert lay[1] == 2 and lay[2] == 2, f"{path}: chunked 1-D layout expected"
        btree = int.from_bytes(lay[3:11], "little")
        clen, es = struct.unpack_from("<II", lay, 11)
        assert es == tsize and clen > 0
        if 0x0B in by:
            f = by[0x0B][0]
            assert f[0] == 1 and 1 <= f[1] <= 4
        if btree == UNDEF:
            return
        entries = []
        lo, hi = self.chunk_node(btree, entries, True)
        offs = [e[0] for e in entries]
        assert offs == sorted(set(offs)), f"{path}: chunk offsets not strictly increasing"
        assert all(o % clen == 0 for o in offs)
        assert offs[0] == 0 and offs[-1] < max(n, 1), f"{path}: chunk offsets outside the dataset"
        assert len(offs) == -(-n // clen), f"{path}: {len(offs)} chunks for {n} rows of {clen}"
        for o, a, nb in entries:
            assert nb > 0 and a + nb <= self.n
        self.chunks += len(entries)

    def chunk_node(self, addr, entries, is_root=False, tree=None):
        assert self.b[addr:addr + 4] == b"TREE", f"chunk B-tree node at {addr}"
        ntype, level, nent = self.b[addr + 4], self.b[addr + 5], self.u(addr + 6, 2)
        assert ntype == 1
        cap = 64
        assert 0 < nent <= cap
        assert addr + 24 + (cap + 1) * 24 + cap * 8 <= self.n, "chunk B-tree node is not allocated at full size"
        tree = {} if tree is None else tree
        self._siblings(tree, level, addr)
        p = addr + 24
        keys, kids = [], []
        for i in range(nent + 1):
            keys.append((self.u(p + 8, 8), self.u(p + 16, 8), self.u(p, 4), self.u(p + 4, 4)))
            p += 24
            if i < nent:
                kids.append(self.u(p, 8))
                p += 8
        order = [(k[0], k[1]) for k in keys]
        assert order == sorted(order), "chunk keys out of order"
        assert all(order[i] < order[i + 1] for i in range(nent)), "chunk keys must increase"
        for i, kid in enumerate(kids):
            if level == 0:
                off, last, nbytes, mask = keys[i]
                assert last == 0 and mask == 0
                entries.append((off, kid, nbytes))
            else:
                lo, hi = self.chunk_node(kid, entries, False, tree)
                assert lo[:2] == keys[i][:2] and hi[:2] == keys[i + 1][:2], "internal chunk keys"
        if is_root:
            self._siblings_done(tree)
        return keys[0], keys[-1]


def check(path):
    return Check(path).run()
