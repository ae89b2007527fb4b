"""HDF5 writer (cooler_b200.h5write) and file-level cooler creation (cooler_b200.create).

CPU only.  h5py / libhdf5 are absent in this image, so validity is established three ways:
(1) a strict structural checker (tests/_h5check.py) calibrated on the reference's own files,
(2) byte-for-byte comparison of the object-header messages with those libhdf5 wrote for the
same dataset / attribute in the reference's files, (3) round trips through h5lite, including
in-place modification of copies of the reference's files.
"""
import os
import shutil
import struct

import numpy as np
import pandas as pd
import pytest

from cooler_b200 import create, h5lite, h5write
from cooler_b200.store import H5Cooler

from ._h5check import check

COOL = os.path.join(os.path.dirname(__file__), "golden", "cool")
REF_FILES = ["hg19.GM12878-MboI.matrix.2000kb.cool", "odd.1.cool", "toy.asymm.2.cool", "toy.symm.upper.2.mcool"]


def _copy(name, tmp_path):
    dst = str(tmp_path / name)
    shutil.copy(os.path.join(COOL, name), dst)
    os.chmod(dst, 0o644)
    return dst


@pytest.mark.parametrize("name", REF_FILES)
def test_checker_accepts_reference_files(name):
    """The structural checker is calibrated on files written by libhdf5."""
    c = check(os.path.join(COOL, name))
    assert c.objects >= 15 and c.chunks > 0


def test_new_file_roundtrip_and_structure(tmp_path):
    p = str(tmp_path / "new.h5")
    rng = np.random.default_rng(0)
    a = np.sort(rng.integers(0, 1 << 40, 300_000)).astype(np.int64)
    c = rng.integers(0, 100, 300_000).astype(np.int32)
    with h5write.File(p, "w") as f:
        g = f.create_group("pixels")
        g.create_dataset("bin1_id", data=a, chunks=(64,))               # 4688 chunks: 3-level B-tree
        ds = g.create_dataset("count", dtype=np.int32, chunks=(1000,))  # streamed, 300 chunks: 2 levels
        for i in range(0, len(c), 7777):
            ds.append(c[i:i + 7777])
        ds.close()
        b = f.create_group("bins")
        b.create_dataset("chrom", data=np.array([0, 0, 1, 1, 1], np.int32), enum={"chr1": 0, "chrX": 1})
        b.create_dataset("weight", data=np.array([1.5, np.nan, 2, 3, 4.0]),
                         attrs={"converged": True, "tol": 1e-5, "ignore_diags": 2, "scale": np.array([1.0, 2.0]),
                                "cv": np.array([True, False]), "who": "me"})
        ch = f.create_group("chroms")
        ch.create_dataset("name", data=np.array(["chr1", "chrX"], dtype="S"))
        ch.create_dataset("empty", data=np.zeros(0, np.int64))
        ch.create_dataset("f32", data=np.arange(5, dtype=np.float32), compression=None, shuffle=False)
        ch.create_dataset("u8", data=np.arange(5, dtype=np.uint8))
        f.attrs.update({"format": "HDF5::Cooler", "format-version": 3, "bin-size": None, "nnz": 300000,
                        "metadata": "{}", "names": ["a", "bc"]})
        for i in range(300):                                            # > 32 links: several SNODs
            f.create_group(f"resolutions/{1000 * (i + 1)}")
    c_ = check(p)
    assert c_.chunks >= 4688 + 300
    r = h5lite.File(p)
    assert sorted(r.keys()) == ["bins", "chroms", "pixels", "resolutions"]
    assert dict(r.attrs) | {"names": None} == {"format": "HDF5::Cooler", "format-version": 3, "bin-size": "null",
                                               "nnz": 300000, "metadata": "{}", "names": None}
    assert list(r.attrs["names"]) == ["a", "bc"]
    assert np.array_equal(r["pixels/bin1_id"][:], a) and r["pixels/bin1_id"].dtype == np.int64
    assert np.array_equal(r["pixels/count"][:], c) and r["pixels/count"].dtype == np.int32
    assert np.array_equal(r["pixels/count"][1234:99999], c[1234:99999])
    assert r["bins/chrom"].enum == {"chr1": 0, "chrX": 1}
    assert np.array_equal(r["bins/chrom"][:], [0, 0, 1, 1, 1])
    w = r["bins/weight"]
    assert np.array_equal(w[:], [1.5, np.nan, 2, 3, 4.0], equal_nan=True)
    at = dict(w.attrs)
    assert at["converged"] is True and at["tol"] == 1e-5 and at["ignore_diags"] == 2 and at["who"] == "me"
    assert np.array_equal(at["scale"], [1.0, 2.0]) and np.array_equal(at["cv"], [True, False])
    assert list(r["chroms/name"][:]) == [b"chr1", b"chrX"]
    assert len(r["chroms/empty"][:]) == 0 and r["chroms/empty"].dtype == np.int64
    assert np.array_equal(r["chroms/f32"][:], np.arange(5, dtype=np.float32)) and r["chroms/f32"].dtype == np.float32
    assert np.array_equal(r["chroms/u8"][:], np.arange(5)) and r["chroms/u8"].dtype == np.uint8
    assert len(r["resolutions"].keys()) == 300 and "299000" in r["resolutions"]


def _msgs(path, dset):
    f = h5lite.File(path)
    return {t: b for t, _f, b in f[dset]._obj.msgs}, f


def test_dataset_header_bytes_match_libhdf5(tmp_path):
    """Same column, same chunking -> the same dataspace / datatype / fill / filter / layout messages."""
    ref = os.path.join(COOL, "toy.asymm.2.cool")
    want, rf = _msgs(ref, "bins/start")
    data = rf["bins/start"][:]
    clen = struct.unpack_from("<I", want[0x08], 11)[0]
    p = str(tmp_path / "x.h5")
    with h5write.File(p, "w") as f:
        f.create_group("bins").create_dataset("start", data=data, chunks=(clen,))
        f["bins"].create_dataset("chrom", data=rf["bins/chrom"][:], enum=rf["bins/chrom"].enum, chunks=(clen,))
        f.create_group("chroms").create_dataset("name", data=rf["chroms/name"][:].astype("S64"), chunks=(2,))
        f.create_group("indexes").create_dataset("bin1_offset", data=rf["indexes/bin1_offset"][:], chunks=(33,))
    for dset in ("bins/start", "bins/chrom", "chroms/name", "indexes/bin1_offset"):
        want, _ = _msgs(ref, dset)
        got, _ = _msgs(p, dset)
        for t in (0x01, 0x03, 0x05, 0x0B):
            assert got[t] == want[t], f"{dset}: message {t:#x} differs from libhdf5's"
        assert got[0x08][:3] == want[0x08][:3] and got[0x08][11:] == want[0x08][11:], "layout message"
    # the stored chunk inflates to the same bytes (zlib streams may differ between zlib builds)
    a = h5lite.File(p)["bins/start"]
    assert np.array_equal(a[:], data)


def test_attribute_bytes_match_libhdf5(tmp_path):
    ref = h5lite.File(os.path.join(COOL, "toy.asymm.2.cool"))
    want = {h5write._attr_name(b): b for t, _f, b in ref._obj.msgs if t == 0x0C}
    p = str(tmp_path / "a.h5")
    with h5write.File(p, "w") as f:
        f.attrs.update({"nbins": int(ref.attrs["nbins"]), "format": ref.attrs["format"]})
    got = {h5write._attr_name(b): b for t, _f, b in h5lite.File(p)._obj.msgs if t == 0x0C}
    assert got["nbins"] == want["nbins"]
    n = len(want["format"])
    assert got["format"][:n - 12] == want["format"][:n - 12]          # all but the heap address and index
    assert dict(h5lite.File(p).attrs) == {"nbins": ref.attrs["nbins"], "format": "HDF5::Cooler"}


def test_store_weight_into_reference_file(tmp_path):
    """In-place: add, overwrite and annotate bins/weight in a copy of the reference's own file."""
    p = _copy("hg19.GM12878-MboI.matrix.2000kb.cool", tmp_path)
    before = h5lite.File(p)
    cols = {k: before[k][:] for k in ("pixels/bin1_id", "pixels/bin2_id", "pixels/count", "bins/start",
                                      "indexes/bin1_offset")}
    root_attrs = dict(before.attrs)
    size0 = os.path.getsize(p)
    w = np.linspace(0, 1, 1561)
    w[5] = np.nan
    stats = {"tol": 1e-5, "min_nnz": 10, "mad_max": 5, "cis_only": False, "ignore_diags": 2,
             "scale": 56.9, "converged": np.bool_(True), "var": 3.7e-6, "divisive_weights": False}
    clr = H5Cooler(p)
    with clr.open("r+") as grp:
        assert "weight" not in grp["bins"]
        grp["bins"].create_dataset("weight", data=w, compression="gzip", compression_opts=6)
        grp["bins"]["weight"].attrs.update(stats)
    check(p)
    with clr.open("r") as grp:                                   # the reader was refreshed
        assert np.array_equal(grp["bins/weight"][:], w, equal_nan=True)
        at = dict(grp["bins/weight"].attrs)
        assert at["converged"] is True and at["cis_only"] is False and at["scale"] == 56.9
    after = h5lite.File(p)
    for k, v in cols.items():
        assert np.array_equal(after[k][:], v)
    assert dict(after.attrs) == root_attrs
    assert set(after["bins"].keys()) == {"chrom", "start", "end", "weight"}
    with open(p, "rb") as f1, open(os.path.join(COOL, "hg19.GM12878-MboI.matrix.2000kb.cool"), "rb") as f0:
        new, old = f1.read(size0), f0.read()
    diff = [i for i in range(size0) if new[i] != old[i]]
    assert diff and max(diff) - min(diff) > 0
    # only the superblock (EOF + root entry) and the 16-byte symbol-table bodies of '/' and 'bins' changed
    assert len(diff) <= 8 + 40 + 16 + 16
    # overwrite (cooler balance --force): del + create
    with clr.open("r+") as grp:
        assert "weight" in grp["bins"]
        del grp["bins"]["weight"]
        assert "weight" not in grp["bins"]
        grp["bins"].create_dataset("weight", data=2 * np.arange(1561.0))
        grp["bins"].create_dataset("KR", data=np.ones(1561))
    check(p)
    r = h5lite.File(p)
    assert np.array_equal(r["bins/weight"][:], 2 * np.arange(1561.0)) and "converged" not in r["bins/weight"].attrs
    assert set(r["bins"].keys()) == {"chrom", "start", "end", "weight", "KR"}


def test_store_into_mcool_level(tmp_path):
    p = _copy("toy.symm.upper.2.mcool", tmp_path)
    clr = H5Cooler(p, "/resolutions/4")
    n = clr.info["nbins"]
    with clr.open("r+") as grp:
        grp["bins"].create_dataset("weight", data=np.ones(n))
    check(p)
    r = h5lite.File(p)
    assert "weight" in r["resolutions/4/bins"] and "weight" not in r["resolutions/2/bins"]
    assert sorted(r["resolutions"].keys()) == ["16", "2", "32", "4", "8"]
    assert dict(r.attrs) == {"format": "HDF5::MCOOL", "format-version": 2}


def test_many_links_split_nodes(tmp_path):
    """More links than one SNOD (2 x leaf K = 8 in libhdf5 files) holds."""
    p = _copy("toy.asymm.2.cool", tmp_path)
    with h5write.File(p, "r+") as f:
        for i in range(40):
            f["bins"].create_dataset(f"col{i:02d}", data=np.full(32, i, dtype=np.int16))
    check(p)
    r = h5lite.File(p)
    assert len(r["bins"].keys()) == 43
    assert np.array_equal(r["bins/col17"][:], np.full(32, 17))


def test_rejects_latest_format_and_bad_modes(tmp_path):
    with pytest.raises(FileNotFoundError):
        h5write.File(str(tmp_path / "missing.h5"), "r+")
    with pytest.raises(ValueError):
        h5write.File(str(tmp_path / "x.h5"), "r")
    p = str(tmp_path / "y.h5")
    with h5write.File(p, "w") as f:
        f.create_group("a")
        with pytest.raises(ValueError):
            f.create_group("a")
        f["a"].create_dataset("d", data=np.arange(3))
        with pytest.raises(ValueError):
            f["a"].create_dataset("d", data=np.arange(3))
        with pytest.raises(KeyError):
            del f["a"]["nope"]
        with pytest.raises(h5write.H5FormatError):
            f["a"].create_dataset("m", data=np.zeros((2, 2)))


def _gm_frames():
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "gm12878_2mb.npz"), allow_pickle=True)
    return d


def test_create_cooler_matches_source(tmp_path):
    """create_cooler(file) from the pixel table of the reference's GM12878 file reproduces it."""
    src = H5Cooler(os.path.join(COOL, "hg19.GM12878-MboI.matrix.2000kb.cool"))
    with src.open("r") as g:
        b1, b2, cnt = g["pixels/bin1_id"][:], g["pixels/bin2_id"][:], g["pixels/count"][:]
    bins = src.bins()[:]
    names = src.chromnames
    bins["chrom"] = pd.Categorical.from_codes(np.asarray(src._load_dset("bins/chrom")), categories=names, ordered=True)
    out = str(tmp_path / "out.cool")
    chunks = ({"bin1_id": b1[i:i + 5000], "bin2_id": b2[i:i + 5000], "count": cnt[i:i + 5000]}
              for i in range(0, len(b1), 5000))
    nnz, total = create.create_cooler(out, bins, chunks)
    assert nnz == 38156 and total == cnt.sum()
    check(out)
    dst = H5Cooler(out)
    for k in ("pixels/bin1_id", "pixels/bin2_id", "pixels/count", "bins/chrom", "bins/start", "bins/end",
              "indexes/chrom_offset", "indexes/bin1_offset", "chroms/length"):
        a, b = src._load_dset(k), dst._load_dset(k)
        assert np.array_equal(a, b), k
        # (the 2016 reference file keeps 32-bit indexes; the current schema, create/_constants.py:8-17, is int64)
        assert b.dtype == (np.int64 if k.startswith(("indexes", "pixels/bin")) else np.int32), k
    assert dst.chromnames == names
    info = dst.info
    assert info["nnz"] == 38156 and info["nbins"] == 1561 and info["nchroms"] == 25
    assert info["bin-size"] == 2000000 and info["bin-type"] == "fixed" and info["sum"] == int(cnt.sum())
    assert info["storage-mode"] == "symmetric-upper" and info["format"] == "HDF5::Cooler"
    assert info["format-version"] == 3 and info["metadata"] == "{}"
    assert dst.binsize == 2000000
    assert h5lite.File(out)["bins/chrom"].enum == {n: i for i, n in enumerate(names)}
    # append a second cooler to the same file under a group, then overwrite it
    create.create_cooler(out + "::resolutions/x", bins, {"bin1_id": b1, "bin2_id": b2, "count": cnt}, mode="a")
    create.create_cooler(out + "::resolutions/x", bins, {"bin1_id": b1[:10], "bin2_id": b2[:10], "count": cnt[:10]},
                         mode="a")
    check(out)
    assert H5Cooler(out, "/resolutions/x").info["nnz"] == 10 and H5Cooler(out).info["nnz"] == 38156


def test_create_cooler_validation(tmp_path):
    bins = pd.DataFrame({"chrom": ["a", "a", "b"], "start": [0, 10, 0], "end": [10, 15, 10]})
    px = {"bin1_id": np.array([0, 1, 2]), "bin2_id": np.array([1, 1, 2]), "count": np.array([1, 2, 3], np.int32)}
    out = str(tmp_path / "v.cool")
    create.create_cooler(out, bins, px)
    c = H5Cooler(out)
    assert c.info["bin-size"] == 10 and list(c.chromsizes.values) == [15, 10] and c.info["sum"] == 6
    with pytest.raises(create.BadInputError, match="exceeds the declared number of bins"):
        create.create_cooler(out, bins, dict(px, bin2_id=np.array([1, 1, 3])))
    with pytest.raises(create.BadInputError, match="bin ID < 0"):
        create.create_cooler(out, bins, dict(px, bin1_id=np.array([-1, 1, 2])))
    with pytest.raises(create.BadInputError, match="bin1_id greater than bin2_id"):
        create.create_cooler(out, bins, dict(px, bin1_id=np.array([0, 2, 2])))
    with pytest.raises(create.BadInputError, match="duplicate pixels"):
        create.create_cooler(out, bins, dict(px, bin1_id=np.array([0, 1, 1]), bin2_id=np.array([1, 1, 1])))
    with pytest.raises(ValueError, match="Missing column from bin table"):
        create.create_cooler(out, bins[["chrom", "start"]], px)
    with pytest.raises(ValueError, match="column not found in input"):
        create.create_cooler(out, bins, {"bin1_id": px["bin1_id"], "bin2_id": px["bin2_id"]})
    # square storage skips the triangularity check; unsorted chunks need ensure_sorted
    create.create_cooler(out, bins, dict(px, bin1_id=np.array([0, 2, 2]), bin2_id=np.array([1, 0, 2])),
                         symmetric_upper=False)
    assert H5Cooler(out).storage_mode == "square"
    with pytest.raises(create.BadInputError, match="sorted"):
        create.create_cooler(out, bins, dict(px, bin1_id=np.array([2, 0, 1]), bin2_id=np.array([2, 1, 1])))
    create.create_cooler(out, bins, dict(px, bin1_id=np.array([2, 0, 1]), bin2_id=np.array([2, 1, 1])),
                         ensure_sorted=True)
    assert list(H5Cooler(out)._load_dset("pixels/bin1_id")) == [0, 1, 2]
    # variable-width bins -> bin-size "null" / bin-type "variable"
    vb = pd.DataFrame({"chrom": ["a", "a", "a"], "start": [0, 10, 30], "end": [10, 30, 35]})
    create.create_cooler(out, vb, px)
    assert H5Cooler(out).binsize is None and H5Cooler(out).info["bin-type"] == "variable"


def test_copy_cooler(tmp_path):
    src = H5Cooler(os.path.join(COOL, "toy.symm.upper.2.mcool"), "/resolutions/2")
    out = str(tmp_path / "c.mcool")
    create.copy_cooler(src, out + "::resolutions/2", mode="w")
    check(out)
    dst = H5Cooler(out, "/resolutions/2")
    for k in ("pixels/bin1_id", "pixels/bin2_id", "pixels/count", "bins/chrom", "bins/start", "bins/end",
              "indexes/chrom_offset", "indexes/bin1_offset", "chroms/length", "chroms/name"):
        assert np.array_equal(src._load_dset(k), dst._load_dset(k)), k
    a, b = src.info, dst.info
    assert a == b


def test_zoomify_file_plumbing_with_oracle_coarsener(tmp_path, monkeypatch):
    """zoomify_cooler(outfile=...) / coarsen_cooler(output_uri=...): base copy, level chaining through
    the file, MCOOL attributes -- with the numpy oracle standing in for the device coarsener
    (the GPU run of the same flow is tests/test_cli_gpu.py::test_coarsen_and_zoomify_cli_write_files)."""
    from cooler_b200 import reduce, store
    from oracle import coarsen_numpy

    from . import _golden

    class OracleCoarsener(reduce.CoolerCoarsener):
        def __init__(self, source, factor, chunksize, columns=None, agg=None, batchsize=1, map=map,
                     device="cuda", table=None, **kw):
            super().__init__(source, factor, chunksize, columns=columns, agg=agg, table=object())

        def __iter__(self):
            clr = self.clr
            px = {k: np.asarray(clr._load_dset("pixels/" + k)) for k in ("bin1_id", "bin2_id", "count")}
            b1, b2, cnt, _ = coarsen_numpy.coarsen_pixels(
                px, np.asarray(clr._load_dset("bins/chrom")), np.asarray(clr._load_dset("bins/start")),
                np.asarray(clr._load_dset("bins/end")), clr.chromsizes.values, self.factor)
            for lo in range(0, len(b1), 7):
                yield {"bin1_id": b1[lo:lo + 7], "bin2_id": b2[lo:lo + 7], "count": cnt[lo:lo + 7]}

    monkeypatch.setattr(reduce, "CoolerCoarsener", OracleCoarsener)
    CO = _golden.coarsen_golden()

    def table(clr):
        return np.stack([np.asarray(clr._load_dset("pixels/" + k)).astype(np.int64)
                         for k in ("bin1_id", "bin2_id", "count")], 1)

    src = os.path.join(COOL, "toy.asymm.2.cool")
    out = str(tmp_path / "toy.4.cool")
    assert reduce.coarsen_cooler(src, out, 2, chunksize=10) is None
    check(out)
    c = store.open_cooler(out)
    np.testing.assert_array_equal(table(c), CO["toy.asymm/4"])
    assert c.binsize == 4 and c.storage_mode == "square"
    mc = str(tmp_path / "toy.mcool")
    assert reduce.zoomify_cooler(src, mc, [4, 8, 16, 32], chunksize=10) is None
    check(mc)
    assert sorted(store.list_coolers(mc)) == sorted(f"/resolutions/{x}" for x in (2, 4, 8, 16, 32))
    assert dict(h5lite.File(mc).attrs) == {"format": "HDF5::MCOOL", "format-version": 2}
    np.testing.assert_array_equal(table(store.open_cooler(mc + "::resolutions/2")), table(store.open_cooler(src)))
    for res in (4, 8, 16, 32):
        lvl = store.open_cooler(f"{mc}::resolutions/{res}")
        np.testing.assert_array_equal(table(lvl), CO[f"toy.asymm/{res}"])
        # one bin per chromosome at 32: util.get_binsize (util.py:393-411) cannot infer a width -> "variable"
        assert lvl.binsize == (res if res < 32 else None) and lvl.info["nnz"] == len(CO[f"toy.asymm/{res}"])
    # the reference's own symmetric-upper .mcool: every re-derived level equals the stored one
    m = os.path.join(COOL, "toy.symm.upper.2.mcool")
    mc2 = str(tmp_path / "symm.mcool")
    reduce.zoomify_cooler(m + "::resolutions/2", mc2, [4, 8, 16, 32], chunksize=1000)
    check(mc2)
    for res in (2, 4, 8, 16, 32):
        a, b = store.open_cooler(f"{m}::resolutions/{res}"), store.open_cooler(f"{mc2}::resolutions/{res}")
        np.testing.assert_array_equal(table(a), table(b))
        for k in ("bins/chrom", "bins/start", "bins/end", "indexes/bin1_offset", "indexes/chrom_offset"):
            np.testing.assert_array_equal(np.asarray(a._load_dset(k)), np.asarray(b._load_dset(k)))
        assert b.storage_mode == "symmetric-upper" and b.binsize == a.binsize


def test_legacy_zoomify_plumbing_with_oracle_coarsener(tmp_path, monkeypatch):
    """legacy_zoomify (_reduce.py:880-944): levels ::n .. ::0, root attributes; oracle as the coarsener."""
    from cooler_b200 import reduce, store
    from oracle import coarsen_numpy

    class OracleCoarsener(reduce.CoolerCoarsener):
        def __init__(self, source, factor, chunksize, columns=None, agg=None, batchsize=1, map=map,
                     device="cuda", table=None, **kw):
            super().__init__(source, factor, chunksize, columns=columns, agg=agg, table=object())

        def __iter__(self):
            clr = self.clr
            px = {k: np.asarray(clr._load_dset("pixels/" + k)) for k in ("bin1_id", "bin2_id", "count")}
            b1, b2, cnt, _ = coarsen_numpy.coarsen_pixels(
                px, np.asarray(clr._load_dset("bins/chrom")), np.asarray(clr._load_dset("bins/start")),
                np.asarray(clr._load_dset("bins/end")), clr.chromsizes.values, self.factor)
            yield {"bin1_id": b1, "bin2_id": b2, "count": cnt}

    monkeypatch.setattr(reduce, "CoolerCoarsener", OracleCoarsener)
    monkeypatch.setattr(reduce, "HIGLASS_TILE_DIM", 4)            # toy genome: 64 bp, 2 bp bins -> 3 zoom levels
    src = os.path.join(COOL, "toy.symm.upper.2.mcool") + "::resolutions/2"
    out = str(tmp_path / "legacy.mcool")
    n_zooms, levels = reduce.legacy_zoomify(src, out, 1, 1000)
    assert n_zooms == 3 and dict(levels) == {"3": 2, "2": 4, "1": 8, "0": 16}
    check(out)
    at = dict(h5lite.File(out).attrs)
    assert at == {"max-zoom": 3, "max-zooms": 3, "3": 2, "2": 4, "1": 8, "0": 16}
    assert sorted(store.list_coolers(out)) == ["/0", "/1", "/2", "/3"]
    m = os.path.join(COOL, "toy.symm.upper.2.mcool")
    for lvl, res in levels.items():
        a, b = store.open_cooler(f"{m}::resolutions/{res}"), store.open_cooler(f"{out}::{lvl}")
        for k in ("pixels/bin1_id", "pixels/bin2_id", "pixels/count", "bins/start", "bins/end"):
            np.testing.assert_array_equal(np.asarray(a._load_dset(k)), np.asarray(b._load_dset(k)))


def test_copy_cooler_is_a_raw_chunk_copy(tmp_path):
    """copy_cooler from a file: header messages and stored chunk bytes are carried over unchanged."""
    m = os.path.join(COOL, "toy.symm.upper.2.mcool")
    src = H5Cooler(m, "/resolutions/2")
    out = str(tmp_path / "raw.cool")
    create.copy_cooler(src, out, mode="w")
    check(out)
    a, b = h5lite.File(m)["resolutions/2"], h5lite.File(out)
    for k in ("bins/chrom", "pixels/bin1_id", "pixels/count", "chroms/name", "indexes/bin1_offset"):
        ma = {t: body for t, _f, body in a[k]._obj.msgs}
        mb = {t: body for t, _f, body in b[k]._obj.msgs}
        for t in (0x01, 0x03, 0x05, 0x0B):
            assert ma[t] == mb[t], (k, hex(t))
        ca, cb = a[k]._chunk_index()[1], b[k]._chunk_index()[1]
        assert [c[0] for c in ca] == [c[0] for c in cb] and [c[2:] for c in ca] == [c[2:] for c in cb]
        assert all(a._rd.bytes(x[1], x[2]) == b._rd.bytes(y[1], y[2]) for x, y in zip(ca, cb))
    assert b["bins/chrom"].enum == a["bins/chrom"].enum


def test_merge_file_plumbing_with_oracle_merger(tmp_path, monkeypatch):
    """merge_coolers(out.cool, [a.cool, b.cool]) / `cooler merge`: bins with chromosome NAMES, attributes,
    append mode -- the numpy oracle stands in for the device merger (GPU run: tests/test_cli_gpu.py)."""
    from click.testing import CliRunner

    from cooler_b200 import reduce, store
    from cooler_b200.cli import cli
    from oracle import coarsen_numpy

    def oracle_iter(self):
        pxs = [{k: np.asarray(c._load_dset("pixels/" + k)) for k in ("bin1_id", "bin2_id", "count")}
               for c in self.coolers]
        b1, b2, cnt = coarsen_numpy.merge_pixels(pxs)
        for lo in range(0, len(b1), self.mergebuf):
            hi = lo + self.mergebuf
            yield {"bin1_id": b1[lo:hi], "bin2_id": b2[lo:hi], "count": cnt[lo:hi]}

    monkeypatch.setattr(reduce.CoolerMerger, "__iter__", oracle_iter)
    a = os.path.join(COOL, "toy.symm.upper.2.mcool") + "::resolutions/4"
    out = str(tmp_path / "merged.cool")
    assert reduce.merge_coolers(out, [a, a, a], mergebuf=5) is None
    check(out)
    src, dst = store.open_cooler(a), store.open_cooler(out)
    for k in ("pixels/bin1_id", "pixels/bin2_id", "bins/chrom", "bins/start", "bins/end", "chroms/length",
              "indexes/bin1_offset", "indexes/chrom_offset"):
        np.testing.assert_array_equal(np.asarray(src._load_dset(k)), np.asarray(dst._load_dset(k)))
    np.testing.assert_array_equal(3 * np.asarray(src._load_dset("pixels/count")), np.asarray(dst._load_dset("pixels/count")))
    assert dst.chromnames == src.chromnames and dst.binsize == 4 and dst.storage_mode == "symmetric-upper"
    assert dst.info["sum"] == 3 * src.info["sum"] and dst.info["nnz"] == src.info["nnz"]
    assert h5lite.File(out)["bins/chrom"].enum == {n: i for i, n in enumerate(src.chromnames)}
    # CLI, appending a second cooler to the same file
    res = CliRunner().invoke(cli, ["merge", out + "::twice", a, a, "-c", "7", "--append"])
    assert res.exit_code == 0, (res.output, res.exception)
    check(out)
    assert sorted(store.list_coolers(out)) == ["/", "/twice"]
    tw = store.open_cooler(out + "::twice")
    np.testing.assert_array_equal(2 * np.asarray(src._load_dset("pixels/count")), np.asarray(tw._load_dset("pixels/count")))
    with pytest.raises(ValueError, match="same resolution"):
        reduce.merge_coolers(out, [a, os.path.join(COOL, "toy.symm.upper.2.mcool") + "::resolutions/8"], mergebuf=5)


def test_copy_cooler_from_memory_store(tmp_path):
    """copy_cooler from an in-memory store: columns are encoded, bins/chrom becomes an enum."""
    from cooler_b200 import MemCooler
    a = np.array([[0, 0, 3], [0, 2, 1], [1, 1, 5], [2, 3, 2]])
    clr = MemCooler.from_chromsizes(["x", "y"], [25, 12], 10, a[:, 0], a[:, 1], a[:, 2].astype(np.int32))
    out = str(tmp_path / "m.cool")
    create.copy_cooler(clr, out)
    check(out)
    dst = H5Cooler(out)
    assert dst.chromnames == ["x", "y"] and dst.binsize == 10 and dst.info["nnz"] == 4
    assert h5lite.File(out)["bins/chrom"].enum == {"x": 0, "y": 1}
    for k in ("pixels/bin1_id", "pixels/bin2_id", "pixels/count", "bins/start", "bins/end", "indexes/bin1_offset"):
        np.testing.assert_array_equal(np.asarray(clr._load_dset(k)), np.asarray(dst._load_dset(k)))


def test_merge_files_with_value_columns_and_aggregators(tmp_path, monkeypatch):
    """`cooler merge --field count --field score:dtype=float32,agg=mean` on real files: both columns are
    read, merged (oracle stand-in for the device merge; GPU run: tests/test_coarsen_gpu.py) and written with
    the requested dtypes."""
    import pandas as pd
    from click.testing import CliRunner

    from cooler_b200 import create, reduce, store
    from cooler_b200.cli import cli
    from oracle import coarsen_numpy

    def oracle_merge_columns(inputs, chrom_offset, columns, agg, device="cuda"):
        tables = []
        for off, b2, cols in inputs:
            b1 = np.repeat(np.arange(len(off) - 1, dtype=np.int64), np.diff(off))
            tables.append(dict(bin1_id=b1, bin2_id=b2, **cols))
        out = coarsen_numpy.merge_columns(tables, list(columns), agg)
        return out.pop("bin1_id"), out.pop("bin2_id"), out

    monkeypatch.setattr(reduce, "merge_columns", oracle_merge_columns)
    bins = pd.DataFrame({"chrom": ["a"] * 4 + ["b"] * 3, "start": [0, 10, 20, 30, 0, 10, 20],
                         "end": [10, 20, 30, 35, 10, 20, 22]})
    rng = np.random.default_rng(5)
    paths, tables = [], []
    for k in range(2):
        i, j = np.triu_indices(7)
        keep = rng.random(len(i)) < 0.6
        px = {"bin1_id": i[keep].astype(np.int64), "bin2_id": j[keep].astype(np.int64),
              "count": rng.integers(1, 50, keep.sum()).astype(np.int32), "score": rng.random(keep.sum())}
        p = str(tmp_path / f"in{k}.cool")
        create.create_cooler(p, bins, px, columns=["count", "score"], dtypes={"score": np.float64})
        check(p)
        paths.append(p)
        tables.append(px)
    out = str(tmp_path / "merged.cool")
    res = CliRunner().invoke(cli, ["merge", out, *paths, "--field", "count", "--field", "score:dtype=float32,agg=mean"])
    assert res.exit_code == 0, (res.output, res.exception)
    check(out)
    want = coarsen_numpy.merge_columns(tables, ["count", "score"], {"score": "mean"})
    got = store.open_cooler(out)
    np.testing.assert_array_equal(np.asarray(got._load_dset("pixels/bin1_id")), want["bin1_id"])
    np.testing.assert_array_equal(np.asarray(got._load_dset("pixels/bin2_id")), want["bin2_id"])
    np.testing.assert_array_equal(np.asarray(got._load_dset("pixels/count")), want["count"])
    score = np.asarray(got._load_dset("pixels/score"))
    assert score.dtype == np.float32
    np.testing.assert_allclose(score, want["score"].astype(np.float32), rtol=1e-6)
    assert got.info["sum"] == int(want["count"].sum()) and got.info["nnz"] == len(want["count"])
