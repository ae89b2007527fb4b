"""The built-in HDF5 subset reader (cooler_b200.h5lite) against real .cool files written by
h5py/libhdf5 (fixtures copied from the reference's tests/data) and against the text twins the
golden tables were made from ("equality with the .cool pixel table is presumed" --
with the reader it is checked)."""
import os

import numpy as np
import pytest

from tests import _golden

HERE = os.path.dirname(os.path.abspath(__file__))
COOL = os.path.join(HERE, "golden", "cool")
GM = os.path.join(COOL, "hg19.GM12878-MboI.matrix.2000kb.cool")


def test_reads_gm12878_cool_and_matches_text_twin():
    from cooler_b200 import h5lite
    f = h5lite.File(GM)
    assert sorted(f.keys()) == ["bins", "chroms", "indexes", "pixels"]
    a = f.attrs
    assert a["nbins"] == 1561 and a["nnz"] == 38156 and a["bin-size"] == 2000000 and a["bin-type"] == "fixed"
    t = _golden.table_of("gm")                       # made from the .coo.txt / .bed.gz twins
    np.testing.assert_array_equal(f["pixels/bin1_id"][:], t["bin1_id"])
    np.testing.assert_array_equal(f["pixels/bin2_id"][:], t["bin2_id"])
    np.testing.assert_array_equal(f["pixels/count"][:], t["count"])
    np.testing.assert_array_equal(f["bins/start"][:], t["bins_start"])
    np.testing.assert_array_equal(f["bins/end"][:], t["bins_end"])
    np.testing.assert_array_equal(f["bins/chrom"][:], t["bins_chrom"])
    np.testing.assert_array_equal(f["chroms/length"][:], t["chromsizes"])
    names = [x.decode() for x in f["chroms/name"][:]]
    assert names == [str(x) for x in t["chromnames"]]
    assert f["bins/chrom"].enum["chr1"] == 0 and f["bins/chrom"].enum["chrM"] == 24
    chrom_offset, bin1_offset = _golden.offsets(t)
    np.testing.assert_array_equal(f["indexes/chrom_offset"][:], chrom_offset)
    np.testing.assert_array_equal(f["indexes/bin1_offset"][:], bin1_offset)
    assert f["pixels/count"].dtype == np.int32 and f["pixels/bin2_id"].dtype == np.int64
    # partial reads decode only the chunks they need
    np.testing.assert_array_equal(f["pixels/bin2_id"][1000:1010], t["bin2_id"][1000:1010])
    np.testing.assert_array_equal(f["pixels/bin2_id"].astype(np.int32)[:], t["bin2_id"].astype(np.int32))
    assert "pixels/count" in f and "pixels/nope" not in f
    with pytest.raises(KeyError):
        f["nope"]


def test_h5cooler_protocol_and_mcool_uri():
    from cooler_b200 import store
    from cooler_b200.table import read_cooler_arrays
    clr = store.H5Cooler(GM)
    assert clr.binsize == 2000000 and clr.storage_mode == "symmetric-upper" and clr.shape == (1561, 1561)
    assert clr.info["nnz"] == 38156 and clr.chromnames[:2] == ["chr1", "chr2"]
    assert list(clr.bins()[:3].columns) == ["chrom", "start", "end"]
    assert list(clr.pixels()[:3].columns) == ["bin1_id", "bin2_id", "count"]
    off, b2, cnt, co = read_cooler_arrays(clr)
    t = _golden.table_of("gm")
    assert b2.dtype == np.int32 and np.array_equal(b2, t["bin2_id"]) and np.array_equal(cnt, t["count"])
    assert len(off) == 1562 and off[-1] == 38156 and co[-1] == 1561
    before = open(GM, "rb").read()
    with clr.open("r+") as g:                           # writable through h5write; untouched -> unchanged
        assert "weight" not in g["bins"]
    assert open(GM, "rb").read() == before
    m = os.path.join(COOL, "toy.symm.upper.2.mcool")
    assert store.list_coolers(m) == [f"/resolutions/{r}" for r in (16, 2, 32, 4, 8)]
    c4 = store.H5Cooler(m, "/resolutions/4")
    assert c4.binsize == 4 and c4.info["nbins"] == 16
    with pytest.raises(KeyError):
        store.H5Cooler(m, "/resolutions/3")
    with pytest.raises(KeyError):
        store.H5Cooler(m)                               # the root of an .mcool is not a cooler


def test_toy_files_match_their_bg2_twins():
    """odd.1.cool / toy.asymm.2.cool pixel tables equal the golden tables made from the text files."""
    from cooler_b200 import h5lite
    gold = _golden.coarsen_golden()
    for fn, key in (("odd.1.cool", "odd/1"), ("toy.asymm.2.cool", "toy.asymm/2")):
        f = h5lite.File(os.path.join(COOL, fn))
        got = np.stack([f["pixels/bin1_id"][:], f["pixels/bin2_id"][:], f["pixels/count"][:].astype(np.int64)], 1)
        np.testing.assert_array_equal(got, gold[key])
    g = h5lite.File(os.path.join(COOL, "toy.asymm.2.cool"))
    assert g.attrs["storage-mode"] == "square" and g.attrs["nnz"] == len(g["pixels/count"])


def test_not_hdf5(tmp_path):
    from cooler_b200 import h5lite
    p = tmp_path / "x.cool"
    p.write_bytes(b"not an hdf5 file" * 100)
    assert not h5lite.is_hdf5(str(p))
    with pytest.raises(h5lite.H5FormatError):
        h5lite.File(str(p))


def _written_cool(tmp_path, n_px=700_000, seed=0, float_counts=False):
    """A multi-chunk .cool written by the built-in writer: (path, bin1, bin2, count)."""
    import pandas as pd

    from cooler_b200 import create
    rng = np.random.default_rng(seed)
    nb = np.array([900, 600, 1, 300])
    n = int(nb.sum())
    i = rng.integers(0, n, n_px)
    j = rng.integers(0, n, n_px)
    i, j = np.minimum(i, j), np.maximum(i, j)
    key = np.unique(i.astype(np.int64) * n + j)
    b1, b2 = key // n, key % n
    cnt = rng.integers(1, 1000, len(b1)).astype(np.int32)
    if float_counts:
        cnt = cnt * rng.random(len(b1))
    names = ["c1", "c2", "c3", "c4"]
    bins = pd.DataFrame({"chrom": pd.Categorical.from_codes(np.repeat(np.arange(4), nb), categories=names, ordered=True),
                         "start": np.concatenate([np.arange(k, dtype=np.int64) * 1000 for k in nb])})
    bins["end"] = bins["start"] + 1000
    path = str(tmp_path / "w.cool")
    create.create_cooler(path, bins, [{"bin1_id": b1, "bin2_id": b2, "count": cnt}], dupcheck=False,
                         dtypes={"count": cnt.dtype})
    return path, b1, b2, cnt


@pytest.mark.parametrize("native", [True, False])
def test_threaded_decode_into_destinations(tmp_path, monkeypatch, native):
    """read_rows_into: the chunks of a column decoded by native host threads (cb200_decode_chunks) or, for
    pipelines it does not cover, by the pool of Python threads -- straight into caller buffers; bin ids
    narrowed to int32 from the low byte planes; unaligned row ranges."""
    from cooler_b200 import h5lite
    if not native:
        monkeypatch.setattr(h5lite, "_native_decode", lambda *a: False)
    path, b1, b2, cnt = _written_cool(tmp_path)
    f = h5lite.File(path)
    d2, dc = f["pixels/bin2_id"], f["pixels/count"]
    assert d2.is_chunked_1d() and d2.dtype == np.int64 and len(d2._chunk_index()[1]) > 4
    n = len(b2)
    for lo, hi in [(0, n), (12345, n - 777), (d2.chunk_rows() - 1, d2.chunk_rows() + 1), (5, 5)]:
        o2 = np.full(hi - lo, -7, np.int32)
        oc = np.full(hi - lo, -7, np.int32)
        o8 = np.full(hi - lo, -7, np.int64)
        h5lite.read_rows_into([(d2, lo, hi, o2), (dc, lo, hi, oc), (d2, lo, hi, o8)])
        np.testing.assert_array_equal(o2, b2[lo:hi].astype(np.int32))
        np.testing.assert_array_equal(o8, b2[lo:hi])
        np.testing.assert_array_equal(oc, cnt[lo:hi])
    np.testing.assert_array_equal(d2.astype(np.int32)[100:n - 100], b2[100:n - 100].astype(np.int32))
    assert d2.astype(np.int32)[:].dtype == np.int32
    # a value that does not fit the narrower type is an error, never a silent wrap
    small = np.zeros(n, np.int8)
    with pytest.raises(OverflowError):
        h5lite.read_rows_into([(d2, 0, n, small)])


@pytest.mark.parametrize("native", [True, False])
def test_threaded_decode_float_column(tmp_path, monkeypatch, native):
    from cooler_b200 import h5lite
    if not native:
        monkeypatch.setattr(h5lite, "_native_decode", lambda *a: False)
    path, b1, b2, val = _written_cool(tmp_path, n_px=200_000, float_counts=True)
    dc = h5lite.File(path)["pixels/count"]
    assert dc.dtype == np.float64
    out = np.empty(len(val))
    h5lite.read_rows_into([(dc, 0, len(val), out)])
    np.testing.assert_array_equal(out, val)


def test_native_decoder_reports_corrupt_chunks(tmp_path):
    """cb200_decode_chunks: a damaged deflate stream is an error (H5FormatError), never silent garbage."""
    from cooler_b200 import h5lite
    path, b1, b2, cnt = _written_cool(tmp_path, n_px=200_000)
    f = h5lite.File(path)
    dc = f["pixels/count"]
    _offs, addr, csize, _mask = dc._chunk_index()[1][0]
    del f, dc
    with open(path, "r+b") as fh:
        fh.seek(addr + csize // 2)
        fh.write(bytes(range(37, 37 + 64)))
    dc = h5lite.File(path)["pixels/count"]
    out = np.empty(len(cnt), np.int32)
    with pytest.raises((h5lite.H5FormatError, Exception)) as ei:
        h5lite.read_rows_into([(dc, 0, len(cnt), out)])
    assert "corrupt" in str(ei.value) or "decompress" in str(ei.value) or "Error" in type(ei.value).__name__
