"""Host-side logic that needs no GPU: resolution planners, bin maps, coarsened bin
tables, the in-memory store, prefilter arithmetic, the synthetic generator."""
import numpy as np
import pytest

from tests import _golden


def test_get_multiplier_sequence_matches_reference_semantics():
    from cooler_b200.reduce import ZOOMS_4DN, get_multiplier_sequence
    resn, pred, mult = get_multiplier_sequence(ZOOMS_4DN)
    assert list(resn) == ZOOMS_4DN
    assert list(mult[1:]) == [2, 5, 2, 5, 2, 2, 5, 2, 2, 5, 2, 2]       # SURVEY 3.3 probe
    assert pred[0] == -1 and list(pred[1:]) == [0, 0, 2, 2, 4, 5, 5, 7, 8, 8, 10, 11]
    resn, pred, mult = get_multiplier_sequence([4, 8, 16, 32], [2])
    assert list(resn) == [2, 4, 8, 16, 32] and list(mult[1:]) == [2, 2, 2, 2]
    with pytest.raises(ValueError):
        get_multiplier_sequence([4, 5, 32], [2])
    # two bases
    resn, pred, mult = get_multiplier_sequence([1000, 5000, 10000, 3000], [1000, 3000])
    assert list(resn) == [1000, 3000, 5000, 10000]
    assert list(pred) == [-1, 0, 0, 2]            # a base that another base divides still gets a predecessor (_reduce.py:481-489)


def test_preferred_sequence():
    from cooler_b200.reduce import preferred_sequence
    assert preferred_sequence(1, 100, "nice") == [1, 2, 5, 10, 20, 50, 100]
    assert preferred_sequence(5, 100, "nice") == [5, 10, 25, 50, 100]
    assert preferred_sequence(1000, 10_000_000, "nice") == [
        1000, 2000, 5000, 10000, 20000, 50000, 100000, 200000, 500000, 1000000, 2000000, 5000000, 10000000]
    assert preferred_sequence(3, 50, "binary") == [3, 6, 12, 24, 48]
    assert preferred_sequence(10, 5) == []


def test_get_quadtree_depth():
    from cooler_b200.reduce import get_quadtree_depth
    assert get_quadtree_depth([1000, 3000], 10, 256) == 1
    assert get_quadtree_depth([248956422, 242193529], 1000, 256) == 11


@pytest.mark.parametrize("factor", [2, 3, 5, 7, 200])
def test_bin_map_equals_coordinate_rebinning(factor):
    """Closed-form id map == the reference's coordinate arithmetic (oracle restates _reduce.py:597-614)."""
    from cooler_b200.reduce import coarsen_bin_map, coarsen_bins
    from oracle import coarsen_numpy
    sizes = np.array([51, 107, 33, 1, 400])
    bc, bs, be = _golden.uniform_bins(sizes, 1)
    off = np.r_[0, np.cumsum(sizes)]
    bmap, new_off = coarsen_bin_map(off, factor)
    n = off[-1]
    px = dict(bin1_id=np.arange(n), bin2_id=np.arange(n), count=np.ones(n, dtype=np.int32))
    b1, b2, c, (nc, ns, ne) = coarsen_numpy.coarsen_pixels(px, bc, bs, be, sizes, factor)
    # every old diagonal pixel i lands on (map[i], map[i]); counts = pooled bins
    want = np.bincount(bmap, minlength=new_off[-1])
    np.testing.assert_array_equal(b1, np.arange(new_off[-1]))
    np.testing.assert_array_equal(c, want)
    gc, gs, ge = coarsen_bins(bc, bs, be, sizes, factor)
    np.testing.assert_array_equal(gc, nc), np.testing.assert_array_equal(gs, ns), np.testing.assert_array_equal(ge, ne)


def test_coarsen_bins_gm12878_golden():
    from cooler_b200.reduce import coarsen_bins
    t = _golden.gm12878()
    CO = _golden.coarsen_golden()
    for f in (2, 5, 10):
        c, s, e = coarsen_bins(t["bins_chrom"], t["bins_start"], t["bins_end"], t["chromsizes"], f)
        np.testing.assert_array_equal(s, CO[f"gm/x{f}/bins_start"])
        np.testing.assert_array_equal(e, CO[f"gm/x{f}/bins_end"])
        np.testing.assert_array_equal(c, CO[f"gm/x{f}/bins_chrom"])
    c, s, e = coarsen_bins(t["bins_chrom"], t["bins_start"], t["bins_end"], t["chromsizes"], 2)
    assert (s[c == 0][-1], e[c == 0][-1]) == (248_000_000, 249_250_621)


def test_prune_edges_never_splits_rows():
    from cooler_b200.reduce import _prune_edges
    ind = np.array([0, 0, 5, 5, 9, 30, 31, 31, 40])
    e = _prune_edges(ind, 10)
    assert e[0] == 0 and e[-1] == 40 and set(e) <= set(ind) and np.all(np.diff(e) > 0)


def test_memcooler_protocol():
    from cooler_b200 import MemCooler
    t = _golden.gm12878()
    clr = MemCooler.from_arrays(t["chromnames"], t["chromsizes"], t["bins_chrom"], t["bins_start"],
                                t["bins_end"], t["bin1_id"], t["bin2_id"], t["count"], binsize=2_000_000)
    assert clr.info["nnz"] == 38156 and clr.info["nbins"] == 1561 and clr.info["sum"] == 100000
    assert clr.binsize == 2_000_000 and clr.storage_mode == "symmetric-upper"
    off = clr._load_dset("indexes/chrom_offset")
    assert list(off[:4]) == [0, 125, 247, 347] and off[-1] == 1561
    assert list(clr.chroms()["name"][:])[:2] == ["chr1", "chr2"]
    assert clr.chromsizes["chr1"] == 249250621
    assert clr.bins()[:].shape == (1561, 3)
    assert "count" in clr.pixels().dtypes
    with clr.open("r+") as grp:
        grp["bins"].create_dataset("weight", data=np.ones(1561), compression="gzip")
        grp["bins"]["weight"].attrs.update({"converged": True})
        assert "weight" in grp["bins"]
        del grp["bins"]["weight"]
        assert "weight" not in grp["bins"]


def test_prefilter_bias_matches_oracle_masks():
    """Host filter arithmetic on exact integer marginals == the reference's masks."""
    from types import SimpleNamespace

    from cooler_b200.balance import prefilter_bias
    t = _golden.gm12878()
    BAL, META = _golden.balance_golden()
    chrom_offset, _ = _golden.offsets(t)
    b1, b2, c = t["bin1_id"], t["bin2_id"], t["count"]
    n = 1561
    for key, nd in (("gm/defaults", 2), ("gm/diag0", 0), ("gm/test_balance_cfg", 1)):
        keep = np.abs(b1 - b2) >= nd
        nnzm = np.bincount(b1[keep], c[keep] != 0, n) + np.bincount(b2[keep], c[keep] != 0, n)
        summ = np.bincount(b1[keep], c[keep], n) + np.bincount(b2[keep], c[keep], n)
        prob = SimpleNamespace(table=SimpleNamespace(n_bins=n), marg_nnz=nnzm.astype(float), marg_sum=summ.astype(float))
        bias = prefilter_bias(prob, chrom_offset, mad_max=5, min_nnz=10, min_count=0, blacklist=None, x0=None)
        np.testing.assert_array_equal(bias == 0, np.isnan(BAL[key]))


def test_synth_generator_shards_tile_the_matrix():
    from cooler_b200 import synth
    plan = synth.SynthPlan(1_000_000, 1_200_000, seed=3)
    synth.ROWS_PER_BLOCK_TARGET, old = 150_000, synth.ROWS_PER_BLOCK_TARGET
    try:
        plan = synth.SynthPlan(1_000_000, 1_200_000, seed=3)
        full = synth.generate(plan, device="cpu", pin=False)
        assert len(plan.block_edges) > 4
        parts = [synth.generate(plan, rows=plan.shard_rows(r, 3), device="cpu", pin=False) for r in range(3)]
    finally:
        synth.ROWS_PER_BLOCK_TARGET = old
    assert sum(p["nnz"] for p in parts) == full["nnz"]
    np.testing.assert_array_equal(np.concatenate([p["bin2_id"].numpy() for p in parts]), full["bin2_id"].numpy())
    np.testing.assert_array_equal(sum(np.diff(p["bin1_offset"].numpy()) for p in parts),
                                  np.diff(full["bin1_offset"].numpy()))
    assert parts[0]["row_range"][1] == parts[1]["row_range"][0]
    off = full["bin1_offset"].numpy()
    b1 = np.repeat(np.arange(full["n_bins"]), np.diff(off))
    key = b1 * full["n_bins"] + full["bin2_id"].numpy()
    assert np.all(np.diff(key) > 0) and np.all(full["bin2_id"].numpy() >= b1)
    assert 0.7 < full["nnz"] / 1_200_000 < 1.3


def test_align_rows_to_factor():
    from cooler_b200.reduce import align_rows_to_factor, coarsen_bin_map
    off = np.array([0, 51, 158, 191])
    for f in (2, 4, 7):
        bmap, _ = coarsen_bin_map(off, f)
        for row in range(0, 192):
            a = align_rows_to_factor(off, f, row)
            assert a <= row and (a == 191 or a == 0 or bmap[a] != bmap[a - 1])   # a starts a coarse bin
            assert row == 191 or bmap[a] == bmap[row]                             # and it is row's coarse bin


def test_choose_layout_policy():
    """Sweep layout by table shape (balance.choose_layout): whole copies while the bias vector fits
    L1/L2 well, column strips for large dense tables, 2-D blocks for huge sparse ones."""
    from cooler_b200.balance import BLOCKS_MIN_BINS, choose_layout
    assert choose_layout(30_895, 200_000_000) == (None, False)                 # 100 kb: whole copies
    assert choose_layout(308_839, 1_000_000_000) == (14, False)                # 10 kb: strips of 16384 bins
    assert choose_layout(247_077, 200_000_000) == (14, False)                  # 12.5 kb shard
    assert choose_layout(3_088_298, 380_000_000) == (None, "allblocks")        # 1 kb shard: blocked kernel
    assert choose_layout(BLOCKS_MIN_BINS - 1, 50_000_000) == (None, False)     # sparse but not huge: whole copies
    assert choose_layout(3_088_298, 380_000_000, strip_bins=0) == (None, False)
    assert choose_layout(3_088_298, 380_000_000, strip_bins=16384, band="blocks") == (14, "blocks")
    assert choose_layout(3_088_298, 380_000_000, strip_bins=16384) == (14, True)
    with pytest.raises(ValueError):
        choose_layout(100_000, 10, strip_bins=1000)


def test_shard_rows_by_pixels_tiles_the_table():
    """Multi-GPU split of balance_cooler(group=...): contiguous row ranges cut where the pixel
    count crosses g * nnz / world, tiling [0, n_bins) exactly."""
    from cooler_b200.table import shard_rows_by_pixels
    rng = np.random.default_rng(3)
    for n, world in [(1, 2), (50, 3), (1000, 8), (7, 8)]:
        counts = rng.integers(0, 40, n) * (rng.random(n) < 0.7)
        off = np.r_[0, np.cumsum(counts)].astype(np.int64)
        cuts = [shard_rows_by_pixels(off, g, world) for g in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(cuts[:-1], cuts[1:]))
        assert all(lo <= hi for lo, hi in cuts)
        if n >= 1000:
            px = np.array([off[hi] - off[lo] for lo, hi in cuts])
            assert px.max() - px.min() <= 2 * counts.max()
    # all pixels in one row: one rank owns everything, the others are empty
    off = np.array([0, 0, 90, 90, 90])
    cuts = [shard_rows_by_pixels(off, g, 4) for g in range(4)]
    assert sum(off[hi] - off[lo] for lo, hi in cuts) == 90 and cuts[-1][1] == 4


def test_read_cooler_arrays_row_slices_tile_the_pixel_table():
    from cooler_b200.store import MemCooler
    from cooler_b200.table import read_cooler_arrays, shard_rows_by_pixels
    t = _golden.table_of("gm")
    chrom_offset, bin1_offset = _golden.offsets(t)
    sizes = np.diff(chrom_offset) * 10
    n = int(chrom_offset[-1])
    b1 = np.repeat(np.arange(n, dtype=np.int64), np.diff(bin1_offset))
    clr = MemCooler.from_chromsizes([f"c{i}" for i in range(len(sizes))], sizes, 10, b1, t["bin2_id"], t["count"])
    off, b2, cnt, co = read_cooler_arrays(clr)
    np.testing.assert_array_equal(off, bin1_offset)
    parts2, partsc, total = [], [], 0
    for g in range(3):
        rows = shard_rows_by_pixels(off, g, 3)
        o, p2, pc, _ = read_cooler_arrays(clr, rows=rows)
        assert o[0] == 0 and o[-1] == len(p2) and len(o) == n + 1
        assert (np.diff(o)[: rows[0]] == 0).all() and (np.diff(o)[rows[1]:] == 0).all()
        np.testing.assert_array_equal(np.diff(o)[rows[0]:rows[1]], np.diff(off)[rows[0]:rows[1]])
        parts2.append(np.asarray(p2)), partsc.append(np.asarray(pc))
        total += len(p2)
    assert total == len(b2)
    np.testing.assert_array_equal(np.concatenate(parts2), np.asarray(b2))
    np.testing.assert_array_equal(np.concatenate(partsc), np.asarray(cnt))


def test_count_columns_are_range_checked_before_narrowing():
    """int64 / uint32 count columns are narrowed to int32 only when every value fits (a silent wrap
    would corrupt the marginals); floating-point columns are refused."""
    import pytest
    import torch
    from cooler_b200.table import _as_tensor, counts_as_int32
    ok = counts_as_int32(_as_tensor(np.array([0, 5, 2**31 - 1], dtype=np.int64)))
    assert ok.dtype == torch.int32 and ok.tolist() == [0, 5, 2**31 - 1]
    assert counts_as_int32(_as_tensor(np.array([7, 9], dtype=np.uint32))).tolist() == [7, 9]
    assert counts_as_int32(_as_tensor(np.array([7, 9], dtype=np.uint16))).tolist() == [7, 9]
    for bad in (np.array([1, 2**31], dtype=np.int64), np.array([3, 2**32 - 1], dtype=np.uint32),
                np.array([-2**31 - 1], dtype=np.int64)):
        with pytest.raises(OverflowError):
            counts_as_int32(_as_tensor(bad))
    with pytest.raises(NotImplementedError):
        counts_as_int32(_as_tensor(np.array([1.5])))


@pytest.mark.parametrize("key", ["sq21/cis", "sq21/cis_diag1", "sq22/cis", "sq22/cis_diag1"])
def test_poisoned_chromosomes_of_square_tables_match_reference(key):
    """cis_only on square-storage tables: the chromosomes the reference leaves NaN after max_iters
    passes (NaN bias of an earlier chromosome times a zeroed trans pixel, _balance.py:57-60,150-151,186)
    are exactly those ``poisoned_chromosomes`` derives from the chromosome-link matrices; the matrices are
    restated here in numpy (the product computes them with cb200_cis_touch)."""
    from cooler_b200.balance import poisoned_chromosomes
    from oracle import ice_numpy
    bal, meta = _golden.balance_golden()
    src, _ = key.split("/")
    t = _golden.table_of(src)
    chrom_offset, _ = _golden.offsets(t)
    n_bins, n_chroms = int(chrom_offset[-1]), len(chrom_offset) - 1
    kw = meta[key]["kwargs"]
    px = {k: t[k] for k in ("bin1_id", "bin2_id", "count")}
    chrom_ids = np.repeat(np.arange(n_chroms, dtype=np.int32), np.diff(chrom_offset))
    spans = [(0, len(px["count"]))]
    base = dict(zero_trans=True, n_diags=kw.get("ignore_diags", 2), zero_cis=False)
    marg_nnz = ice_numpy._marginal(px, chrom_ids, n_bins, spans, map, binarize=True, vec=None, **base)
    marg = ice_numpy._marginal(px, chrom_ids, n_bins, spans, map, binarize=False, vec=None, **base)
    bias0 = ice_numpy.initial_bias(n_bins, marg_nnz, marg, chrom_offset)
    first = ice_numpy._marginal(px, chrom_ids, n_bins, spans, map, binarize=False, vec=bias0, **base)
    seg_empty = np.array([not np.any(first[lo:hi]) for lo, hi in zip(chrom_offset[:-1], chrom_offset[1:])])
    c1, c2 = chrom_ids[t["bin1_id"]], chrom_ids[t["bin2_id"]]
    early = c2 < c1
    any_ = np.zeros((n_chroms, n_chroms), bool)
    masked = np.zeros_like(any_)
    any_[c1[early], c2[early]] = True
    m = early & (bias0[t["bin2_id"]] == 0)
    masked[c1[m], c2[m]] = True
    got = poisoned_chromosomes((any_, masked), seg_empty)
    scale = np.array([np.nan if x is None else x for x in meta[key]["stats"]["scale"]], dtype=float)
    want = np.isnan(scale) & (np.array(meta[key]["passes"]) == 200)
    assert want.any() and not want.all()
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(seg_empty & ~got, np.array(meta[key]["passes"]) == 0)


def test_general_merge_host_logic_with_oracle_stand_in(monkeypatch):
    """CoolerMerger / merge_coolers with value columns and aggregators: column checks, chunking by
    ``mergebuf`` without splitting a row, column order of ``agg``, dtype promotion -- the numpy oracle
    stands in for the device merge (GPU run: tests/test_coarsen_gpu.py)."""
    from cooler_b200 import MemCooler, reduce
    from oracle import coarsen_numpy
    from tests import _merge_cases

    def oracle_merge_columns(inputs, chrom_offset, columns, agg, device="cuda"):
        tables = []
        for off, b2, cols in inputs:
            b1 = np.repeat(np.arange(len(off) - 1, dtype=np.int64), np.diff(off))
            tables.append(dict(bin1_id=b1, bin2_id=b2, **cols))
        out = coarsen_numpy.merge_columns(tables, list(columns), agg)
        return out.pop("bin1_id"), out.pop("bin2_id"), out

    monkeypatch.setattr(reduce, "merge_columns", oracle_merge_columns)
    gm = _golden.gm12878()
    clrs = []
    for sel, cols in _merge_cases.parts(gm):
        extra = {c: v for c, v in cols.items() if c != "count"}
        clrs.append(MemCooler.from_arrays([str(x) for x in gm["chromnames"]], gm["chromsizes"], gm["bins_chrom"],
                                          gm["bins_start"], gm["bins_end"], gm["bin1_id"][sel], gm["bin2_id"][sel],
                                          cols["count"], binsize=2_000_000, extra_pixel_columns=extra))
    G = _golden._npz("merge_agg_golden.npz")
    for key, (columns, agg) in _merge_cases.CASES.items():
        m = reduce.CoolerMerger(clrs, 5000, columns=columns, agg=agg)
        assert not m._fast
        chunks = list(m)
        assert len(chunks) > 3
        for a, b in zip(chunks[:-1], chunks[1:]):
            assert a["bin1_id"][-1] < b["bin1_id"][0]                  # a row is never split
        got = {k: np.concatenate([c[k] for c in chunks]) for k in chunks[0]}
        assert list(got) == [k.split("/", 1)[1] for k in G if k.startswith(key + "/")]
        for k, v in got.items():
            want = G[f"{key}/{k}"]
            assert v.dtype == want.dtype
            np.testing.assert_allclose(v, want, rtol=1e-12)
    assert reduce.CoolerMerger(clrs, 5000)._fast                       # plain integer count + sum
    with pytest.raises(ValueError, match="not found"):
        list(reduce.CoolerMerger(clrs, 5000, columns=["count", "nope"]))
    with pytest.raises(NotImplementedError):
        reduce.CoolerMerger(clrs, 5000, agg={"count": "median"})
    out = reduce.merge_coolers(None, clrs, mergebuf=10_000, columns=["count", "fcount"], dtypes={"fcount": np.float32})
    with out.open() as g:
        assert g["pixels/fcount"][:].dtype == np.float32 and g["pixels/count"][:].dtype == np.int32
        np.testing.assert_array_equal(g["pixels/count"][:], G["cols_sum/count"])


def test_host_planners_and_region_parser_agree_with_the_live_reference():
    """get_multiplier_sequence, preferred_sequence, the chunk-edge pruning, util.parse_region (2 200 region
    strings / tuples, malformed ones included), the coarsened bin table and the closed-form old -> new bin map
    (fixed and variable bins, random factors, against the live CoolerCoarsener), create's chunk validation, the
    general-merge oracle (random
    aggregators against the live CoolerMerger) and the matrix-fetch oracle (240 random dense / sparse windows,
    symmetric-upper and square storage, NaN weights, against the live Cooler.matrix) on seeded random inputs.  Runs in a subprocess; skipped where no copy of the reference exists."""
    import os
    import subprocess
    import sys

    from oracle import refshim
    if not refshim.available():
        pytest.skip("reference sources not present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tests", "_fuzz_host_vs_reference.py"), "200", "0"],
                         capture_output=True, text=True, timeout=900, cwd=root,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    assert res.stdout.strip().splitlines()[-1].startswith("ok:")


def test_balance_cli_options_agree_with_the_live_reference_cli():
    """180 `cooler balance` runs (--check / --name / mode flags / --ignore-diags / --ignore-dist / bin filters /
    BED blacklists with and without header, unknown names and out-of-bounds regions included / --tol /
    --max-iters / the four convergence policies / --stdout) on seeded random coolers through cooler_b200/cli.py
    (the numpy oracle standing in for the device balancer) and through the live unmodified reference CLI: same exit
    codes, same printed weights (1 350 more runs with other seeds were made once by hand).  Subprocess; skipped
    without the reference."""
    import os
    import subprocess
    import sys

    from oracle import refshim
    if not refshim.available():
        pytest.skip("reference sources not present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tests", "_fuzz_balance_cli_vs_reference.py"), "60", "0"],
                         capture_output=True, text=True, timeout=900, cwd=root,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    assert res.stdout.strip().splitlines()[-1].startswith("ok: 180 balance CLI runs agree")


def test_zoomify_and_coarsen_cli_arguments_agree_with_the_live_reference_cli():
    """600 `cooler zoomify` / `cooler coarsen` command lines (resolution strings incl. N / B / 4DN / <k>N
    progressions, --field specifications, -o, -k, -c) on random bin tables: both back ends replaced by recorders
    in cooler_b200/cli.py and in the live unmodified reference CLIs, the recorded calls compared (3 600 more
    command lines with other seeds were compared once by hand).  Subprocess; skipped without the reference."""
    import os
    import subprocess
    import sys

    from oracle import refshim
    if not refshim.available():
        pytest.skip("reference sources not present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tests", "_fuzz_zoomify_cli_vs_reference.py"), "100", "0"],
                         capture_output=True, text=True, timeout=900, cwd=root,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    assert res.stdout.strip().splitlines()[-1].startswith("ok: 600 zoomify / coarsen command lines")


def test_prefilter_bias_equals_the_oracle_on_random_marginals():
    """balance.prefilter_bias (the host half of the bin filters, _balance.py:366-413) against
    oracle.ice_numpy.initial_bias -- itself bit-identical to the live reference -- on 400 random marginal vectors
    with random chromosome layouts, filter settings, blacklists and warm starts: the bias vectors are identical bit
    for bit (NaN-free by construction), and the warm-start array is mutated in place as the reference does."""
    from types import SimpleNamespace

    from cooler_b200.balance import prefilter_bias
    from oracle import ice_numpy
    rng = np.random.default_rng(12)
    for case in range(400):
        nb = [int(rng.integers(1, 60)) for _ in range(int(rng.integers(1, 6)))]
        chrom_offset = np.r_[0, np.cumsum(nb)].astype(np.int64)
        n = int(chrom_offset[-1])
        marg_sum = np.floor(rng.lognormal(4, 1.5, n) * (rng.random(n) > 0.1))
        marg_nnz = np.minimum(marg_sum, np.floor(rng.uniform(0, 30, n)))
        if rng.random() < 0.1:
            lo, hi = chrom_offset[0], chrom_offset[1]
            marg_sum[lo:hi] = 0                               # an empty chromosome: median of nothing
            marg_nnz[lo:hi] = 0
        kw = dict(mad_max=int(rng.integers(0, 7)), min_nnz=int(rng.integers(0, 12)),
                  min_count=int(rng.integers(0, 2)) * int(rng.integers(0, 200)),
                  blacklist=(sorted({int(x) for x in rng.integers(0, n, 3)}) if rng.random() < 0.3 else None))
        x0 = None
        if rng.random() < 0.3:
            x0 = rng.lognormal(0, 0.3, n)
            x0[rng.random(n) < 0.1] = np.nan
        prob = SimpleNamespace(table=SimpleNamespace(n_bins=n, device=None), marg_nnz=marg_nnz.copy(), marg_sum=marg_sum.copy())
        mine_x0 = None if x0 is None else x0.copy()
        got = prefilter_bias(prob, chrom_offset, x0=mine_x0, **kw)
        want = ice_numpy.initial_bias(n, marg_nnz.copy(), marg_sum.copy(), chrom_offset,
                                      x0=None if x0 is None else x0.copy(), **kw)
        np.testing.assert_array_equal(got, want, err_msg=str((case, kw)))
        if x0 is not None:
            assert got is mine_x0                             # _balance.py:368-370: x0 is the bias vector
        np.testing.assert_array_equal(prob.marg_sum, marg_sum)   # the stored marginal itself is left alone


def test_balance_table_glue_agrees_with_the_live_reference_with_a_numpy_device():
    """300 balance_table runs on seeded random tables (symmetric-upper / square, integer / float counts, all three
    modes, random filters, warm starts, blacklists, iteration limits) with a numpy restatement of the device loop in
    place of the kernels: everything above the C ABI -- bin filters, mode dispatch, logging, empty-segment and
    iteration-limit branches, NaN conversion, rescaling, stats, warnings, x0 aliasing, the square-storage cascade --
    against the live unmodified reference (masks and pass counts exact, bias / scale 1e-9).  5 600 more runs with
    other seeds were made once by hand.  Subprocess; skipped without the reference."""
    import os
    import subprocess
    import sys

    from oracle import refshim
    if not refshim.available():
        pytest.skip("reference sources not present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tests", "_fuzz_glue_vs_reference.py"), "150", "0"],
                         capture_output=True, text=True, timeout=900, cwd=root,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    assert res.stdout.strip().splitlines()[-1].startswith("ok: 300 balance_table runs agree")


def test_reduce_glue_agrees_with_the_live_reference_with_numpy_device_functions():
    """CoolerCoarsener (fast and general value-column paths: chunking, dtypes, new bin tables), the in-memory
    zoomify_cooler chain and CoolerMerger on 60 seeded random tables (symmetric-upper / square; int16 / int32 / int64 /
    float32 / float64 counts; extra columns; random aggregators, factors and chunk sizes) with numpy restatements in
    place of the device functions: chunk streams equal the live unmodified reference's (ids and integers exact, dtypes
    equal, floats 1e-12).  1 200 more tables with other seeds were run once by hand.  Subprocess; skipped without
    the reference."""
    import os
    import subprocess
    import sys

    from oracle import refshim
    if not refshim.available():
        pytest.skip("reference sources not present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tests", "_fuzz_reduce_glue_vs_reference.py"), "60", "0"],
                         capture_output=True, text=True, timeout=900, cwd=root,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    assert res.stdout.strip().splitlines()[-1].startswith("ok: {'coarsen': 60")
