"""CLI behaviour (flags, exit codes, convergence policies) mirrored from the reference's
tests/test_cli_ops.py:37-231, run against in-memory stores."""
import numpy as np
import pytest
from click.testing import CliRunner

from tests import _golden

pytestmark = pytest.mark.gpu


@pytest.fixture()
def gm():
    from cooler_b200 import MemCooler, store
    t = _golden.gm12878()
    clr = MemCooler.from_arrays(t["chromnames"], t["chromsizes"], t["bins_chrom"], t["bins_start"],
                                t["bins_end"], t["bin1_id"], t["bin2_id"], t["count"], binsize=2_000_000)
    store.REGISTRY.clear()
    store.register("gm.cool", clr)
    return clr


def _toy():
    from cooler_b200 import MemCooler, store
    CO = _golden.coarsen_golden()
    sizes = np.array([32, 32])
    bc, bs, be = _golden.uniform_bins(sizes, 2)
    a = CO["toy.asymm/2"]
    clr = MemCooler.from_arrays(["chr1", "chr2"], sizes, bc, bs, be, a[:, 0], a[:, 1], a[:, 2].astype(np.int32),
                                binsize=2, symmetric_upper=False)
    store.register("toy.asymm.2.cool", clr)
    return clr, CO


def test_balance_cli_flags_and_exit_codes(gm, tmp_path):
    from cooler_b200.cli import cli
    BAL, META = _golden.balance_golden()
    r = CliRunner()
    res = r.invoke(cli, ["balance", "gm.cool", "--check"])
    assert res.exit_code == 1 and "No 'weight' column found" in res.output
    res = r.invoke(cli, ["balance", "gm.cool", "--cis-only", "--trans-only"])
    assert res.exit_code != 0
    res = r.invoke(cli, ["balance", "-p", "2", "gm.cool"])
    assert res.exit_code == 0
    with gm.open() as g:
        np.testing.assert_allclose(g["bins/weight"][:], BAL["gm/defaults"], rtol=1e-9, equal_nan=True)
        assert bool(g["bins/weight"].attrs["converged"])
    res = r.invoke(cli, ["balance", "gm.cool", "--check"])
    assert res.exit_code == 0 and "is balanced" in res.output
    res = r.invoke(cli, ["balance", "gm.cool"])                      # exists, no --force
    assert res.exit_code == 1
    res = r.invoke(cli, ["balance", "--force", "--cis-only", "--ignore-dist", "3000000", "gm.cool"])
    assert res.exit_code == 0
    with gm.open() as g:
        assert g["bins/weight"].attrs["cis_only"] and g["bins/weight"].attrs["ignore_diags"] == 2
    res = r.invoke(cli, ["balance", "--name", "w2", "--trans-only", "gm.cool"])
    assert res.exit_code == 0
    with gm.open() as g:
        np.testing.assert_allclose(g["bins/w2"][:], BAL["gm/trans_only"], rtol=1e-9, equal_nan=True)
    # blacklist BED == explicit bin list
    bed = tmp_path / "bad.bed"
    bed.write_text("chr1\t0\t2000000\nchr1\t8000000\t10000000\nchr2\t150000000\t156000000\nchrX\t96000000\t98000000\n")
    res = r.invoke(cli, ["balance", "--name", "w3", "--blacklist", str(bed), "gm.cool"])
    assert res.exit_code == 0
    with gm.open() as g:
        np.testing.assert_allclose(g["bins/w3"][:], BAL["gm/blacklist"], rtol=1e-9, equal_nan=True)


def test_balance_cli_convergence_policies(gm):
    from cooler_b200.cli import cli
    r = CliRunner()
    common = ["balance", "gm.cool", "--max-iters", "1", "--force"]
    res = r.invoke(cli, common + ["--convergence-policy", "store_final"])
    assert res.exit_code == 0
    with gm.open() as g:
        assert np.isfinite(g["bins/weight"][:]).any() and not bool(g["bins/weight"].attrs["converged"])
    res = r.invoke(cli, common + ["--convergence-policy", "store_nan"])
    assert res.exit_code == 0
    with gm.open() as g:
        assert np.isnan(g["bins/weight"][:]).all()
    res = r.invoke(cli, common + ["--convergence-policy", "discard"])
    assert res.exit_code == 0
    with gm.open() as g:
        assert "weight" not in g["bins"]
    res = r.invoke(cli, common + ["--convergence-policy", "error"])
    assert res.exit_code == 1
    res = r.invoke(cli, ["balance", "gm.cool", "--stdout"])
    assert res.exit_code == 0 and len(res.output.splitlines()) >= 1561 - 118


def test_coarsen_and_zoomify_cli():
    from cooler_b200 import store
    from cooler_b200.cli import cli
    clr, CO = _toy()
    r = CliRunner()
    res = r.invoke(cli, ["coarsen", "toy.asymm.2.cool", "-k", "2", "-o", "toy.4.cool", "-p", "2"])
    assert res.exit_code == 0, res.output
    out = store.open_cooler("toy.4.cool")
    with out.open() as g:
        got = np.stack([g["pixels/bin1_id"][:], g["pixels/bin2_id"][:], g["pixels/count"][:]], 1)
    np.testing.assert_array_equal(got, CO["toy.asymm/4"])
    res = r.invoke(cli, ["zoomify", "toy.asymm.2.cool", "-r", "4,8,16,32", "-o", "toy.mcool"])
    assert res.exit_code == 0, res.output
    for res_ in (4, 8, 16, 32):
        c = store.open_cooler(f"toy.mcool::resolutions/{res_}")
        with c.open() as g:
            got = np.stack([g["pixels/bin1_id"][:], g["pixels/bin2_id"][:], g["pixels/count"][:]], 1)
        np.testing.assert_array_equal(got, CO[f"toy.asymm/{res_}"])
    with pytest.raises(KeyError):
        store.open_cooler("toy.mcool::resolutions/5")
    # --balance on every level that lacks weights (cli/zoomify.py:20-46)
    res = r.invoke(cli, ["zoomify", "toy.asymm.2.cool", "-r", "4,8", "-o", "toy2.mcool", "--balance",
                         "--balance-args", "--ignore-diags 0 --min-nnz 0 --mad-max 0"])
    assert res.exit_code == 0, res.output
    for res_ in (4, 8):
        with store.open_cooler(f"toy2.mcool::resolutions/{res_}").open() as g:
            assert "weight" in g["bins"]


# ---- real .cool files through the built-in HDF5 reader (no h5py in the image) --------------
import os  # noqa: E402

COOL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cool")
GM_COOL = os.path.join(COOL, "hg19.GM12878-MboI.matrix.2000kb.cool")


def test_balance_bundled_gm12878_cool_file_matches_reference():
    """BASELINE config 1: `cooler balance` on the bundled GM12878 2 Mb .cool, defaults."""
    import cooler_b200
    from cooler_b200 import store
    BAL, META = _golden.balance_golden()
    clr = store.open_cooler(GM_COOL)
    assert type(clr).__name__ in ("H5Cooler", "Cooler")
    bias, stats = cooler_b200.balance_cooler(clr)
    np.testing.assert_array_equal(np.isnan(bias), np.isnan(BAL["gm/defaults"]))
    np.testing.assert_allclose(bias, BAL["gm/defaults"], rtol=1e-9, atol=0, equal_nan=True)
    assert stats["converged"] and abs(stats["scale"] / 56.94014353567062 - 1) < 1e-9
    bias, stats = cooler_b200.balance_cooler(clr, cis_only=True)
    np.testing.assert_allclose(bias, BAL["gm/cis_only"], rtol=1e-9, atol=0, equal_nan=True)


def test_balance_cli_stdout_on_cool_file():
    from cooler_b200.cli import cli
    BAL, META = _golden.balance_golden()
    r = CliRunner()
    res = r.invoke(cli, ["balance", GM_COOL, "--check"])
    assert res.exit_code == 1 and "No 'weight' column found" in res.output
    try:
        r2 = CliRunner(mix_stderr=False)                  # click < 8.2: keep the log lines out of stdout
    except TypeError:
        r2 = CliRunner()
    res = r2.invoke(cli, ["balance", "--stdout", GM_COOL])
    assert res.exit_code == 0, res.output
    lines = [x for x in res.stdout.split("\n") if not x.startswith(("INFO", "ERROR", "WARNING"))]
    vals = np.array([float(x) if x.strip() else np.nan for x in lines[:1561]])
    want = BAL["gm/defaults"]
    np.testing.assert_array_equal(np.isnan(vals), np.isnan(want))
    np.testing.assert_allclose(vals, want, rtol=1e-5, equal_nan=True)      # printed with %g (6 significant digits)


def _copy_cool(name, tmp_path):
    import shutil
    dst = str(tmp_path / name)
    shutil.copy(os.path.join(COOL, name), dst)
    os.chmod(dst, 0o644)
    return dst


def test_balance_cli_stores_weight_in_cool_file(tmp_path):
    """BASELINE config 1 end to end: `cooler balance file.cool` writes bins/weight + stats into the file
    (cli/balance.py:283-289), then --check / the exists-guard / --force behave like the reference."""
    from cooler_b200 import store
    from cooler_b200.cli import cli
    from ._h5check import check
    BAL, META = _golden.balance_golden()
    path = _copy_cool("hg19.GM12878-MboI.matrix.2000kb.cool", tmp_path)
    r = CliRunner()
    res = r.invoke(cli, ["balance", path])
    assert res.exit_code == 0, (res.output, res.exception)
    check(path)
    clr = store.open_cooler(path)
    with clr.open("r") as g:
        w = g["bins/weight"][:]
        at = dict(g["bins/weight"].attrs)
        px = g["pixels/count"][:]
    np.testing.assert_array_equal(np.isnan(w), np.isnan(BAL["gm/defaults"]))
    np.testing.assert_allclose(w, BAL["gm/defaults"], rtol=1e-9, atol=0, equal_nan=True)
    assert at["converged"] is True and at["ignore_diags"] == 2 and at["mad_max"] == 5 and at["min_nnz"] == 10
    assert at["cis_only"] is False and at["divisive_weights"] is False
    assert abs(at["scale"] / 56.94014353567062 - 1) < 1e-9 and abs(at["var"] / 3.7801217049722795e-06 - 1) < 1e-6
    assert len(px) == 38156 and px.sum() == 100000
    res = r.invoke(cli, ["balance", path, "--check"])
    assert res.exit_code == 0 and "is balanced" in res.output
    res = r.invoke(cli, ["balance", path])
    assert res.exit_code == 1
    res = r.invoke(cli, ["balance", path, "--force", "--cis-only", "--name", "weight"])
    assert res.exit_code == 0, (res.output, res.exception)
    check(path)
    with store.open_cooler(path).open("r") as g:
        np.testing.assert_allclose(g["bins/weight"][:], BAL["gm/cis_only"], rtol=1e-9, atol=0, equal_nan=True)
        at = dict(g["bins/weight"].attrs)
    assert at["cis_only"] is True and len(at["scale"]) == 25 and len(at["converged"]) == 25
    # balance_cooler(store=True) through the API, second column
    import cooler_b200
    bias, stats = cooler_b200.balance_cooler(store.open_cooler(path), trans_only=True, store=True, store_name="wt")
    with store.open_cooler(path).open("r") as g:
        np.testing.assert_array_equal(g["bins/wt"][:], bias)
        assert sorted(g["bins"].keys()) == ["chrom", "end", "start", "weight", "wt"]


def test_coarsen_and_zoomify_cli_write_files(tmp_path):
    """`cooler coarsen -o out.cool` and `cooler zoomify -o out.mcool --balance` on real files."""
    from cooler_b200 import store
    from cooler_b200.cli import cli
    from ._h5check import check
    CO = _golden.coarsen_golden()
    src = _copy_cool("toy.asymm.2.cool", tmp_path)
    out = str(tmp_path / "toy.4.cool")
    r = CliRunner()
    res = r.invoke(cli, ["coarsen", src, "-k", "2", "-o", out])
    assert res.exit_code == 0, (res.output, res.exception)
    check(out)
    c = store.open_cooler(out)

    def table(clr):
        return np.stack([np.asarray(clr._load_dset("pixels/" + k)).astype(np.int64)
                         for k in ("bin1_id", "bin2_id", "count")], 1)

    np.testing.assert_array_equal(table(c), CO["toy.asymm/4"])
    assert c.binsize == 4 and c.storage_mode == "square" and c.info["nnz"] == len(CO["toy.asymm/4"])
    assert c.info["sum"] == CO["toy.asymm/4"][:, 2].sum()
    ind = np.asarray(c._load_dset("indexes/bin1_offset"))
    np.testing.assert_array_equal(ind, np.searchsorted(CO["toy.asymm/4"][:, 0], np.arange(len(ind))))
    assert c._load_dset("pixels/count").dtype == np.int32 and c._load_dset("pixels/bin1_id").dtype == np.int64
    # zoomify the same file into an .mcool, balancing every level that can converge
    mc = str(tmp_path / "toy.mcool")
    res = r.invoke(cli, ["zoomify", src, "-r", "4,8,16,32", "-o", mc, "--balance",
                         "--balance-args", "--mad-max 0 --min-nnz 0 --ignore-diags 0 --max-iters 500"])
    assert res.exit_code == 0, (res.output, res.exception)
    check(mc)
    assert sorted(store.list_coolers(mc)) == sorted(f"/resolutions/{x}" for x in (2, 4, 8, 16, 32))
    from cooler_b200 import h5lite
    assert dict(h5lite.File(mc).attrs) == {"format": "HDF5::MCOOL", "format-version": 2}
    np.testing.assert_array_equal(table(store.open_cooler(mc + "::resolutions/2")), table(store.open_cooler(src)))
    for res_ in (4, 8, 16, 32):
        lvl = store.open_cooler(f"{mc}::resolutions/{res_}")
        np.testing.assert_array_equal(table(lvl), CO[f"toy.asymm/{res_}"])
        assert lvl.binsize == (res_ if res_ < 32 else None)     # one bin per chromosome: util.get_binsize -> None
        with lvl.open("r") as g:
            assert "weight" in g["bins"]
    # the symmetric-upper .mcool of the reference: re-derive its levels and compare with the file's own
    m = os.path.join(COOL, "toy.symm.upper.2.mcool")
    mc2 = str(tmp_path / "toy.symm.mcool")
    res = r.invoke(cli, ["zoomify", m + "::resolutions/2", "-r", "4,8,16,32", "-o", mc2])
    assert res.exit_code == 0, (res.output, res.exception)
    check(mc2)
    for res_ in (2, 4, 8, 16, 32):
        a = store.open_cooler(f"{m}::resolutions/{res_}")
        b = store.open_cooler(f"{mc2}::resolutions/{res_}")
        np.testing.assert_array_equal(table(a), table(b))
        for k in ("bins/chrom", "bins/start", "bins/end", "indexes/bin1_offset", "indexes/chrom_offset"):
            np.testing.assert_array_equal(np.asarray(a._load_dset(k)), np.asarray(b._load_dset(k)))
        assert b.storage_mode == "symmetric-upper"


def test_coarsen_real_cool_files_match_reference_goldens():
    """CoolerCoarsener fed from .cool files: toy.asymm 2 -> 4, odd 1 -> 4, mcool 2 -> 4."""
    from cooler_b200 import store
    from cooler_b200.reduce import CoolerCoarsener
    CO = _golden.coarsen_golden()

    def run(clr, factor):
        chunks = list(CoolerCoarsener(clr, factor, chunksize=1000))
        return np.stack([np.concatenate([c[k] for c in chunks]).astype(np.int64)
                         for k in ("bin1_id", "bin2_id", "count")], 1)

    np.testing.assert_array_equal(run(store.open_cooler(os.path.join(COOL, "toy.asymm.2.cool")), 2), CO["toy.asymm/4"])
    np.testing.assert_array_equal(run(store.open_cooler(os.path.join(COOL, "odd.1.cool")), 4), CO["odd/4"])
    m = os.path.join(COOL, "toy.symm.upper.2.mcool")
    got = run(store.open_cooler(m + "::resolutions/2"), 2)
    np.testing.assert_array_equal(got, CO["toy.symm.upper/4"])
    want4 = store.open_cooler(m + "::resolutions/4")
    with want4.open() as g:
        np.testing.assert_array_equal(got[:, 2], g["pixels/count"][:])


def test_zoomify_cli_legacy_layout(tmp_path):
    """`cooler zoomify --legacy`: integer-labeled levels ::3 (base) .. ::0 (_reduce.py:880-944)."""
    from cooler_b200 import h5lite, store
    from cooler_b200.cli import cli
    from ._h5check import check
    CO = _golden.coarsen_golden()
    src = _copy_cool("hg19.GM12878-MboI.matrix.2000kb.cool", tmp_path)
    out = str(tmp_path / "legacy.mcool")
    r = CliRunner()
    res = r.invoke(cli, ["zoomify", src, "--legacy", "-o", out])
    assert res.exit_code == 0, (res.output, res.exception)
    check(out)
    at = dict(h5lite.File(out).attrs)
    assert at == {"max-zoom": 3, "max-zooms": 3, "3": 2000000, "2": 4000000, "1": 8000000, "0": 16000000}
    assert sorted(store.list_coolers(out)) == ["/0", "/1", "/2", "/3"]
    lvl2 = store.open_cooler(out + "::2")
    for k in ("bin1_id", "bin2_id", "count"):
        np.testing.assert_array_equal(np.asarray(lvl2._load_dset("pixels/" + k)), CO[f"gm/x2/{k}"])
    np.testing.assert_array_equal(np.asarray(lvl2._load_dset("bins/start")), CO["gm/x2/bins_start"])
    np.testing.assert_array_equal(np.asarray(lvl2._load_dset("bins/end")), CO["gm/x2/bins_end"])
    sums = [int(np.asarray(store.open_cooler(f"{out}::{i}")._load_dset("pixels/count")).sum()) for i in range(4)]
    assert sums == [100000] * 4


def test_zoomify_and_coarsen_never_truncate_their_input(tmp_path):
    """`zoomify x.mcool::resolutions/2` without -o derives the output name x.mcool == the input;
    opening it with mode 'w' would destroy the (memory-mapped) source.  The reference is protected
    because h5py refuses to truncate an open file; here the call fails before anything is opened."""
    import hashlib
    import shutil

    from cooler_b200.cli import cli
    src = str(tmp_path / "toy.mcool")
    shutil.copy(os.path.join(COOL, "toy.symm.upper.2.mcool"), src)
    digest = hashlib.sha256(open(src, "rb").read()).hexdigest()
    r = CliRunner()
    res = r.invoke(cli, ["zoomify", src + "::resolutions/2", "-r", "4,8"])          # no -o
    assert res.exit_code != 0 and isinstance(res.exception, OSError), (res.output, res.exception)
    res = r.invoke(cli, ["zoomify", src + "::resolutions/2", "-r", "4,8", "-o", src])
    assert res.exit_code != 0 and isinstance(res.exception, OSError)
    res = r.invoke(cli, ["zoomify", src + "::resolutions/2", "--legacy", "-o", src])
    assert res.exit_code != 0 and isinstance(res.exception, OSError)
    one = _copy_cool("toy.asymm.2.cool", tmp_path)
    d1 = hashlib.sha256(open(one, "rb").read()).hexdigest()
    res = r.invoke(cli, ["coarsen", one, "-k", "2", "-o", one])
    assert res.exit_code != 0 and isinstance(res.exception, OSError)
    assert hashlib.sha256(open(src, "rb").read()).hexdigest() == digest
    assert hashlib.sha256(open(one, "rb").read()).hexdigest() == d1


def test_merge_cli_writes_file(tmp_path):
    """`cooler merge out.cool a b c` on real files (cli/merge.py:8-80)."""
    from cooler_b200 import h5lite, store
    from cooler_b200.cli import cli
    from ._h5check import check
    a = os.path.join(COOL, "toy.symm.upper.2.mcool") + "::resolutions/4"
    out = str(tmp_path / "merged.cool")
    r = CliRunner()
    res = r.invoke(cli, ["merge", out, a, a, a, "-c", "5"])
    assert res.exit_code == 0, (res.output, res.exception)
    check(out)
    src, dst = store.open_cooler(a), store.open_cooler(out)
    for k in ("pixels/bin1_id", "pixels/bin2_id", "bins/chrom", "bins/start", "bins/end", "chroms/length",
              "indexes/bin1_offset", "indexes/chrom_offset"):
        np.testing.assert_array_equal(np.asarray(src._load_dset(k)), np.asarray(dst._load_dset(k)))
    np.testing.assert_array_equal(3 * np.asarray(src._load_dset("pixels/count")),
                                  np.asarray(dst._load_dset("pixels/count")))
    assert dst.chromnames == src.chromnames and dst.binsize == 4 and dst.storage_mode == "symmetric-upper"
    assert dst.info["sum"] == 3 * src.info["sum"] and dst.info["nnz"] == src.info["nnz"]
    assert h5lite.File(out)["bins/chrom"].enum == {n: i for i, n in enumerate(src.chromnames)}
