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
    try:
        import h5py  # noqa: F401
    except ImportError:
        res = r.invoke(cli, ["balance", GM_COOL])                        # storing needs h5py
        assert res.exit_code != 0 and isinstance(res.exception, OSError)


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
