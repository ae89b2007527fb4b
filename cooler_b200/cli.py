"""``cooler balance | coarsen | zoomify`` command-line entry points on the B200 path.

Same flags, defaults, messages and exit codes as the reference's click commands
(open2c/cooler src/cooler/cli/balance.py, coarsen.py, zoomify.py); ``--nproc`` and
``--chunksize`` are accepted and ignored (no process pool: the GPU path must not
fork after CUDA initialisation).  COOL_PATH is resolved by
``cooler_b200.store.open_cooler``: a registered in-memory store, else the real
``cooler`` package (h5py) if installed, else the built-in HDF5 subset reader / writer
(``cooler_b200.h5lite`` / ``h5write``): ``balance`` stores ``bins/<name>`` + stats in the
file, ``coarsen -o out.cool`` and ``zoomify -o out.mcool [--balance]`` write real files.

    python -m cooler_b200.cli balance sample.cool --cis-only
"""
from __future__ import annotations

import logging
import os
import shlex
import sys
import warnings
from math import ceil

import click
import numpy as np

from . import store
from .balance import balance_cooler
from .reduce import coarsen_cooler, legacy_zoomify, merge_coolers, preferred_sequence, zoomify_cooler

HIGLASS_TILE_DIM = 256      # _reduce.py:28
logger = logging.getLogger("cooler")


@click.group()
def cli():
    """B200-native subset of the cooler CLI: balance, coarsen, zoomify, merge, cload pairs, load, dump.

    Files are written through the real `cooler` + h5py when they are installed; otherwise through the
    built-in HDF5 writer, which is EXPERIMENTAL (its output has not been opened by libhdf5)."""
    if not logger.handlers:
        h = logging.StreamHandler(sys.stderr)
        h.setFormatter(logging.Formatter("%(levelname)s:%(name)s:%(message)s"))
        logger.addHandler(h)
        logger.setLevel(logging.INFO)


def parse_field_param(arg):
    """'<name>[:dtype=<dtype>][,agg=<agg>]' (cli/_util.py:53-97, includes_colnum=False)."""
    parts = arg.split(":")
    if len(parts) > 2:
        raise click.BadParameter(arg)
    name, dtype, agg = parts[0], None, None
    if len(parts) == 2:
        for item in parts[1].split(","):
            try:
                prop, value = item.split("=")
            except ValueError as e:
                raise click.BadParameter(arg) from e
            if prop == "dtype":
                dtype = np.dtype(value)
            elif prop == "agg":
                agg = value
            else:
                raise click.BadParameter(f"Invalid property: '{prop}'.", param_hint=arg)
    return name, dtype, agg


def _fields(field):
    if len(field):
        specs = [parse_field_param(a) for a in field]
        columns = [s[0] for s in specs]
        dtypes = {s[0]: s[1] for s in specs if s[1] is not None}
        agg = {s[0]: s[2] for s in specs if s[2] is not None}
        return columns, dtypes, agg
    return ["count"], None, None


def _blacklist_bins(clr, path):
    """BED regions -> concatenated bin ids (cli/balance.py:211-232; util.bedslice)."""
    import csv

    import pandas as pd
    with open(path) as f:
        has_header = csv.Sniffer().has_header(f.read(1024))
    bad = pd.read_csv(path, sep="\t", header=0 if has_header else None, usecols=[0, 1, 2],
                      names=["chrom", "start", "end"], dtype={"chrom": str})
    bins = clr.bins()[:]
    names = clr.chromnames
    chrom = np.asarray(bins["chrom"])
    if chrom.dtype.kind in "iu":
        chrom = np.asarray(names, dtype=object)[chrom]
    else:
        chrom = chrom.astype(str)
    start, end = bins["start"].values, bins["end"].values
    from .api import parse_region
    chromsizes = clr.chromsizes
    out = []
    for _, reg in bad.iterrows():
        c, s, e = parse_region((reg.chrom, reg.start, reg.end), chromsizes)    # util.bedslice: unknown names and
        sel = np.flatnonzero((chrom == c) & (end > s) & (start < e))            # out-of-bounds regions are errors
        out.append(sel)
    return np.concatenate(out)                                     # no region at all: ValueError, as in the reference


def _dist_group():
    """(process group, rank) when launched by torchrun with several ranks (one per GPU, NCCL),
    else (None, 0).  Binds this process to GPU LOCAL_RANK."""
    import os
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return None, 0
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lg = logging.getLogger("cooler")
    if dist.get_rank() != 0:
        lg.setLevel(logging.ERROR)                     # one log stream: rank 0's
    return dist.group.WORLD, dist.get_rank()


def _on_rank0(group, rank, fn):
    """Run ``fn`` on rank 0 and hand its result (or its exception) to every rank."""
    if group is None:
        return fn()
    import torch.distributed as dist
    box = [None]
    if rank == 0:
        try:
            box[0] = ("ok", fn())
        except BaseException as e:  # noqa: BLE001 - re-raised on every rank
            box[0] = ("err", e)
    dist.broadcast_object_list(box, src=0, group=group)
    kind, val = box[0]
    if kind == "err":
        if rank == 0:
            raise val
        sys.exit(2)
    return val


_KEEP_DIST = 0     # > 0 while an outer command (zoomify --balance) still needs the process group


def _finish_dist(group):
    if group is not None:
        import torch.distributed as dist
        dist.barrier(group)
        if _KEEP_DIST == 0:
            dist.destroy_process_group()


@cli.command()
@click.argument("cool_uri", type=str, metavar="COOL_PATH")
@click.option("--cis-only", is_flag=True, default=False,
              help="Calculate weights against intra-chromosomal data only instead of genome-wide.")
@click.option("--trans-only", is_flag=True, default=False,
              help="Calculate weights against inter-chromosomal data only instead of genome-wide.")
@click.option("--ignore-diags", type=int, default=2, show_default=True,
              help="Number of diagonals of the contact matrix to ignore, including the main diagonal.")
@click.option("--ignore-dist", type=int,
              help="Distance from the diagonal in bp to ignore (max with --ignore-diags is used).")
@click.option("--mad-max", type=int, default=5, show_default=True, help="MAD-max bin filter.")
@click.option("--min-nnz", type=int, default=10, show_default=True,
              help="Ignore bins whose marginal number of nonzeros is less than this number.")
@click.option("--min-count", type=int, default=0, show_default=True,
              help="Ignore bins whose marginal count is less than this number.")
@click.option("--blacklist", type=click.Path(exists=True),
              help="3-column BED file of genomic regions to mask out.")
@click.option("--nproc", "-p", type=int, default=8, show_default=True,
              help="Accepted for compatibility; ignored (no process pool on the GPU path).")
@click.option("--chunksize", "-c", type=int, default=int(10e6), show_default=True,
              help="Accepted for compatibility; ignored (the table is resident in HBM).")
@click.option("--tol", type=float, default=1e-5, show_default=True,
              help="Threshold value of variance of the marginals for the algorithm to converge.")
@click.option("--max-iters", type=int, default=200, show_default=True, help="Iteration limit.")
@click.option("--name", type=str, default="weight", show_default=True, help="Name of column to write to.")
@click.option("--force", "-f", is_flag=True, default=False,
              help="Overwrite the target dataset, 'weight', if it already exists.")
@click.option("--check", is_flag=True, default=False,
              help="Check whether a data column 'weight' already exists.")
@click.option("--stdout", is_flag=True, default=False,
              help="Print weight column to stdout instead of saving to file.")
@click.option("--convergence-policy", default="store_final", show_default=True,
              type=click.Choice(["store_final", "store_nan", "discard", "error"]),
              help="What to do with weights when balancing doesn't converge in max_iters.")
def balance(cool_uri, nproc, chunksize, mad_max, min_nnz, min_count, blacklist, ignore_diags, tol,
            cis_only, trans_only, max_iters, name, force, check, stdout, convergence_policy,
            ignore_dist):
    """Matrix balancing on the GPU.  COOL_PATH : Path to a COOL file.

    Multi-GPU: launch one process per GPU with torchrun, e.g.
    ``torchrun --nproc-per-node 8 -m cooler_b200.cli balance FILE.cool``; every rank reads and keeps
    its own row range of the pixel table (where the reference forks ``--nproc`` workers,
    cli/balance.py:237-243), rank 0 checks, reports and writes."""
    group, rank = _dist_group()
    cool_path, group_path = store.parse_cooler_uri(cool_uri)

    def pre():
        """Checks and column clean-up (rank 0).  Returns an exit code or None to go on."""
        clr = store.open_cooler(cool_uri)
        if check:                                                 # cli/balance.py:180-188
            with clr.open("r") as grp:
                if name not in grp["bins"]:
                    click.echo(f"{cool_path}: No '{name}' column found.")
                    return 1
                click.echo(f"{cool_path}::{group_path} is balanced.")
                return 0
        if cis_only and trans_only:                               # 190-193
            raise click.UsageError("Provide at most one of --cis-only and --trans-only flags")
        # 195-206 (the reference opens "r+" here even with --stdout; read-only stores -- a .cool read
        # by the built-in HDF5 reader -- can still print their weights)
        with clr.open("r" if stdout else "r+") as grp:
            if name in grp["bins"] and not stdout:
                if not force:
                    print(f"'{name}' column already exists. Use --force option to overwrite.", file=sys.stderr)
                    return 1
                del grp["bins"][name]
        return None

    code = _on_rank0(group, rank, pre)
    if code is not None:
        _finish_dist(group)
        sys.exit(code)
    clr = store.open_cooler(cool_uri)

    if rank == 0:
        logger.info(f'Balancing "{cool_uri}"')
    bad_bins = _blacklist_bins(clr, blacklist) if blacklist is not None else None
    if ignore_dist is not None:                                   # 234-235
        ignore_diags = max(ignore_diags, int(np.ceil(ignore_dist / clr.binsize)))

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                           # the CLI reports through the logger
        bias, stats = balance_cooler(
            clr, chunksize=chunksize, cis_only=cis_only, trans_only=trans_only, tol=tol,
            min_nnz=min_nnz, min_count=min_count, blacklist=bad_bins, mad_max=mad_max,
            max_iters=max_iters, ignore_diags=ignore_diags, rescale_marginals=True, use_lock=False,
            group=group)

    if rank != 0:                                                 # every rank holds the same result
        _finish_dist(group)
        return
    if not np.all(stats["converged"]):                            # 265-277
        logger.error("Iteration limit reached without convergence")
        if convergence_policy == "store_final":
            logger.error("Storing final result. Check log to assess convergence.")
        elif convergence_policy == "store_nan":
            logger.error("Saving weights as NaN.")
            bias[:] = np.nan
        elif convergence_policy == "discard":
            logger.error("Discarding result and aborting.")
            _finish_dist(group)
            sys.exit(0)
        elif convergence_policy == "error":
            logger.error("Discarding result and aborting.")
            _finish_dist(group)
            sys.exit(1)

    if stdout:                                                    # 279-282
        import pandas as pd
        pd.Series(bias).to_string(sys.stdout, header=False, index=False, na_rep="", float_format="%g")
    else:                                                         # 283-289
        with clr.open("r+") as grp:
            h5opts = {"compression": "gzip", "compression_opts": 6}
            grp["bins"].create_dataset(name, data=bias, **h5opts)
            grp["bins"][name].attrs.update(stats)
    _finish_dist(group)


@cli.command()
@click.argument("cool_uri", metavar="COOL_PATH")
@click.option("--factor", "-k", type=int, default=2, show_default=True, help="Gridding factor.")
@click.option("--nproc", "-n", "-p", default=1, type=int, help="Accepted for compatibility; ignored.")
@click.option("--chunksize", "-c", type=int, default=int(10e6), show_default=True,
              help="Number of pixels per output chunk.")
@click.option("--field", type=str, multiple=True,
              help="Value columns to merge as '<name>[:dtype=<dtype>][,agg=<agg>]'.")
@click.option("--out", "-o", required=True, help="Output file or URI")
@click.option("--append", "-a", is_flag=True, default=False,
              help="Append the output cooler to an existing file instead of overwriting the file.")
def coarsen(cool_uri, factor, nproc, chunksize, field, out, append):
    """Coarsen a cooler to a lower resolution (k-by-k pooling per chromosomal block).

    Multi-GPU: ``torchrun --nproc-per-node N -m cooler_b200.cli coarsen FILE -k 2 -o OUT``: every
    rank coarsens its own bin1 row range, rank 0 writes (the reference's ``--nproc`` pool,
    cli/coarsen.py)."""
    columns, dtypes, agg = _fields(field)
    clr = store.open_cooler(cool_uri)
    if store.is_memory_uri(cool_uri):
        result = coarsen_cooler(clr, None, factor, chunksize=chunksize, nproc=nproc, columns=columns,
                                dtypes=dtypes, agg=agg)
        store.register(out, result)
    else:
        group, _ = _dist_group()
        coarsen_cooler(clr, out, factor, chunksize=chunksize, nproc=nproc, columns=columns,
                       dtypes=dtypes, agg=agg, mode="a" if append else "w", group=group)
        _finish_dist(group)


@cli.command()
@click.argument("out_path", type=click.Path(exists=False))
@click.argument("in_paths", nargs=-1, type=click.Path(exists=False))
@click.option("--chunksize", "-c", type=int, default=int(20e6), show_default=True,
              help="Size of the merge buffer in number of pixel table rows.")
@click.option("--field", type=str, multiple=True,
              help="Value columns to merge as '<name>[,dtype=<dtype>][,agg=<agg>]'.")
@click.option("--append", "-a", is_flag=True, default=False,
              help="Append the output cooler to an existing file instead of overwriting the file.")
def merge(out_path, in_paths, chunksize, field, append):
    """Merge multiple coolers with identical axes (cli/merge.py:8-80) on the GPU."""
    columns, dtypes, agg = _fields(field)
    clrs = [store.open_cooler(p) for p in in_paths]
    if in_paths and all(store.is_memory_uri(p) for p in in_paths):      # in-memory stores in, in-memory store out
        store.register(out_path, merge_coolers(None, clrs, mergebuf=chunksize, columns=columns,
                                               dtypes=dtypes, agg=agg))
    else:
        merge_coolers(out_path, clrs, mergebuf=chunksize, columns=columns, dtypes=dtypes, agg=agg,
                      mode="a" if append else "w")


def parse_bins(arg):
    """BINS argument: '<chromsizes file>:<binsize>' or a bins BED file (cli/_util.py:100-143).
    Returns (chromsizes Series in file order, bins DataFrame)."""
    import os.path as op

    import pandas as pd
    if ":" in op.splitdrive(arg)[1]:
        chromsizes_file, binsize = arg.rsplit(":", maxsplit=1)
        if not op.exists(chromsizes_file):
            raise ValueError(f'File "{chromsizes_file}" not found')
        try:
            binsize = int(binsize)
        except ValueError as e:
            raise ValueError(f'Expected integer binsize argument (bp), got "{binsize}"') from e
        tab = pd.read_csv(chromsizes_file, sep="\t", usecols=[0, 1], names=["name", "length"], dtype={"name": str})
        chromsizes = pd.Series(tab["length"].values, index=tab["name"].values)        # read_chromsizes(all_names=True)
        parts = []
        for name, clen in chromsizes.items():                                             # util.binnify
            nb = int(np.ceil(clen / binsize))
            edges = np.arange(0, nb + 1) * binsize
            edges[-1] = clen
            parts.append(pd.DataFrame({"chrom": [name] * nb, "start": edges[:-1], "end": edges[1:]}))
        bins = pd.concat(parts, axis=0, ignore_index=True)
        bins["chrom"] = pd.Categorical(bins["chrom"], categories=list(chromsizes.index), ordered=True)
    elif op.exists(arg):
        bins = pd.read_csv(arg, sep="\t", names=["chrom", "start", "end"], usecols=[0, 1, 2], dtype={"chrom": str})
        last = bins.drop_duplicates(["chrom"], keep="last")
        chromsizes = pd.Series(last["end"].values, index=last["chrom"].values)
    else:
        raise ValueError("Expected BINS to be either <Path to bins file> or "
                         "<Path to chromsizes file>:<binsize in bp>.")
    return chromsizes, bins


@cli.group()
def cload():
    """Create a cooler from genomic pairs and bins (GPU binning; `pairs` only)."""


@cload.command()
@click.argument("bins", metavar="BINS")
@click.argument("pairs_path", metavar="PAIRS_PATH")
@click.argument("cool_path", metavar="COOL_PATH")
@click.option("--metadata", type=str, help="Path to JSON file containing user metadata.")
@click.option("--assembly", type=str, help="Name of genome assembly (e.g. hg19, mm10)")
@click.option("--chrom1", "-c1", type=int, required=True, help="chrom1 field number (one-based)")
@click.option("--pos1", "-p1", type=int, required=True, help="pos1 field number (one-based)")
@click.option("--chrom2", "-c2", type=int, required=True, help="chrom2 field number (one-based)")
@click.option("--pos2", "-p2", type=int, required=True, help="pos2 field number (one-based)")
@click.option("--zero-based", "-0", is_flag=True, default=False, show_default=True,
              help="Positions are zero-based")
@click.option("--comment-char", type=str, default="#", show_default=True, help="Comment character that skips lines")
@click.option("--no-symmetric-upper", "-N", is_flag=True, default=False,
              help="Create a complete square matrix without implicit symmetry.")
@click.option("--input-copy-status", type=click.Choice(["unique", "duplex"]), default="unique", show_default=True,
              help="Copy status of input data when using symmetric-upper storage.")
@click.option("--field", type=str, multiple=True,
              help="Value columns next to the pair counts: '<name>=<field number>[:dtype=<dtype>][,agg=<agg>]'.")
@click.option("--chunksize", "-c", type=int, default=15_000_000, help="Lines of the input parsed at a time.")
@click.option("--mergebuf", type=int, help="Accepted for compatibility; ignored (one device-wide sort).")
@click.option("--max-merge", type=int, default=200, help="Accepted for compatibility; ignored.")
@click.option("--temp-dir", type=str, help="Accepted for compatibility; ignored (no temporary coolers).")
@click.option("--no-delete-temp", is_flag=True, default=False, help="Accepted for compatibility; ignored.")
@click.option("--storage-options", help="HDF5 filter options 'k1=v1,k2=v2,...'.")
@click.option("--append", "-a", is_flag=True, default=False,
              help="Append the output cooler to an existing file instead of overwriting the file.")
def pairs(bins, pairs_path, cool_path, metadata, assembly, chrom1, pos1, chrom2, pos2, zero_based, comment_char,
          no_symmetric_upper, input_copy_status, field, chunksize, mergebuf, max_merge, temp_dir, no_delete_temp,
          storage_options, append):
    """Bin any text file or stream of pairs (cli/cload.py:484-666).  PAIRS_PATH '-' reads stdin.

    The reference bins chunk by chunk into temporary coolers and merges them; here the parsed
    columns are binned and counted by ONE device-wide sort / reduce-by-key (cooler_b200.ingest)."""
    import json

    import pandas as pd

    from . import ingest
    from .create import create_cooler
    bins_spec = bins
    chromsizes, bins = parse_bins(bins)
    from .create import _binsize, _bins_arrays
    names, ids, start, end, _ = _bins_arrays(bins)
    binsize = _binsize(ids, start, end)
    if binsize is None and ":" in os.path.splitdrive(bins_spec)[1]:
        binsize = int(bins_spec.rsplit(":", maxsplit=1)[1])      # every chromosome shorter than one bin
    if len(field):
        return _cload_pairs_with_fields(bins, pairs_path, cool_path, metadata, assembly, chrom1, pos1, chrom2, pos2,
                                        zero_based, comment_char, no_symmetric_upper, input_copy_status, field,
                                        chunksize, mergebuf, storage_options, append)
    uniform = binsize is not None and np.array_equal(
        np.asarray(start), np.concatenate([np.arange(k) * binsize for k in np.bincount(ids)]))
    if not uniform:
        raise click.UsageError("the B200 ingest path needs fixed-size bins (<chromsizes>:<binsize>)")
    symmetric_upper = not no_symmetric_upper
    tril_action = None
    if symmetric_upper:                                          # 517-523
        tril_action = "reflect" if input_copy_status == "unique" else "drop"
    if metadata is not None:
        with open(metadata) as f:
            metadata = json.load(f)
    cols = {}
    for name, num in (("chrom1", chrom1), ("pos1", pos1), ("chrom2", chrom2), ("pos2", pos2)):
        if num == 0:
            raise click.BadParameter("Field numbers start at 1", param_hint=name)
        cols[name] = num - 1
    h5opts = None
    if storage_options is not None:
        h5opts = {}
        for kv in storage_options.split(","):
            k, v = kv.split("=", 1)
            h5opts[k.strip()] = int(v) if v.strip().lstrip("-").isdigit() else v.strip()
    order = sorted(cols, key=cols.get)                            # read_csv returns usecols in file order
    reader = pd.read_csv(sys.stdin if pairs_path == "-" else pairs_path, sep="\t", comment=comment_char or None,
                         header=None, usecols=sorted(cols.values()), names=order,
                         dtype={"chrom1": str, "pos1": np.int64, "chrom2": str, "pos2": np.int64},
                         iterator=True, chunksize=chunksize, compression="infer")
    idx = pd.Index(names)
    parts = {k: [] for k in cols}
    for chunk in reader:                                          # names -> ids; unknown contigs get -1 and are dropped
        parts["chrom1"].append(idx.get_indexer(chunk["chrom1"]).astype(np.int32))
        parts["chrom2"].append(idx.get_indexer(chunk["chrom2"]).astype(np.int32))
        parts["pos1"].append(chunk["pos1"].to_numpy(np.int64))
        parts["pos2"].append(chunk["pos2"].to_numpy(np.int64))
    cat = {k: (np.concatenate(v) if v else np.zeros(0, np.int32 if k.startswith("chrom") else np.int64))
           for k, v in parts.items()}
    lengths = np.asarray(end)[np.r_[ids[1:] != ids[:-1], True]]
    px = ingest.bin_pairs(cat["chrom1"], cat["pos1"], cat["chrom2"], cat["pos2"], lengths, binsize,
                          is_one_based=not zero_based, tril_action=tril_action, validate=True)
    create_cooler(cool_path, bins, px, columns=["count"], metadata=metadata, assembly=assembly,
                  boundscheck=False, triucheck=False, dupcheck=False, ensure_sorted=False,
                  symmetric_upper=symmetric_upper, h5opts=h5opts, mode="a" if append else "w")


def _cload_pairs_with_fields(bins, pairs_path, cool_path, metadata, assembly, chrom1, pos1, chrom2, pos2, zero_based,
                             comment_char, no_symmetric_upper, input_copy_status, field, chunksize, mergebuf,
                             storage_options, append):
    """``cooler cload pairs --field <name>=<number>[:dtype=..,agg=..]`` (cli/cload.py:540-666): value columns
    next to the pair counts.  As in the reference every chunk of records is sanitized and aggregated per pixel
    with the requested functions (``aggregate_records``; the count is the group size unless a field is named
    ``count``), and the chunk tables are then merged by SUMMING every column (``create_from_unordered`` ->
    ``CoolerMerger``).  Both aggregations run on the device (``ingest.UnorderedPixels``); the records are
    sanitized on the host (``ingest.sanitize_records``, the ``pairs`` preset)."""
    import json

    import pandas as pd

    from . import ingest
    from .create import _bins_arrays, create_from_unordered
    symmetric_upper = not no_symmetric_upper
    tril_action = None
    if symmetric_upper:
        tril_action = "reflect" if input_copy_status == "unique" else "drop"
    if metadata is not None:
        with open(metadata) as f:
            metadata = json.load(f)
    in_numbers = {}
    for name, num in (("chrom1", chrom1), ("pos1", pos1), ("chrom2", chrom2), ("pos2", pos2)):
        if num == 0:
            raise click.BadParameter("Field numbers start at 1", param_hint=name)
        in_numbers[name] = num - 1
    in_dtypes = {"chrom1": str, "pos1": np.int64, "chrom2": str, "pos2": np.int64}
    out_names, out_dtypes, aggregations = [], {}, {}
    for arg in field:                                             # cli/cload.py:546-585
        name, colnum, dtype, agg = _parse_input_field(arg, includes_agg=True)
        if colnum is None:
            if agg is None and dtype is not None and name in ("bin1_id", "bin2_id", "count"):
                out_dtypes[name] = dtype
                continue
            raise click.BadParameter("A field number is required.", param_hint=arg)
        if name not in out_names:
            out_names.append(name)
        in_numbers[name] = colnum
        if dtype is not None:
            in_dtypes[name] = out_dtypes[name] = dtype
        aggregations[name] = agg if agg is not None else "sum"
    if "count" not in out_names:
        out_names.append("count")
    h5opts = None
    if storage_options is not None:
        h5opts = {}
        for kv in storage_options.split(","):
            k, v = kv.split("=", 1)
            h5opts[k.strip()] = int(v) if v.strip().lstrip("-").isdigit() else v.strip()
    order = sorted(in_numbers, key=in_numbers.get)                # read_csv hands usecols back in file order
    reader = pd.read_csv(sys.stdin if pairs_path == "-" else pairs_path, sep="\t", comment=comment_char or None,
                         header=None, usecols=[in_numbers[n] for n in order], names=order,
                         dtype={n: in_dtypes[n] for n in order if n in in_dtypes}, iterator=True,
                         chunksize=chunksize, compression="infer")
    names, ids, start, end, _ = _bins_arrays(bins)
    lengths = np.asarray(end)[np.r_[ids[1:] != ids[:-1], True]] if len(ids) else np.zeros(0, np.int64)
    n_bins = len(ids)

    def pipeline():
        for chunk in reader:
            rec = ingest.sanitize_records({k: chunk[k].to_numpy() for k in chunk.columns}, names, lengths, ids,
                                          start, end, anchor_field="pos", is_one_based=not zero_based,
                                          tril_action=tril_action)
            vals = {f: rec[f] for f in aggregations}
            agg = dict(aggregations)
            if "count" not in agg:                                # aggregate_records(count=True): the group size
                vals["count"] = np.ones(len(rec["bin1_id"]), dtype=np.int64)
                agg["count"] = "sum"
            acc = ingest.UnorderedPixels(n_bins)
            acc.add(rec["bin1_id"], rec["bin2_id"], dupcheck=False)
            b1, b2, out = acc.merge(vals, agg)
            yield {"bin1_id": b1, "bin2_id": b2, **out}

    create_from_unordered(cool_path, bins, pipeline(), columns=out_names, dtypes=out_dtypes, metadata=metadata,
                          assembly=assembly, mergebuf=mergebuf or chunksize, boundscheck=False, triucheck=False,
                          dupcheck=False, symmetric_upper=symmetric_upper, h5opts=h5opts,
                          mode="a" if append else "w")


def _parse_input_field(arg, includes_agg=False):
    """'<name>[=<field number>][:dtype=<dtype>][,agg=<agg>]' (cli/_util.py:53-97): (name, 0-based column |
    None, dtype | None, agg | None)."""
    parts = arg.split(":")
    if len(parts) > 2:
        raise click.BadParameter(arg)
    head = parts[0].split("=")
    if len(head) > 2:
        raise click.BadParameter(arg)
    name, colnum = head[0], None
    if len(head) == 2:
        try:
            colnum = int(head[1]) - 1
        except ValueError as e:
            raise click.BadParameter(f"Not a number: '{head[1]}'", param_hint=arg) from e
        if colnum < 0:
            raise click.BadParameter("Field numbers start at 1.", param_hint=arg)
    dtype = agg = None
    if len(parts) == 2:
        for item in parts[1].split(","):
            try:
                prop, value = item.split("=")
            except ValueError as e:
                raise click.BadParameter(arg) from e
            if prop == "dtype":
                dtype = np.dtype(value)
            elif prop == "agg" and includes_agg:
                agg = value
            else:
                raise click.BadParameter(f"Invalid property: '{prop}'.", param_hint=arg)
    return name, colnum, dtype, agg


@cli.command()
@click.argument("bins_path", metavar="BINS_PATH")
@click.argument("pixels_path", metavar="PIXELS_PATH")
@click.argument("cool_path", metavar="COOL_PATH")
@click.option("--format", "-f", "format_", type=click.Choice(["coo", "bg2"]), required=True,
              help="'coo': tab-delimited sparse triplets (bin1, bin2, count); 'bg2': 2D bedGraph-like file "
                   "(chrom1, start1, end1, chrom2, start2, end2, count).")
@click.option("--metadata", help="Path to JSON file containing user metadata.")
@click.option("--assembly", help="Name of genome assembly (e.g. hg19, mm10)")
@click.option("--field", type=str, multiple=True,
              help="Supplemental value fields or other field numbers: '<name>=<field number>[:dtype=<dtype>]' "
                   "(field numbers are 1-based); repeat for each field.")
@click.option("--count-as-float", is_flag=True, default=False, help="Store the 'count' column as floating point.")
@click.option("--one-based", is_flag=True, default=False,
              help="The bin IDs of a COO file (the start coordinates of a BG2 file) are one-based.")
@click.option("--comment-char", type=str, default="#", show_default=True, help="Comment character that skips lines.")
@click.option("--no-symmetric-upper", "-N", is_flag=True, default=False,
              help="Create a complete square matrix without implicit symmetry.")
@click.option("--input-copy-status", type=click.Choice(["unique", "duplex"]), default="unique", show_default=True,
              help="'unique': lower-triangle records are mirrored; 'duplex': they are dropped (symmetric-upper only).")
@click.option("--chunksize", "-c", type=int, default=int(20e6), show_default=True,
              help="Lines of the input parsed (and duplicate-checked) at a time.")
@click.option("--mergebuf", type=int, help="Rows of the merged table written at a time. Default: CHUNKSIZE.")
@click.option("--temp-dir", type=str, help="Accepted for compatibility; ignored (no temporary coolers).")
@click.option("--max-merge", type=int, default=200, help="Accepted for compatibility; ignored.")
@click.option("--no-delete-temp", is_flag=True, default=False, help="Accepted for compatibility; ignored.")
@click.option("--storage-options", help="HDF5 filter options 'k1=v1,k2=v2,...'.")
@click.option("--append", "-a", is_flag=True, default=False,
              help="Append the output cooler to an existing file instead of overwriting the file.")
def load(bins_path, pixels_path, cool_path, format_, metadata, assembly, field, count_as_float, one_based,
         comment_char, no_symmetric_upper, input_copy_status, chunksize, mergebuf, temp_dir, max_merge,
         no_delete_temp, storage_options, append):
    """Create a cooler from a pre-binned matrix in COO or BG2 text form, records in any order
    (cli/load.py:15-400).  PIXELS_PATH '-' reads stdin; the file may be compressed.

    The reference sorts every chunk into a temporary cooler and merges the files; here the chunks are checked,
    parked on the GPU and merged by one device-wide sort / reduce-by-key (create.create_from_unordered)."""
    import json

    import pandas as pd

    from . import ingest
    from .create import _bins_arrays, create_from_unordered
    chromsizes, bins = parse_bins(bins_path)
    if mergebuf is None:
        mergebuf = chunksize
    symmetric_upper = not no_symmetric_upper
    tril_action = None
    if symmetric_upper:                                           # cli/load.py:199-204
        tril_action = "reflect" if input_copy_status == "unique" else "drop"
    if metadata is not None:
        with open(metadata) as f:
            metadata = json.load(f)
    out_names = ["bin1_id", "bin2_id"]
    out_dtypes = {"bin1_id": np.int64, "bin2_id": np.int64, "count": np.int32}
    if format_ == "bg2":
        in_names = ["chrom1", "start1", "end1", "chrom2", "start2", "end2"]
        in_dtypes = {"chrom1": str, "start1": int, "end1": int, "chrom2": str, "start2": int, "end2": int,
                     "count": out_dtypes["count"]}
        in_numbers = {"chrom1": 0, "start1": 1, "end1": 2, "chrom2": 3, "start2": 4, "end2": 5, "count": 6}
    else:
        in_names = ["bin1_id", "bin2_id"]
        in_dtypes = {"bin1_id": int, "bin2_id": int, "count": out_dtypes["count"]}
        in_numbers = {"bin1_id": 0, "bin2_id": 1, "count": 2}
    if len(field):                                                # cli/load.py:270-305
        for arg in field:
            name, colnum, dtype, _ = _parse_input_field(arg)
            if colnum is None:
                if name in ("bin1_id", "bin2_id") and dtype is not None:
                    out_dtypes[name] = dtype
                    continue
                if name == "count" and dtype is not None:
                    in_names.append("count")
                    out_names.append("count")
                    in_dtypes[name] = out_dtypes[name] = dtype
                    continue
                raise click.BadParameter("A field number is required.", param_hint=arg)
            if name not in in_names:
                in_names.append(name)
            if name not in out_names:
                out_names.append(name)
            in_numbers[name] = colnum
            if dtype is not None:
                in_dtypes[name] = out_dtypes[name] = dtype
    else:
        in_names.append("count")
        out_names.append("count")
    if "count" in in_names and count_as_float:
        in_dtypes["count"] = out_dtypes["count"] = np.float64
    h5opts = None
    if storage_options is not None:
        h5opts = {}
        for kv in storage_options.split(","):
            k, v = kv.split("=", 1)
            h5opts[k.strip()] = int(v) if v.strip().lstrip("-").isdigit() else v.strip()
    order = sorted(in_names, key=in_numbers.get)                  # read_csv hands usecols back in file order
    reader = pd.read_csv(sys.stdin if pixels_path == "-" else pixels_path, sep="\t", comment=comment_char or None,
                         header=None, usecols=[in_numbers[n] for n in order], names=order,
                         dtype={n: in_dtypes[n] for n in order if n in in_dtypes}, iterator=True,
                         chunksize=chunksize, compression="infer")
    names, ids, start, end, _ = _bins_arrays(bins)
    lengths = np.asarray(end)[np.r_[ids[1:] != ids[:-1], True]] if len(ids) else np.zeros(0, np.int64)

    def pipeline():
        for chunk in reader:
            cols = {k: chunk[k].to_numpy() for k in chunk.columns}
            if format_ == "bg2":
                yield ingest.sanitize_bg2(cols, names, lengths, ids, start, end, is_one_based=one_based,
                                          tril_action=tril_action)
            else:
                yield ingest.sanitize_pixels(cols, is_one_based=one_based, tril_action=tril_action)

    create_from_unordered(cool_path, bins, pipeline(), columns=out_names, dtypes=out_dtypes, metadata=metadata,
                          assembly=assembly, mergebuf=mergebuf, triucheck=symmetric_upper,
                          symmetric_upper=symmetric_upper, h5opts=h5opts, mode="a" if append else "w")


@cli.command()
@click.argument("cool_uri", metavar="COOL_PATH")
@click.option("--table", "-t", type=click.Choice(["chroms", "bins", "pixels"]), default="pixels", show_default=True,
              help="Which table to dump.")
@click.option("--columns", "-c", type=str, help="Restrict output to a comma-separated subset of columns.")
@click.option("--header", "-H", is_flag=True, default=False, help="Print the header of column names as the first row.")
@click.option("--na-rep", default="", help="Missing data representation. Default is empty ''.")
@click.option("--float-format", default="g", show_default=True, help="Format string for floating point numbers.")
@click.option("--range", "-r", "range_", type=str, help="Genomic region along the row dimension (UCSC notation).")
@click.option("--range2", "-r2", type=str, help="Genomic region along the column dimension.")
@click.option("--fill-lower", "-f", is_flag=True, default=False,
              help="For symmetric-upper coolers, generate the lower triangle pixels inside the query box.")
@click.option("--balanced/--no-balance", "-b", is_flag=True, default=False,
              help="Apply balancing weights to data (extra column `balanced`).")
@click.option("--join", is_flag=True, default=False, help="Print chromosome bin coordinates instead of bin IDs.")
@click.option("--annotate", type=str, help="Join additional bin-table columns (comma-separated) against the pixels.")
@click.option("--one-based-ids", is_flag=True, default=False, help="Print bin IDs as one-based.")
@click.option("--one-based-starts", is_flag=True, default=False, help="Print start coordinates as one-based.")
@click.option("--chunksize", "-k", type=int, default=1_000_000, show_default=True,
              help="Number of pixel records per query task.")
@click.option("--out", "-o", help="Output text file (.gz is compressed). Default: stdout.")
def dump(cool_uri, table, columns, header, na_rep, float_format, range_, range2, fill_lower, balanced, join,
         annotate, one_based_ids, one_based_starts, chunksize, out):
    """Dump a cooler's data to a text stream (cli/dump.py:57-304); pixel queries and balancing on the GPU."""
    from .dump import dump as _dump
    split = lambda v: tuple(x for x in v.split(",") if x) if v is not None else None   # noqa: E731
    _dump(store.open_cooler(cool_uri), out, table=table, columns=split(columns), header=header, na_rep=na_rep,
          float_format=float_format, range=range_, range2=range2, fill_lower=fill_lower, balanced=balanced,
          join=join, annotate=split(annotate), one_based_ids=one_based_ids, one_based_starts=one_based_starts,
          chunksize=chunksize)


def invoke_balance(args, resolutions, outfile):
    """Balance every zoom level that lacks a 'weight' column (cli/zoomify.py:20-46)."""
    args = [] if args is None else shlex.split(args)
    for res in resolutions:
        uri = outfile + "::resolutions/" + str(res)
        with store.open_cooler(uri).open("r") as grp:
            if "weight" in grp["bins"]:
                continue
        logger.info(f"Balancing zoom level with bin size {res}")
        try:
            balance.main(args=[uri, *args], prog_name="cooler")
        except SystemExit as e:
            if (e.code or 0) != 0:
                raise e


@cli.command()
@click.argument("cool_uri", metavar="COOL_PATH")
@click.option("--nproc", "-n", "-p", default=1, type=int, help="Accepted for compatibility; ignored.")
@click.option("--chunksize", "-c", type=int, default=int(10e6), show_default=True)
@click.option("--resolutions", "-r",
              help="Comma-separated target resolutions; suffix B (binary) or N (nice) for a "
                   "progression; 4DN = 1000,2000,5000N [default: B]")
@click.option("--balance", "do_balance", is_flag=True, default=False, help="Apply balancing to each zoom level.")
@click.option("--balance-args", type=str, help="Additional arguments to pass to cooler balance.")
@click.option("--base-uri", "-i", multiple=True, help="Additional base coolers to aggregate from.")
@click.option("--out", "-o", help="Output file or URI")
@click.option("--field", type=str, multiple=True, help="Value columns to merge.")
@click.option("--legacy", is_flag=True, default=False,
              help="Use the legacy layout of integer-labeled zoom levels.")
def zoomify(cool_uri, nproc, chunksize, resolutions, do_balance, balance_args, field, legacy, base_uri, out):
    """Generate a multi-resolution cooler by coarsening on the GPU."""
    infile, _ = store.parse_cooler_uri(cool_uri)
    outfile = infile.replace(".cool", ".mcool") if out is None else store.parse_cooler_uri(out)[0]
    logger.info(f'Recursively aggregating "{cool_uri}"')
    logger.info(f'Writing to "{outfile}"')
    if legacy:                                                    # cli/zoomify.py:139-173
        if store.is_memory_uri(cool_uri):
            raise click.UsageError("--legacy writes an .mcool file; it needs a file input")
        n_zooms, zoom_levels = legacy_zoomify(cool_uri, outfile, nproc, chunksize)
        if do_balance:
            args = [] if balance_args is None else shlex.split(balance_args)
            for level, res in reversed(list(zoom_levels.items())):
                uri = outfile + "::" + str(level)
                if level == str(n_zooms):
                    with store.open_cooler(uri).open("r") as grp:
                        if "weight" in grp["bins"]:
                            continue
                logger.info(f"Balancing zoom level {level}, bin size {res}")
                try:
                    balance.main(args=[uri, *args], prog_name="cooler")
                except SystemExit as e:
                    if (e.code or 0) != 0:
                        raise e
        return
    clr = store.open_cooler(cool_uri)
    genome_length = clr.chromsizes.values.sum()
    if clr.binsize:                                               # cli/zoomify.py:176-186
        maxres = ceil(genome_length / HIGLASS_TILE_DIM)
        curres = clr.binsize
    else:
        b = clr.bins()[["start", "end"]][:]
        maxres = ceil(genome_length / (b["end"] - b["start"]).mean() / HIGLASS_TILE_DIM)
        curres = 1
    if resolutions is None:
        resolutions = "b"
    resolutions, rstring = [], resolutions
    for res in [s.strip().lower() for s in rstring.split(",")]:   # 192-217
        if ("n" in res or "b" in res) and maxres < curres:
            warnings.warn("Map is already < 256 x 256. Provide resolutions explicitly if you want "
                          "to coarsen more.", stacklevel=1)
        if res == "n":
            r = preferred_sequence(curres, maxres, "nice")
        elif res == "b":
            r = preferred_sequence(curres, maxres, "binary")
        elif res == "4dn":
            r = [1000, 2000, *preferred_sequence(5000, maxres, "nice")]
        elif res.endswith("n"):
            r = preferred_sequence(int(res.split("n")[0]), maxres, "nice")
        elif res.endswith("b"):   # the reference repeats endswith("n") here and cannot reach this branch
            r = preferred_sequence(int(res.split("b")[0]), maxres, "binary")
        else:
            r = [int(res)]
        resolutions.extend(r)
    columns, dtypes, agg = _fields(field)
    bases = [clr] + [store.open_cooler(u) for u in base_uri]
    if store.is_memory_uri(cool_uri):
        levels = zoomify_cooler(bases, None, resolutions, chunksize, nproc=nproc, columns=columns,
                                dtypes=dtypes, agg=agg)
        store.register(outfile, {int(r): c for r, c in levels.items()})
        if do_balance:
            invoke_balance(balance_args, resolutions, outfile)
        return
    # files: under torchrun every rank keeps its row shard of every level, rank 0 writes; the per-level
    # `cooler balance` calls below then run on the same process group
    global _KEEP_DIST
    group, _ = _dist_group()
    _KEEP_DIST += 1
    try:
        zoomify_cooler([cool_uri, *list(base_uri)], outfile, resolutions, chunksize, nproc=nproc,
                       columns=columns, dtypes=dtypes, agg=agg, group=group)
        if do_balance:
            invoke_balance(balance_args, resolutions, outfile)
    finally:
        _KEEP_DIST -= 1
    _finish_dist(group)


if __name__ == "__main__":
    cli(prog_name="cooler")
