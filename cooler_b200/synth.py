"""Synthetic power-law-decay Hi-C pixel tables (bench / tests input).

hg38 chr1-22,X,Y,M binned at ``binsize``; candidate pixels are drawn per row
block (cis: log-uniform distance from the diagonal, i.e. occupancy ~ 1/(1+d);
trans/far: uniform over the upper triangle), de-duplicated and emitted in cooler
order (sorted by bin1, bin2).  Counts are ``1 + Poisson(A*b_i*b_j/(1+d))`` (cis)
or ``1 + Poisson(0.05*b_i*b_j)`` (trans) with a hidden log-normal bias b and 1 %
"bad" bins (b*0.01) so the min_nnz / MAD filters fire and ICE needs tens of passes.

Row blocks are seeded independently, so any contiguous row range (a shard) can be
generated on its own and shards tile the whole matrix exactly.  torch is used as
the array engine (CUDA if available); this is data generation, not the hot path.
"""
from __future__ import annotations

import math

import numpy as np
import torch

HG38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636,
        138394717, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345,
        83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415, 16569]
HG38_NAMES = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY", "chrM"]

ROWS_PER_BLOCK_TARGET = 24_000_000      # candidates per generation block


def genome_bins(binsize, chromsizes=None):
    chromsizes = np.asarray(HG38 if chromsizes is None else chromsizes, dtype=np.int64)
    nb = -(-chromsizes // int(binsize))
    return np.r_[0, np.cumsum(nb)].astype(np.int64), chromsizes


def hidden_bias(n_bins, seed):
    rng = np.random.default_rng(seed)
    b = rng.lognormal(0.0, 0.3, n_bins)
    b[rng.random(n_bins) < 0.01] *= 0.01
    return b


def cis_fraction(binsize):
    return float(np.clip(0.12 + 0.34 * math.log10(100_000 / binsize), 0.1, 0.85))


class SynthPlan:
    """Deterministic plan of row blocks for (binsize, nnz_target, seed)."""

    def __init__(self, binsize, nnz_target, seed=0, chromsizes=None, oversample=None):
        self.binsize, self.seed = int(binsize), int(seed)
        self.chrom_offset, self.chromsizes = genome_bins(binsize, chromsizes)
        self.n_bins = n = int(self.chrom_offset[-1])
        self.nnz_target = int(nnz_target)
        cells = n * (n + 1) // 2
        fill = min(self.nnz_target / cells, 0.9)
        f_cis = cis_fraction(binsize)
        # uniform candidates needed so that ~ (1-f_cis) * target distinct far cells are hit
        thin = 0.78                       # mean of min(1, b_i*b_j) for b ~ LogNormal(0, 0.3)
        far_target = min((1 - f_cis) * self.nnz_target / thin, 0.95 * cells)
        self.m_far = int(-cells * math.log(max(1 - far_target / cells, 1e-3)))
        rows = np.arange(n, dtype=np.float64)
        nb_ = np.diff(self.chrom_offset)
        room_ = np.repeat(self.chrom_offset[1:], nb_) - rows           # bins left in the chromosome
        lnr_ = np.log1p(room_)

        def expected_cis(c):
            # candidates are log-uniform in distance: cells with >= 1 expected candidate saturate
            dstar = np.clip(c / lnr_ - 1.0, 0.0, room_)
            return np.minimum(dstar + c * (lnr_ - np.log1p(dstar)) / lnr_, room_)

        want = f_cis * self.nnz_target / thin                          # distinct cis cells before thinning
        c = want / n
        for _ in range(30):                                            # fixed point on the candidate count
            got = float(expected_cis(c).sum())
            if got <= 0 or abs(got - want) < 1e-3 * want:
                break
            c *= min(4.0, want / got)
        self.cis_per_row = c * (oversample or 1.0)
        # per-row candidate weights -> block boundaries of ~equal candidate count
        far_w = (n - rows) / cells * self.m_far
        w = far_w + self.cis_per_row
        cw = np.r_[0.0, np.cumsum(w)]
        n_blocks = max(1, int(math.ceil(cw[-1] / ROWS_PER_BLOCK_TARGET)))
        cuts = np.searchsorted(cw, np.linspace(0, cw[-1], n_blocks + 1)[1:-1])
        self.block_edges = np.unique(np.r_[0, cuts, n]).astype(np.int64)
        # expected DISTINCT pixels per row (candidates collide near the diagonal), used to
        # balance shards by pixels rather than by candidates
        rho = self.m_far / cells
        uniq_far = (n - rows) * (1.0 - math.exp(-rho))
        self.row_weight_cum = np.r_[0.0, np.cumsum(uniq_far + expected_cis(self.cis_per_row))]

    def shard_rows(self, rank, world):
        """Contiguous row range [lo, hi) of ``rank`` holding ~1/world of the expected pixels."""
        cw = self.row_weight_cum
        n = self.n_bins
        lo = int(np.searchsorted(cw, cw[-1] * rank / world, side="left")) if rank else 0
        hi = int(np.searchsorted(cw, cw[-1] * (rank + 1) / world, side="left")) if rank + 1 < world else n
        lo, hi = min(lo, n), min(max(hi, lo), n)
        return lo, hi


def shard_rows_by_share(plan: SynthPlan, shares):
    """Contiguous row ranges [(lo, hi), ...] holding the given fractions of the expected pixels
    (``shares`` sums to 1): the cost-balanced variant of ``SynthPlan.shard_rows``, whose shares are
    all 1 / world."""
    cw = plan.row_weight_cum
    n = plan.n_bins
    cum = np.r_[0.0, np.cumsum(np.asarray(shares, dtype=np.float64))]
    cum /= cum[-1]
    cuts = np.searchsorted(cw, cw[-1] * cum[1:-1], side="left")
    cuts = np.maximum.accumulate(np.minimum(np.r_[0, cuts, n], n))
    cuts[0], cuts[-1] = 0, n
    return [(int(a), int(b)) for a, b in zip(cuts[:-1], cuts[1:])]


def _gen_block(plan: SynthPlan, blk, bias_t, row_lo_t, row_hi_t, device):
    n = plan.n_bins
    r0, r1 = int(plan.block_edges[blk]), int(plan.block_edges[blk + 1])
    g = torch.Generator(device=device)
    g.manual_seed((plan.seed * 1_000_003 + blk) & 0x7FFFFFFF)
    nrows = r1 - r0
    cells = n * (n + 1) // 2
    # far / uniform candidates: row weighted by its number of upper-triangle cells
    roww = torch.arange(n - r0, n - r1, -1, device=device, dtype=torch.float64)
    m_far = int(plan.m_far * float(roww.sum()) / cells)
    keys = []
    if m_far > 0:
        cdf = torch.cumsum(roww, 0)
        u = torch.rand(m_far, device=device, generator=g, dtype=torch.float64) * cdf[-1]
        i = torch.searchsorted(cdf, u).clamp_(max=nrows - 1) + r0
        span = (n - i).to(torch.float64)
        j = i + (torch.rand(m_far, device=device, generator=g, dtype=torch.float64) * span).to(torch.int64)
        keys.append(i * n + j.clamp_(max=n - 1))
    m_cis = int(plan.cis_per_row * nrows)
    if m_cis > 0:
        i = torch.randint(r0, r1, (m_cis,), device=device, generator=g)
        room = (row_hi_t[i] - i).to(torch.float64)            # bins left in the chromosome (>= 1)
        u = torch.rand(m_cis, device=device, generator=g, dtype=torch.float64)
        d = (torch.exp(u * torch.log(room + 1.0)) - 1.0).to(torch.int64)
        d = torch.minimum(d, (row_hi_t[i] - i - 1).to(torch.int64))
        keys.append(i * n + i + d)
    if not keys:
        z = torch.zeros(0, dtype=torch.int64, device=device)
        return z, z, z.to(torch.int32)
    key = torch.unique(torch.cat(keys), sorted=True)
    b1 = key // n
    b2 = key - b1 * n
    # occupancy thinning by the hidden bias: low-visibility bins hold few pixels
    keep = torch.rand(key.numel(), device=device, generator=g, dtype=torch.float64) < bias_t[b1] * bias_t[b2]
    b1, b2 = b1[keep], b2[keep]
    cis = b2 < row_hi_t[b1]
    dist = (b2 - b1).to(torch.float64)
    lam = torch.where(cis, 200.0 * bias_t[b1] * bias_t[b2] / (1.0 + dist), 0.05 * bias_t[b1] * bias_t[b2])
    cnt = (1 + torch.poisson(lam.to(torch.float32), generator=g)).to(torch.int32)
    return b1, b2, cnt


def generate(plan: SynthPlan, rows=None, device=None, pin=True):
    """Generate the pixels of the row range ``rows=(lo, hi)`` (default: all rows).

    Row blocks are the unit of random generation; a shard that starts or ends inside a
    block generates that block and keeps its own rows, so shards tile the matrix exactly.

    Returns a dict of HOST tensors in the reference's on-disk dtypes
    (docs/schema_v3.rst): bin2_id int64, count int32, bin1_offset int64 [n_bins+1]
    (local offsets: rows outside the shard are empty), plus row_range, nnz.
    Host tensors are pinned when a CUDA device is present.
    """
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    device = torch.device(device)
    n = plan.n_bins
    r_lo, r_hi = rows if rows is not None else (0, n)
    edges = plan.block_edges
    lo_b = int(np.searchsorted(edges, r_lo, side="right")) - 1
    hi_b = int(np.searchsorted(edges, r_hi, side="left"))
    lo_b, hi_b = max(lo_b, 0), min(max(hi_b, lo_b), len(edges) - 1)
    if r_hi <= r_lo:
        hi_b = lo_b
    bias_t = torch.from_numpy(hidden_bias(n, plan.seed)).to(device)
    nb = np.diff(plan.chrom_offset)
    row_lo_t = torch.from_numpy(np.repeat(plan.chrom_offset[:-1], nb)).to(device)
    row_hi_t = torch.from_numpy(np.repeat(plan.chrom_offset[1:], nb)).to(device)
    parts = []
    row_counts = torch.zeros(n, dtype=torch.int64, device=device)
    for blk in range(lo_b, hi_b):
        b1, b2, cnt = _gen_block(plan, blk, bias_t, row_lo_t, row_hi_t, device)
        if int(edges[blk]) < r_lo or int(edges[blk + 1]) > r_hi:      # block shared with a neighbour shard
            keep = (b1 >= r_lo) & (b1 < r_hi)
            b1, b2, cnt = b1[keep], b2[keep], cnt[keep]
        if b1.numel():
            row_counts += torch.bincount(b1, minlength=n)
        parts.append((b2.cpu(), cnt.cpu()))
        del b1, b2, cnt
    nnz = int(sum(p[0].numel() for p in parts))
    use_pin = pin and torch.cuda.is_available()
    bin2 = torch.empty(nnz, dtype=torch.int64, pin_memory=use_pin)
    count = torch.empty(nnz, dtype=torch.int32, pin_memory=use_pin)
    o = 0
    for b2, cnt in parts:
        k = b2.numel()
        bin2[o:o + k] = b2
        count[o:o + k] = cnt
        o += k
    off = torch.zeros(n + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(row_counts.cpu(), 0)
    if use_pin:
        off = off.pin_memory()
    return {"bin2_id": bin2, "count": count, "bin1_offset": off, "nnz": nnz,
            "row_range": (int(r_lo), int(r_hi)),
            "chrom_offset": plan.chrom_offset, "n_bins": n}
