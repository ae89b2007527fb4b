// ICE iteration kernels (K3 sweep, combine, K4 update).
//
// Reference semantics (open2c/cooler src/cooler/_balance.py:87-124):
//   marg_i = sum over stored pixels touching bin i of  w_bin1 * w_bin2 * count
//   nzmarg = marg[marg != 0];  mean = nzmarg.mean()
//   bias  /= where(marg/mean == 0, 1, marg/mean);  var = nzmarg.var();  stop if var < tol
// Here marg_i = w_i * s_i with s_i = sum_j w_j x_ij computed as two
// deterministic segmented reductions (rows of the bin1-major copy + rows of the
// bin2-major copy), float64, no atomics.
#include "walk.cuh"
#include <cstdlib>
#include "far.cuh"
#include "../../include/cooler_b200.h"

namespace cb200 {

// Exact int32 -> double without the (quarter-rate, XU pipe) I2F.F64 instruction:
// the bit pattern {0x43300000, count ^ 0x80000000} is 2^52 + 2^31 + count.
__device__ __forceinline__ double i32_to_f64(int v) {
  return __hiloint2double(0x43300000, v ^ 0x80000000) - 4503601774854144.0;   // 2^52 + 2^31
}

// Gather of w[other] through L1/L2.  For bias vectors far larger than L1 only the band
// around the current row is worth caching in L1: gathers farther than kNearBins from the
// row bypass L1 allocation so that they do not evict the band (the stream loads bypass
// L1 as well).  The result is the same either way.
constexpr int kNearBins = 12288;      // 96 KB of fp64 on either side of the diagonal

__device__ __forceinline__ double ld_gather_far(const double* p) {
  double r;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}

template <bool kSelective>
struct IceOpT {
  typedef double T;
  const double* __restrict__ w;
  int row_base;      // bin id of local row 0 of the sub-copy
  int center;
  __device__ __forceinline__ void at_row(int row) { if (kSelective) center = row + row_base; }
  __device__ __forceinline__ double value(int other, int count) const {
    double wv;
    if (kSelective && (unsigned)(other - center + kNearBins) > (unsigned)(2 * kNearBins))
      wv = ld_gather_far(w + other);
    else
      wv = __ldg(w + other);
    return wv * i32_to_f64(count);
  }
};
typedef IceOpT<false> IceOp;

// Floating-point count columns (cb200_ice.float_counts): the count field of a record indexes the
// sub-copy's value array (cb200_sub.val) -- its own stream position after cb200_attach_values, so the
// loads are coalesced and bypass L1 like the record stream.
template <bool kSelective>
struct IceOpValT {
  typedef double T;
  const double* __restrict__ w;
  const double* __restrict__ val;
  int row_base;
  int center;
  __device__ __forceinline__ void at_row(int row) { if (kSelective) center = row + row_base; }
  __device__ __forceinline__ double value(int other, int idx) const {
    double wv;
    if (kSelective && (unsigned)(other - center + kNearBins) > (unsigned)(2 * kNearBins))
      wv = ld_gather_far(w + other);
    else
      wv = __ldg(w + other);
    return wv * ld_gather_far(val + idx);
  }
};

// 128-bit stream loads kept in flight per lane: 4 (whole copies: 3 blocks of 8 warps per SM; strips: one
// 768-thread CTA per SM, 80 registers, no spills).  The variants with 1, 2 and 8 loads and 512- / 1024-thread
// strip CTAs lost the round-1 / round-2 measurements (profiles/r01_tune_*.log) and were removed.
constexpr int ICE_U = 4;
constexpr int STRIP_THREADS_DEFAULT = 768;
static int g_selective = -1;   // cb200_set_tuning: -1 = automatic

template <int U, bool kSelective, int MINB, bool kFloat = false>
__global__ void __launch_bounds__(BLOCK_THREADS, MINB)
ice_sweep_kernel(const cb200_sub* __restrict__ subs, int n_subs, int span, int64_t total_warps,
                 const double* __restrict__ w, const int32_t* __restrict__ all_done) {
  if (*all_done) return;
  const int64_t wg = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (wg >= total_warps) return;
  int lo = 0, hi = n_subs;                         // largest k with subs[k].first_warp <= wg
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (subs[mid].first_warp <= wg) lo = mid; else hi = mid;
  }
  const cb200_sub sub = subs[lo];
  const int64_t n_tiles = (sub.nnz + WARP_TILE - 1) / WARP_TILE;
  const int64_t t0 = (wg - sub.first_warp) * span;
  if (t0 >= n_tiles) return;
  if constexpr (kFloat) {
    IceOpValT<kSelective> op{w, sub.val, sub.row_base, 0};
    walk_reduce_span<IceOpValT<kSelective>, U>(reinterpret_cast<const int2*>(sub.px), sub.indptr, sub.tile_row,
                                               sub.row_count, sub.nnz, t0, min(t0 + span, n_tiles), op, sub.direct, sub.pieces);
  } else {
    IceOpT<kSelective> op{w, sub.row_base, 0};
    walk_reduce_span<IceOpT<kSelective>, U>(reinterpret_cast<const int2*>(sub.px), sub.indptr, sub.tile_row,
                                            sub.row_count, sub.nnz, t0, min(t0 + span, n_tiles), op, sub.direct, sub.pieces);
  }
}

// Shared-memory form of K3 for column strips: a persistent CTA per SM owns a
// contiguous range of the global tile sequence (sub-copies concatenated); for every
// strip it touches it stages that strip's slice of w (<= smem_bins doubles) in shared
// memory once and its 32 warps gather from there.  Random 8-byte gathers cost a few
// bank-conflict cycles in shared memory instead of one L1 tag lookup per lane.
struct IceOpSmem {
  typedef double T;
  const double* ws;     // shared memory slice
  int base;
  __device__ __forceinline__ void at_row(int) {}
  __device__ __forceinline__ double value(int other, int count) const {
    return ws[other - base] * i32_to_f64(count);
  }
};
struct IceOpSmemVal {                               // floating-point count columns, see IceOpValT
  typedef double T;
  const double* ws;
  int base;
  const double* __restrict__ val;
  __device__ __forceinline__ void at_row(int) {}
  __device__ __forceinline__ double value(int other, int idx) const {
    return ws[other - base] * ld_gather_far(val + idx);
  }
};


// Work distribution: one persistent CTA per SM.  Every sub-copy has a claim counter.
// A CTA starts at the sub-copy where an even split of all tiles would put it and
// visits the sub-copies cyclically; on a sub-copy that still has work it stages the
// strip's slice of w once, then each of its WARPS claims `span` tiles at a time until
// the sub-copy is exhausted (no CTA barrier per claim).  Claims only decide WHO
// computes a tile, never the arithmetic, so results stay bit-reproducible.  The last
// CTA to finish re-arms the counters.
template <int U, int STRIP_THREADS, bool kFloat = false>
__global__ void __launch_bounds__(STRIP_THREADS, 1)
ice_sweep_strips_kernel(const cb200_sub* __restrict__ subs, int n_subs, int span,
                        int64_t total_tiles, int smem_bins, const double* __restrict__ w,
                        unsigned long long* __restrict__ claim,
                        const int32_t* __restrict__ all_done) {
  extern __shared__ double w_s[];
  __shared__ int s_has_work;
  if (*all_done) return;
  const int lane = threadIdx.x & 31;
  const int64_t g0 = total_tiles * (int64_t)blockIdx.x / gridDim.x;
  int lo = 0, hi = n_subs;                         // sub-copy that holds global tile g0
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (subs[mid].first_tile <= g0) lo = mid; else hi = mid;
  }
  int k = lo;
  for (int visited = 0; visited < n_subs; ++visited, k = (k + 1 == n_subs) ? 0 : k + 1) {
    const cb200_sub sub = subs[k];
    const int64_t n_tiles = (sub.nnz + WARP_TILE - 1) / WARP_TILE;
    __syncthreads();                               // everyone is done with the previous strip's w_s
    if (threadIdx.x == 0)
      s_has_work = (int64_t)(*reinterpret_cast<volatile unsigned long long*>(claim + k)) < n_tiles;
    __syncthreads();
    if (!s_has_work) continue;                     // CTA-uniform
    const bool in_smem = sub.gather_len <= smem_bins;
    if (in_smem) {                                 // stage this strip's slice of w
      for (int i = threadIdx.x; i < sub.gather_len; i += STRIP_THREADS) w_s[i] = w[sub.gather_base + i];
      __syncthreads();
    }
    while (true) {
      // guided self-scheduling: claims shrink towards the end of a sub-copy so that the
      // warps of a CTA reach the end-of-strip barrier at nearly the same time
      unsigned long long c = 0;
      int take = span;
      if (lane == 0) {
        const int64_t rem = n_tiles - (int64_t)(*reinterpret_cast<volatile unsigned long long*>(claim + k));
        take = (int)max((int64_t)1, min((int64_t)span, rem / 512));
        c = atomicAdd(&claim[k], (unsigned long long)take);
      }
      const int64_t t0 = (int64_t)__shfl_sync(0xffffffffu, c, 0);
      take = __shfl_sync(0xffffffffu, take, 0);
      if (t0 >= n_tiles) break;
      const int64_t t1 = min(t0 + take, n_tiles);
      if constexpr (kFloat) {
        if (in_smem) {
          IceOpSmemVal op{w_s, sub.gather_base, sub.val};
          walk_reduce_span<IceOpSmemVal, U>(reinterpret_cast<const int2*>(sub.px), sub.indptr, sub.tile_row,
                                            sub.row_count, sub.nnz, t0, t1, op, sub.direct, sub.pieces);
        } else {
          IceOpValT<false> op{w, sub.val, 0, 0};
          walk_reduce_span<IceOpValT<false>, U>(reinterpret_cast<const int2*>(sub.px), sub.indptr, sub.tile_row,
                                                sub.row_count, sub.nnz, t0, t1, op, sub.direct, sub.pieces);
        }
      } else if (in_smem) {
        IceOpSmem op{w_s, sub.gather_base};
        walk_reduce_span<IceOpSmem, U>(reinterpret_cast<const int2*>(sub.px), sub.indptr, sub.tile_row,
                                       sub.row_count, sub.nnz, t0, t1, op, sub.direct, sub.pieces);
      } else {                                     // far-field remainder: gather through L1/L2
        IceOp op{w};
        walk_reduce_span<IceOp, U>(reinterpret_cast<const int2*>(sub.px), sub.indptr, sub.tile_row,
                                   sub.row_count, sub.nnz, t0, t1, op, sub.direct, sub.pieces);
      }
    }
  }
  // re-arm the claim counters once every CTA is through
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long ticket = atomicAdd(&claim[n_subs], 1ULL);
    if (ticket == gridDim.x - 1) {
      for (int i = 0; i <= n_subs; ++i) claim[i] = 0ULL;
      __threadfence();
    }
  }
}

// s[i] = sum over sub-copies (fixed order) of the row total of bin i, plus the far-field units.
__global__ void ice_combine_kernel(const cb200_ice st) {
  if (*st.all_done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= st.n_bins) return;
  double v = 0.0;
  for (int k = 0; k < st.n_subs; ++k) {
    const unsigned r = (unsigned)(i - st.subs[k].row_base);
    if (r < (unsigned)st.subs[k].row_count)
      v += row_total<double>((int)r, st.subs[k].indptr, st.subs[k].nnz, st.subs[k].direct, st.subs[k].pieces);
  }
  if (st.n_far_units > 0) v += far_row_total(st, i);
  st.s[i] = v;
}

// Reduce-scatter half of the two-stage all-reduce: this rank sums, in rank order, the slice of
// the per-bin vector it owns over every rank's local s (peer loads) into its own `red` buffer.
__global__ void ice_reduce_slice_kernel(const int32_t* __restrict__ all_done,
                                        const double* const* __restrict__ peer_s, int world, int64_t peer_off,
                                        double* __restrict__ red, int lo, int hi) {
  if (*all_done) return;
  // one bin per thread, `world` independent peer loads in flight (two bins per thread in half as many
  // threads measured slower at 8 GPUs: 0.075 vs 0.045 ms for the 3.1e6-bin vector)
  const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hi) return;
  double v = 0.0;
  for (int r = 0; r < world; ++r) v += __ldcv(peer_s[r] + peer_off + i);
  red[i] = v;
}

constexpr int UPD_THREADS = 256;

__device__ __forceinline__ double block_sum(double v, double* sm) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int k = 0; k < UPD_THREADS / 32; ++k) t += sm[k];
  return t;  // valid in thread 0
}

// Deterministic "last block finishes" pattern: every block publishes its partials,
// takes a ticket, and the block that draws the last ticket reduces all partials in a
// fixed order.  The ticket only elects WHO reduces; the summation order is fixed.
__device__ __forceinline__ bool last_block_ticket(int32_t* ticket) {
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const int t = atomicAdd(ticket, 1);
    s_last = (t == (int)gridDim.x - 1);
    if (s_last) *ticket = 0;                       // re-arm for the next launch
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

// K4a: [s = row totals over sub-copies] -> marg = w * s -> per-block sum / count of the
// nonzero marginals -> (last block) per-segment mean.
// kMode 0: s is given (after an NCCL all-reduce); 1: s = row totals of the local
// sub-copies (single GPU); 2: s = sum over ranks, in rank order, of every rank's local s
// read through peer (NVLink) pointers -- the all-reduce fused into K4a, identical bits
// on every rank; 3: s was reduced slice-wise by ice_reduce_slice_kernel (rank q owns the bins
// [q, q+1) * ceil(n_bins / world)) and is gathered here from the owners' buffers -- the
// all-gather half of a two-stage all-reduce, (world - 1) / world * 8 B per bin over NVLink
// instead of (world - 1) * 8 B.
template <int kMode>
__global__ void __launch_bounds__(UPD_THREADS)
ice_marg_kernel(cb200_ice st, const double* const* __restrict__ peer_s, int world, int64_t peer_off) {
  __shared__ double sm[UPD_THREADS / 32];
  if (*st.all_done) return;
  const int b = blockIdx.x;
  const int seg = st.blk_seg[b];
  if (!st.seg_done[seg]) {
    const int lo = st.blk_lo[b], hi = st.blk_hi[b];
    double sum = 0.0, cnt = 0.0;
    // kMode 3: a block's bins nearly always have ONE owner -> its pointer is fetched once per block
    const int slice = (kMode == 3) ? (st.n_bins + world - 1) / world : 1;
    const bool one_owner = (kMode == 3) && (lo / slice == (hi - 1) / slice);
    const double* __restrict__ owner = one_owner ? peer_s[lo / slice] + peer_off : nullptr;
    // The loads of a batch of bins are issued before anything is stored: the compiler cannot hoist a
    // load over the store to st.marg of the previous bin (the pointers may alias), and a peer load
    // over NVLink takes microseconds -- one bin at a time made this kernel latency-bound.
    constexpr int B = 8;
    for (int i0 = lo + threadIdx.x; i0 < hi; i0 += UPD_THREADS * B) {
      double sv[B], wv[B];
#pragma unroll
      for (int u = 0; u < B; ++u) {
        const int i = i0 + u * UPD_THREADS;
        sv[u] = 0.0;
        wv[u] = 0.0;
        if (i < hi) {
          if (kMode == 2) {
            for (int r = 0; r < world; ++r) sv[u] += __ldcv(peer_s[r] + peer_off + i);
          } else if (kMode == 3) {
            sv[u] = one_owner ? __ldcv(owner + i) : __ldcv(peer_s[i / slice] + peer_off + i);
          } else if (kMode == 0) {
            sv[u] = st.s[i];
          }
          wv[u] = st.w[i];
        }
      }
      if (kMode == 1) {
        // row totals over the sub-copies, in sub-copy order per bin.  A row total is two dependent
        // loads (row pointers, then the sum); per sub-copy the row pointers of the whole batch are
        // fetched first, then the sums -- bin after bin, sub-copy after sub-copy (38 strips at 10 kb)
        // was a chain of ~150 dependent load latencies per thread (0.13 ms per pass).
        for (int k = 0; k < st.n_subs; ++k) {
          const cb200_sub sb = st.subs[k];
          int64_t ra[B], rb[B];
#pragma unroll
          for (int u = 0; u < B; ++u) {
            const int i = i0 + u * UPD_THREADS;
            const unsigned r = (unsigned)(i - sb.row_base);
            ra[u] = rb[u] = 0;
            if (i < hi && r < (unsigned)sb.row_count) {
              ra[u] = sb.indptr[r];
              rb[u] = sb.indptr[r + 1];
            }
          }
#pragma unroll
          for (int u = 0; u < B; ++u) {
            const int i = i0 + u * UPD_THREADS;
            if (ra[u] < rb[u])
              sv[u] += row_total_ab<double>(i - sb.row_base, ra[u], rb[u], sb.nnz, sb.direct, sb.pieces);
          }
        }
        if (st.n_far_units > 0) {
#pragma unroll
          for (int u = 0; u < B; ++u) {
            const int i = i0 + u * UPD_THREADS;
            if (i < hi) sv[u] += far_row_total(st, i);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < B; ++u) {
        const int i = i0 + u * UPD_THREADS;
        if (i < hi) {
          const double m = wv[u] * sv[u];
          st.marg[i] = m;
          if (m != 0.0) { sum += m; cnt += 1.0; }
        }
      }
    }
    const double bs = block_sum(sum, sm);
    const double bc = block_sum(cnt, sm);
    if (threadIdx.x == 0) { st.blk_sum[b] = bs; st.blk_cnt[b] = bc; }
  }
  if (!last_block_ticket(st.tickets)) return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int seg2 = wid; seg2 < st.n_segs; seg2 += UPD_THREADS / 32) {
    if (st.seg_done[seg2]) continue;
    double sum = 0.0, cnt = 0.0;
    for (int k = st.seg_blk_offset[seg2] + lane; k < st.seg_blk_offset[seg2 + 1]; k += 32) {
      sum += __ldcg(st.blk_sum + k);
      cnt += __ldcg(st.blk_cnt + k);
    }
    sum = warp_sum(sum);
    cnt = warp_sum(cnt);
    if (lane == 0) {
      st.seg_nnz[seg2] = cnt;
      st.seg_mean[seg2] = (cnt > 0.0) ? sum / cnt : nan("");
    }
  }
}

// K4b: bias /= marg/mean (0 -> 1), refresh w, per-block squared deviations ->
// (last block) per-segment variance, iteration counters, convergence flags.
__global__ void __launch_bounds__(UPD_THREADS)
ice_update_kernel(cb200_ice st, int iter) {
  __shared__ double sm[UPD_THREADS / 32];
  if (*st.all_done) return;
  const int b = blockIdx.x;
  const int seg = st.blk_seg[b];
  const double nz = st.seg_nnz[seg];
  if (!st.seg_done[seg] && nz != 0.0) {          // empty segments are finalised below
    const double mean = st.seg_mean[seg];
    const int lo = st.blk_lo[b], hi = st.blk_hi[b];
    double dev = 0.0;
    constexpr int B = 8;                           // loads of a batch of bins before its stores (see ice_marg_kernel)
    for (int i0 = lo + threadIdx.x; i0 < hi; i0 += UPD_THREADS * B) {
      double mv[B], bv[B], cw[B];
#pragma unroll
      for (int u = 0; u < B; ++u) {
        const int i = i0 + u * UPD_THREADS;
        mv[u] = 0.0;
        bv[u] = 0.0;
        cw[u] = 1.0;
        if (i < hi) {
          mv[u] = st.marg[i];
          bv[u] = st.bias[i];
          if (st.cweights) cw[u] = st.cweights[i];
        }
      }
#pragma unroll
      for (int u = 0; u < B; ++u) {
        const int i = i0 + u * UPD_THREADS;
        if (i < hi) {
          const double m = mv[u];
          if (m != 0.0) { const double d = m - mean; dev += d * d; }
          double q = m / mean;
          if (q == 0.0) q = 1.0;
          const double nb = bv[u] / q;
          st.bias[i] = nb;
          if (st.cweights) st.w[i] = nb * cw[u];
          else if (st.w != st.bias) st.w[i] = nb;
        }
      }
    }
    const double bd = block_sum(dev, sm);
    if (threadIdx.x == 0) st.blk_dev[b] = bd;
  }
  if (!last_block_ticket(st.tickets + 1)) return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int my_all = 1;
  for (int seg2 = wid; seg2 < st.n_segs; seg2 += UPD_THREADS / 32) {
    if (!st.seg_done[seg2]) {
      const double nz2 = st.seg_nnz[seg2];
      double dev = 0.0;
      if (nz2 > 0.0)
        for (int k = st.seg_blk_offset[seg2] + lane; k < st.seg_blk_offset[seg2 + 1]; k += 32)
          dev += __ldcg(st.blk_dev + k);
      dev = warp_sum(dev);
      if (lane == 0) {
        if (nz2 == 0.0) {                          // _balance.py:98-102 / 166-170
          st.seg_scale[seg2] = nan("");
          st.seg_var[seg2] = 0.0;
          st.seg_empty[seg2] = 1;
          st.seg_done[seg2] = 1;
        } else {
          const double var = dev / nz2;
          st.seg_scale[seg2] = st.seg_mean[seg2];
          st.seg_var[seg2] = var;
          st.seg_iters[seg2] += 1;
          if (st.var_hist) st.var_hist[(int64_t)iter * st.n_segs + seg2] = var;
          if (var < st.tol) st.seg_done[seg2] = 1;
        }
      }
      __syncwarp();
    }
    if (!((volatile int32_t*)st.seg_done)[seg2]) my_all = 0;   // lane 0's write, ordered by __syncwarp
  }
  const int all = __syncthreads_and(my_all);
  if (threadIdx.x == 0 && all) *st.all_done = 1;
}

}  // namespace cb200

using namespace cb200;

extern "C" {

int cb200_set_tuning(int32_t ice_unroll, int32_t selective_l1) {
  if (ice_unroll != 0 && ice_unroll != ICE_U) {
    set_error_msg("set_tuning: the sweep kernels keep 4 stream loads in flight (ice_unroll 0 or 4)");
    return -1;
  }
  g_selective = selective_l1 < 0 ? -1 : (selective_l1 ? 1 : 0);
  return 0;
}

int cb200_far_row_shift(void) { return CB200_FAR_ROW_SHIFT; }

// K3 over the far-field blocks (far.cuh): one persistent CTA per SM claims units.
static int ice_sweep_far(const cb200_ice* st, void* stream) {
  if (st->n_far_units <= 0) return 0;
  if (st->float_counts) {
    set_error_msg("ice_sweep: far-field blocks hold integer counts only");
    return -1;
  }
  if (st->far_col_shift < 1 || st->far_col_shift > 13 || st->far_warps != 16 ||
      (st->far_rec_bytes != 8 && !(st->far_rec_bytes == 4 && CB200_FAR_ROW_SHIFT <= 12))) {
    set_error_msg("ice_sweep: far_col_shift must be in 1..13, far_warps 16, far_rec_bytes 8 (or 4 with <= 4096-row blocks)");
    return -1;
  }
  const size_t smem = ((size_t)FAR_ACC + FAR_BUFS * ((size_t)1 << st->far_col_shift)) * sizeof(double);
  constexpr size_t kSmemLimit = 227 * 1024;        // opt-in maximum of dynamic shared memory per CTA on sm_100
  if (smem + 64 > kSmemLimit) {
    set_error_msg("ice_sweep: far-field row sums + two column tiles exceed 227 KB of shared memory");
    return -1;
  }
  static size_t attr_bytes = 0;
  if (smem > attr_bytes) {
    CB_CHECK(cudaFuncSetAttribute(ice_sweep_far_kernel<16, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CB_CHECK(cudaFuncSetAttribute(ice_sweep_far_kernel<16, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_bytes = smem;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = st->n_far_units < sms ? st->n_far_units : sms;
  // 16 warps x 8 128-bit loads in flight per lane (measured 1.70 vs 2.15 ms with 4 on the 1 kb rank-0 shard)
  // compact records, 8 x 64-bit loads in flight per lane: 1.33 ms on the 1 kb rank-0 shard (4: 1.57, 16: 1.35)
  if (st->far_rec_bytes == 4)
    ice_sweep_far_kernel<16, 8, true><<<grid, 512, smem, (cudaStream_t)stream>>>(*st);
  else
    ice_sweep_far_kernel<16, 8, false><<<grid, 512, smem, (cudaStream_t)stream>>>(*st);
  CB_LAUNCH_CHECK("ice_sweep_far");
  return 0;
}

static int ice_sweep_near(const cb200_ice* st, void* stream);

int cb200_ice_sweep(const cb200_ice* st, void* stream) {
  const int rc = ice_sweep_near(st, stream);
  if (rc != 0) return rc;
  return ice_sweep_far(st, stream);
}

static int ice_sweep_near(const cb200_ice* st, void* stream) {
  if (st->total_warps <= 0 || st->n_subs <= 0) return 0;
  if (st->smem_bins > 0) {                          // column strips: persistent shared-memory kernel
    const size_t smem = (size_t)st->smem_bins * sizeof(double);
    static size_t attr_bytes = 0;
    if (smem > attr_bytes) {
      if (smem > 226 * 1024) {
        set_error_msg("ice_sweep: strip slice does not fit shared memory");
        return -1;
      }
      CB_CHECK(cudaFuncSetAttribute(ice_sweep_strips_kernel<ICE_U, STRIP_THREADS_DEFAULT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CB_CHECK(cudaFuncSetAttribute(ice_sweep_strips_kernel<ICE_U, STRIP_THREADS_DEFAULT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_bytes = smem;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (st->float_counts)
      ice_sweep_strips_kernel<ICE_U, STRIP_THREADS_DEFAULT, true><<<sms, STRIP_THREADS_DEFAULT, smem, (cudaStream_t)stream>>>(
          st->subs, st->n_subs, st->span, st->total_tiles, st->smem_bins, st->w,
          reinterpret_cast<unsigned long long*>(st->claim), st->all_done);
    else
      ice_sweep_strips_kernel<ICE_U, STRIP_THREADS_DEFAULT><<<sms, STRIP_THREADS_DEFAULT, smem, (cudaStream_t)stream>>>(
          st->subs, st->n_subs, st->span, st->total_tiles, st->smem_bins, st->w,
          reinterpret_cast<unsigned long long*>(st->claim), st->all_done);
    CB_LAUNCH_CHECK("ice_sweep_strips");
    return 0;
  }
  const int64_t blocks = ceil_div64(st->total_warps, WARPS_PER_BLOCK);
  // selective L1 allocation only pays when w is much larger than L1 (tools/tune_sweep.py)
  const int selective = g_selective < 0 ? (st->n_bins > 4 * kNearBins ? 1 : 0) : g_selective;
  const cudaStream_t cs = (cudaStream_t)stream;
  const unsigned grid = (unsigned)blocks;
  if (st->float_counts) {
    if (selective)
      ice_sweep_kernel<ICE_U, true, 2, true><<<grid, BLOCK_THREADS, 0, cs>>>(st->subs, st->n_subs, st->span, st->total_warps, st->w, st->all_done);
    else
      ice_sweep_kernel<ICE_U, false, 2, true><<<grid, BLOCK_THREADS, 0, cs>>>(st->subs, st->n_subs, st->span, st->total_warps, st->w, st->all_done);
  } else if (selective) {
    ice_sweep_kernel<ICE_U, true, 3><<<grid, BLOCK_THREADS, 0, cs>>>(st->subs, st->n_subs, st->span, st->total_warps, st->w, st->all_done);
  } else {
    ice_sweep_kernel<ICE_U, false, 3><<<grid, BLOCK_THREADS, 0, cs>>>(st->subs, st->n_subs, st->span, st->total_warps, st->w, st->all_done);
  }
  CB_LAUNCH_CHECK("ice_sweep");
  return 0;
}

int cb200_ice_combine(const cb200_ice* st, void* stream) {
  if (st->n_bins <= 0) return 0;
  ice_combine_kernel<<<(unsigned)ceil_div64(st->n_bins, 256), 256, 0, (cudaStream_t)stream>>>(*st);
  CB_LAUNCH_CHECK("ice_combine");
  return 0;
}

int cb200_ice_update_peer(const cb200_ice* st, int32_t iter, const double* const* peer_s,
                          int32_t world, int64_t peer_offset, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (st->n_blocks <= 0) return 0;
  ice_marg_kernel<2><<<st->n_blocks, UPD_THREADS, 0, s>>>(*st, peer_s, world, peer_offset);
  CB_LAUNCH_CHECK("ice_marg_peer");
  ice_update_kernel<<<st->n_blocks, UPD_THREADS, 0, s>>>(*st, iter);
  CB_LAUNCH_CHECK("ice_update");
  return 0;
}

int cb200_ice_reduce_peer(const cb200_ice* st, const double* const* peer_s, int32_t world, int32_t rank,
                          int64_t peer_offset, double* red_local, void* stream) {
  if (st->n_bins <= 0 || world <= 0 || rank < 0 || rank >= world) return 0;
  const int slice = (st->n_bins + world - 1) / world;
  const int lo = rank * slice < st->n_bins ? rank * slice : st->n_bins;
  const int hi = lo + slice < st->n_bins ? lo + slice : st->n_bins;
  if (hi <= lo) return 0;
  ice_reduce_slice_kernel<<<(unsigned)ceil_div64(hi - lo, 256), 256, 0, (cudaStream_t)stream>>>(
      st->all_done, peer_s, world, peer_offset, red_local, lo, hi);
  CB_LAUNCH_CHECK("ice_reduce_slice");
  return 0;
}

int cb200_ice_update_gather(const cb200_ice* st, int32_t iter, const double* const* peer_red,
                            int32_t world, int64_t peer_offset, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (st->n_blocks <= 0) return 0;
  ice_marg_kernel<3><<<st->n_blocks, UPD_THREADS, 0, s>>>(*st, peer_red, world, peer_offset);
  CB_LAUNCH_CHECK("ice_marg_gather");
  ice_update_kernel<<<st->n_blocks, UPD_THREADS, 0, s>>>(*st, iter);
  CB_LAUNCH_CHECK("ice_update");
  return 0;
}

int cb200_ice_update(const cb200_ice* st, int32_t iter, int32_t fused_combine, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (st->n_blocks <= 0) return 0;
  if (fused_combine)
    ice_marg_kernel<1><<<st->n_blocks, UPD_THREADS, 0, s>>>(*st, nullptr, 0, 0);
  else
    ice_marg_kernel<0><<<st->n_blocks, UPD_THREADS, 0, s>>>(*st, nullptr, 0, 0);
  CB_LAUNCH_CHECK("ice_marg");
  ice_update_kernel<<<st->n_blocks, UPD_THREADS, 0, s>>>(*st, iter);
  CB_LAUNCH_CHECK("ice_update");
  return 0;
}

}  // extern "C"
