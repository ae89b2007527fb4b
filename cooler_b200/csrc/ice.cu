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
#include "../../include/cooler_b200.h"

namespace cb200 {

struct IceOp {
  typedef double T;
  const double* __restrict__ w;
  __device__ __forceinline__ double value(int other, int count) const {
    return __ldg(w + other) * (double)count;
  }
};

constexpr int ICE_UNROLL = 4;

__global__ void __launch_bounds__(BLOCK_THREADS)
ice_sweep_kernel(const int2* __restrict__ px_a, const int64_t* __restrict__ indptr_a,
                 const int32_t* __restrict__ tile_row_a, int64_t nnz_a, int64_t n_tiles_a,
                 double* __restrict__ direct_a, double* __restrict__ pieces_a,
                 const int2* __restrict__ px_b, const int64_t* __restrict__ indptr_b,
                 const int32_t* __restrict__ tile_row_b, int64_t nnz_b, int64_t n_tiles_b,
                 double* __restrict__ direct_b, double* __restrict__ pieces_b,
                 const double* __restrict__ w, const int32_t* __restrict__ all_done) {
  if (*all_done) return;
  int64_t t = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  IceOp op{w};
  if (t < n_tiles_a) {
    walk_reduce_tile<IceOp, ICE_UNROLL>(px_a, indptr_a, nnz_a, t, tile_row_a[t], op, direct_a, pieces_a);
  } else {
    t -= n_tiles_a;
    if (t < n_tiles_b)
      walk_reduce_tile<IceOp, ICE_UNROLL>(px_b, indptr_b, nnz_b, t, tile_row_b[t], op, direct_b, pieces_b);
  }
}

__global__ void ice_combine_kernel(const int64_t* __restrict__ indptr_a, int64_t nnz_a,
                                   const double* __restrict__ direct_a, const double* __restrict__ pieces_a,
                                   const int64_t* __restrict__ indptr_b, int64_t nnz_b,
                                   const double* __restrict__ direct_b, const double* __restrict__ pieces_b,
                                   int n_bins, double* __restrict__ s, const int32_t* __restrict__ all_done) {
  if (*all_done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_bins) return;
  s[i] = row_total<double>(i, indptr_a, nnz_a, direct_a, pieces_a) +
         row_total<double>(i, indptr_b, nnz_b, direct_b, pieces_b);
}

constexpr int UPD_THREADS = 256;
constexpr int UPD_BINS = 1024;   // bins per block of the per-bin kernels (host block table uses the same)

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

// K4a: marg = w * s; per-block sum and count of the nonzero marginals.
__global__ void __launch_bounds__(UPD_THREADS)
ice_marg_kernel(cb200_ice st) {
  __shared__ double sm[UPD_THREADS / 32];
  if (*st.all_done) return;
  const int b = blockIdx.x;
  const int seg = st.blk_seg[b];
  if (st.seg_done[seg]) return;
  const int lo = st.blk_lo[b], hi = st.blk_hi[b];
  double sum = 0.0, cnt = 0.0;
  for (int i = lo + threadIdx.x; i < hi; i += UPD_THREADS) {
    const double m = st.w[i] * st.s[i];
    st.marg[i] = m;
    if (m != 0.0) { sum += m; cnt += 1.0; }
  }
  const double bs = block_sum(sum, sm);
  const double bc = block_sum(cnt, sm);
  if (threadIdx.x == 0) { st.blk_sum[b] = bs; st.blk_cnt[b] = bc; }
}

// one block; warp per segment
__global__ void __launch_bounds__(1024)
ice_mean_kernel(cb200_ice st) {
  if (*st.all_done) return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int seg = wid; seg < st.n_segs; seg += 32) {
    if (st.seg_done[seg]) continue;
    double sum = 0.0, cnt = 0.0;
    for (int b = st.seg_blk_offset[seg] + lane; b < st.seg_blk_offset[seg + 1]; b += 32) {
      sum += st.blk_sum[b];
      cnt += st.blk_cnt[b];
    }
    sum = warp_sum(sum);
    cnt = warp_sum(cnt);
    if (lane == 0) {
      st.seg_nnz[seg] = cnt;
      st.seg_mean[seg] = (cnt > 0.0) ? sum / cnt : nan("");
    }
  }
}

// K4b: bias /= marg/mean (0 -> 1), refresh w, per-block squared deviations.
__global__ void __launch_bounds__(UPD_THREADS)
ice_update_kernel(cb200_ice st) {
  __shared__ double sm[UPD_THREADS / 32];
  if (*st.all_done) return;
  const int b = blockIdx.x;
  const int seg = st.blk_seg[b];
  if (st.seg_done[seg]) return;
  const double nz = st.seg_nnz[seg];
  if (nz == 0.0) return;                       // empty segment: handled in ice_var_kernel
  const double mean = st.seg_mean[seg];
  const int lo = st.blk_lo[b], hi = st.blk_hi[b];
  double dev = 0.0;
  for (int i = lo + threadIdx.x; i < hi; i += UPD_THREADS) {
    const double m = st.marg[i];
    if (m != 0.0) { const double d = m - mean; dev += d * d; }
    double q = m / mean;
    if (q == 0.0) q = 1.0;
    const double nb = st.bias[i] / q;
    st.bias[i] = nb;
    if (st.cweights) st.w[i] = nb * st.cweights[i];
    else if (st.w != st.bias) st.w[i] = nb;
  }
  const double bd = block_sum(dev, sm);
  if (threadIdx.x == 0) st.blk_dev[b] = bd;
}

__global__ void __launch_bounds__(1024)
ice_var_kernel(cb200_ice st, int iter) {
  if (*st.all_done) return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int my_all = 1;
  for (int seg = wid; seg < st.n_segs; seg += 32) {
    if (!st.seg_done[seg]) {
      const double nz = st.seg_nnz[seg];
      double dev = 0.0;
      if (nz > 0.0)
        for (int b = st.seg_blk_offset[seg] + lane; b < st.seg_blk_offset[seg + 1]; b += 32)
          dev += st.blk_dev[b];
      dev = warp_sum(dev);
      if (lane == 0) {
        if (nz == 0.0) {                       // _balance.py:98-102 / 166-170
          st.seg_scale[seg] = nan("");
          st.seg_var[seg] = 0.0;
          st.seg_empty[seg] = 1;
          st.seg_done[seg] = 1;
        } else {
          const double var = dev / nz;
          st.seg_scale[seg] = st.seg_mean[seg];
          st.seg_var[seg] = var;
          st.seg_iters[seg] += 1;
          if (st.var_hist) st.var_hist[(int64_t)iter * st.n_segs + seg] = var;
          if (var < st.tol) st.seg_done[seg] = 1;
        }
      }
      __syncwarp();
    }
    if (!((volatile int32_t*)st.seg_done)[seg]) my_all = 0;   // lane 0's write, ordered by __syncwarp
  }
  const int all = __syncthreads_and(my_all);
  if (threadIdx.x == 0 && all) *st.all_done = 1;
}

}  // namespace cb200

using namespace cb200;

extern "C" {

int cb200_ice_sweep(const cb200_ice* st, void* stream) {
  const int64_t ta = ceil_div64(st->csr.nnz, WARP_TILE), tb = ceil_div64(st->csc.nnz, WARP_TILE);
  const int64_t blocks = ceil_div64(ta + tb, WARPS_PER_BLOCK);
  if (blocks <= 0) return 0;
  ice_sweep_kernel<<<(unsigned)blocks, BLOCK_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const int2*>(st->csr.px), st->csr.indptr, st->csr.tile_row, st->csr.nnz, ta,
      st->direct_row, st->pieces_row, reinterpret_cast<const int2*>(st->csc.px), st->csc.indptr,
      st->csc.tile_row, st->csc.nnz, tb, st->direct_col, st->pieces_col, st->w, st->all_done);
  CB_LAUNCH_CHECK("ice_sweep");
  return 0;
}

int cb200_ice_combine(const cb200_ice* st, void* stream) {
  if (st->n_bins <= 0) return 0;
  ice_combine_kernel<<<(unsigned)ceil_div64(st->n_bins, 256), 256, 0, (cudaStream_t)stream>>>(
      st->csr.indptr, st->csr.nnz, st->direct_row, st->pieces_row, st->csc.indptr, st->csc.nnz,
      st->direct_col, st->pieces_col, st->n_bins, st->s, st->all_done);
  CB_LAUNCH_CHECK("ice_combine");
  return 0;
}

int cb200_ice_update(const cb200_ice* st, int32_t iter, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (st->n_blocks > 0) {
    ice_marg_kernel<<<st->n_blocks, UPD_THREADS, 0, s>>>(*st);
    CB_LAUNCH_CHECK("ice_marg");
  }
  ice_mean_kernel<<<1, 1024, 0, s>>>(*st);
  CB_LAUNCH_CHECK("ice_mean");
  if (st->n_blocks > 0) {
    ice_update_kernel<<<st->n_blocks, UPD_THREADS, 0, s>>>(*st);
    CB_LAUNCH_CHECK("ice_update");
  }
  ice_var_kernel<<<1, 1024, 0, s>>>(*st, iter);
  CB_LAUNCH_CHECK("ice_var");
  return 0;
}

}  // extern "C"
