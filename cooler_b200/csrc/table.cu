// Pixel-table build kernels: packing, warp-tile row index, indptr from sorted
// ids, static filters by stable compaction, integer (prefilter) marginals.
#include <cstdio>
#include "walk.cuh"
#include "far.cuh"
#include "../../include/cooler_b200.h"

namespace cb200 {

static thread_local char g_err[512] = "";

void set_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), where);
}
void set_error_msg(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg); }

// ------------------------------------------------------------------ pack ----
template <typename TBin>
__global__ void pack_pixels_kernel(const TBin* __restrict__ bin2, const int32_t* __restrict__ count,
                                   int64_t nnz, int2* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride)
    out[i] = make_int2((int)bin2[i], count[i]);
  // zero the pad element so that 128-bit loads of the last pair are defined
  if (blockIdx.x == 0 && threadIdx.x == 0 && (nnz & 1)) out[nnz] = make_int2(0, 0);
}

// ------------------------------------------------------------- tile rows ----
__global__ void tile_rows_kernel(const int64_t* __restrict__ indptr, int n_rows, int64_t nnz,
                                 int64_t n_tiles, int32_t* __restrict__ tile_row) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  const int64_t p = t * (int64_t)WARP_TILE;
  // largest r in [0, n_rows) with indptr[r] <= p   (upper_bound - 1)
  int lo = 0, hi = n_rows;  // invariant: indptr[lo] <= p, indptr[hi] > p  (indptr[n_rows] = nnz > p)
  while (hi - lo > 1) {
    int mid = lo + ((hi - lo) >> 1);
    if (indptr[mid] <= p) lo = mid; else hi = mid;
  }
  tile_row[t] = lo;
}

// --------------------------------------------------- indptr from sorted ----
__global__ void indptr_from_sorted_kernel(const int32_t* __restrict__ keys, int64_t n, int n_rows,
                                          int key_base, int shift, int64_t* __restrict__ indptr) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p <= n; p += stride) {
    const int prev = (p == 0) ? -1 : ((keys[p - 1] - key_base) >> shift);
    const int cur = (p == n) ? n_rows : ((keys[p] - key_base) >> shift);
    for (int r = prev + 1; r <= cur; ++r) indptr[r] = p;
  }
}

// ---------------------------------------------------------- strip helpers ----
// (key, {a, v}) -> (a, {key, v}): makes the gathered index the sort key.
__global__ void swap_key_other_kernel(const int32_t* __restrict__ keys, const int2* __restrict__ vals,
                                      int64_t n, int32_t* __restrict__ keys_out,
                                      int2* __restrict__ vals_out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int2 v = vals[i];
    keys_out[i] = v.x;
    vals_out[i] = make_int2(keys[i], v.y);
  }
}

// Elements are grouped by strip = key >> shift (tight packing, strip s starts at
// strip_start[s]).  Writes px {key, count} and the row id of every element at
// p + pad_before[s], which makes every strip start at an even (16-byte aligned)
// position (the pad record between strips is never used: walkers stop at nnz).
__global__ void strip_finish_kernel(const int32_t* __restrict__ keys, const int2* __restrict__ vals,
                                    int64_t n, int shift,
                                    const int64_t* __restrict__ pad_before,
                                    int2* __restrict__ px_out, int32_t* __restrict__ rows_out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int k = keys[i];
    const int s = k >> shift;
    const int2 v = vals[i];
    const int64_t o = i + pad_before[s];
    px_out[o] = make_int2(k, v.y);
    rows_out[o] = v.x;
  }
}

// Far-field blocks: see cb200_far_copy in the header.  The input is sorted by
// (row >> RS, other >> col_shift, row, other); every element becomes an 8-byte record with
// block-local coordinates plus its slot id (row block, column tile, warp row slice).
__global__ void far_finish_kernel(const int32_t* __restrict__ rows, const int2* __restrict__ vals,
                                  int64_t n, int col_shift, int rb0, int j0, int n_j, int warps,
                                  int2* __restrict__ px_out, int32_t* __restrict__ slot_out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int cmask = (1 << col_shift) - 1;
  constexpr int RS = CB200_FAR_ROW_SHIFT;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int r = rows[i];
    const int2 v = vals[i];
    const int lrow = r & ((1 << RS) - 1);
    px_out[i] = make_int2((far_swizzle(lrow, warps) << 16) | (v.x & cmask), v.y);
    slot_out[i] = (((r >> RS) - rb0) * n_j + ((v.x >> col_shift) - j0)) * warps + ((lrow * warps) >> RS);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {       // defined pad records (never used: guarded by range checks)
    px_out[n] = make_int2(0, 0);
    px_out[n + 1] = make_int2(0, 0);
  }
}

// Slot s holds records [ptr_rec[s], ptr_rec[s+1]) of the sorted stream; it becomes
// nch[s] = ceil(len / CHUNK) lane-interleaved chunks.
__global__ void far_slot_chunks_kernel(const int64_t* __restrict__ ptr_rec, int64_t n_slots,
                                       int32_t* __restrict__ nch) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += stride)
    nch[s] = (int32_t)((ptr_rec[s + 1] - ptr_rec[s] + CHUNK - 1) / CHUNK);
}

__global__ void far_fill_kernel(int2* __restrict__ px_out, int64_t n_total) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_total; i += stride)
    px_out[i] = make_int2((int)((1u << CB200_FAR_ROW_SHIFT) << 16), 0);   // padding record: row id R (dummy), count 0
}

// Record i of slot s (j-th of the slot) goes to lane j / L, position j % L of that lane's
// piece, L = 2 * nch(s) records per lane; lane pieces are interleaved in units of one int4
// (2 records): chunk c of the slot holds records 2c, 2c+1 of every lane.
__global__ void far_interleave_kernel(const int2* __restrict__ rec, const int32_t* __restrict__ slot,
                                      const int64_t* __restrict__ ptr_rec,
                                      const int64_t* __restrict__ cptr, int64_t n,
                                      int2* __restrict__ px_out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int s = slot[i];
    const int j = (int)(i - ptr_rec[s]);
    const int64_t c0 = cptr[s];
    const int L = 2 * (int)(cptr[s + 1] - c0);
    const int lane = j / L, idx = j - lane * L;
    px_out[(c0 + (idx >> 1)) * CHUNK + 2 * lane + (idx & 1)] = rec[i];
  }
}

// ---- compact (4-byte) far-field records: {swz(row & (R - 1)) : 12 | other & (C - 1) : 13 | count : 7}
// (R <= 4096, C <= 8192).  98 % of the counts of a 1 kb matrix are <= 3; a count above 127 becomes
// ceil(count / 127) consecutive records of the same cell (same row, same column: they extend the
// lane's run and add up to the same product), a count of 0 is a record that adds nothing.
// pieces[i] = records that element i needs; flags[0] |= 1 if any count > 127, flags[1] |= 1 if any
// count < 0 (such a stream keeps the 8-byte records).
__global__ void far_pieces_kernel(const int2* __restrict__ vals, int64_t n, int32_t* __restrict__ pieces,
                                  int32_t* __restrict__ flags) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int big = 0, neg = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int c = vals[i].y;
    neg |= (c < 0);
    const int p = (c <= 127) ? 1 : (int)(((int64_t)c + 126) / 127);
    big |= (p > 1);
    pieces[i] = p;
  }
  big = __syncthreads_or(big);
  neg = __syncthreads_or(neg);
  if (threadIdx.x == 0) {
    if (big) atomicOr(flags, 1);
    if (neg) atomicOr(flags + 1, 1);
  }
}

// element i -> records [off[i], off[i+1]) of the expanded stream (order preserved)
__global__ void far_expand_kernel(const int32_t* __restrict__ rows, const int2* __restrict__ vals, int64_t n,
                                  const int64_t* __restrict__ off, int32_t* __restrict__ rows_out,
                                  int2* __restrict__ vals_out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int r = rows[i];
    const int2 v = vals[i];
    int left = v.y;
    for (int64_t o = off[i]; o < off[i + 1]; ++o) {
      const int c = left > 127 ? 127 : left;
      left -= c;
      rows_out[o] = r;
      vals_out[o] = make_int2(v.x, c);
    }
  }
}

__global__ void far_finish4_kernel(const int32_t* __restrict__ rows, const int2* __restrict__ vals,
                                   int64_t n, int col_shift, int rb0, int j0, int n_j, int warps,
                                   uint32_t* __restrict__ rec_out, int32_t* __restrict__ slot_out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int cmask = (1 << col_shift) - 1;
  constexpr int RS = CB200_FAR_ROW_SHIFT;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int r = rows[i];
    const int2 v = vals[i];
    const int lrow = r & ((1 << RS) - 1);
    rec_out[i] = ((uint32_t)far_swizzle(lrow, warps) << 20) | ((uint32_t)(v.x & cmask) << 7) | (uint32_t)(v.y & 127);
    slot_out[i] = (((r >> RS) - rb0) * n_j + ((v.x >> col_shift) - j0)) * warps + ((lrow * warps) >> RS);
  }
}

// Same lane-interleaved layout as far_interleave_kernel with 4-byte records (a chunk is 256 bytes,
// one 64-bit load per lane).
__global__ void far_interleave4_kernel(const uint32_t* __restrict__ rec, const int32_t* __restrict__ slot,
                                       const int64_t* __restrict__ ptr_rec, const int64_t* __restrict__ cptr,
                                       int64_t n, uint32_t* __restrict__ px_out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int s = slot[i];
    const int j = (int)(i - ptr_rec[s]);
    const int64_t c0 = cptr[s];
    const int L = 2 * (int)(cptr[s + 1] - c0);
    const int lane = j / L, idx = j - lane * L;
    px_out[(c0 + (idx >> 1)) * CHUNK + 2 * lane + (idx & 1)] = rec[i];
  }
}

// Padding of the last chunk(s) of every slot: count 0 on the row of the slot's LAST record, so a
// padding record neither closes a run nor opens a new one (there is no spare row id in 12 bits).
__global__ void far_pad4_kernel(const uint32_t* __restrict__ rec, const int64_t* __restrict__ ptr_rec,
                                const int64_t* __restrict__ cptr, int64_t n_slots, uint32_t* __restrict__ px_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t s = warp; s < n_slots; s += n_warps) {
    const int64_t r0 = ptr_rec[s], r1 = ptr_rec[s + 1];
    const int64_t c0 = cptr[s];
    const int nch = (int)(cptr[s + 1] - c0);
    const int n_s = (int)(r1 - r0);
    if (nch == 0 || n_s == nch * CHUNK) continue;
    const uint32_t pad = rec[r1 - 1] & 0xFFF00000u;
    const int L = 2 * nch;
    for (int j = n_s + lane; j < nch * CHUNK; j += 32) {
      const int ln = j / L, idx = j - ln * L;
      px_out[(c0 + (idx >> 1)) * CHUNK + 2 * ln + (idx & 1)] = pad;
    }
  }
}

// ---------------------------------------------------------------- filter ----
struct KeepRule {
  const int32_t* row_lo;
  const int32_t* row_hi;
  int ndiag;
  int keep;
  int band_shift;
  int row_base;      // bin id of local row 0 (row chunks / shards)
  __device__ __forceinline__ bool operator()(int r_local, int c) const {
    const int r = r_local + row_base;
    if (ndiag > 0) {
      int d = r - c;
      if (d < 0) d = -d;
      if (d < ndiag) return false;
    }
    if (keep == CB200_KEEP_ALL) return true;
    if (keep == CB200_KEEP_EARLIER_CHROM) return c < row_lo[r];
    if (keep == CB200_KEEP_BAND_NEAR || keep == CB200_KEEP_BAND_FAR) {
      int ds = (c >> band_shift) - (r >> band_shift);
      if (ds < 0) ds = -ds;
      return (keep == CB200_KEEP_BAND_NEAR) ? (ds <= 1) : (ds > 1);
    }
    const bool cis = (c >= row_lo[r]) && (c < row_hi[r]);
    return (keep == CB200_KEEP_CIS) ? cis : !cis;
  }
};

struct CountKept {
  KeepRule rule;
  int n;
  __device__ __forceinline__ void operator()(int64_t, bool ok0, bool ok1, int r0, int r1, int4 q) {
    n += (ok0 && rule(r0, q.x)) ? 1 : 0;
    n += (ok1 && rule(r1, q.z)) ? 1 : 0;
  }
};

__global__ void __launch_bounds__(BLOCK_THREADS)
filter_count_kernel(const int2* __restrict__ px, const int64_t* __restrict__ indptr,
                    const int32_t* __restrict__ tile_row, int64_t nnz, int64_t n_tiles,
                    KeepRule rule, int32_t* __restrict__ tile_count) {
  const int64_t t = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (t >= n_tiles) return;
  CountKept f{rule, 0};
  walk_map_tile(px, indptr, nnz, t, tile_row[t], f);
  int tot = warp_sum(f.n);
  if ((threadIdx.x & 31) == 0) tile_count[t] = tot;
}

struct WriteKept {
  KeepRule rule;
  int64_t base;      // next output position (warp-uniform)
  int2* px_out;
  int32_t* row_out;
  int32_t* tkey_out;
  int2* tval_out;
  __device__ __forceinline__ void emit(int64_t o, int r_local, int c, int v) {
    const int r = r_local + rule.row_base;         // absolute bin id
    px_out[o] = make_int2(c, v);
    if (row_out) row_out[o] = r;
    if (tkey_out) tkey_out[o] = c;
    if (tval_out) tval_out[o] = make_int2(r, v);
  }
  __device__ __forceinline__ void operator()(int64_t, bool ok0, bool ok1, int r0, int r1, int4 q) {
    const int lane = threadIdx.x & 31;
    const bool k0 = ok0 && rule(r0, q.x);
    const bool k1 = ok1 && rule(r1, q.z);
    const unsigned b0 = __ballot_sync(0xffffffffu, k0);
    const unsigned b1 = __ballot_sync(0xffffffffu, k1);
    const unsigned lt = (1u << lane) - 1u;
    int64_t o = base + __popc(b0 & lt) + __popc(b1 & lt);
    if (k0) emit(o, r0, q.x, q.y);
    if (k1) emit(o + (k0 ? 1 : 0), r1, q.z, q.w);
    base += __popc(b0) + __popc(b1);
  }
};

__global__ void __launch_bounds__(BLOCK_THREADS)
filter_write_kernel(const int2* __restrict__ px, const int64_t* __restrict__ indptr,
                    const int32_t* __restrict__ tile_row, int64_t nnz, int64_t n_tiles,
                    KeepRule rule, const int64_t* __restrict__ tile_off, int2* __restrict__ px_out,
                    int32_t* __restrict__ row_out, int32_t* __restrict__ tkey_out,
                    int2* __restrict__ tval_out) {
  const int64_t t = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (t >= n_tiles) return;
  WriteKept f{rule, tile_off[t], px_out, row_out, tkey_out, tval_out};
  walk_map_tile(px, indptr, nnz, t, tile_row[t], f);
  // pad element after the last kept pixel (total = tile_off[n_tiles])
  if (t == n_tiles - 1 && (threadIdx.x & 31) == 0 && (f.base & 1)) px_out[f.base] = make_int2(0, 0);
}

// ---------------------------------------------------- integer marginals ----
struct NnzSumOp {
  typedef I64x2 T;
  __device__ __forceinline__ void at_row(int) {}
  __device__ __forceinline__ I64x2 value(int, int count) const {
    return I64x2{count != 0 ? 1 : 0, (long long)count};
  }
};

__global__ void __launch_bounds__(BLOCK_THREADS)
marginals_sweep_kernel(const int2* __restrict__ px, const int64_t* __restrict__ indptr,
                       const int32_t* __restrict__ tile_row, int n_rows, int64_t nnz, int64_t n_tiles,
                       I64x2* __restrict__ direct, I64x2* __restrict__ pieces) {
  const int64_t t = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (t >= n_tiles) return;
  NnzSumOp op;
  walk_reduce_tile<NnzSumOp, 4>(px, indptr, n_rows, nnz, t, tile_row, op, direct, pieces);
}

// direct is laid over (nnz_out, sum_out) as scratch is not possible (interleaved
// I64x2), so the sweep writes direct rows into pieces-side scratch `direct`.
__global__ void marginals_finish_kernel(const int64_t* __restrict__ indptr, int64_t nnz, int n_rows,
                                        const I64x2* __restrict__ direct,
                                        const I64x2* __restrict__ pieces,
                                        int64_t* __restrict__ nnz_out, int64_t* __restrict__ sum_out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  I64x2 v = row_total<I64x2>(r, indptr, nnz, direct, pieces);
  nnz_out[r] = v.a;
  sum_out[r] = v.b;
}

}  // namespace cb200

namespace cb200 {

// Floating-point count columns: the build carries the position of every pixel in the stored table in
// the count field; this attaches the values in stream order and relabels the field with the position
// inside the sub-copy (see cb200_attach_values in the header).
__global__ void attach_values_kernel(int2* __restrict__ px, int64_t n, const double* __restrict__ values,
                                     int64_t n_values, double* __restrict__ val_out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int2 v = px[k];
  val_out[k] = ((unsigned long long)(long long)v.y < (unsigned long long)n_values) ? values[v.y] : 0.0;
  px[k] = make_int2(v.x, (int)k);
}

__global__ void values_nonzero_kernel(const double* __restrict__ values, int64_t n, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = values[i] != 0.0 ? 1.0 : 0.0;
}

// cis_only on tables that store pixels (r, c) with c in a chromosome BEFORE r's (square storage): which
// (chromosome of r, chromosome of c) pairs occur at all, and which occur with a masked column bin (see
// cb200_cis_touch in the header).  Every thread stores the same value 1: no atomics needed.
__global__ void cis_touch_kernel(const int2* __restrict__ px, const int32_t* __restrict__ rows, int64_t n,
                                 const int32_t* __restrict__ bin_chrom, int n_chroms,
                                 const double* __restrict__ bias, uint8_t* __restrict__ touch_any,
                                 uint8_t* __restrict__ touch_masked) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    const int c = px[k].x;
    const int64_t cell = (int64_t)bin_chrom[rows[k]] * n_chroms + bin_chrom[c];
    touch_any[cell] = 1;
    const double b = bias[c];
    if (b == 0.0 || b != b) touch_masked[cell] = 1;
  }
}

}  // namespace cb200

using namespace cb200;

extern "C" {

int cb200_abi_version(void) { return 3; }
const char* cb200_last_error(void) { return g_err; }
int cb200_set_device(int device) {
  CB_CHECK(cudaSetDevice(device));
  return 0;
}
int cb200_warp_tile(void) { return WARP_TILE; }

int cb200_pack_pixels(const int64_t* bin2_id, const int32_t* count, int64_t nnz, cb200_px* px_out,
                      void* stream) {
  if (nnz <= 0) return 0;
  int64_t blocks = ceil_div64(nnz, 256 * 8);
  if (blocks > 148 * 32) blocks = 148 * 32;
  pack_pixels_kernel<int64_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      bin2_id, count, nnz, reinterpret_cast<int2*>(px_out));
  CB_LAUNCH_CHECK("pack_pixels");
  return 0;
}

int cb200_pack_pixels_i32(const int32_t* bin2_id, const int32_t* count, int64_t nnz,
                          cb200_px* px_out, void* stream) {
  if (nnz <= 0) return 0;
  int64_t blocks = ceil_div64(nnz, 256 * 8);
  if (blocks > 148 * 32) blocks = 148 * 32;
  pack_pixels_kernel<int32_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      bin2_id, count, nnz, reinterpret_cast<int2*>(px_out));
  CB_LAUNCH_CHECK("pack_pixels_i32");
  return 0;
}

int cb200_tile_rows(const int64_t* indptr, int32_t n_rows, int64_t nnz, int32_t* tile_row,
                    void* stream) {
  const int64_t n_tiles = ceil_div64(nnz, WARP_TILE);
  if (n_tiles <= 0) return 0;
  tile_rows_kernel<<<(unsigned)ceil_div64(n_tiles, 256), 256, 0, (cudaStream_t)stream>>>(
      indptr, n_rows, nnz, n_tiles, tile_row);
  CB_LAUNCH_CHECK("tile_rows");
  return 0;
}

int cb200_indptr_from_sorted(const int32_t* keys, int64_t n, int32_t n_rows, int32_t key_base,
                             int32_t shift, int64_t* indptr, void* stream) {
  int64_t blocks = ceil_div64(n + 1, 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  indptr_from_sorted_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      keys, n, n_rows, key_base, shift, indptr);
  CB_LAUNCH_CHECK("indptr_from_sorted");
  return 0;
}

int cb200_swap_key_other(const int32_t* keys, const cb200_px* vals, int64_t n, int32_t* keys_out,
                         cb200_px* vals_out, void* stream) {
  if (n <= 0) return 0;
  int64_t blocks = ceil_div64(n, 256 * 4);
  if (blocks > 148 * 32) blocks = 148 * 32;
  swap_key_other_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      keys, reinterpret_cast<const int2*>(vals), n, keys_out, reinterpret_cast<int2*>(vals_out));
  CB_LAUNCH_CHECK("swap_key_other");
  return 0;
}

int cb200_strip_finish(const int32_t* keys, const cb200_px* vals, int64_t n, int32_t shift,
                       const int64_t* pad_before, cb200_px* px_out, int32_t* rows_out,
                       void* stream) {
  if (n <= 0) return 0;
  int64_t blocks = ceil_div64(n, 256 * 4);
  if (blocks > 148 * 32) blocks = 148 * 32;
  strip_finish_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      keys, reinterpret_cast<const int2*>(vals), n, shift, pad_before,
      reinterpret_cast<int2*>(px_out), rows_out);
  CB_LAUNCH_CHECK("strip_finish");
  return 0;
}

int cb200_far_finish(const int32_t* rows, const cb200_px* vals, int64_t n, int32_t col_shift,
                     int32_t rb0, int32_t j0, int32_t n_j, int32_t warps, cb200_px* px_out,
                     int32_t* slot_out, void* stream) {
  if (col_shift < 1 || col_shift > 16 || warps != 16 || n_j <= 0) {
    set_error_msg("far_finish: bad arguments");
    return -1;
  }
  int64_t blocks = ceil_div64(n > 0 ? n : 1, 256 * 4);
  if (blocks > 148 * 32) blocks = 148 * 32;
  far_finish_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      rows, reinterpret_cast<const int2*>(vals), n, col_shift, rb0, j0, n_j, warps,
      reinterpret_cast<int2*>(px_out), slot_out);
  CB_LAUNCH_CHECK("far_finish");
  return 0;
}

int cb200_far_slot_chunks(const int64_t* ptr_rec, int64_t n_slots, int32_t* nch_out, void* stream) {
  if (n_slots <= 0) return 0;
  int64_t blocks = ceil_div64(n_slots, 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  far_slot_chunks_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(ptr_rec, n_slots, nch_out);
  CB_LAUNCH_CHECK("far_slot_chunks");
  return 0;
}

int cb200_far_interleave(const cb200_px* rec, const int32_t* slot, const int64_t* ptr_rec,
                         const int64_t* chunk_ptr, int64_t n, int64_t total_chunks, cb200_px* px_out,
                         void* stream) {
  const int64_t n_total = total_chunks * CHUNK + 2;
  int64_t blocks = ceil_div64(n_total, 256 * 4);
  if (blocks > 148 * 32) blocks = 148 * 32;
  far_fill_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<int2*>(px_out), n_total);
  CB_LAUNCH_CHECK("far_fill");
  if (n <= 0) return 0;
  blocks = ceil_div64(n, 256 * 4);
  if (blocks > 148 * 32) blocks = 148 * 32;
  far_interleave_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const int2*>(rec), slot, ptr_rec, chunk_ptr, n, reinterpret_cast<int2*>(px_out));
  CB_LAUNCH_CHECK("far_interleave");
  return 0;
}

int cb200_far_pieces(const cb200_px* vals, int64_t n, int32_t* pieces_out, int32_t* flags, void* stream) {
  CB_CHECK(cudaMemsetAsync(flags, 0, 2 * sizeof(int32_t), (cudaStream_t)stream));
  if (n <= 0) return 0;
  int64_t blocks = ceil_div64(n, 256 * 4);
  if (blocks > 148 * 32) blocks = 148 * 32;
  far_pieces_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const int2*>(vals), n,
                                                                       pieces_out, flags);
  CB_LAUNCH_CHECK("far_pieces");
  return 0;
}

int cb200_far_expand(const int32_t* rows, const cb200_px* vals, int64_t n, const int64_t* offsets,
                     int32_t* rows_out, cb200_px* vals_out, void* stream) {
  if (n <= 0) return 0;
  int64_t blocks = ceil_div64(n, 256 * 4);
  if (blocks > 148 * 32) blocks = 148 * 32;
  far_expand_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      rows, reinterpret_cast<const int2*>(vals), n, offsets, rows_out, reinterpret_cast<int2*>(vals_out));
  CB_LAUNCH_CHECK("far_expand");
  return 0;
}

int cb200_far_finish4(const int32_t* rows, const cb200_px* vals, int64_t n, int32_t col_shift,
                      int32_t rb0, int32_t j0, int32_t n_j, int32_t warps, uint32_t* rec_out,
                      int32_t* slot_out, void* stream) {
  if (col_shift < 1 || col_shift > 13 || CB200_FAR_ROW_SHIFT > 12 || warps != 16 || n_j <= 0) {
    set_error_msg("far_finish4: compact records need col_shift <= 13 and 4096-row blocks at most");
    return -1;
  }
  int64_t blocks = ceil_div64(n > 0 ? n : 1, 256 * 4);
  if (blocks > 148 * 32) blocks = 148 * 32;
  far_finish4_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      rows, reinterpret_cast<const int2*>(vals), n, col_shift, rb0, j0, n_j, warps, rec_out, slot_out);
  CB_LAUNCH_CHECK("far_finish4");
  return 0;
}

int cb200_far_interleave4(const uint32_t* rec, const int32_t* slot, const int64_t* ptr_rec,
                          const int64_t* chunk_ptr, int64_t n, int64_t n_slots, int64_t total_chunks,
                          uint32_t* px_out, void* stream) {
  (void)total_chunks;
  if (n <= 0) return 0;
  int64_t blocks = ceil_div64(n, 256 * 4);
  if (blocks > 148 * 32) blocks = 148 * 32;
  far_interleave4_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rec, slot, ptr_rec, chunk_ptr, n, px_out);
  CB_LAUNCH_CHECK("far_interleave4");
  blocks = ceil_div64(n_slots, 8);
  if (blocks > 148 * 32) blocks = 148 * 32;
  far_pad4_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rec, ptr_rec, chunk_ptr, n_slots, px_out);
  CB_LAUNCH_CHECK("far_pad4");
  return 0;
}

int cb200_filter_count(const cb200_px* px, const int64_t* indptr, const int32_t* tile_row,
                       int32_t n_rows, int64_t nnz, const int32_t* row_lo, const int32_t* row_hi,
                       int32_t ndiag, int32_t keep, int32_t band_shift, int32_t row_base,
                       int32_t* tile_count, void* stream) {
  (void)n_rows;
  const int64_t n_tiles = ceil_div64(nnz, WARP_TILE);
  if (n_tiles <= 0) return 0;
  KeepRule rule{row_lo, row_hi, ndiag, keep, band_shift, row_base};
  filter_count_kernel<<<(unsigned)ceil_div64(n_tiles, WARPS_PER_BLOCK), BLOCK_THREADS, 0,
                        (cudaStream_t)stream>>>(reinterpret_cast<const int2*>(px), indptr, tile_row,
                                                nnz, n_tiles, rule, tile_count);
  CB_LAUNCH_CHECK("filter_count");
  return 0;
}

int cb200_filter_write(const cb200_px* px, const int64_t* indptr, const int32_t* tile_row,
                       int32_t n_rows, int64_t nnz, const int32_t* row_lo, const int32_t* row_hi,
                       int32_t ndiag, int32_t keep, int32_t band_shift, int32_t row_base,
                       const int64_t* tile_off, cb200_px* px_out, int32_t* row_out, int32_t* tkey_out,
                       cb200_px* tval_out, void* stream) {
  (void)n_rows;
  const int64_t n_tiles = ceil_div64(nnz, WARP_TILE);
  if (n_tiles <= 0) return 0;
  KeepRule rule{row_lo, row_hi, ndiag, keep, band_shift, row_base};
  filter_write_kernel<<<(unsigned)ceil_div64(n_tiles, WARPS_PER_BLOCK), BLOCK_THREADS, 0,
                        (cudaStream_t)stream>>>(
      reinterpret_cast<const int2*>(px), indptr, tile_row, nnz, n_tiles, rule, tile_off,
      reinterpret_cast<int2*>(px_out), row_out, tkey_out, reinterpret_cast<int2*>(tval_out));
  CB_LAUNCH_CHECK("filter_write");
  return 0;
}

int cb200_marginals_i64(const cb200_copy* copy, int32_t n_rows, int64_t* nnz_out, int64_t* sum_out,
                        int64_t* pieces, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nnz = copy->nnz;
  const int64_t n_tiles = ceil_div64(nnz, WARP_TILE);
  // scratch layout: pieces[2*n_tiles] (I64x2) then direct[n_rows] (I64x2)
  I64x2* pc = reinterpret_cast<I64x2*>(pieces);
  I64x2* direct = pc + 2 * n_tiles;
  if (n_tiles > 0) {
    marginals_sweep_kernel<<<(unsigned)ceil_div64(n_tiles, WARPS_PER_BLOCK), BLOCK_THREADS, 0, st>>>(
        reinterpret_cast<const int2*>(copy->px), copy->indptr, copy->tile_row, n_rows, nnz, n_tiles,
        direct, pc);
    CB_LAUNCH_CHECK("marginals_sweep");
  }
  if (n_rows > 0) {
    marginals_finish_kernel<<<(unsigned)ceil_div64(n_rows, 256), 256, 0, st>>>(
        copy->indptr, nnz, n_rows, direct, pc, nnz_out, sum_out);
    CB_LAUNCH_CHECK("marginals_finish");
  }
  return 0;
}

int cb200_attach_values(cb200_px* px, int64_t n, const double* values, int64_t n_values, double* val_out,
                        void* stream) {
  if (n <= 0) return 0;
  if (n >= ((int64_t)1 << 31)) {
    set_error_msg("attach_values: a sub-copy with floating-point values must hold fewer than 2^31 pixels");
    return -1;
  }
  attach_values_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<int2*>(px), n, values, n_values, val_out);
  CB_LAUNCH_CHECK("attach_values");
  return 0;
}

int cb200_values_nonzero(const double* values, int64_t n, double* out, void* stream) {
  if (n <= 0) return 0;
  values_nonzero_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(values, n, out);
  CB_LAUNCH_CHECK("values_nonzero");
  return 0;
}

int cb200_cis_touch(const cb200_px* px, const int32_t* rows, int64_t n, const int32_t* bin_chrom,
                    int32_t n_chroms, const double* bias, uint8_t* touch_any, uint8_t* touch_masked,
                    void* stream) {
  if (n <= 0) return 0;
  if (!px || !rows || !bin_chrom || !bias || !touch_any || !touch_masked || n_chroms <= 0) {
    set_error_msg("cis_touch: bad arguments");
    return -1;
  }
  int64_t blocks = ceil_div64(n, 256 * 4);
  if (blocks > 148 * 16) blocks = 148 * 16;
  cis_touch_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const int2*>(px), rows, n, bin_chrom, n_chroms, bias, touch_any, touch_masked);
  CB_LAUNCH_CHECK("cis_touch");
  return 0;
}

}  // extern "C"
