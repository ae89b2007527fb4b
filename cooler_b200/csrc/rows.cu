// Build of the "lane rows" layout swept by ice_sweep_rows_kernel (rows.cuh).
//
// Input: the pixels of one copy as (row, {other, count}) sorted by (row >> RS, other >> col_shift,
// row, other) -- the order the two radix passes of table.far_blocks produce.  A *run* is the set of
// records of one row inside one column tile; runs of one slot (row block, warp row slice, column
// tile) are ordered by decreasing length (two stable radix sorts of the runs) and cut into slices
// of 32: run number 32 g + l of a slot becomes lane l of slice g, and record k of that run lands in
// step k of the slice.  A slice has as many steps as its first (longest) run; shorter runs are
// padded with count-0 records of the same row.  Slots are laid out in (row block, warp, tile) order
// so that the steps of one warp are contiguous over the tiles of a row block.
#include "rows.cuh"
#include "../../include/cooler_b200.h"

namespace cb200 {

constexpr int RS = CB200_FAR_ROW_SHIFT;

__device__ __forceinline__ int rows_warp_of(int row) { return ((row & (FAR_ROWS - 1)) * ROWS_WARPS) >> RS; }

// flags[i] = 1 where a run starts
__global__ void rows_flags_kernel(const int32_t* __restrict__ rows, const int2* __restrict__ vals, int64_t n,
                                  int col_shift, int32_t* __restrict__ flags) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int f = 1;
    if (i > 0) f = (rows[i] != rows[i - 1]) || ((vals[i].x >> col_shift) != (vals[i - 1].x >> col_shift));
    flags[i] = f;
  }
}

// run_start[r] = position of the first record of run r; run_start[n_runs] = n
__global__ void rows_run_starts_kernel(const int32_t* __restrict__ flags, const int64_t* __restrict__ run_id,
                                       int64_t n, int64_t* __restrict__ run_start) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += stride) {
    if (i == n) run_start[run_id[n]] = n;
    else if (flags[i]) run_start[run_id[i]] = i;
  }
}

// stream-order slot id of a run: ((row block - rb0) * n_j + tile - j0) * WARPS + warp (non-decreasing
// along the sorted stream).  key = C - length (ascending = longest first), val = {slot, run}.
__global__ void rows_run_keys_kernel(const int32_t* __restrict__ rows, const int2* __restrict__ vals,
                                     const int64_t* __restrict__ run_start, int64_t n_runs, int col_shift,
                                     int rb0, int j0, int n_j, int32_t* __restrict__ key_out,
                                     int2* __restrict__ val_out, int32_t* __restrict__ err_out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int C = 1 << col_shift;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_runs; r += stride) {
    const int64_t i = run_start[r];
    const int64_t len = run_start[r + 1] - i;
    const int row = rows[i];
    const int j = vals[i].x >> col_shift;
    const int slot = (((row >> RS) - rb0) * n_j + (j - j0)) * ROWS_WARPS + rows_warp_of(row);
    if (len > C) *err_out = 1;                     // duplicate pixels: more records than columns in a tile
    key_out[r] = len >= C ? 0 : (int)(C - len);
    val_out[r] = make_int2(slot, (int)r);
  }
}

__device__ __forceinline__ int64_t rows_storage_slot(int sid, int n_j) {
  // stream order (rb, j, w) -> storage order (rb, w, j)
  const int w = sid % ROWS_WARPS;
  const int bj = sid / ROWS_WARPS;
  const int j = bj % n_j, rb = bj / n_j;
  return ((int64_t)rb * ROWS_WARPS + w) * n_j + j;
}

// Steps of every slot = sum over its slices of the length of the slice's first run.  vals2[p] =
// {C - length, run} of the p-th run in (slot, length descending) order.
__global__ void rows_slot_steps_kernel(const int64_t* __restrict__ slot_run_ptr, const int2* __restrict__ vals2,
                                       int64_t n_slots, int col_shift, int n_j,
                                       int32_t* __restrict__ steps_out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int C = 1 << col_shift;
  for (int64_t sid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; sid < n_slots; sid += stride) {
    const int64_t a = slot_run_ptr[sid], b = slot_run_ptr[sid + 1];
    int steps = 0;
    for (int64_t p = a; p < b; p += 32) steps += C - vals2[p].x;
    steps_out[rows_storage_slot((int)sid, n_j)] = steps;
  }
}

// Destination of every run: record k of the run goes to dst + 32 k.  Also writes the run's padding
// records (same row, column 0 of the tile, count 0) up to the slice length and marks the last step.
__global__ void rows_run_dst_kernel(const int32_t* __restrict__ keys2, const int2* __restrict__ vals2,
                                    const int64_t* __restrict__ slot_run_ptr, const int64_t* __restrict__ slot_base,
                                    const int32_t* __restrict__ rows, const int64_t* __restrict__ run_start,
                                    int64_t n_runs, int col_shift, int n_j, int64_t* __restrict__ dst_out,
                                    int32_t* __restrict__ slice_len_out, int2* __restrict__ px_out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int C = 1 << col_shift;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_runs; p += stride) {
    const int sid = keys2[p];
    const int64_t a = slot_run_ptr[sid];
    const int rank = (int)(p - a);
    const int g = rank >> 5, lane = rank & 31;
    int off = 0;
    for (int gg = 0; gg < g; ++gg) off += C - vals2[a + 32 * gg].x;
    const int L = C - vals2[a + 32 * g].x;
    const int len = C - vals2[p].x;                // clamped lengths (>= C) never need padding
    const int run = vals2[p].y;
    const int64_t dst = (slot_base[rows_storage_slot(sid, n_j)] + off) * 32 + lane;
    dst_out[run] = dst;
    slice_len_out[run] = L;
    if (len < L) {
      const int row = rows[run_start[run]];
      const unsigned x = (unsigned)(row & (FAR_ROWS - 1)) << 16;
      for (int k = len; k < L; ++k)
        px_out[dst + 32 * (int64_t)k] = make_int2((int)(x | (k == L - 1 ? 0x80000000u : 0u)), 0);
    }
    if (p + 1 == slot_run_ptr[sid + 1] && lane < 31) {
      // last run of the slot: the remaining lanes of its slice own no row -> dummy accumulator
      // R + warp, column 0 of the tile, count 0
      const int dummy = (int)((unsigned)(FAR_ROWS + sid % ROWS_WARPS) << 16);
      for (int k = 0; k < L; ++k)
        for (int l = lane + 1; l < 32; ++l) px_out[dst + 32 * (int64_t)k + (l - lane)] = make_int2(dummy, 0);
    }
  }
}

__global__ void rows_scatter_kernel(const int32_t* __restrict__ rows, const int2* __restrict__ vals,
                                    const int64_t* __restrict__ run_id, const int64_t* __restrict__ run_start,
                                    const int64_t* __restrict__ dst, const int32_t* __restrict__ slice_len,
                                    int64_t n, int col_shift, int2* __restrict__ px_out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const unsigned cmask = (1u << col_shift) - 1u;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t r = run_id[i + 1] - 1;           // inclusive count of run starts up to i, minus 1
    const int64_t k = i - run_start[r];
    const int2 v = vals[i];
    unsigned x = ((unsigned)(rows[i] & (FAR_ROWS - 1)) << 16) | (((unsigned)v.x & cmask) << 3);   // byte offset into the w tile
    if (k == slice_len[r] - 1) x |= 0x80000000u;
    px_out[dst[r] + 32 * k] = make_int2((int)x, v.y);
  }
}

}  // namespace cb200

using namespace cb200;

static inline unsigned grid_for(int64_t n, int per_block) {
  int64_t blocks = ceil_div64(n > 0 ? n : 1, per_block);
  if (blocks > 148 * 32) blocks = 148 * 32;
  return (unsigned)blocks;
}

extern "C" {

int cb200_rows_warps(void) { return ROWS_WARPS; }

int cb200_rows_flags(const int32_t* rows, const cb200_px* vals, int64_t n, int32_t col_shift,
                     int32_t* flags_out, void* stream) {
  if (n <= 0) return 0;
  rows_flags_kernel<<<grid_for(n, 1024), 256, 0, (cudaStream_t)stream>>>(
      rows, reinterpret_cast<const int2*>(vals), n, col_shift, flags_out);
  CB_LAUNCH_CHECK("rows_flags");
  return 0;
}

int cb200_rows_run_starts(const int32_t* flags, const int64_t* run_id, int64_t n, int64_t* run_start_out,
                          void* stream) {
  rows_run_starts_kernel<<<grid_for(n + 1, 1024), 256, 0, (cudaStream_t)stream>>>(flags, run_id, n, run_start_out);
  CB_LAUNCH_CHECK("rows_run_starts");
  return 0;
}

int cb200_rows_run_keys(const int32_t* rows, const cb200_px* vals, const int64_t* run_start, int64_t n_runs,
                        int32_t col_shift, int32_t rb0, int32_t j0, int32_t n_j, int32_t* key_out,
                        cb200_px* val_out, int32_t* err_out, void* stream) {
  if (n_runs <= 0) return 0;
  rows_run_keys_kernel<<<grid_for(n_runs, 256), 256, 0, (cudaStream_t)stream>>>(
      rows, reinterpret_cast<const int2*>(vals), run_start, n_runs, col_shift, rb0, j0, n_j, key_out,
      reinterpret_cast<int2*>(val_out), err_out);
  CB_LAUNCH_CHECK("rows_run_keys");
  return 0;
}

int cb200_rows_slot_steps(const int64_t* slot_run_ptr, const cb200_px* vals2, int64_t n_slots, int32_t col_shift,
                          int32_t n_j, int32_t* steps_out, void* stream) {
  if (n_slots <= 0) return 0;
  rows_slot_steps_kernel<<<grid_for(n_slots, 256), 256, 0, (cudaStream_t)stream>>>(
      slot_run_ptr, reinterpret_cast<const int2*>(vals2), n_slots, col_shift, n_j, steps_out);
  CB_LAUNCH_CHECK("rows_slot_steps");
  return 0;
}

int cb200_rows_place(const int32_t* keys2, const cb200_px* vals2, const int64_t* slot_run_ptr,
                     const int64_t* slot_base, const int32_t* rows,
                     const cb200_px* vals, const int64_t* run_id, const int64_t* run_start, int64_t n_runs,
                     int64_t n, int32_t col_shift, int32_t n_j, int64_t* dst_scratch, int32_t* len_scratch,
                     cb200_px* px_out, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  int2* out = reinterpret_cast<int2*>(px_out);
  if (n_runs <= 0) return 0;
  rows_run_dst_kernel<<<grid_for(n_runs, 256), 256, 0, s>>>(
      keys2, reinterpret_cast<const int2*>(vals2), slot_run_ptr, slot_base, rows, run_start, n_runs, col_shift,
      n_j, dst_scratch, len_scratch, out);
  CB_LAUNCH_CHECK("rows_run_dst");
  rows_scatter_kernel<<<grid_for(n, 1024), 256, 0, s>>>(rows, reinterpret_cast<const int2*>(vals), run_id,
                                                        run_start, dst_scratch, len_scratch, n, col_shift, out);
  CB_LAUNCH_CHECK("rows_scatter");
  return 0;
}

}  // extern "C"
