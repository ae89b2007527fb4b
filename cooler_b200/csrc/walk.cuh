// Row-walking building blocks shared by the ICE, filter and coarsen kernels.
//
// A copy of the pixel table is an array px[nnz] of int2 {other bin, count}
// ordered by row, plus indptr[n_rows+1].  Work is split by PIXELS, not rows:
// warp tile t owns pixels [t*WARP_TILE, (t+1)*WARP_TILE).  Rows are located by
// walking indptr from tile_row[t] (the row holding the tile's first pixel).
//
//   walk_reduce_tile : deterministic warp-level segmented reduction (no atomics).
//       Rows strictly inside the tile are written to direct[row]; the first and
//       last row of the tile (which may continue in neighbouring tiles) are
//       written as pieces[2t], pieces[2t+1]; row_total() reassembles a row.
//   walk_map_tile    : calls a functor for every pixel with its row known.
#pragma once
#include "common.cuh"

namespace cb200 {

// ---- accumulator types ------------------------------------------------------
struct I64x2 {
  long long a, b;
};
__device__ __forceinline__ double acc_zero(double*) { return 0.0; }
__device__ __forceinline__ I64x2 acc_zero(I64x2*) { return I64x2{0, 0}; }
__device__ __forceinline__ double acc_add(double x, double y) { return x + y; }
__device__ __forceinline__ I64x2 acc_add(I64x2 x, I64x2 y) { return I64x2{x.a + y.a, x.b + y.b}; }
__device__ __forceinline__ double acc_wsum(double x) { return warp_sum(x); }
__device__ __forceinline__ I64x2 acc_wsum(I64x2 x) { return I64x2{warp_sum(x.a), warp_sum(x.b)}; }

// Op: struct with  typedef T;  __device__ T value(int other, int count) const;
//
// One warp streams through the consecutive tiles [t_begin, t_end) of a copy.
// The pixel stream is software-pipelined in registers: the U 128-bit loads of
// group g+1 are issued before group g is consumed, across tile boundaries, so
// a warp always has 512*U bytes of the stream in flight next to its gathers.
// A "group" is U chunks (U*64 pixels); WARP_TILE is a multiple of the group.
template <typename Op, int U>
__device__ __forceinline__ void walk_reduce_span(const int2* __restrict__ px,
                                                 const int64_t* __restrict__ indptr,
                                                 const int32_t* __restrict__ tile_row, int n_rows,
                                                 int64_t nnz, int64_t t_begin, int64_t t_end,
                                                 Op& op, typename Op::T* __restrict__ direct,
                                                 typename Op::T* __restrict__ pieces) {
  typedef typename Op::T T;
  constexpr int G = CHUNK * U;                     // pixels per group
  static_assert(WARP_TILE % G == 0, "tile must be a whole number of groups");
  const int lane = threadIdx.x & 31;
  const int64_t pbeg = t_begin * (int64_t)WARP_TILE;
  // all positions below are 32-bit offsets relative to pbeg (a span is a few tiles)
  const int n = (int)(min(t_end * (int64_t)WARP_TILE, nnz) - pbeg);
  const int4* __restrict__ src = reinterpret_cast<const int4*>(px + pbeg) + lane;
  const T zero = acc_zero((T*)nullptr);

  int4 qn[U];                                      // next group (in flight)
#pragma unroll
  for (int u = 0; u < U; ++u)
    qn[u] = (u * CHUNK + 2 * lane < n) ? ld_stream_int4(src + u * (CHUNK / 2)) : make_int4(0, 0, 0, 0);

  int row = tile_row[t_begin];
  // Row ends are read through a register window: lane l holds the end of row wbase + l
  // (relative to pbeg, clamped), the next 32 are prefetched.  Moving to the next row is a
  // ballot + shuffle instead of a dependent global load per row, which matters when rows
  // are short (a few chunks) or mostly empty.
  constexpr unsigned FULL = 0xffffffffu;
  auto load_end = [&](int r) -> int {
    if (r >= n_rows) return 0x7fffffff;
    const int64_t e = indptr[r + 1] - pbeg;
    return e > (int64_t)0x7fffffff ? 0x7fffffff : (int)e;
  };
  int wbase = row;
  int wcur = load_end(wbase + lane);
  int wnext = load_end(wbase + 32 + lane);
  int re = __shfl_sync(FULL, wcur, 0);
  auto advance = [&](int done_to) {                // first later row whose end is > done_to
    while (true) {
      const int k = row - wbase + 1;               // first candidate lane of the window
      unsigned m = __ballot_sync(FULL, wcur > done_to);
      m = (k < 32) ? (m & (FULL << k)) : 0u;
      if (m) {
        const int kk = __ffs(m) - 1;
        row = wbase + kk;
        re = __shfl_sync(FULL, wcur, kk);
        return;
      }
      wbase += 32;
      wcur = wnext;
      wnext = load_end(wbase + 32 + lane);
      row = wbase - 1;
    }
  };
  T acc = zero;
  bool first = true;
  int64_t t = t_begin;
  op.at_row(row);

  for (int c = 0; c < n; c += G) {
    int4 q[U];
#pragma unroll
    for (int u = 0; u < U; ++u) q[u] = qn[u];
    const int cn = c + G;
    if (cn + G <= n) {                             // next group is full: unpredicated loads
#pragma unroll
      for (int u = 0; u < U; ++u) qn[u] = ld_stream_int4(src + (cn >> 1) + u * (CHUNK / 2));
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (cn + u * CHUNK + 2 * lane < n) qn[u] = ld_stream_int4(src + (cn >> 1) + u * (CHUNK / 2));
    }
    T v0[U], v1[U];
    if (cn <= n) {                                 // this group is full
#pragma unroll
      for (int u = 0; u < U; ++u) {
        v0[u] = op.value(q[u].x, q[u].y);
        v1[u] = op.value(q[u].z, q[u].w);
      }
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pos = c + u * CHUNK + 2 * lane;
        v0[u] = (pos < n) ? op.value(q[u].x, q[u].y) : zero;
        v1[u] = (pos + 1 < n) ? op.value(q[u].z, q[u].w) : zero;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int cbeg = c + u * CHUNK;
      if (cbeg >= n) break;
      const int cend = min(cbeg + CHUNK, n);
      if (cend <= re) {                            // whole chunk inside the current row
        acc = acc_add(acc, acc_add(v0[u], v1[u]));
      } else {                                     // one or more rows end inside this chunk
        const int pos = cbeg + 2 * lane;
        T a0 = v0[u], a1 = v1[u];
        while (true) {
          if (pos < re) { acc = acc_add(acc, a0); a0 = zero; }
          if (pos + 1 < re) { acc = acc_add(acc, a1); a1 = zero; }
          if (re >= cend) break;
          T tot = acc_wsum(acc);
          if (lane == 0) {
            if (first) pieces[2 * t] = tot; else direct[row] = tot;
          }
          first = false;
          acc = zero;
          advance(re);                             // next non-empty row
          op.at_row(row);
        }
      }
    }
    if ((cn & (WARP_TILE - 1)) == 0 || cn >= n) {  // tile finished: emit its closing piece
      T tot = acc_wsum(acc);
      if (lane == 0) pieces[2 * t + (first ? 0 : 1)] = tot;
      acc = zero;
      first = true;
      ++t;
      if (cn < n && re <= cn) {                    // the row ended exactly at the tile boundary
        advance(cn);
        op.at_row(row);
      }
    }
  }
}

// Single-tile form (integer marginals and other one-time passes).
template <typename Op, int U>
__device__ __forceinline__ void walk_reduce_tile(const int2* __restrict__ px,
                                                 const int64_t* __restrict__ indptr, int n_rows,
                                                 int64_t nnz, int64_t t,
                                                 const int32_t* __restrict__ tile_row,
                                                 Op& op, typename Op::T* __restrict__ direct,
                                                 typename Op::T* __restrict__ pieces) {
  walk_reduce_span<Op, U>(px, indptr, tile_row, n_rows, nnz, t, t + 1, op, direct, pieces);
}

// Total of row r = pixels [a, b) from the outputs of walk_reduce_tile.
template <typename T>
__device__ __forceinline__ T row_total_ab(int r, int64_t a, int64_t b, int64_t nnz,
                                          const T* __restrict__ direct,
                                          const T* __restrict__ pieces) {
  if (a >= b) return acc_zero((T*)nullptr);
  const int64_t ta = a / WARP_TILE, tb = (b - 1) / WARP_TILE;
  if (ta == tb) {
    if (a == ta * WARP_TILE) return pieces[2 * ta];
    if (b == min((ta + 1) * (int64_t)WARP_TILE, nnz)) return pieces[2 * ta + 1];
    return direct[r];
  }
  T v = (a == ta * WARP_TILE) ? pieces[2 * ta] : pieces[2 * ta + 1];
  for (int64_t t = ta + 1; t <= tb; ++t) v = acc_add(v, pieces[2 * t]);
  return v;
}

template <typename T>
__device__ __forceinline__ T row_total(int r, const int64_t* __restrict__ indptr, int64_t nnz,
                                       const T* __restrict__ direct,
                                       const T* __restrict__ pieces) {
  return row_total_ab<T>(r, indptr[r], indptr[r + 1], nnz, direct, pieces);
}

// F: __device__ void operator()(int64_t pos, bool ok0, bool ok1, int r0, int r1, int4 q)
// called by all 32 lanes for every chunk (warp-collectives allowed inside).
// Lane l sees pixels pos = chunk_begin + 2l (q.x, q.y) and pos + 1 (q.z, q.w).
template <typename F>
__device__ __forceinline__ void walk_map_tile(const int2* __restrict__ px,
                                              const int64_t* __restrict__ indptr, int64_t nnz,
                                              int64_t t, int row0, F& f) {
  const int lane = threadIdx.x & 31;
  const int64_t p0 = t * (int64_t)WARP_TILE;
  const int64_t p1 = min(p0 + (int64_t)WARP_TILE, nnz);
  int row = row0;
  int64_t row_end = indptr[row + 1];
  for (int64_t cbeg = p0; cbeg < p1; cbeg += CHUNK) {
    const int64_t cend = min(cbeg + (int64_t)CHUNK, p1);
    const int64_t pos = cbeg + 2 * lane;
    const bool ok0 = pos < p1, ok1 = pos + 1 < p1;
    int4 q = ok0 ? ld_stream_int4(reinterpret_cast<const int4*>(px + pos)) : make_int4(0, 0, 0, 0);
    int r0 = row, r1 = row;
    if (cend > row_end) {                          // rows change inside the chunk
      int r = row;
      int64_t e = row_end;
      if (ok0) {
        while (pos >= e) { ++r; e = indptr[r + 1]; }
        r0 = r;
        if (ok1) while (pos + 1 >= e) { ++r; e = indptr[r + 1]; }
        r1 = r;
      }
      // continue from the row of the chunk's last pixel
      const int last_lane = (int)((cend - 1 - cbeg) >> 1);
      row = __shfl_sync(0xffffffffu, r, last_lane);
      row_end = indptr[row + 1];
    }
    f(pos, ok0, ok1, r0, r1, q);
  }
}

}  // namespace cb200
