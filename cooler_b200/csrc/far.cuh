// K3 for far-field blocks (see cb200_far_copy in include/cooler_b200.h).
//
// Very large, very sparse matrices (1 kb human Hi-C: 3.1e6 bins, ~1e3 pixels per row spread
// over the whole genome) defeat both forms of the row-walking sweep: gathers of w[other]
// through L1/L2 cost one 32-byte L2 sector per 8-byte pixel (L2-bound at ~0.5 of the HBM
// rate), and column strips need one partial sum per (row, strip) although such a segment
// holds 1-2 pixels.  Here the far pixels are blocked in BOTH dimensions: a persistent CTA owns
// R = 1 << CB200_FAR_ROW_SHIFT rows (4096) whose sums stay in shared memory while it walks along the column tiles of that
// row block, staging the C-wide slice of w of the next tile with cp.async while the current
// one is consumed.  Records carry block-local (row, column) ids, so there is no row pointer
// per (row, tile).
//
// Warp k owns rows [k, k+1) * R / WARPS of the block; the records of one (tile, warp
// slice) = "slot" are sorted by row and cut into 32 equal contiguous pieces, one per lane,
// stored lane-interleaved so that every warp load is one coalesced 512-byte chunk.  A lane
// walks through its piece sequentially: it sums a run of equal rows in a register and adds
// the run total to the row's accumulator when the row changes.  Only the LAST run of a lane
// can continue in the next lane; it is left open and merged once per slot by a warp-level
// segmented scan over the 32 lanes.  A run that ends inside a lane's piece ends nowhere
// else, so every accumulator is updated by exactly one lane per phase (pieces, then the
// merge, separated by __syncwarp): no atomics, fixed summation order,
// bit-reproducible.  Padding records (row id R = a dummy accumulator) fill the last chunk
// of a slot.
//
// Two record formats (cb200_ice.far_rec_bytes): 8 bytes {swz row << 16 | column, count} and the
// compact 4 bytes {swz row : 12 | column : 13 | count : 7} (counts above 127 are split into several
// records of the same cell when the table is built).  The kernel is bound by the shared-memory
// data pipe, not by HBM (ncu, 8-byte records, 1 kb rank-0 shard: 262 M shared wavefronts per
// launch = 6.0 per random 64-bit gather instruction, ideal 2, + 1.5 per predicated accumulator
// load / store; l1tex data pipe 78 % busy, 13 points of that the record stream itself), so halving
// the stream takes load off the limiting pipe: 1.40 -> 1.33 ms on that shard, 1.33 -> 1.23 and
// 1.42 -> 1.25 ms on the rank-3 and rank-7 shards (0.73 / 0.78 / 0.74 of the measured HBM rate in
// terms of the algorithmic 16 B per pixel).  A variant without the per-tile CTA barrier (mbarrier
// ring of three / two tiles, warps drifting up to a tile apart, thread 0 or a 17th warp as the
// producer) was measured SLOWER on the same shard (1.59-1.96 ms): waiting warps spin on
// mbarrier.try_wait through the same shared-memory pipe that bounds the kernel, whereas a warp
// parked at __syncthreads costs nothing -- so the per-tile barrier stays.
#pragma once
#include <type_traits>
#include "walk.cuh"
#include "../../include/cooler_b200.h"

namespace cb200 {

constexpr int FAR_ROWS = 1 << CB200_FAR_ROW_SHIFT;
constexpr unsigned FAR_NOKEY = FAR_ROWS;           // row id of padding records: a dummy accumulator slot
constexpr int FAR_PAD_X = (int)(FAR_NOKEY << 16);  // first word of a padding record
constexpr int FAR_ACC = FAR_ROWS + 8;              // accumulators incl. the dummy slot (keeps 64-byte alignment)

__device__ __forceinline__ double far_i32_to_f64(int v) {      // exact, see i32_to_f64 in ice.cu
  return __hiloint2double(0x43300000, v ^ 0x80000000) - 4503601774854144.0;
}

// mbarrier + 1-D bulk copy (TMA) staging of a column tile of w: one elected thread issues a single
// cp.async.bulk for the whole contiguous slice, which keeps the staging off the LSU instruction
// stream that bounds this kernel.
__device__ __forceinline__ void mbar_init(unsigned bar_s, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_s), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar_s, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar_s, unsigned phase) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar_s), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst_s, const void* src, unsigned bytes, unsigned bar_s) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_s), "l"(src), "r"(bytes), "r"(bar_s) : "memory");
}
// shared-memory accesses by 32-bit shared-space address (keeps address arithmetic to one shift/add)
__device__ __forceinline__ double lds_f64(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f64(unsigned a, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}

// Bank swizzle of the block-local row id stored in the records (applied by cb200_far_finish):
// a lane's piece covers about RW/32 consecutive rows of the warp's slice of RW rows, lane l
// roughly rows [l, l+1) * RW/32.  Row lr of the slice is kept at accumulator
// (lr % (RW/32)) * 32 + lr / (RW/32), so lanes that are in step hit 32 different words and
// the run-end updates are (nearly) free of shared-memory bank conflicts.
__host__ __device__ inline int far_swizzle(int lrow, int warps) {
  const int rw = FAR_ROWS / warps, g = rw / 32;
  const int lr = lrow & (rw - 1);
  return (lrow - lr) + ((lr & (g - 1)) << 5) + lr / g;
}

// One record of a lane's piece: (k, v) = (block-local row, w[col] * count).  `key`/`val` hold
// the run being summed.  When the row changes the finished run is added to its accumulator;
// the initial state and padding records use row id FAR_ROWS, whose accumulator is a dummy
// slot behind the real ones, so the update needs no special cases.
__device__ __forceinline__ void far_add(unsigned& key, double& val, const unsigned k, const double v,
                                        const unsigned acc_s) {
  const bool changed = (k != key);
  if (changed) {
    const unsigned a = acc_s + (key << 3);
    sts_f64(a, lds_f64(a) + val);
  }
  val = changed ? v : val + v;
  key = k;
}

template <int U>
__device__ __forceinline__ void far_first_group(const int4* __restrict__ px4, const int64_t c0,
                                                const int64_t c1, int4 (&qn)[U], const int lane) {
  const int4* __restrict__ src = px4 + c0 * 32 + lane;
  const int nch = (int)(c1 - c0);
#pragma unroll
  for (int u = 0; u < U; ++u)
    qn[u] = (u < nch) ? ld_stream_int4(src + u * 32) : make_int4(FAR_PAD_X, 0, FAR_PAD_X, 0);
}

// Chunks [c0, c1) of one slot; qn holds the first group (already in flight).
template <int U>
__device__ __forceinline__ void far_slot(const int4* __restrict__ px4, const int64_t c0, const int64_t c1,
                                         int4 (&qn)[U], const unsigned ws_s, const unsigned acc_s,
                                         const int lane) {
  constexpr unsigned F = 0xffffffffu;
  const int nch = (int)(c1 - c0);
  if (nch <= 0) return;                            // warp-uniform
  const int4* __restrict__ src = px4 + c0 * 32 + lane;
  unsigned key = FAR_NOKEY;
  double val = 0.0;
  for (int c = 0; c < nch; c += U) {
    int4 q[U];
#pragma unroll
    for (int u = 0; u < U; ++u) q[u] = qn[u];
    const int cn = c + U;
    if (cn < nch) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        qn[u] = (cn + u < nch) ? ld_stream_int4(src + (cn + u) * 32) : make_int4(FAR_PAD_X, 0, FAR_PAD_X, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      // Never walk into a chunk past the end: its padding would make EVERY lane close its
      // last run at the same instruction, and adjacent lanes may share that run.
      if (c + u >= nch) break;                     // warp-uniform
      const double g0 = lds_f64(ws_s + (((unsigned)q[u].x & 0xFFFFu) << 3));
      const double g1 = lds_f64(ws_s + (((unsigned)q[u].z & 0xFFFFu) << 3));
      far_add(key, val, (unsigned)q[u].x >> 16, g0 * far_i32_to_f64(q[u].y), acc_s);
      far_add(key, val, (unsigned)q[u].z >> 16, g1 * far_i32_to_f64(q[u].w), acc_s);
    }
  }
  // The LAST run of every lane is still open and may continue in the next lanes (whose first
  // run was added above, or is their last run too): merge equal adjacent keys with a
  // segmented scan over the lanes and let the last lane of each run add the total.
  __syncwarp();
  const unsigned pk = __shfl_up_sync(F, key, 1);
  const unsigned nk = __shfl_down_sync(F, key, 1);
  const bool head = (lane == 0) || (pk != key);
  const unsigned heads = __ballot_sync(F, head);
  const int h = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
  const int span = (key != FAR_NOKEY) ? lane - h : 0;
  const int ms = __reduce_max_sync(F, span);
  double x = val;
  for (int d = 1; d <= ms; d <<= 1) {
    const double y = __shfl_up_sync(F, x, d);
    if (lane - d >= h) x += y;
  }
  if (key != FAR_NOKEY && (lane == 31 || nk != key)) {
    const unsigned a = acc_s + (key << 3);
    sts_f64(a, lds_f64(a) + x);
  }
  __syncwarp();
}

// ---- compact records {swz row : 12 | column : 13 | count : 7} (cb200_far_finish4): a chunk of 64
// records is 256 bytes, one 64-bit load per lane.  Padding records repeat the row of the slot's last
// record with count 0, so they extend a run instead of closing one.
__device__ __forceinline__ void far_add4(unsigned& key, double& val, const unsigned x, const double g,
                                         const unsigned acc_s) {
  const unsigned k = x >> 20;
  const unsigned c = x & 127u;
  // exact small int -> double: {0x43300000, c} is 2^52 + c.  A record with count 0 adds exactly
  // nothing, whatever the gathered value is.
  const double v = c ? g * (__hiloint2double(0x43300000, (int)c) - 4503599627370496.0) : 0.0;
  const bool changed = (k != key);
  if (changed) {
    const unsigned a = acc_s + (key << 3);
    sts_f64(a, lds_f64(a) + val);
  }
  val = changed ? v : val + v;
  key = k;
}

template <int U>
__device__ __forceinline__ void far_first_group4(const int2* __restrict__ px2, const int64_t c0,
                                                 const int64_t c1, int2 (&qn)[U], const int lane) {
  const int2* __restrict__ src = px2 + c0 * 32 + lane;
  const int nch = (int)(c1 - c0);
#pragma unroll
  for (int u = 0; u < U; ++u) qn[u] = (u < nch) ? ld_stream_int2(src + u * 32) : make_int2(0, 0);
}

template <int U>
__device__ __forceinline__ void far_slot4(const int2* __restrict__ px2, const int64_t c0, const int64_t c1,
                                          int2 (&qn)[U], const unsigned ws_s, const unsigned acc_s,
                                          const int lane) {
  constexpr unsigned F = 0xffffffffu;
  const int nch = (int)(c1 - c0);
  if (nch <= 0) return;                            // warp-uniform
  const int2* __restrict__ src = px2 + c0 * 32 + lane;
  unsigned key = FAR_NOKEY;
  double val = 0.0;
  for (int c = 0; c < nch; c += U) {
    int2 q[U];
#pragma unroll
    for (int u = 0; u < U; ++u) q[u] = qn[u];
    const int cn = c + U;
    if (cn < nch) {
#pragma unroll
      for (int u = 0; u < U; ++u) qn[u] = (cn + u < nch) ? ld_stream_int2(src + (cn + u) * 32) : make_int2(0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (c + u >= nch) break;                     // warp-uniform
      // byte offset of w[column] in the tile: ((x >> 7) & 0x1FFF) << 3
      const double g0 = lds_f64(ws_s + (((unsigned)q[u].x >> 4) & 0xFFF8u));
      const double g1 = lds_f64(ws_s + (((unsigned)q[u].y >> 4) & 0xFFF8u));
      far_add4(key, val, (unsigned)q[u].x, g0, acc_s);
      far_add4(key, val, (unsigned)q[u].y, g1, acc_s);
    }
  }
  // merge the open last runs of the lanes (see far_slot); every lane holds a real row here
  __syncwarp();
  const unsigned pk = __shfl_up_sync(F, key, 1);
  const unsigned nk = __shfl_down_sync(F, key, 1);
  const bool head = (lane == 0) || (pk != key);
  const unsigned heads = __ballot_sync(F, head);
  const int h = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
  const int span = lane - h;
  const int ms = __reduce_max_sync(F, span);
  double x = val;
  for (int d = 1; d <= ms; d <<= 1) {
    const double y = __shfl_up_sync(F, x, d);
    if (lane - d >= h) x += y;
  }
  if (lane == 31 || nk != key) {
    const unsigned a = acc_s + (key << 3);
    sts_f64(a, lds_f64(a) + x);
  }
  __syncwarp();
}

constexpr int FAR_BUFS = 2;                       // shared-memory buffers of C doubles: the tile in use + the next one

// One persistent CTA per SM.  The slice of w of the NEXT column tile is staged with ONE bulk copy
// (TMA, cp.async.bulk + mbarrier) issued by an elected thread while the current tile is consumed.
// Measured on the 1 kb rank-0 shard (3.87e8 px): per-thread 8-byte cp.async 1.56 ms, one bulk copy
// per tile 1.40 ms, two tiles in flight (three buffers) 1.63 ms -- the second copy competes with
// the record stream for the SM's L2 port -- so the distance stays at one tile.
template <int WARPS, int U, bool kCompact>
__global__ void __launch_bounds__(WARPS * 32, 1)
ice_sweep_far_kernel(const cb200_ice st) {
  extern __shared__ double far_sm[];               // acc[FAR_ACC] | w tile [FAR_BUFS][C]
  __shared__ int s_unit;
  __shared__ unsigned long long s_bar[FAR_BUFS];   // "tile staged" barrier of each w buffer
  if (*st.all_done) return;
  constexpr int RW = FAR_ROWS / WARPS;             // rows per warp slice
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int cs = st.far_col_shift, C = 1 << cs;
  double* __restrict__ acc = far_sm;
  const unsigned acc_s = (unsigned)__cvta_generic_to_shared(far_sm);
  const unsigned wb_s = acc_s + FAR_ACC * 8;
  unsigned long long* claim = reinterpret_cast<unsigned long long*>(st.far_claim);

  const unsigned bar_s = (unsigned)__cvta_generic_to_shared(s_bar);
  unsigned phases = 0;                             // bit b = parity of the phase barrier b is in (uniform over the CTA)
  if (threadIdx.x == 0) {
    for (int b = 0; b < FAR_BUFS; ++b) mbar_init(bar_s + 8 * b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // stage the slice of w of column tile j into buffer b
  auto stage = [&](int j, int b) {
    if (threadIdx.x == 0) {
      const unsigned dst_s = wb_s + (unsigned)b * (unsigned)C * 8u;
      const int64_t base = (int64_t)j << cs;
      const int64_t left = (int64_t)st.n_bins - base;
      const int nv = (int)(left < C ? left : C);               // valid bins of this tile (>= 1)
      const unsigned bytes = (unsigned)(nv & ~1) * 8u;         // bulk copies move multiples of 16 bytes
      if (nv & 1) sts_f64(dst_s + (unsigned)(nv - 1) * 8u, st.w[base + nv - 1]);   // odd tail, before the arrive
      mbar_expect_tx(bar_s + 8 * b, bytes);                    // 0 bytes: the arrive alone completes the phase
      if (bytes) bulk_g2s(dst_s, st.w + base, bytes, bar_s + 8 * b);
    }
  };
  auto staged = [&](int b) {                       // wait until buffer b holds its tile (followed by __syncthreads)
    mbar_wait(bar_s + 8 * b, (phases >> b) & 1u);
    phases ^= 1u << b;
  };

  while (true) {
    __syncthreads();                               // the previous unit is finished by every warp
    if (threadIdx.x == 0) s_unit = (int)atomicAdd(claim, 1ULL);
    __syncthreads();
    const int u = s_unit;
    if (u >= st.n_far_units) break;                // CTA-uniform
    const cb200_far_unit un = st.far_units[u];
    const cb200_far_copy fc = st.far_copies[un.copy];
    for (int i = lane; i < RW; i += 32) acc[wid * RW + i] = 0.0;
    // slot (j, wid) of this row block: chunks tp[j * WARPS] .. tp[j * WARPS + 1]
    const int64_t* __restrict__ tp = fc.ptr + ((int64_t)(un.rb - fc.rb0) * fc.n_j - fc.j0) * WARPS + wid;
    typedef typename std::conditional<kCompact, int2, int4>::type Q;    // one lane's share of a 64-record chunk
    const Q* __restrict__ pxq = reinterpret_cast<const Q*>(fc.px);
    const int j_first = st.far_tiles[un.t0];
    int jn = (un.t0 + 1 < un.t1) ? st.far_tiles[un.t0 + 1] : -1;
    stage(j_first, 0);
    int64_t a = tp[(int64_t)j_first * WARPS], b = tp[(int64_t)j_first * WARPS + 1];
    Q qn[U];
    if constexpr (kCompact) far_first_group4<U>(pxq, a, b, qn, lane);
    else far_first_group<U>(pxq, a, b, qn, lane);
    staged(0);
    __syncthreads();                               // tile 0 staged; accumulators zeroed
    int buf = 0;
    for (int t = un.t0; t < un.t1; ++t) {
      const int jnn = (t + 2 < un.t1) ? st.far_tiles[t + 2] : -1;     // in flight during this tile
      int64_t a2 = 0, b2 = 0;
      if (jn >= 0) {
        stage(jn, buf ^ 1);                        // every warp left that buffer at the barrier that ended tile t - 1
        a2 = tp[(int64_t)jn * WARPS];
        b2 = tp[(int64_t)jn * WARPS + 1];
      }
      if constexpr (kCompact) far_slot4<U>(pxq, a, b, qn, wb_s + (unsigned)buf * (unsigned)C * 8u, acc_s, lane);
      else far_slot<U>(pxq, a, b, qn, wb_s + (unsigned)buf * (unsigned)C * 8u, acc_s, lane);
      if (jn >= 0) {
        if constexpr (kCompact) far_first_group4<U>(pxq, a2, b2, qn, lane);
        else far_first_group<U>(pxq, a2, b2, qn, lane);
        staged(buf ^ 1);                           // (a barrier is only waited on when a copy was issued to it)
      }
      __syncthreads();                             // next tile staged, this one released
      buf ^= 1;
      a = a2;
      b = b2;
      jn = jnn;
    }
    double* __restrict__ part = st.far_parts + (int64_t)u * FAR_ROWS + wid * RW;
    for (int i = lane; i < RW; i += 32)            // undo the bank swizzle of the row ids (see far_swizzle)
      part[i] = acc[wid * RW + ((i & (RW / 32 - 1)) << 5) + i / (RW / 32)];
  }
  // re-arm the claim counter once every CTA is through
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long ticket = atomicAdd(&claim[1], 1ULL);
    if (ticket == gridDim.x - 1) {
      claim[0] = 0ULL;
      claim[1] = 0ULL;
      __threadfence();
    }
  }
}

// Far-field part of s[i]: the units of i's row block, in unit order.
__device__ __forceinline__ double far_row_total(const cb200_ice& st, int i) {
  double v = 0.0;
  const int rb = i >> CB200_FAR_ROW_SHIFT, lr = i & (FAR_ROWS - 1);
  for (int c = 0; c < st.n_far_copies; ++c) {
    const cb200_far_copy& fc = st.far_copies[c];
    const unsigned k = (unsigned)(rb - fc.rb0);
    if (k < (unsigned)fc.n_rb)
      for (int u = fc.rb_unit[k]; u < fc.rb_unit[k + 1]; ++u) v += st.far_parts[(int64_t)u * FAR_ROWS + lr];
  }
  return v;
}

}  // namespace cb200
