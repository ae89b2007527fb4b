// K3 for the 2-D blocked layout with ONE ROW PER LANE ("lane rows", far_layout = 1).
//
// Same blocking as far.cuh -- row blocks of R = 1 << CB200_FAR_ROW_SHIFT rows whose sums live in
// shared memory, column tiles of C = 1 << far_col_shift bins whose slice of w is staged in shared
// memory -- but a different arrangement of the records inside a slot (row block, warp row slice,
// column tile) and a different pipeline:
//
//  * sliced-ELL arrangement.  The rows of a slot that hold records are ordered by decreasing
//    record count and cut into slices of 32 rows; lane l of the warp owns row l of the slice and
//    walks through that row's records sequentially: record k of all 32 rows forms "step" k, one
//    coalesced 256-byte load.  A lane accumulates its row in a register -- no shuffles, no key
//    compares -- and adds the total to the row's accumulator once, when the slice ends (bit 31 of
//    the record's first word marks the last step of a slice).  Because the rows of a slice have
//    (nearly) equal lengths the padding is small (measured < 2 %).  Every accumulator has exactly
//    one writer per slice and the slices of a warp are processed in order: no atomics, fixed
//    summation order, bit-reproducible.
//  * the steps of one warp are contiguous in memory over all column tiles of a row block (slots
//    are stored in (row block, warp, tile) order), so the record stream is prefetched with a flat
//    pointer, U steps ahead, across tile boundaries.
//  * no CTA barrier per tile: the tiles of w travel through a ring of shared-memory stages filled
//    by a producer warp with one bulk copy (TMA, cp.async.bulk) per tile; full / empty mbarriers
//    let the 16 consumer warps drift up to STAGES - 1 tiles apart, which absorbs the imbalance
//    between the row slices of a tile (the dense diagonal) that stalled far.cuh at its barriers.
#pragma once
#include "far.cuh"

namespace cb200 {

constexpr int ROWS_WARPS = 16;                     // consumer warps = row slices per block
constexpr int ROWS_THREADS = (ROWS_WARPS + 1) * 32;   // + the producer warp
constexpr int ROWS_RW = FAR_ROWS / ROWS_WARPS;     // rows per warp slice
constexpr int ROWS_ACC = FAR_ROWS + ROWS_WARPS;    // accumulators + one dummy slot per warp (padding lanes)
constexpr int ROWS_MAX_STAGES = 8;
constexpr unsigned ROWS_KEY8_MASK = 0x3FFF8u;      // (x >> 13) & mask = 8 * block-local row id (or dummy id R + warp), bits 16..30 of x
static_assert(ROWS_RW % 32 == 0 && FAR_ROWS + ROWS_WARPS <= 32768, "row ids (incl. the dummy ids) take 15 bits");

__device__ __forceinline__ void mbar_arrive(unsigned bar_s) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_s) : "memory");
}
// Wait for the phase with parity `phase` to complete.  The spin is bounded: a protocol error traps
// (the launch fails with an error) instead of hanging the device.
__device__ __forceinline__ void mbar_wait_rows(unsigned bar_s, unsigned phase) {
  for (unsigned spins = 0;; ++spins) {
    unsigned ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}" : "=r"(ok) : "r"(bar_s), "r"(phase) : "memory");
    if (ok) return;
    if (spins > (1u << 24)) __trap();
  }
}

template <int U>
__global__ void __launch_bounds__(ROWS_THREADS, 1)
ice_sweep_rows_kernel(const cb200_ice st, const int n_stages, const int tma_chunk, const int pf_steps) {
  extern __shared__ double rows_sm[];              // acc[ROWS_ACC] | ring[n_stages][C]
  __shared__ unsigned long long s_full[ROWS_MAX_STAGES], s_empty[ROWS_MAX_STAGES];
  __shared__ int s_unit[ROWS_MAX_STAGES];          // unit of the tile in each stage (-1: no more work)
  if (*st.all_done) return;
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int cs = st.far_col_shift, C = 1 << cs;
  const unsigned acc_s = (unsigned)__cvta_generic_to_shared(rows_sm);
  const unsigned ring_s = acc_s + ROWS_ACC * 8;
  const unsigned full_s = (unsigned)__cvta_generic_to_shared(s_full);
  const unsigned empty_s = (unsigned)__cvta_generic_to_shared(s_empty);
  unsigned long long* claim = reinterpret_cast<unsigned long long*>(st.far_claim);

  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(full_s + 8 * s, 1);
      mbar_init(empty_s + 8 * s, ROWS_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();                                 // the only CTA-wide barrier of the kernel

  if (wid == ROWS_WARPS) {
    // ------------------------------------------------------------------ producer (one thread)
    if (lane != 0) return;
    int s = 0;
    unsigned k = 0;                                // how often stage s has been used so far
    while (true) {
      const int u = (int)atomicAdd(claim, 1ULL);
      const bool more = u < st.n_far_units;
      int t0 = 0, t1 = 1;                          // the terminal item occupies one stage
      if (more) {
        const cb200_far_unit un = st.far_units[u];
        t0 = un.t0;
        t1 = un.t1;
      }
      for (int t = t0; t < t1; ++t) {
        if (k > 0) mbar_wait_rows(empty_s + 8 * s, (k - 1) & 1u);     // every consumer warp released the stage
        *reinterpret_cast<volatile int*>(&s_unit[s]) = more ? u : -1;
        if (more) {
          const int64_t base = (int64_t)st.far_tiles[t] << cs;
          const int64_t left = (int64_t)st.n_bins - base;
          const int nv = (int)(left < C ? left : C);             // valid bins of this tile (>= 1)
          const unsigned bytes = (unsigned)(nv & ~1) * 8u;       // bulk copies move multiples of 16 bytes
          const unsigned dst_s = ring_s + (unsigned)s * (unsigned)C * 8u;
          if (nv & 1) sts_f64(dst_s + (unsigned)(nv - 1) * 8u, st.w[base + nv - 1]);   // odd tail, before the arrive
          mbar_expect_tx(full_s + 8 * s, bytes);                 // 0 bytes: the arrive alone completes the phase
          for (unsigned o = 0; o < bytes; o += (unsigned)tma_chunk) {   // several bulk copies in flight per tile
            const unsigned nb = bytes - o < (unsigned)tma_chunk ? bytes - o : (unsigned)tma_chunk;
            bulk_g2s(dst_s + o, reinterpret_cast<const char*>(st.w + base) + o, nb, full_s + 8 * s);
          }
        } else {
          mbar_arrive(full_s + 8 * s);
        }
        if (++s == n_stages) { s = 0; ++k; }
      }
      if (!more) break;
    }
    // re-arm the claim counter once the producer of every CTA is through
    __threadfence();
    const unsigned long long ticket = atomicAdd(&claim[1], 1ULL);
    if (ticket == gridDim.x - 1) {
      claim[0] = 0ULL;
      claim[1] = 0ULL;
      __threadfence();
    }
    return;
  }

  // -------------------------------------------------------------------- consumers (16 warps)
  int s = 0;
  unsigned ph = 0;                                 // parity of the phase stage s is in
  auto next_stage = [&]() { if (++s == n_stages) { s = 0; ph ^= 1u; } };
  auto release = [&]() {                           // this warp is done with stage s
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_s + 8 * s);
    next_stage();
  };
  const unsigned my_acc = acc_s + (unsigned)(wid * ROWS_RW) * 8u;
  while (true) {
    mbar_wait_rows(full_s + 8 * s, ph);
    const int u = *reinterpret_cast<volatile int*>(&s_unit[s]);
    if (u < 0) break;
    const cb200_far_unit un = st.far_units[u];
    const cb200_far_copy fc = st.far_copies[un.copy];
    for (int i = lane; i < ROWS_RW; i += 32) sts_f64(my_acc + 8u * i, 0.0);
    // tp[j] = first step of slot (row block, this warp, column tile j)
    const int64_t* __restrict__ tp = fc.ptr + ((int64_t)(un.rb - fc.rb0) * ROWS_WARPS + wid) * fc.n_j - fc.j0;
    const int32_t* __restrict__ tl = st.far_tiles;
    const int64_t i0 = tp[tl[un.t0]];
    const int n_steps = (int)(tp[tl[un.t1 - 1] + 1] - i0);
    // window of slot ends (relative steps): lane l holds the end of tile t_base + l, the next 32 are prefetched
    auto load_end = [&](int t) -> int { return (t < un.t1) ? (int)(tp[tl[t] + 1] - i0) : 0x7fffffff; };
    int t = un.t0, t_base = un.t0;
    int wcur = load_end(t_base + lane);
    int wnext = load_end(t_base + 32 + lane);
    int tile_end = __shfl_sync(FULL, wcur, 0);
    unsigned ws = ring_s + (unsigned)s * (unsigned)C * 8u;
    auto switch_tile = [&]() {                     // this warp's slot of the current tile is exhausted
      release();
      ++t;
      if (t - t_base == 32) {
        t_base += 32;
        wcur = wnext;
        wnext = load_end(t_base + 32 + lane);
      }
      tile_end = __shfl_sync(FULL, wcur, t - t_base);
      mbar_wait_rows(full_s + 8 * s, ph);
      ws = ring_s + (unsigned)s * (unsigned)C * 8u;
    };
    const int2* __restrict__ src = reinterpret_cast<const int2*>(fc.px) + i0 * 32 + lane;
    const int2 pad = make_int2(0, 0);
    int2 qn[U];
#pragma unroll
    for (int v = 0; v < U; ++v) qn[v] = (v < n_steps) ? ld_stream_int2(src + v * 32) : pad;
    __syncwarp();                                  // accumulators zeroed
    double val = 0.0;                              // running sum of this lane's current row
    for (int base = 0; base < n_steps; base += U) {
      int2 q[U];
#pragma unroll
      for (int v = 0; v < U; ++v) q[v] = qn[v];
      const int nb = base + U;
      if (pf_steps > 0 && nb + pf_steps + U <= n_steps) {
        // pull the stream pf_steps ahead into L2: the register pipeline below then only has to cover the
        // L2 latency, not the HBM latency (16 warps x U x 256 B in flight are too few for the latter)
#pragma unroll
        for (int v = 0; v < U; ++v)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (nb + pf_steps + v) * 32));
      }
      if (nb + U <= n_steps) {
#pragma unroll
        for (int v = 0; v < U; ++v) qn[v] = ld_stream_int2(src + (nb + v) * 32);
      } else {
#pragma unroll
        for (int v = 0; v < U; ++v)
          if (nb + v < n_steps) qn[v] = ld_stream_int2(src + (nb + v) * 32);
      }
      const int cnt = min(U, n_steps - base);      // valid steps of this group
      int v0 = 0;
      while (v0 < cnt) {
        while (base + v0 == tile_end) switch_tile();
        const int v1 = min(cnt, tile_end - base);  // steps [v0, v1) of the group gather from the current tile
        // phase 1: the U gathers of w (independent loads: one shared-memory latency per group, not per step)
        double g[U];
        unsigned flags = 0;
#pragma unroll
        for (int v = 0; v < U; ++v) {
          const bool on = (v >= v0) && (v < v1);
          const unsigned x = (unsigned)q[v].x;
          g[v] = on ? lds_f64(ws + (x & 0xFFFFu)) : 0.0;
          flags |= on ? x : 0u;
        }
        if (!__any_sync(FULL, (int)flags < 0)) {
          // no row of any lane ends inside these steps: a plain chain of fused multiply-adds
#pragma unroll
          for (int v = 0; v < U; ++v) val = fma(g[v], far_i32_to_f64(q[v].y), val);
        } else {
          // phase 2, four steps at a time: load the accumulators of the rows that end here (distinct
          // rows, each owned by one lane of the slot), then the multiply-add chain with the updates
          constexpr int H = U < 4 ? U : 4;
#pragma unroll
          for (int h = 0; h < U; h += H) {
            double a[H];
#pragma unroll
            for (int v = 0; v < H; ++v) {
              const bool end = (h + v >= v0) && (h + v < v1) && (q[h + v].x < 0);
              a[v] = end ? lds_f64(acc_s + (((unsigned)q[h + v].x >> 13) & ROWS_KEY8_MASK)) : 0.0;
            }
#pragma unroll
            for (int v = 0; v < H; ++v) {
              val = fma(g[h + v], far_i32_to_f64(q[h + v].y), val);
              if ((h + v >= v0) && (h + v < v1) && (q[h + v].x < 0)) {
                sts_f64(acc_s + (((unsigned)q[h + v].x >> 13) & ROWS_KEY8_MASK), a[v] + val);
                val = 0.0;
              }
            }
          }
        }
        v0 = v1;
      }
    }
    // release the current tile and the remaining tiles of the unit (empty slots of this warp)
    while (true) {
      release();
      if (++t >= un.t1) break;
      mbar_wait_rows(full_s + 8 * s, ph);
    }
    __syncwarp();
    double* __restrict__ part = st.far_parts + (int64_t)u * FAR_ROWS + wid * ROWS_RW;
    for (int i = lane; i < ROWS_RW; i += 32) part[i] = lds_f64(my_acc + 8u * i);
    __syncwarp();                                  // before the next unit zeroes the accumulators
  }
}

}  // namespace cb200
