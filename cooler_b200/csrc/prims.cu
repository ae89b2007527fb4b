// Device-wide primitives: exclusive scan and stable LSD radix sort of
// (int32 key, 8-byte value) pairs.  Hand-written for sm_100a; no CUB/Thrust.
#include "common.cuh"
#include "../../include/cooler_b200.h"

namespace cb200 {

// ---------------------------------------------------------------- scan ----
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename TIn>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_sums(const TIn* __restrict__ in, int64_t n, int64_t* __restrict__ tile_sum) {
  __shared__ long long wsum[SCAN_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  long long v = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + (int64_t)k * SCAN_THREADS + threadIdx.x;
    if (i < n) v += (long long)in[i];
  }
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int w = 0; w < SCAN_THREADS / 32; ++w) t += wsum[w];
    tile_sum[blockIdx.x] = t;
  }
}

// Each thread owns SCAN_ITEMS consecutive elements (blocked layout).
template <typename TIn>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply(const TIn* __restrict__ in, int64_t n, const int64_t* __restrict__ tile_off,
           int64_t* __restrict__ out) {
  __shared__ long long wsum[SCAN_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  long long x[SCAN_ITEMS];
  long long tsum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + k;
    x[k] = (i < n) ? (long long)in[i] : 0;
    tsum += x[k];
  }
  // inclusive warp scan of the thread sums
  long long incl = tsum;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    long long y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) wsum[wid] = incl;
  __syncthreads();
  long long woff = 0;
  for (int w = 0; w < wid; ++w) woff += wsum[w];
  long long run = (long long)tile_off[blockIdx.x] + woff + incl - tsum;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + k;
    if (i < n) out[i] = run;
    run += x[k];
    if (i == n - 1) out[n] = run;
  }
}

__global__ void scan_zero(int64_t* out) { out[0] = 0; }

static size_t scan_ws_bytes(int64_t n) {
  size_t total = 0;
  while (n > 1) {
    int64_t tiles = ceil_div64(n, SCAN_TILE);
    total += (size_t)(tiles + 1) * sizeof(int64_t) * 2;  // sums + scanned sums
    n = tiles;
  }
  return total + 64;
}

template <typename TIn>
static int scan_rec(const TIn* in, int64_t* out, int64_t n, char* ws, cudaStream_t st) {
  if (n <= 0) {
    scan_zero<<<1, 1, 0, st>>>(out);
    CB_LAUNCH_CHECK("scan_zero");
    return 0;
  }
  int64_t tiles = ceil_div64(n, SCAN_TILE);
  int64_t* sums = reinterpret_cast<int64_t*>(ws);
  int64_t* offs = sums + (tiles + 1);
  char* next_ws = reinterpret_cast<char*>(offs + (tiles + 1));
  scan_tile_sums<TIn><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n, sums);
  CB_LAUNCH_CHECK("scan_tile_sums");
  if (tiles == 1) {
    scan_zero<<<1, 1, 0, st>>>(offs);
    CB_LAUNCH_CHECK("scan_zero");
  } else {
    int rc = scan_rec<int64_t>(sums, offs, tiles, next_ws, st);
    if (rc) return rc;
  }
  scan_apply<TIn><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n, offs, out);
  CB_LAUNCH_CHECK("scan_apply");
  return 0;
}

int exclusive_scan_i32(const int32_t* in, int64_t* out, int64_t n, void* ws, size_t ws_bytes,
                       cudaStream_t st) {
  if (ws_bytes < scan_ws_bytes(n)) {
    set_error_msg("scan workspace too small");
    return -1;
  }
  return scan_rec<int32_t>(in, out, n, reinterpret_cast<char*>(ws), st);
}

// ---------------------------------------------------------- radix sort ----
constexpr int SORT_WARPS = 16;
constexpr int SORT_THREADS = SORT_WARPS * 32;
constexpr int SORT_STEPS = 8;                        // elements per thread
constexpr int SORT_WARP_SPAN = SORT_STEPS * 32;      // consecutive elements owned by a warp
constexpr int SORT_TILE = SORT_THREADS * SORT_STEPS; // 4096 elements per CTA
constexpr int RADIX = 256;

__global__ void __launch_bounds__(SORT_THREADS)
radix_hist(const int32_t* __restrict__ keys, int64_t n, int shift, int mask,
           int32_t* __restrict__ hist, int nblocks) {
  __shared__ int h[RADIX];
  for (int d = threadIdx.x; d < RADIX; d += SORT_THREADS) h[d] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * SORT_TILE;
#pragma unroll
  for (int s = 0; s < SORT_STEPS; ++s) {
    int64_t e = base + (int64_t)s * SORT_THREADS + threadIdx.x;
    if (e < n) atomicAdd(&h[(keys[e] >> shift) & mask], 1);   // integer: order-independent
  }
  __syncthreads();
  for (int d = threadIdx.x; d <= mask; d += SORT_THREADS)
    hist[(int64_t)d * nblocks + blockIdx.x] = h[d];
}

// Stable scatter: a warp owns SORT_WARP_SPAN consecutive elements and ranks
// them 32 at a time in element order with match.any; warps are then chained
// per digit in warp order, so equal digits keep their input order.  Elements
// are first placed at their block-local sorted position in shared memory and
// then written out in that order, so every digit run leaves the block as one
// contiguous, coalesced segment.
struct ScatterSmem {
  unsigned wcount[SORT_WARPS][RADIX];
  unsigned lbase[RADIX];          // block-local exclusive offset of every digit
  long long gbase[RADIX];         // global base of the digit minus lbase
  int32_t skey[SORT_TILE];
  int2 sval[SORT_TILE];
};

__global__ void __launch_bounds__(SORT_THREADS)
radix_scatter(const int32_t* __restrict__ keys_in, const int2* __restrict__ vals_in,
              int32_t* __restrict__ keys_out, int2* __restrict__ vals_out, int64_t n,
              int shift, int mask, const int64_t* __restrict__ off, int nblocks) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScatterSmem& sm = *reinterpret_cast<ScatterSmem*>(smem_raw);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&sm.wcount[0][0])[i] = 0u;
  __syncthreads();

  const int64_t tbase = (int64_t)blockIdx.x * SORT_TILE;
  const int64_t wbase = tbase + (int64_t)wid * SORT_WARP_SPAN;
  int32_t key[SORT_STEPS];
  int2 val[SORT_STEPS];
  unsigned rank[SORT_STEPS];
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int s = 0; s < SORT_STEPS; ++s) {
    int64_t e = wbase + s * 32 + lane;
    bool ok = e < n;
    key[s] = ok ? keys_in[e] : 0;
    val[s] = ok ? vals_in[e] : make_int2(0, 0);
  }
#pragma unroll
  for (int s = 0; s < SORT_STEPS; ++s) {
    int64_t e = wbase + s * 32 + lane;
    bool ok = e < n;
    int d = ok ? ((key[s] >> shift) & mask) : (RADIX + lane);   // invalid lanes match nobody
    unsigned peers = __match_any_sync(0xffffffffu, d);
    unsigned old = 0;
    if (ok) old = sm.wcount[wid][d];
    __syncwarp();
    if (ok && (peers & lt) == 0) sm.wcount[wid][d] = old + __popc(peers);
    __syncwarp();
    rank[s] = old + __popc(peers & lt);
  }
  __syncthreads();
  // per digit: chain the warps; then an exclusive scan over the digit totals (warp 0)
  if (threadIdx.x < RADIX) {
    const int d = threadIdx.x;
    unsigned run = 0;
    if (d <= mask) {
      for (int w = 0; w < SORT_WARPS; ++w) {
        unsigned c = sm.wcount[w][d];
        sm.wcount[w][d] = run;
        run += c;
      }
    }
    sm.lbase[d] = run;                              // digit total, scanned below
  }
  __syncthreads();
  if (wid == 0) {
    unsigned v[RADIX / 32];
    unsigned tsum = 0;
#pragma unroll
    for (int k = 0; k < RADIX / 32; ++k) { v[k] = sm.lbase[lane * (RADIX / 32) + k]; tsum += v[k]; }
    unsigned incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    unsigned run = incl - tsum;
#pragma unroll
    for (int k = 0; k < RADIX / 32; ++k) {
      const int d = lane * (RADIX / 32) + k;
      sm.lbase[d] = run;
      if (d <= mask) sm.gbase[d] = off[(int64_t)d * nblocks + blockIdx.x] - (long long)run;
      run += v[k];
    }
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < SORT_STEPS; ++s) {
    int64_t e = wbase + s * 32 + lane;
    if (e < n) {
      const int d = (key[s] >> shift) & mask;
      const unsigned lp = sm.lbase[d] + sm.wcount[wid][d] + rank[s];
      sm.skey[lp] = key[s];
      sm.sval[lp] = val[s];
    }
  }
  __syncthreads();
  const int n_tile = (int)min((int64_t)SORT_TILE, n - tbase);
  for (int i = threadIdx.x; i < n_tile; i += SORT_THREADS) {
    const int32_t k = sm.skey[i];
    const int64_t pos = sm.gbase[(k >> shift) & mask] + i;
    keys_out[pos] = k;
    vals_out[pos] = sm.sval[i];
  }
}

static size_t sort_ws_bytes(int64_t n) {
  int64_t nblocks = ceil_div64(n > 0 ? n : 1, SORT_TILE);
  int64_t m = nblocks * RADIX;
  return (size_t)m * sizeof(int32_t) + (size_t)(m + 1) * sizeof(int64_t) + scan_ws_bytes(m) + 256;
}

int sort_pairs(const int32_t* keys_in, const int2* vals_in, int32_t* keys_out, int2* vals_out,
               int32_t* keys_tmp, int2* vals_tmp, int64_t n, int key_shift, int key_bits, void* ws,
               size_t ws_bytes, cudaStream_t st) {
  if (ws_bytes < sort_ws_bytes(n)) {
    set_error_msg("sort workspace too small");
    return -1;
  }
  if (n <= 0) return 0;
  if (key_shift < 0) key_shift = 0;
  if (key_bits < 1) key_bits = 1;
  if (key_shift + key_bits > 31) key_bits = 31 - key_shift;
  const int passes = (key_bits + 7) / 8;
  const int bits = (key_bits + passes - 1) / passes;
  const int64_t nblocks = ceil_div64(n, SORT_TILE);
  if (nblocks > 0x7fffffff) {
    set_error_msg("sort: too many elements");
    return -1;
  }
  char* p = reinterpret_cast<char*>(ws);
  int32_t* hist = reinterpret_cast<int32_t*>(p);
  size_t hist_bytes = ((size_t)nblocks * RADIX * sizeof(int32_t) + 127) / 128 * 128;
  int64_t* off = reinterpret_cast<int64_t*>(p + hist_bytes);
  size_t off_bytes = (((size_t)nblocks * RADIX + 1) * sizeof(int64_t) + 127) / 128 * 128;
  char* scan_ws = p + hist_bytes + off_bytes;
  size_t scan_bytes = ws_bytes - hist_bytes - off_bytes;

  static bool attr_set = false;
  if (!attr_set) {
    CB_CHECK(cudaFuncSetAttribute(radix_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)sizeof(ScatterSmem)));
    attr_set = true;
  }
  const int32_t* ksrc = keys_in;
  const int2* vsrc = vals_in;
  for (int k = 0; k < passes; ++k) {
    const int shift = key_shift + k * bits;
    const int nb = (k == passes - 1) ? (key_bits - k * bits) : bits;
    const int mask = (1 << nb) - 1;
    const bool to_out = ((passes - k) % 2) == 1;
    int32_t* kdst = to_out ? keys_out : keys_tmp;
    int2* vdst = to_out ? vals_out : vals_tmp;
    radix_hist<<<(unsigned)nblocks, SORT_THREADS, 0, st>>>(ksrc, n, shift, mask, hist, (int)nblocks);
    CB_LAUNCH_CHECK("radix_hist");
    int rc = exclusive_scan_i32(hist, off, (int64_t)(mask + 1) * nblocks, scan_ws, scan_bytes, st);
    if (rc) return rc;
    radix_scatter<<<(unsigned)nblocks, SORT_THREADS, sizeof(ScatterSmem), st>>>(
        ksrc, vsrc, kdst, vdst, n, shift, mask, off, (int)nblocks);
    CB_LAUNCH_CHECK("radix_scatter");
    ksrc = kdst;
    vsrc = vdst;
  }
  return 0;
}

}  // namespace cb200

extern "C" {

size_t cb200_scan_workspace_bytes(int64_t n) { return cb200::scan_ws_bytes(n); }

int cb200_exclusive_scan_i32(const int32_t* in, int64_t* out, int64_t n, void* workspace,
                             size_t workspace_bytes, void* stream) {
  return cb200::exclusive_scan_i32(in, out, n, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t cb200_sort_workspace_bytes(int64_t n) { return cb200::sort_ws_bytes(n); }

int cb200_sort_pairs(const int32_t* keys_in, const cb200_px* vals_in, int32_t* keys_out,
                     cb200_px* vals_out, int32_t* keys_tmp, cb200_px* vals_tmp, int64_t n,
                     int key_shift, int key_bits, void* workspace, size_t workspace_bytes,
                     void* stream) {
  return cb200::sort_pairs(keys_in, reinterpret_cast<const int2*>(vals_in), keys_out,
                           reinterpret_cast<int2*>(vals_out), keys_tmp,
                           reinterpret_cast<int2*>(vals_tmp), n, key_shift, key_bits, workspace,
                           workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"

// ------------------------------------------------------ exact medians ----
// Order statistics of float64 values by radix selection (8 passes of 8 bits from the most
// significant byte of an order-preserving 64-bit key).  Used for the medians of the bin filters
// (_balance.py:400-409, util.mad): selection returns elements of the input, so the result is
// bit-identical to numpy's partition-based np.median whatever the order of evaluation.
// Query q = 2 s + h asks for the element of rank (cnt_s - 1) / 2 (h = 0) or cnt_s / 2 (h = 1)
// among the valid elements of segment s -- the two middle elements whose mean is the median.
namespace cb200 {

__device__ __forceinline__ unsigned long long f64_order_key(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);      // negatives reversed below the positives
}
__device__ __forceinline__ double f64_from_order_key(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
  return __longlong_as_double((long long)b);
}
__device__ __forceinline__ bool median_valid(double v, int positive_only) {
  return positive_only ? (v > 0.0) : (v == v);               // NaNs never take part
}

__global__ void median_count_kernel(const double* __restrict__ x, const int64_t* __restrict__ seg_offset,
                                    int n_segs, int positive_only, unsigned long long* __restrict__ cnt) {
  const int s = blockIdx.y;
  const int64_t lo = seg_offset[s], hi = seg_offset[s + 1];
  unsigned long long c = 0;
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x)
    c += median_valid(x[i], positive_only) ? 1 : 0;
  c = (unsigned long long)warp_sum((long long)c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&cnt[s], c);
}

// state[q] = {prefix, rank}; hist[q][256]
__global__ void median_init_kernel(const unsigned long long* __restrict__ cnt, int n_segs,
                                   unsigned long long* __restrict__ state) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= 2 * n_segs) return;
  const unsigned long long c = cnt[q >> 1];
  state[2 * q] = 0ULL;
  state[2 * q + 1] = c ? ((q & 1) ? c / 2 : (c - 1) / 2) : 0ULL;
}

__global__ void median_hist_kernel(const double* __restrict__ x, const int64_t* __restrict__ seg_offset,
                                   int n_segs, int positive_only, int pass,
                                   const unsigned long long* __restrict__ state,
                                   unsigned int* __restrict__ hist) {
  __shared__ unsigned int sh[2][256];
  const int s = blockIdx.y;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) (&sh[0][0])[i] = 0;
  __syncthreads();
  const int shift = 56 - 8 * pass;
  const unsigned long long p0 = state[2 * (2 * s)], p1 = state[2 * (2 * s + 1)];
  const int64_t lo = seg_offset[s], hi = seg_offset[s + 1];
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = x[i];
    if (!median_valid(v, positive_only)) continue;
    const unsigned long long k = f64_order_key(v);
    const unsigned long long high = pass ? (k >> (shift + 8)) : 0ULL;
    const unsigned d = (unsigned)(k >> shift) & 255u;
    if (high == p0) atomicAdd(&sh[0][d], 1u);
    if (high == p1) atomicAdd(&sh[1][d], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    const unsigned v = (&sh[0][0])[i];
    if (v) atomicAdd(&hist[(size_t)(2 * s) * 256 + i], v);
  }
}

// one warp per query: find the digit whose bucket holds the wanted rank, descend into it
__global__ void median_step_kernel(int n_queries, unsigned long long* __restrict__ state,
                                   unsigned int* __restrict__ hist) {
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= n_queries) return;
  const int lane = threadIdx.x & 31;
  unsigned int* h = hist + (size_t)q * 256;
  unsigned long long c[8], tot = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) { c[k] = h[lane * 8 + k]; tot += c[k]; h[lane * 8 + k] = 0; }   // re-armed for the next pass
  unsigned long long incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  const unsigned long long rank = state[2 * q + 1];
  const unsigned long long before = incl - tot;
  if (rank >= before && rank < incl) {             // exactly one lane (none when the segment is empty)
    unsigned long long r = rank - before;
    int k = 0;
    while (r >= c[k]) { r -= c[k]; ++k; }
    state[2 * q] = (state[2 * q] << 8) | (unsigned long long)(lane * 8 + k);
    state[2 * q + 1] = r;
  }
}

__global__ void median_finish_kernel(const unsigned long long* __restrict__ state,
                                     const unsigned long long* __restrict__ cnt, int n_segs,
                                     double* __restrict__ mid, int64_t* __restrict__ count_out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= 2 * n_segs) return;
  const unsigned long long c = cnt[q >> 1];
  mid[q] = c ? f64_from_order_key(state[2 * q]) : __longlong_as_double(0x7ff8000000000000LL);
  if (!(q & 1)) count_out[q >> 1] = (int64_t)c;
}

}  // namespace cb200

extern "C" {

size_t cb200_median_workspace_bytes(int32_t n_segs) {
  return (size_t)n_segs * (8 + 2 * 16 + 2 * 256 * 4);
}

int cb200_segment_medians(const double* x, const int64_t* seg_offset, int32_t n_segs, int64_t max_seg_len,
                          int32_t positive_only, double* mid_out, int64_t* count_out, void* workspace,
                          size_t workspace_bytes, void* stream) {
  using namespace cb200;
  if (n_segs <= 0) return 0;
  if (workspace_bytes < cb200_median_workspace_bytes(n_segs)) {
    set_error_msg("segment_medians: workspace too small");
    return -1;
  }
  cudaStream_t s = (cudaStream_t)stream;
  CB_CHECK(cudaMemsetAsync(workspace, 0, cb200_median_workspace_bytes(n_segs), s));
  unsigned long long* cnt = reinterpret_cast<unsigned long long*>(workspace);
  unsigned long long* state = cnt + n_segs;
  unsigned int* hist = reinterpret_cast<unsigned int*>(state + 4 * (size_t)n_segs);
  int64_t bx = ceil_div64(max_seg_len > 0 ? max_seg_len : 1, 256 * 8);
  if (bx > 1024) bx = 1024;
  const dim3 grid((unsigned)bx, (unsigned)n_segs);
  median_count_kernel<<<grid, 256, 0, s>>>(x, seg_offset, n_segs, positive_only, cnt);
  CB_LAUNCH_CHECK("median_count");
  median_init_kernel<<<(unsigned)ceil_div64(2 * n_segs, 128), 128, 0, s>>>(cnt, n_segs, state);
  CB_LAUNCH_CHECK("median_init");
  for (int pass = 0; pass < 8; ++pass) {
    median_hist_kernel<<<grid, 256, 0, s>>>(x, seg_offset, n_segs, positive_only, pass, state, hist);
    CB_LAUNCH_CHECK("median_hist");
    median_step_kernel<<<(unsigned)ceil_div64(2 * n_segs, 4), 128, 0, s>>>(2 * n_segs, state, hist);
    CB_LAUNCH_CHECK("median_step");
  }
  median_finish_kernel<<<(unsigned)ceil_div64(2 * n_segs, 128), 128, 0, s>>>(state, cnt, n_segs, mid_out, count_out);
  CB_LAUNCH_CHECK("median_finish");
  return 0;
}

}  // extern "C"
