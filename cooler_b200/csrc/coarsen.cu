// Coarsening kernels (K5): fused bin remap feeding the stable radix sort, then
// reduce-by-key.  Replaces CoolerCoarsener._aggregate's annotate-join + pandas
// groupby(sort=True).sum() (open2c/cooler src/cooler/_reduce.py:582-620).
//
// Chain for one level (host side: cooler_b200/reduce.py):
//   CSR(bin1,bin2) --remap--> key=bin2', val={bin1',v}  --stable sort by bin2'-->
//   runs of equal (bin2',bin1') are adjacent --runs_write(swap)--> key=bin1', val={bin2',sum}
//   --stable sort by bin1'--> CSR' sorted by (bin1',bin2'), duplicate-free.
#include "walk.cuh"
#include "../../include/cooler_b200.h"

namespace cb200 {

struct RemapOut {
  const int32_t* __restrict__ bin_map;
  int32_t* key_out;
  int2* val_out;
  __device__ __forceinline__ void operator()(int64_t pos, bool ok0, bool ok1, int r0, int r1, int4 q) {
    if (ok0) {
      key_out[pos] = __ldg(bin_map + q.x);
      val_out[pos] = make_int2(__ldg(bin_map + r0), q.y);
    }
    if (ok1) {
      key_out[pos + 1] = __ldg(bin_map + q.z);
      val_out[pos + 1] = make_int2(__ldg(bin_map + r1), q.w);
    }
  }
};

__global__ void __launch_bounds__(BLOCK_THREADS)
coarsen_remap_kernel(const int2* __restrict__ px, const int64_t* __restrict__ indptr,
                     const int32_t* __restrict__ tile_row, int64_t nnz, int64_t n_tiles,
                     const int32_t* __restrict__ bin_map, int32_t* __restrict__ key_out,
                     int2* __restrict__ val_out) {
  const int64_t t = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (t >= n_tiles) return;
  RemapOut f{bin_map, key_out, val_out};
  walk_map_tile(px, indptr, nnz, t, tile_row[t], f);
}

// A run head is an element whose (key, other) differs from its predecessor's.
__device__ __forceinline__ bool is_head(const int32_t* __restrict__ keys, const int2* __restrict__ vals,
                                        int64_t p) {
  if (p == 0) return true;
  return keys[p] != keys[p - 1] || vals[p].x != vals[p - 1].x;
}

__global__ void __launch_bounds__(BLOCK_THREADS)
runs_count_kernel(const int32_t* __restrict__ keys, const int2* __restrict__ vals, int64_t n,
                  int64_t n_tiles, int32_t* __restrict__ tile_count) {
  const int64_t t = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (t >= n_tiles) return;
  const int lane = threadIdx.x & 31;
  const int64_t p0 = t * (int64_t)WARP_TILE, p1 = min(p0 + (int64_t)WARP_TILE, n);
  int c = 0;
  for (int64_t p = p0 + lane; p < p1; p += 32) c += is_head(keys, vals, p) ? 1 : 0;
  c = warp_sum(c);
  if (lane == 0) tile_count[t] = c;
}

__global__ void __launch_bounds__(BLOCK_THREADS)
runs_write_kernel(const int32_t* __restrict__ keys, const int2* __restrict__ vals, int64_t n,
                  int64_t n_tiles, const int64_t* __restrict__ tile_off, int swap,
                  int32_t* __restrict__ key_out, int2* __restrict__ val_out,
                  int64_t* __restrict__ sum64_out, int32_t* __restrict__ overflow) {
  const int64_t t = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (t >= n_tiles) return;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int64_t p0 = t * (int64_t)WARP_TILE, p1 = min(p0 + (int64_t)WARP_TILE, n);
  int64_t base = tile_off[t];
  for (int64_t c = p0; c < p1; c += 32) {
    const int64_t p = c + lane;
    const bool head = (p < p1) && is_head(keys, vals, p);
    const unsigned hb = __ballot_sync(0xffffffffu, head);
    if (head) {
      const int k = keys[p];
      const int o = vals[p].x;
      long long sum = vals[p].y;
      for (int64_t e = p + 1; e < n && keys[e] == k && vals[e].x == o; ++e) sum += vals[e].y;   // run may cross tiles
      const int64_t out = base + __popc(hb & lt);
      const int lo32 = (int)sum;
      if ((long long)lo32 != sum) *overflow = 1;
      key_out[out] = swap ? o : k;
      val_out[out] = make_int2(swap ? k : o, lo32);
      if (sum64_out) sum64_out[out] = sum;
    }
    base += __popc(hb);
  }
}

// ---- K5 as a merge (no global sort) ---------------------------------------------------------
// The input copy is sorted by (bin1, bin2) and the bin map is monotone, so the f fine rows pooled
// into one coarse row are f lists ALREADY sorted by the coarse column: the coarse row is their
// f-way merge, done here as ceil(log2 f) rounds of pairwise merges.  In round r a "list" is the
// concatenation of 2^r consecutive fine rows of one coarse row (merged by the previous rounds);
// lists 2q and 2q+1 are merged into the same range of the output buffer, so no descriptor other
// than the fine indptr and the first fine row of every coarse row is needed.
//
// One warp per (coarse row, pair), claimed dynamically.  The warp walks through the merged output
// in chunks of 32 * VT elements: it stages the next chunk of either list in shared memory (coalesced
// loads; in round 0 the column ids are remapped on the fly), every lane finds the start of its VT
// outputs with a merge-path binary search and merges them sequentially (ties: the earlier list
// first, i.e. stable).  Keys and payloads are staged in separate 4-byte arrays and VT is odd: a
// lane's accesses are VT (or about VT / 2) words apart from its neighbour's, which keeps shared-memory
// bank conflicts low -- with 8-byte elements and VT = 8 the kernel was bound by the shared-memory
// pipe (8- to 16-way conflicts on every merge step and on the staging of the outputs).
//
// kReduce (the LAST round, one pair per coarse row): adjacent equal coarse columns are summed on the
// fly (int64) -- the reduce-by-key of `groupby(...).sum()` fused into the merge.  A lane closes a run
// when it meets the next column; a run that continues from earlier lanes gets their open sum through
// a warp-level segmented scan, a run that continues into the next chunk is carried in registers.
// The reduced row g is written at the row's INPUT offset (it is never longer than the input) and its
// length to row_count[g]; cb200_coarsen_compact then moves the rows to their final places.
constexpr int MG_WARPS = 8;
constexpr int MG_KEY_MAX = 0x7fffffff;

// Column remap of round 0.  A gather from the full bin_map costs one L1 tag lookup per lane (the columns
// of a row are spread over the whole genome: ~20 distinct lines per warp load, which made the gathers
// 90 % of the kernel's L1 requests).  The map is piecewise arithmetic, new = (col + pad[chromosome]) / factor
// with pad = new_off * factor - off, so a table with one entry per BLOCK of 2^shift bins -- {first chromosome
// boundary inside the block or INT_MAX, pad before it, pad behind it} -- answers it from a few cached
// lines; blocks holding several boundaries (tiny scaffolds) are marked and fall back to the gather.
struct RemapFn {
  const int32_t* __restrict__ bin_map;
  const int4* __restrict__ blocks;                 // nullptr: always gather
  int shift;
  unsigned magic;
  int mshift;
  __device__ __forceinline__ int operator()(int x) const {
    if (blocks == nullptr) return __ldg(bin_map + x);
    const int4 e = __ldg(blocks + (x >> shift));
    if (e.x < 0) return __ldg(bin_map + x);
    const unsigned y = (unsigned)x + (unsigned)(x >= e.x ? e.z : e.y);
    return (int)(__umulhi(y, magic) >> mshift);    // y / factor, exact for y < 2^31
  }
};

template <int VT, bool kReduce>
__global__ void __launch_bounds__(MG_WARPS * 32, 4)
coarsen_merge_kernel(const int2* __restrict__ in, int2* __restrict__ out,
                     const int64_t* __restrict__ indptr, const int32_t* __restrict__ group_first,
                     int n_groups, int round, int pairs_per_group,
                     const bool do_remap, const RemapFn remap,  // round 0: column remap
                     int32_t* __restrict__ key_out,            // coarse row of every output element, or nullptr
                     int32_t* __restrict__ row_count,          // kReduce: outputs of every coarse row
                     int32_t* __restrict__ overflow,           // kReduce: set when a sum leaves int32
                     unsigned long long* __restrict__ claim) {
  constexpr int NOUT = 32 * VT;                    // outputs per warp and chunk
  __shared__ int32_t sm[MG_WARPS][4][NOUT];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int32_t* kA = sm[wid][0];
  int32_t* vA = sm[wid][1];
  int32_t* kB = sm[wid][2];
  int32_t* vB = sm[wid][3];
  constexpr unsigned FULL = 0xffffffffu;
  const int64_t n_items = (int64_t)n_groups * pairs_per_group;
  const int span = 1 << round;                     // fine rows per list
  int ovf = 0;
  while (true) {
    unsigned long long w = 0;
    if (lane == 0) w = atomicAdd(claim, 1ULL);
    w = __shfl_sync(FULL, w, 0);
    if ((int64_t)w >= n_items) break;
    const int g = (int)(w / pairs_per_group), q = (int)(w % pairs_per_group);
    const int g0 = group_first[g], g1 = group_first[g + 1];
    const int a0 = g0 + 2 * q * span;
    if (a0 >= g1) {                                // this coarse row has fewer lists
      if (kReduce && lane == 0) row_count[g] = 0;
      continue;
    }
    const int a1 = min(a0 + span, g1), a2 = min(a1 + span, g1);
    const int64_t pa = indptr[a0], pb = indptr[a1], pe = indptr[a2];
    const int na = (int)(pb - pa), nb = (int)(pe - pb);
    const int2* __restrict__ A = in + pa;
    const int2* __restrict__ B = in + pb;
    int2* __restrict__ Y = out + pa;
    int i0 = 0, j0 = 0;
    int obase = 0;                                 // elements of this item written so far
    int ckey = 0;                                  // the run left open by the previous chunk
    long long csum = 0;
    bool cvalid = false;
    while (i0 + j0 < na + nb) {
      const int la = min(NOUT, na - i0), lb = min(NOUT, nb - j0);
      const int nout = min(NOUT, la + lb);
      // 2 x VT independent loads per lane in flight (a loop with a run-time bound would serialise the
      // load -> remap-gather chains)
      int2 va[VT], vb[VT];
#pragma unroll
      for (int t = 0; t < VT; ++t) {
        const int k = lane + 32 * t;
        va[t] = (k < la) ? ld_stream_int2(A + i0 + k) : make_int2(0, 0);
        vb[t] = (k < lb) ? ld_stream_int2(B + j0 + k) : make_int2(0, 0);
      }
      if (do_remap) {
#pragma unroll
        for (int t = 0; t < VT; ++t) {
          va[t].x = remap(va[t].x);                // (padding lanes remap column 0)
          vb[t].x = remap(vb[t].x);
        }
      }
#pragma unroll
      for (int t = 0; t < VT; ++t) {
        const int k = lane + 32 * t;
        if (k < la) { kA[k] = va[t].x; vA[k] = va[t].y; }
        if (k < lb) { kB[k] = vb[t].x; vB[k] = vb[t].y; }
      }
      __syncwarp();
      // merge path: the split (ia, ib), ia + ib = d, with A[ia-1] <= B[ib] and B[ib-1] < A[ia]
      const int d = min(lane * VT, nout);
      int lo = max(0, d - lb), hi = min(d, la);
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (kA[mid] <= kB[d - 1 - mid]) lo = mid + 1; else hi = mid;
      }
      int ia = lo, ib = d - lo;
      int ka = ia < la ? kA[ia] : MG_KEY_MAX, kb = ib < lb ? kB[ib] : MG_KEY_MAX;
      int key[VT], val[VT];
#pragma unroll
      for (int v = 0; v < VT; ++v) {
        key[v] = 0;
        val[v] = 0;
        if (d + v < nout) {
          const bool take_a = ka <= kb;            // an exhausted list shows MG_KEY_MAX; ties: A first
          key[v] = take_a ? ka : kb;
          val[v] = take_a ? vA[ia] : vB[ib];
          ia += take_a ? 1 : 0;
          ib += take_a ? 0 : 1;
          const int nxt = take_a ? ia : ib;        // only the consumed list shows a new head
          const int nk = (nxt < (take_a ? la : lb)) ? (take_a ? kA : kB)[nxt] : MG_KEY_MAX;
          ka = take_a ? nk : ka;
          kb = take_a ? kb : nk;
        }
      }
      i0 += __shfl_sync(FULL, ia, 31);             // lane 31 ends at the split of diagonal nout
      j0 += __shfl_sync(FULL, ib, 31);
      __syncwarp();                                // every lane is done reading the staged lists
      // ---- close runs.  hm: bit v = element v starts a run (kReduce: its column differs from the
      // previous element's; else every element).  A run is emitted by the lane that meets the NEXT head.
      const int cnt = min(VT, nout - d);
      int pk = __shfl_up_sync(FULL, key[VT - 1], 1);    // lane l has elements only if lane l-1 is full
      const bool pvalid = lane > 0 || cvalid;
      if (lane == 0) pk = ckey;
      unsigned hm = 0;
      long long open = 0;                          // sum behind the lane's last head (or of the whole lane)
      int lastk = pk;
      {
        int prev = pk;
        bool pv = pvalid;
#pragma unroll
        for (int v = 0; v < VT; ++v) {
          if (v < cnt) {
            const bool head = kReduce ? (!pv || key[v] != prev) : true;
            if (head) { hm |= 1u << v; open = 0; }
            open += val[v];
            prev = key[v];
            pv = true;
          }
        }
        lastk = prev;
      }
      const bool has_head = hm != 0;
      const int n_emit = __popc(hm) - (((hm & 1u) && !pvalid) ? 1 : 0);
      if (lane == 0 && !has_head) open += csum;    // the whole lane continues the previous chunk's run
      long long S = open;                          // inclusive segmented scan: open sum at the end of lane l
      bool F = has_head;
      int incl = n_emit;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long ts = __shfl_up_sync(FULL, S, o);
        const int tf = __shfl_up_sync(FULL, (int)F, o);
        const int tn = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) {
          if (!F) S += ts;
          F = F || (tf != 0);
          incl += tn;
        }
      }
      long long cin = __shfl_up_sync(FULL, S, 1);
      if (lane == 0) cin = csum;
      const int total = __shfl_sync(FULL, incl, 31);
      {
        int pos = incl - n_emit;
        long long run = cin;
        int prev = pk;
        bool have = pvalid;
#pragma unroll
        for (int v = 0; v < VT; ++v) {
          if (v < cnt) {
            if ((hm >> v) & 1u) {
              if (have) {
                const int lo32 = (int)run;
                if ((long long)lo32 != run) ovf = 1;
                kA[pos] = prev;
                vA[pos] = lo32;
                ++pos;
              }
              run = 0;
            }
            run += val[v];
            prev = key[v];
            have = true;
          }
        }
      }
      csum = __shfl_sync(FULL, S, 31);
      ckey = __shfl_sync(FULL, lastk, (nout - 1) / VT);
      cvalid = true;
      __syncwarp();
      // the closed runs leave through shared memory: the warp writes contiguous 256-byte rows
#pragma unroll
      for (int t = 0; t < VT; ++t) {
        const int k = lane + 32 * t;
        if (k < total) {
          Y[obase + k] = make_int2(kA[k], vA[k]);
          if (key_out) key_out[pa + obase + k] = g;
        }
      }
      obase += total;
      __syncwarp();
    }
    if (cvalid) {                                  // the last run of the item
      if (lane == 0) {
        const int lo32 = (int)csum;
        if ((long long)lo32 != csum) ovf = 1;
        Y[obase] = make_int2(ckey, lo32);
        if (key_out) key_out[pa + obase] = g;
      }
      ++obase;
    }
    if (kReduce && lane == 0) row_count[g] = obase;
  }
  if (kReduce && ovf) *overflow = 1;
}

// Moves the reduced coarse rows (written at their input offsets by the kReduce round) to their final,
// contiguous places: row g = new_indptr[g+1] - new_indptr[g] elements from in + indptr[group_first[g]].
__global__ void __launch_bounds__(256)
coarsen_compact_kernel(const int2* __restrict__ in, const int64_t* __restrict__ indptr,
                       const int32_t* __restrict__ group_first, const int64_t* __restrict__ new_indptr,
                       int n_groups, int2* __restrict__ out, int32_t* __restrict__ key_out) {
  const int lane = threadIdx.x & 31;
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < n_groups; g += n_warps) {
    const int2* __restrict__ src = in + indptr[group_first[g]];
    const int64_t o0 = new_indptr[g];
    const int n = (int)(new_indptr[g + 1] - o0);
    for (int k0 = 0; k0 < n; k0 += 128) {
      int2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + lane + 32 * u;
        v[u] = (k < n) ? ld_stream_int2(src + k) : make_int2(0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + lane + 32 * u;
        if (k < n) {
          out[o0 + k] = v[u];
          if (key_out) key_out[o0 + k] = (int)g;
        }
      }
    }
  }
}

// bin_map[i] = new_off[c] + (i - off[c]) / factor for bin i of chromosome c (closed form of the
// reference's coordinate re-binning, _reduce.py:597-614), and group_first[g] = first fine bin of
// coarse bin g (g = n_new: n_bins).
__global__ void coarsen_maps_kernel(const int64_t* __restrict__ off, const int64_t* __restrict__ new_off,
                                    int n_chroms, int factor, int n_bins, int n_new,
                                    int32_t* __restrict__ bin_map, int32_t* __restrict__ group_first) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_bins) {
    int lo = 0, hi = n_chroms;                     // largest c with off[c] <= i
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (off[mid] <= i) lo = mid; else hi = mid;
    }
    bin_map[i] = (int)(new_off[lo] + (i - off[lo]) / factor);
  }
  if (i <= n_new) {
    int g = n_bins;
    if (i < n_new) {
      int lo = 0, hi = n_chroms;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (new_off[mid] <= i) lo = mid; else hi = mid;
      }
      g = (int)(off[lo] + (i - new_off[lo]) * (int64_t)factor);
    }
    group_first[i] = g;
  }
}

// Block table of RemapFn: entry b describes the bins [b << shift, (b + 1) << shift).
__global__ void coarsen_block_table_kernel(const int64_t* __restrict__ off, const int64_t* __restrict__ new_off,
                                           int n_chroms, int factor, int n_bins, int shift, int n_blocks,
                                           int4* __restrict__ blocks) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  const int64_t i0 = (int64_t)b << shift;
  const int64_t end = min(i0 + ((int64_t)1 << shift), (int64_t)n_bins);
  auto chrom_of = [&](int64_t i) {                 // largest c with off[c] <= i: the non-empty chromosome holding bin i
    int lo = 0, hi = n_chroms;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (off[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
  };
  const int c0 = chrom_of(i0);
  const int pad0 = (int)(new_off[c0] * factor - off[c0]);
  const int64_t bnd = off[c0 + 1];
  int4 e = make_int4(0x7fffffff, pad0, pad0, 0);
  if (bnd < end) {
    const int c1 = chrom_of(bnd);
    if (off[c1 + 1] < end) e = make_int4(-1, 0, 0, 0);          // a second boundary: gather instead
    else e = make_int4((int)bnd, pad0, (int)(new_off[c1] * factor - off[c1]), 0);
  }
  blocks[b] = e;
}

// Generic aggregation of one value column over the runs of equal (key, other) of a sorted stream
// whose vals are {other, index into the source column}: what `groupby([bin1_id, bin2_id],
// sort=True).aggregate({col: agg})` of CoolerCoarsener._aggregate (_reduce.py:616-620) computes for
// any numeric column and agg in {sum, min, max, first, last, mean}.  Every run is walked by one
// thread in stream order (= pixel order inside the group, as pandas does); integers accumulate in
// int64, floating-point columns in float64.
enum { AGG_SUM = 0, AGG_MIN = 1, AGG_MAX = 2, AGG_FIRST = 3, AGG_LAST = 4, AGG_MEAN = 5 };

template <typename T, typename ACC>
__global__ void __launch_bounds__(BLOCK_THREADS)
runs_aggregate_kernel(const int32_t* __restrict__ keys, const int2* __restrict__ vals, int64_t n,
                      int64_t n_tiles, const int64_t* __restrict__ tile_off, const T* __restrict__ src,
                      int op, int32_t* __restrict__ key_out, int32_t* __restrict__ other_out,
                      ACC* __restrict__ acc_out, double* __restrict__ mean_out) {
  const int64_t t = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (t >= n_tiles) return;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int64_t p0 = t * (int64_t)WARP_TILE, p1 = min(p0 + (int64_t)WARP_TILE, n);
  int64_t base = tile_off[t];
  for (int64_t c = p0; c < p1; c += 32) {
    const int64_t p = c + lane;
    const bool head = (p < p1) && is_head(keys, vals, p);
    const unsigned hb = __ballot_sync(0xffffffffu, head);
    if (head) {
      const int k = keys[p];
      const int o = vals[p].x;
      ACC a = (ACC)src[vals[p].y];
      long long len = 1;
      for (int64_t e = p + 1; e < n && keys[e] == k && vals[e].x == o; ++e, ++len) {   // run may cross tiles
        const ACC v = (ACC)src[vals[e].y];
        if (op == AGG_SUM || op == AGG_MEAN) a += v;
        else if (op == AGG_MIN) a = v < a ? v : a;
        else if (op == AGG_MAX) a = v > a ? v : a;
        else if (op == AGG_LAST) a = v;
      }
      const int64_t out = base + __popc(hb & lt);
      if (key_out) key_out[out] = k;
      if (other_out) other_out[out] = o;
      if (op == AGG_MEAN) mean_out[out] = (double)a / (double)len;
      else acc_out[out] = a;
    }
    base += __popc(hb);
  }
}

__global__ void unpack_pixels_kernel(const int32_t* __restrict__ bin1, const int2* __restrict__ val,
                                     int64_t n, int64_t* __restrict__ b1, int64_t* __restrict__ b2,
                                     int32_t* __restrict__ cnt) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int2 v = val[i];
    b1[i] = bin1[i];
    b2[i] = v.x;
    cnt[i] = v.y;
  }
}

}  // namespace cb200

using namespace cb200;

// outputs per lane and chunk of the merge kernel: 7 (measured on the 10 kb table, factor 2 / 5: VT 5: 1.68 / 1.05 ms,
// 7: 1.59 / 0.96, 9: 1.65 / 1.00, 11: 1.85 / 1.18 for the fused round; profiles/r02d_profile_coarsen_fused_10kb.log)
constexpr int MG_VT = 7;

template <bool kReduce>
static int launch_merge(const cb200_px* in, cb200_px* out, const int64_t* indptr, const int32_t* group_first,
                        int32_t n_groups, int32_t round, int32_t pairs_per_group, const cb200_remap* rm,
                        int32_t* key_out, int32_t* row_count, int32_t* overflow, int64_t* claim, void* stream) {
  if (n_groups <= 0 || pairs_per_group <= 0) return 0;
  if (round < 0 || round > 30) {
    set_error_msg("coarsen_merge: bad round");
    return -1;
  }
  cudaStream_t s = (cudaStream_t)stream;
  CB_CHECK(cudaMemsetAsync(claim, 0, sizeof(int64_t), s));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t items = (int64_t)n_groups * pairs_per_group;
  int64_t blocks = ceil_div64(items, MG_WARPS);
  if (blocks > (int64_t)sms * 6) blocks = (int64_t)sms * 6;      // up to 48 warps per SM claim the pairs dynamically
  RemapFn fn{nullptr, nullptr, 0, 0u, 0};
  if (rm) {
    if (rm->bin_map == nullptr || (rm->blocks != nullptr && (rm->block_shift < 0 || rm->block_shift > 30 ||
                                                             rm->magic_shift < 0 || rm->magic_shift > 31))) {
      set_error_msg("coarsen_merge: bad remap descriptor");
      return -1;
    }
    fn = RemapFn{rm->bin_map, reinterpret_cast<const int4*>(rm->blocks), rm->block_shift, rm->magic, rm->magic_shift};
  }
  coarsen_merge_kernel<MG_VT, kReduce><<<(unsigned)blocks, MG_WARPS * 32, 0, s>>>(
      reinterpret_cast<const int2*>(in), reinterpret_cast<int2*>(out), indptr, group_first, n_groups, round,
      pairs_per_group, rm != nullptr, fn, key_out, row_count, overflow,
      reinterpret_cast<unsigned long long*>(claim));
  CB_LAUNCH_CHECK("coarsen_merge");
  return 0;
}

extern "C" {


int cb200_coarsen_remap(const cb200_copy* csr, int32_t n_rows, const int32_t* bin_map,
                        int32_t* key_out, cb200_px* val_out, void* stream) {
  (void)n_rows;
  const int64_t n_tiles = ceil_div64(csr->nnz, WARP_TILE);
  if (n_tiles <= 0) return 0;
  coarsen_remap_kernel<<<(unsigned)ceil_div64(n_tiles, WARPS_PER_BLOCK), BLOCK_THREADS, 0,
                         (cudaStream_t)stream>>>(reinterpret_cast<const int2*>(csr->px), csr->indptr,
                                                 csr->tile_row, csr->nnz, n_tiles, bin_map, key_out,
                                                 reinterpret_cast<int2*>(val_out));
  CB_LAUNCH_CHECK("coarsen_remap");
  return 0;
}

int cb200_coarsen_maps(const int64_t* chrom_offset, const int64_t* new_chrom_offset, int32_t n_chroms,
                       int32_t factor, int32_t n_bins, int32_t n_new, int32_t* bin_map,
                       int32_t* group_first, void* stream) {
  if (n_chroms <= 0 || factor <= 0) {
    set_error_msg("coarsen_maps: bad arguments");
    return -1;
  }
  const int n = (n_bins > n_new + 1) ? n_bins : n_new + 1;
  coarsen_maps_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(
      chrom_offset, new_chrom_offset, n_chroms, factor, n_bins, n_new, bin_map, group_first);
  CB_LAUNCH_CHECK("coarsen_maps");
  return 0;
}

int cb200_coarsen_block_table(const int64_t* chrom_offset, const int64_t* new_chrom_offset, int32_t n_chroms,
                              int32_t factor, int32_t n_bins, int32_t block_shift, int32_t* blocks_out,
                              void* stream) {
  if (n_chroms <= 0 || factor <= 0 || block_shift < 0 || block_shift > 30) {
    set_error_msg("coarsen_block_table: bad arguments");
    return -1;
  }
  if (n_bins <= 0) return 0;
  const int n_blocks = (int)(((int64_t)n_bins - 1) >> block_shift) + 1;
  coarsen_block_table_kernel<<<(unsigned)ceil_div64(n_blocks, 128), 128, 0, (cudaStream_t)stream>>>(
      chrom_offset, new_chrom_offset, n_chroms, factor, n_bins, block_shift, n_blocks,
      reinterpret_cast<int4*>(blocks_out));
  CB_LAUNCH_CHECK("coarsen_block_table");
  return 0;
}

int cb200_coarsen_merge_round(const cb200_px* in, cb200_px* out, const int64_t* indptr,
                              const int32_t* group_first, int32_t n_groups, int32_t round,
                              int32_t pairs_per_group, const cb200_remap* remap, int32_t* key_out,
                              int64_t* claim, void* stream) {
  return launch_merge<false>(in, out, indptr, group_first, n_groups, round, pairs_per_group, remap, key_out,
                             nullptr, nullptr, claim, stream);
}

int cb200_coarsen_merge_reduce(const cb200_px* in, cb200_px* out, const int64_t* indptr,
                               const int32_t* group_first, int32_t n_groups, int32_t round,
                               const cb200_remap* remap, int32_t* row_count, int32_t* overflow,
                               int64_t* claim, void* stream) {
  return launch_merge<true>(in, out, indptr, group_first, n_groups, round, 1, remap, nullptr, row_count,
                            overflow, claim, stream);
}

int cb200_coarsen_compact(const cb200_px* in, const int64_t* indptr, const int32_t* group_first,
                          const int64_t* new_indptr, int32_t n_groups, cb200_px* out, int32_t* key_out,
                          void* stream) {
  if (n_groups <= 0) return 0;
  int64_t blocks = ceil_div64(n_groups, 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  coarsen_compact_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const int2*>(in), indptr, group_first, new_indptr, n_groups,
      reinterpret_cast<int2*>(out), key_out);
  CB_LAUNCH_CHECK("coarsen_compact");
  return 0;
}

int cb200_runs_count(const int32_t* keys, const cb200_px* vals, int64_t n, int32_t* tile_count,
                     void* stream) {
  const int64_t n_tiles = ceil_div64(n, WARP_TILE);
  if (n_tiles <= 0) return 0;
  runs_count_kernel<<<(unsigned)ceil_div64(n_tiles, WARPS_PER_BLOCK), BLOCK_THREADS, 0,
                      (cudaStream_t)stream>>>(keys, reinterpret_cast<const int2*>(vals), n, n_tiles,
                                              tile_count);
  CB_LAUNCH_CHECK("runs_count");
  return 0;
}

int cb200_runs_write(const int32_t* keys, const cb200_px* vals, int64_t n, const int64_t* tile_off,
                     int32_t swap, int32_t* key_out, cb200_px* val_out, int64_t* sum64_out,
                     int32_t* overflow, void* stream) {
  const int64_t n_tiles = ceil_div64(n, WARP_TILE);
  if (n_tiles <= 0) return 0;
  runs_write_kernel<<<(unsigned)ceil_div64(n_tiles, WARPS_PER_BLOCK), BLOCK_THREADS, 0,
                      (cudaStream_t)stream>>>(keys, reinterpret_cast<const int2*>(vals), n, n_tiles,
                                              tile_off, swap, key_out,
                                              reinterpret_cast<int2*>(val_out), sum64_out, overflow);
  CB_LAUNCH_CHECK("runs_write");
  return 0;
}

int cb200_runs_aggregate(const int32_t* keys, const cb200_px* vals, int64_t n, const int64_t* tile_off,
                         const void* src, int32_t src_dtype, int32_t op, int32_t* key_out,
                         int32_t* other_out, int64_t* int_out, double* float_out, void* stream) {
  const int64_t n_tiles = ceil_div64(n, WARP_TILE);
  if (n_tiles <= 0) return 0;
  if (op < AGG_SUM || op > AGG_MEAN || src_dtype < 0 || src_dtype > 3) {
    set_error_msg("runs_aggregate: bad op or dtype");
    return -1;
  }
  const unsigned grid = (unsigned)ceil_div64(n_tiles, WARPS_PER_BLOCK);
  const int2* v = reinterpret_cast<const int2*>(vals);
  cudaStream_t s = (cudaStream_t)stream;
  switch (src_dtype) {      // 0: int32, 1: int64 (int64 accumulation); 2: float32, 3: float64 (float64 accumulation)
    case 0: runs_aggregate_kernel<int32_t, long long><<<grid, BLOCK_THREADS, 0, s>>>(
                keys, v, n, n_tiles, tile_off, (const int32_t*)src, op, key_out, other_out, (long long*)int_out, float_out); break;
    case 1: runs_aggregate_kernel<long long, long long><<<grid, BLOCK_THREADS, 0, s>>>(
                keys, v, n, n_tiles, tile_off, (const long long*)src, op, key_out, other_out, (long long*)int_out, float_out); break;
    case 2: runs_aggregate_kernel<float, double><<<grid, BLOCK_THREADS, 0, s>>>(
                keys, v, n, n_tiles, tile_off, (const float*)src, op, key_out, other_out, float_out, float_out); break;
    default: runs_aggregate_kernel<double, double><<<grid, BLOCK_THREADS, 0, s>>>(
                keys, v, n, n_tiles, tile_off, (const double*)src, op, key_out, other_out, float_out, float_out); break;
  }
  CB_LAUNCH_CHECK("runs_aggregate");
  return 0;
}

int cb200_unpack_pixels(const int32_t* bin1, const cb200_px* val, int64_t n, int64_t* bin1_out,
                        int64_t* bin2_out, int32_t* count_out, void* stream) {
  if (n <= 0) return 0;
  int64_t blocks = ceil_div64(n, 256 * 4);
  if (blocks > 148 * 32) blocks = 148 * 32;
  unpack_pixels_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      bin1, reinterpret_cast<const int2*>(val), n, bin1_out, bin2_out, count_out);
  CB_LAUNCH_CHECK("unpack_pixels");
  return 0;
}

}  // extern "C"

// ---- ingest binning: pairs -> (bin1, bin2) records -----------------------------
// Reference: open2c/cooler src/cooler/create/_ingest.py:68-214 (_sanitize_records: drop records on
// unknown chromosomes, 1-based -> 0-based, bounds checks, reflect / drop / raise lower-triangle
// records, bin id = chrom_binoffset[chrom] + pos // binsize); the groupby(bin1, bin2).size() of
// aggregate_records (:452-504) is then the same sort + reduce-by-key as coarsening.
namespace cb200 {

__global__ void bin_pairs_kernel(const int32_t* __restrict__ chrom1, const int64_t* __restrict__ pos1,
                                 const int32_t* __restrict__ chrom2, const int64_t* __restrict__ pos2,
                                 int64_t n, const int32_t* __restrict__ chrom_binoffset,
                                 const int64_t* __restrict__ chromsizes, int64_t binsize, int one_based,
                                 int tril_action, int n_bins, int32_t* __restrict__ key_out,
                                 int2* __restrict__ val_out, int32_t* __restrict__ flags) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int f_neg = 0, f_exc = 0, f_tril = 0, dropped = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int c1 = chrom1[i], c2 = chrom2[i];
    int key = n_bins, other = n_bins, cnt = 0;            // dropped records sort behind every bin
    if (c1 >= 0 && c2 >= 0) {
      int64_t a1 = pos1[i] - one_based, a2 = pos2[i] - one_based;
      if (a1 < 0 || a2 < 0) f_neg = 1;
      else if (a1 > chromsizes[c1] || a2 > chromsizes[c2]) f_exc |= 1;
      else {
        bool keep = true;
        if (tril_action != CB200_TRIL_NONE && (c1 > c2 || (c1 == c2 && a1 > a2))) {
          if (tril_action == CB200_TRIL_REFLECT) {
            const int tc = c1; c1 = c2; c2 = tc;
            const int64_t ta = a1; a1 = a2; a2 = ta;
          } else if (tril_action == CB200_TRIL_DROP) keep = false;
          else { f_tril = 1; keep = false; }
        }
        if (keep) {
          other = chrom_binoffset[c1] + (int)(a1 / binsize);   // bin1
          key = chrom_binoffset[c2] + (int)(a2 / binsize);     // bin2: first sort key
          cnt = 1;
          // a position equal to the length of the LAST chromosome passes the reference's anchor check
          // (`> chromsize`) but lands on bin n_bins: its bounds check then fails
          // (create/_ingest.py:401-410).  Never let it collide with the dropped-record sentinel.
          if (other >= n_bins || key >= n_bins) { f_exc = 2; other = key = n_bins; cnt = 0; }
        }
      }
    }
    if (cnt == 0) ++dropped;
    key_out[i] = key;
    val_out[i] = make_int2(other, cnt);
  }
  if (f_neg) atomicOr(&flags[0], 1);
  if (f_exc) atomicOr(&flags[1], f_exc);
  if (f_tril) atomicOr(&flags[2], 1);
  dropped = warp_sum(dropped);
  if ((threadIdx.x & 31) == 0 && dropped) atomicAdd(&flags[3], dropped);
}

}  // namespace cb200

extern "C" int cb200_bin_pairs(const int32_t* chrom1, const int64_t* pos1, const int32_t* chrom2,
                               const int64_t* pos2, int64_t n, const int32_t* chrom_binoffset,
                               const int64_t* chromsizes, int64_t binsize, int32_t one_based,
                               int32_t tril_action, int32_t n_bins, int32_t* key_out,
                               cb200_px* val_out, int32_t* flags, void* stream) {
  if (binsize <= 0 || tril_action < 0 || tril_action > 3) {
    cb200::set_error_msg("bin_pairs: bad arguments");
    return -1;
  }
  if (n <= 0) return 0;
  int64_t blocks = cb200::ceil_div64(n, 256 * 4);
  if (blocks > 148 * 32) blocks = 148 * 32;
  cb200::bin_pairs_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      chrom1, pos1, chrom2, pos2, n, chrom_binoffset, chromsizes, binsize, one_based, tril_action, n_bins,
      key_out, reinterpret_cast<int2*>(val_out), flags);
  CB_LAUNCH_CHECK("bin_pairs");
  return 0;
}
