/*
 * cooler_b200 -- C ABI of the B200-native (sm_100a) ICE balancing + coarsening
 * hot path of open2c/cooler.  Shared library: cooler_b200/libcooler_b200.so.
 *
 * The reference (open2c/cooler v0.10.4) is pure Python and has no FFI; the
 * entry points below are what a binding for its hot path would call.  Each one
 * cites the reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); all
 *     calls are asynchronous on that stream and never synchronise;
 *   - return value 0 = success, otherwise a cudaError_t (or -1 for argument
 *     errors); cb200_last_error() returns a thread-local description;
 *   - the caller owns every buffer (in Python: torch tensors); the library
 *     allocates nothing;
 *   - pixel copies are arrays of cb200_px {other, count} (8 bytes), 16-byte
 *     aligned and padded to an even number of elements;
 *   - a "warp tile" is cb200_warp_tile() consecutive pixels of one copy;
 *     n_tiles = ceil(nnz / warp_tile).
 */
#ifndef COOLER_B200_H
#define COOLER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int32_t other; int32_t count; } cb200_px;

/* ---- library ---------------------------------------------------------- */
int         cb200_abi_version(void);
const char* cb200_last_error(void);
int         cb200_set_device(int device);
int         cb200_warp_tile(void);

/* ---- primitives ------------------------------------------------------- */
/* out[0..n] = exclusive prefix sum of in[0..n-1] (out has n+1 entries). */
size_t cb200_scan_workspace_bytes(int64_t n);
int cb200_exclusive_scan_i32(const int32_t* in, int64_t* out, int64_t n,
                             void* workspace, size_t workspace_bytes, void* stream);

/* Stable LSD radix sort of (key, 8-byte value) pairs on key bits
 * [key_shift, key_shift + key_bits).  Result lands in keys_out/vals_out; *_tmp are scratch of
 * the same size.  Replaces nothing in the reference directly: it builds the
 * transposed (bin2-major) copy that the reference re-derives every pass with
 * np.bincount(bin2_id) (src/cooler/_balance.py:63-69), and is the sort of the
 * groupby(sort=True) in src/cooler/_reduce.py:616-620. */
size_t cb200_sort_workspace_bytes(int64_t n);
int cb200_sort_pairs(const int32_t* keys_in, const cb200_px* vals_in,
                     int32_t* keys_out, cb200_px* vals_out,
                     int32_t* keys_tmp, cb200_px* vals_tmp,
                     int64_t n, int key_shift, int key_bits,
                     void* workspace, size_t workspace_bytes, void* stream);

/* Exact medians of float64 segments by radix selection (the medians of the bin filters,
 * src/cooler/_balance.py:400-409 and util.mad, util.py:513-514, on O(n_bins) vectors -- 3.1e6 bins at
 * 1 kb made numpy's three np.median calls the largest single item of the end-to-end time).  For
 * every segment s = [seg_offset[s], seg_offset[s+1]) of x: count_out[s] = number of valid elements
 * (positive_only != 0: x > 0, as in `c_marg[c_marg > 0]`; else: not NaN) and mid_out[2s], mid_out[2s+1]
 * = the valid elements of rank (count-1)/2 and count/2 (NaN when count == 0); the median is their
 * mean.  Selection returns input elements, hence bit-identical to np.median.  max_seg_len sizes the
 * grid. */
size_t cb200_median_workspace_bytes(int32_t n_segs);
int cb200_segment_medians(const double* x, const int64_t* seg_offset, int32_t n_segs, int64_t max_seg_len,
                          int32_t positive_only, double* mid_out, int64_t* count_out, void* workspace,
                          size_t workspace_bytes, void* stream);

/* ---- pixel table build (replaces parallel.chunkgetter re-reading the file
 *      every pass, src/cooler/parallel.py:276-304, and the per-chunk filter
 *      functions _zero_diags/_zero_trans/_zero_cis, src/cooler/_balance.py:34-54,
 *      which here are applied ONCE by compaction) --------------------------- */
/* (bin2_id int64, count int32) -> packed px {bin2, count}. */
int cb200_pack_pixels(const int64_t* bin2_id, const int32_t* count, int64_t nnz,
                      cb200_px* px_out, void* stream);
/* Same for a bin2_id column already narrowed to int32 by the reader (h5py:
 * dset.astype("int32")[:]), which cuts the host->device copy from 12 to 8 B/pixel. */
int cb200_pack_pixels_i32(const int32_t* bin2_id, const int32_t* count, int64_t nnz,
                          cb200_px* px_out, void* stream);
/* tile_row[t] = row that contains pixel t*warp_tile (n_tiles entries). */
int cb200_tile_rows(const int64_t* indptr, int32_t n_rows, int64_t nnz,
                    int32_t* tile_row, void* stream);
/* indptr[r] = first position p with ((keys[p] - key_base) >> shift) >= r, r = 0..n_rows
 * (keys sorted, all >= key_base). */
int cb200_indptr_from_sorted(const int32_t* keys, int64_t n, int32_t n_rows, int32_t key_base,
                             int32_t shift, int64_t* indptr, void* stream);
/* Column strips ("cache blocking" of the gathered bias vector): a copy can be
 * split into strips by other >> shift, each strip again a row-ordered copy, so
 * that the slice of the bias vector one strip gathers from stays L1-resident.
 * swap: (key, {a, v}) -> (a, {key, v}).  finish: elements grouped by strip
 * (key >> shift) -> px {key, count} and row ids, element i written at
 * i + pad_before[strip] so that every strip starts 16-byte aligned. */
int cb200_swap_key_other(const int32_t* keys, const cb200_px* vals, int64_t n,
                         int32_t* keys_out, cb200_px* vals_out, void* stream);
int cb200_strip_finish(const int32_t* keys, const cb200_px* vals, int64_t n, int32_t shift,
                       const int64_t* pad_before, cb200_px* px_out, int32_t* rows_out,
                       void* stream);
/* Far-field blocks: elements (key = row, val = {other, count}) sorted by (row >> RS,
 * other >> col_shift, row, other) -> records {swz(row & (R - 1)) << 16 | (other & (C-1)), count}
 * (swz: bank swizzle of the row id inside its warp slice, see far_swizzle in csrc/far.cuh)
 * and the slot id g = (((row >> RS) - rb0) * n_j + (other >> col_shift) - j0) * warps
 * + ((row & (R - 1)) * warps >> RS) of every element (non-decreasing; cb200_indptr_from_sorted
 * on g gives cb200_far_copy.ptr). */
int cb200_far_finish(const int32_t* rows, const cb200_px* vals, int64_t n, int32_t col_shift,
                     int32_t rb0, int32_t j0, int32_t n_j, int32_t warps,
                     cb200_px* px_out, int32_t* slot_out, void* stream);
/* nch_out[s] = ceil((ptr_rec[s+1] - ptr_rec[s]) / 64): 64-record chunks of every slot. */
int cb200_far_slot_chunks(const int64_t* ptr_rec, int64_t n_slots, int32_t* nch_out, void* stream);
/* Lane-interleaved layout of the far-field records: slot s (records [ptr_rec[s], ptr_rec[s+1])
 * of `rec`, sorted by row) occupies chunks [chunk_ptr[s], chunk_ptr[s+1]) of px_out; its
 * records are cut into 32 contiguous pieces of L = 2 * n_chunks(s) records (lane l gets records
 * [l L, (l+1) L)), and chunk c holds records 2c, 2c+1 of every lane at positions 2 lane,
 * 2 lane + 1 -- one coalesced 128-bit load per lane.  Unused positions hold padding records
 * {R << 16, 0}.  px_out has total_chunks * 64 + 2 records. */
int cb200_far_interleave(const cb200_px* rec, const int32_t* slot, const int64_t* ptr_rec,
                         const int64_t* chunk_ptr, int64_t n, int64_t total_chunks, cb200_px* px_out,
                         void* stream);

/* Compact (4-byte) far-field records {swz(row & (R - 1)) : 12 | other & (C - 1) : 13 | count : 7}
 * (R <= 4096, col_shift <= 13): half the bytes per record of the blocked sweep.
 * cb200_far_pieces: pieces_out[i] = ceil(count_i / 127) (>= 1) = records element i needs;
 * flags[0] = 1 if any count > 127, flags[1] = 1 if any count < 0 (then the 8-byte records are kept).
 * cb200_far_expand: element i -> records [offsets[i], offsets[i+1]) (offsets = exclusive scan of the
 * pieces) with counts <= 127 that add up to count_i; order is preserved.
 * cb200_far_finish4 / cb200_far_interleave4: as cb200_far_finish / cb200_far_interleave for 4-byte
 * records; chunks are 64 records = 256 bytes, padding = count 0 on the row of the slot's last record. */
int cb200_far_pieces(const cb200_px* vals, int64_t n, int32_t* pieces_out, int32_t* flags, void* stream);
int cb200_far_expand(const int32_t* rows, const cb200_px* vals, int64_t n, const int64_t* offsets,
                     int32_t* rows_out, cb200_px* vals_out, void* stream);
int cb200_far_finish4(const int32_t* rows, const cb200_px* vals, int64_t n, int32_t col_shift,
                      int32_t rb0, int32_t j0, int32_t n_j, int32_t warps, uint32_t* rec_out,
                      int32_t* slot_out, void* stream);
int cb200_far_interleave4(const uint32_t* rec, const int32_t* slot, const int64_t* ptr_rec,
                          const int64_t* chunk_ptr, int64_t n, int64_t n_slots, int64_t total_chunks,
                          uint32_t* px_out, void* stream);

/* Static pixel filter.  A pixel (row r, other c) is KEPT iff
 *   !(ndiag > 0 && |r - c| < ndiag)                       (_zero_diags)
 *   && (keep == CB200_KEEP_ALL
 *       || (keep == CB200_KEEP_CIS   &&  row_lo[r] <= c < row_hi[r])   (_zero_trans)
 *       || (keep == CB200_KEEP_TRANS && !(row_lo[r] <= c < row_hi[r])))(_zero_cis)
 * where [row_lo[r], row_hi[r]) is the bin range of r's chromosome.  The copy may be a row
 * chunk: its local row 0 is bin `row_base`; r above, row_lo/row_hi indices and the emitted
 * row ids are absolute bin ids (local row + row_base). */
enum { CB200_KEEP_ALL = 0, CB200_KEEP_CIS = 1, CB200_KEEP_TRANS = 2,
       /* layout-only rules (no reference counterpart): split a copy into the band
        * |(c >> band_shift) - (r >> band_shift)| <= 1 around the diagonal and the rest */
       CB200_KEEP_BAND_NEAR = 3, CB200_KEEP_BAND_FAR = 4,
       /* c < row_lo[r]: the pixels of a row that lie in a chromosome BEFORE the row's own (they exist
        * only in square storage); input of cb200_cis_touch */
       CB200_KEEP_EARLIER_CHROM = 5 };
int cb200_filter_count(const cb200_px* px, const int64_t* indptr, const int32_t* tile_row,
                       int32_t n_rows, int64_t nnz,
                       const int32_t* row_lo, const int32_t* row_hi, int32_t ndiag, int32_t keep,
                       int32_t band_shift, int32_t row_base, int32_t* tile_count, void* stream);
/* Second pass: writes kept pixels, order preserved, at tile_off[t] + rank.
 * Optional outputs (NULL to skip): row_out[q] = row of kept pixel q;
 * tkey_out[q] = other, tval_out[q] = {row, count} (input of the transposition sort). */
int cb200_filter_write(const cb200_px* px, const int64_t* indptr, const int32_t* tile_row,
                       int32_t n_rows, int64_t nnz,
                       const int32_t* row_lo, const int32_t* row_hi, int32_t ndiag, int32_t keep,
                       int32_t band_shift, int32_t row_base, const int64_t* tile_off,
                       cb200_px* px_out, int32_t* row_out, int32_t* tkey_out, cb200_px* tval_out,
                       void* stream);

/* ---- marginals -------------------------------------------------------- */
/* One copy of the table as the row-walking kernels see it. */
typedef struct {
  const cb200_px* px;        /* [nnz] (+pad) */
  const int64_t*  indptr;    /* [n_rows + 1] */
  const int32_t*  tile_row;  /* [n_tiles] */
  int64_t         nnz;
} cb200_copy;

/* Integer marginals of one copy (prefilter passes A and B of balance_cooler,
 * src/cooler/_balance.py:375-393): for every row r
 *   nnz_out[r] = #{p in row r : count != 0}     (_binarize then _marginalize)
 *   sum_out[r] = sum of count over row r
 * exact in int64.  The marginal of the reference is row(CSR) + row(CSC).
 * scratch: [(2 * n_tiles + n_rows) * 2] int64.  Outputs are fully written. */
int cb200_marginals_i64(const cb200_copy* copy, int32_t n_rows,
                        int64_t* nnz_out, int64_t* sum_out, int64_t* scratch, void* stream);

/* Floating-point count columns (_init copies the column in its own dtype, src/cooler/_balance.py:25-26;
 * every later product is float64).  The table is built with px[k].count = position of the pixel in the
 * stored table, so filters, transposition and strip partition carry the values along as indices.
 *   cb200_attach_values: val_out[k] = values[px[k].count] (0 when the index is outside [0, n_values)),
 *     then px[k].count = k: the sweep reads val_out in stream order.
 *   cb200_values_nonzero: out[i] = values[i] != 0 ? 1.0 : 0.0   (_binarize, _balance.py:29-31). */
int cb200_attach_values(cb200_px* px, int64_t n, const double* values, int64_t n_values,
                        double* val_out, void* stream);
int cb200_values_nonzero(const double* values, int64_t n, double* out, void* stream);

/* cis_only on square-storage tables.  _balance_cisonly (src/cooler/_balance.py:127-196) balances the
 * chromosomes one after the other, scanning the ROWS of a chromosome (spans from bin1_offset, :150-151)
 * with `_zero_trans` piped.  A stored pixel (r, c) whose column lies in an EARLIER chromosome is zeroed,
 * but its weight is still bias[r] * bias[c] * 0 (_timesouterproduct, :57-60), and bias[c] is NaN once
 * c's chromosome is finished and c is a masked bin (:186 `bias[lo:hi][bias == 0] = nan`), or when that
 * whole chromosome came out NaN: the row marginal of r turns NaN and r's chromosome never converges
 * (NaN bias, scale and variance, max_iters passes, ConvergenceWarning).  Symmetric-upper tables hold no
 * such pixel.  To return what the reference returns, the host needs to know which chromosome pairs are
 * linked by such pixels:  for the n pixels selected by CB200_KEEP_EARLIER_CHROM (px[k].other = c,
 * rows[k] = r, absolute bin ids) set
 *   touch_any   [chrom(r) * n_chroms + chrom(c)] = 1
 *   touch_masked[chrom(r) * n_chroms + chrom(c)] = 1   if bias[c] == 0 or bias[c] is NaN
 * (bias = the vector after the bin filters, 0 = masked; bin_chrom[b] = chromosome index of bin b).
 * The flag matrices are caller-zeroed uint8[n_chroms * n_chroms]. */
int cb200_cis_touch(const cb200_px* px, const int32_t* rows, int64_t n, const int32_t* bin_chrom,
                    int32_t n_chroms, const double* bias, uint8_t* touch_any, uint8_t* touch_masked,
                    void* stream);

/* ---- ICE iteration (replaces the PASS C map-reduce + the O(n_bins) numpy
 *      update of _balance_genomewide/_cisonly/_transonly,
 *      src/cooler/_balance.py:72-260) ------------------------------------- */
/* A sub-copy swept by K3: a whole copy or one column strip of it.  It covers the
 * rows [row_base, row_base + row_count); indptr, tile_row and direct are local to
 * that range. */
typedef struct {
  const cb200_px* px; const int64_t* indptr; const int32_t* tile_row; int64_t nnz;
  double*  direct;           /* [row_count] rows closed strictly inside a warp tile */
  double*  pieces;           /* [2 * n_tiles] first / last row of every warp tile */
  int64_t  first_warp;       /* prefix sum over sub-copies of ceil(n_tiles / span) */
  int64_t  first_tile;       /* prefix sum over sub-copies of n_tiles */
  int32_t  gather_base;      /* `other` of this sub-copy lies in [gather_base, gather_base + gather_len) */
  int32_t  gather_len;
  int32_t  row_base;
  int32_t  row_count;
  const double* val;         /* NULL: integer counts in px[k].count.  Floating-point count columns (cb200_ice.float_counts
                                != 0): the pixel's value is val[px[k].count] -- px[k].count is an INDEX: k itself after
                                cb200_attach_values (sweeps read val in stream order), or the pixel's position in the
                                stored table while val is the raw column (prefilter marginals) */
} cb200_sub;

/* Far-field blocks (large, very sparse matrices: the gathered bias vector is far larger
 * than L1 and (row, column strip) segments hold only a few pixels).  The far pixels of a
 * copy are re-ordered by (row block = row >> RS, column tile = other >> col_shift, row,
 * other), RS = CB200_FAR_ROW_SHIFT, R = 1 << RS rows per block, and stored as records {(row & (R - 1)) << 16 | (other & (C - 1)), count},
 * C = 1 << col_shift.  A persistent CTA owns a unit = (row block, run of non-empty column
 * tiles): it keeps the R row sums in shared memory and stages the C-wide slice of the
 * bias vector of one tile after the other (one TMA bulk copy per tile, issued one tile ahead).  Warp k owns the rows
 * [k, k+1) * R / far_warps of the block; the records of one (row block, tile, warp slice)
 * = slot are stored as lane-interleaved 64-record chunks (cb200_far_interleave): every lane
 * sums the runs of equal rows of its own contiguous piece and adds them to the accumulators;
 * lane-boundary runs are merged once per slot by a warp-level segmented scan -- no atomics,
 * fixed summation order.  Unit u writes its row sums to far_parts[u * R ...]; the
 * per-bin kernels add the units of a row block in order. */
#ifndef CB200_FAR_ROW_SHIFT_VALUE
#define CB200_FAR_ROW_SHIFT_VALUE 12       /* build-time: log2 rows per far-field block (11..14); 12 measured best */
#endif
enum { CB200_FAR_ROW_SHIFT = CB200_FAR_ROW_SHIFT_VALUE };
/* The value this library was built with (hosts size far_parts and sort keys from it). */
int cb200_far_row_shift(void);
typedef struct {
  const void*     px;        /* [n_chunks * 64] (+2 pad) lane-interleaved records, see above: cb200_px, or uint32_t when cb200_ice.far_rec_bytes == 4 */
  const int64_t*  ptr;       /* [n_rb * n_j * far_warps + 1] first CHUNK of every slot (row block, column tile, warp row slice) */
  int64_t  nnz;              /* records without padding */
  int32_t  rb0, n_rb;        /* row blocks [rb0, rb0 + n_rb) */
  int32_t  j0, n_j;          /* column tiles [j0, j0 + n_j) */
  const int32_t* rb_unit;    /* [n_rb + 1] units (indices into far_units) of each row block */
} cb200_far_copy;
typedef struct {
  int32_t copy;              /* index into far_copies */
  int32_t rb;                /* absolute row block */
  int32_t t0, t1;            /* this unit handles the non-empty column tiles far_tiles[t0 .. t1) */
} cb200_far_unit;

typedef struct {
  const cb200_sub* subs;     /* DEVICE array: bin1-major sub-copies then bin2-major ones (static filters applied) */
  int32_t  n_subs;
  int32_t  span;             /* warp tiles per warp in K3 */
  int64_t  total_warps;      /* sum over sub-copies of ceil(n_tiles / span) */
  int64_t  total_tiles;      /* sum over sub-copies of n_tiles */
  int32_t  smem_bins;        /* 0: gather through L1/L2; else max gather_len (column strips staged in shared memory) */
  int64_t* claim;            /* [n_subs + 1] zero-initialised work-claim counters of the strip kernel (self re-arming) */
  int32_t  n_bins;
  int32_t  n_segs;           /* 1 (genome-wide / trans) or n_chroms (cis_only) */
  const int32_t* seg_offset; /* [n_segs + 1] bin range of each independently balanced segment */
  /* block table of the per-bin kernels: block b covers bins [blk_lo[b], blk_hi[b]) of segment blk_seg[b] */
  int32_t  n_blocks;
  const int32_t* blk_seg; const int32_t* blk_lo; const int32_t* blk_hi;
  const int32_t* seg_blk_offset;  /* [n_segs + 1] blocks of each segment */
  double*  bias;             /* [n_bins] in/out */
  const double* cweights;    /* [n_bins] or NULL (trans_only: 1/(1 - nbins_chrom/n_bins), _balance.py:214-220) */
  double*  w;                /* [n_bins] gathered vector = bias * cweights (may alias bias when cweights == NULL) */
  double*  s;                /* [n_bins] per-bin sum  sum_j w_j x_ij  (row + column part); allreduced across ranks */
  double*  marg;             /* [n_bins] scratch */
  double*  blk_sum; double* blk_cnt; double* blk_dev;  /* [n_blocks] scratch */
  /* per-segment state */
  double*  seg_mean; double* seg_nnz;        /* [n_segs] scratch */
  double*  seg_scale; double* seg_var;       /* [n_segs] out: nzmarg.mean() and nzmarg.var() of the last pass */
  int32_t* seg_iters;        /* [n_segs] out: number of "variance is" passes */
  int32_t* seg_done;         /* [n_segs] out: 1 = converged or empty */
  int32_t* seg_empty;        /* [n_segs] out: 1 = no nonzero marginal (reference sets bias to NaN) */
  int32_t* all_done;         /* [1] out: every segment done -> later launches are no-ops */
  int32_t* tickets;          /* [2] zero-initialised; elect the block that finishes a reduction (self re-arming) */
  double*  var_hist;         /* [max_iters * n_segs] variance per pass (NaN where not run) or NULL */
  double   tol;
  int32_t  max_iters;
  /* far-field blocks (n_far_units == 0: none) */
  int32_t  n_far_copies;
  const cb200_far_copy* far_copies;   /* DEVICE array */
  const cb200_far_unit* far_units;    /* DEVICE array, ordered by (copy, row block, t0) */
  const int32_t* far_tiles;           /* column tile of every non-empty (row block, tile), grouped by unit */
  double*  far_parts;                 /* [n_far_units * R], R = 1 << CB200_FAR_ROW_SHIFT */
  int64_t* far_claim;                 /* [2] zero-initialised (self re-arming) */
  int32_t  n_far_units;
  int32_t  far_col_shift;             /* log2 C, 1..13; (rows + 2 C) * 8 B must fit the 227 KB of shared memory */
  int32_t  far_warps;                 /* 16: warps per CTA = row slices per block */
  int32_t  far_rec_bytes;             /* 8: cb200_px records; 4: compact records (cb200_far_finish4), every far copy alike */
  int32_t  float_counts;              /* != 0: every sub-copy carries cb200_sub.val (float64 values, _balance.py:25-26 copies
                                         whatever dtype the count column has); far-field blocks are not available then */
} cb200_ice;

/* Tuning of K3 (optional).  ice_unroll: 0 or 4 -- the sweep kernels keep 4 128-bit stream loads in flight
 * per lane (the 1 / 2 / 8 variants lost the measurements and were removed; any other value is an error).
 * selective_l1: -1 automatic, 0/1 = whether gathers far from the diagonal bypass L1 allocation in the
 * L1/L2-gather kernel. */
int cb200_set_tuning(int32_t ice_unroll, int32_t selective_l1);
/* K3: the marginal sweeps of every sub-copy (bin1-major rows, bin2-major columns), one
 * launch, followed by one launch over the far-field blocks if there are any.  No-op once
 * *all_done. */
int cb200_ice_sweep(const cb200_ice* st, void* stream);
/* s[i] = rowpart(i) + colpart(i) -- the vector ranks all-reduce.  No-op once *all_done. */
int cb200_ice_combine(const cb200_ice* st, void* stream);
/* K4 (two launches): marg = w*s, per-segment nonzero mean/variance, bias /= marg/mean,
 * convergence flags, iteration counters.  `iter` = 0-based index of this pass.
 * fused_combine != 0: s is computed from the sweep outputs inside the first kernel
 * (single GPU; cb200_ice_combine is then not needed); 0: st->s is read as given
 * (multi-GPU: after the all-reduce). */
int cb200_ice_update(const cb200_ice* st, int32_t iter, int32_t fused_combine, void* stream);
/* Multi-GPU K4 with the all-reduce fused in: peer_s is a DEVICE array of `world`
 * pointers, peer_s[r] + peer_offset = rank r's local s vector (written by
 * cb200_ice_combine into peer-mapped symmetric memory; the caller places a
 * cross-rank barrier between the two calls).  Every rank sums the `world` vectors in
 * rank order over NVLink, so all ranks get identical bits without an NCCL call. */
int cb200_ice_update_peer(const cb200_ice* st, int32_t iter, const double* const* peer_s,
                          int32_t world, int64_t peer_offset, void* stream);

/* Two-stage form of the same fused all-reduce for long per-bin vectors (1 kb matrices: 25 MB per
 * rank): cb200_ice_reduce_peer sums -- in rank order -- the slice [rank, rank+1) * ceil(n_bins /
 * world) of every rank's local s (peer_s[r] + peer_offset) into red_local[bin]; after a second
 * cross-rank barrier cb200_ice_update_gather reads every bin from its owner's buffer
 * (peer_red[owner] + peer_offset) and runs K4.  NVLink traffic per rank and pass drops from
 * (world - 1) to 2 (world - 1) / world vectors; every rank still sees identical bits. */
int cb200_ice_reduce_peer(const cb200_ice* st, const double* const* peer_s, int32_t world, int32_t rank,
                          int64_t peer_offset, double* red_local, void* stream);
int cb200_ice_update_gather(const cb200_ice* st, int32_t iter, const double* const* peer_red,
                            int32_t world, int64_t peer_offset, void* stream);

/* ---- coarsen (replaces CoolerCoarsener._aggregate's join + pandas
 *      groupby(sort=True).sum(), src/cooler/_reduce.py:582-620) ------------ */
/* Fused bin remap of one CSR copy: for every pixel (r, c, v)
 *   key_out = bin_map[c], val_out = {bin_map[r], v}   (sort input). */
int cb200_coarsen_remap(const cb200_copy* csr, int32_t n_rows, const int32_t* bin_map,
                        int32_t* key_out, cb200_px* val_out, void* stream);
/* Bin maps of one coarsening step on the device: bin_map[n_bins] (old bin -> new bin = new_off[c] +
 * (bin - off[c]) / factor, the closed form of src/cooler/_reduce.py:597-614 for chromosome c) and
 * group_first[n_new + 1] (first old bin of every new bin).  chrom_offset / new_chrom_offset: DEVICE
 * int64 [n_chroms + 1]; chromosomes without bins are allowed. */
int cb200_coarsen_maps(const int64_t* chrom_offset, const int64_t* new_chrom_offset, int32_t n_chroms,
                       int32_t factor, int32_t n_bins, int32_t n_new, int32_t* bin_map,
                       int32_t* group_first, void* stream);
/* Column remap of merge round 0.  bin_map (cb200_coarsen_maps) is always valid; with `blocks`
 * (cb200_coarsen_block_table: 4 int32 per block of 2^block_shift bins, 16-byte aligned) the kernels compute
 * new = (col + pad) / factor from the block's entry instead of gathering bin_map[col] -- a row's columns are
 * spread over the whole genome, so the gathers cost one L1 tag lookup per lane.  magic / magic_shift:
 * x / factor == umulhi(x, magic) >> magic_shift for 0 <= x < 2^31 (l = ceil(log2 factor),
 * magic = ceil(2^(31+l) / factor), magic_shift = l - 1); requires n_new * factor < 2^31. */
typedef struct {
  const int32_t* bin_map;
  const int32_t* blocks;     /* NULL: gather from bin_map */
  int32_t  block_shift;
  uint32_t magic;
  int32_t  magic_shift;
} cb200_remap;
int cb200_coarsen_block_table(const int64_t* chrom_offset, const int64_t* new_chrom_offset, int32_t n_chroms,
                              int32_t factor, int32_t n_bins, int32_t block_shift, int32_t* blocks_out,
                              void* stream);
/* One round of the merge form of K5.  `in` = pixels {column, count} of a bin1-major copy (round 0:
 * fine column ids, remapped on the fly (remap != NULL); later rounds: coarse column ids, remap == NULL); the
 * fine rows group_first[g] .. group_first[g+1] pool into coarse row g.  In round r a list is the
 * run of 2^r consecutive fine rows of one coarse row; lists 2q and 2q+1 (q < pairs_per_group) are
 * merged by coarse column (stable) into the same element range of `out`.  After
 * ceil(log2(max rows per coarse row)) rounds every coarse row is sorted by coarse column with
 * equal columns adjacent (cb200_runs_count / cb200_runs_write then sum them).  key_out (last
 * round, else NULL): coarse row of every element.  claim: one int64 of scratch. */
int cb200_coarsen_merge_round(const cb200_px* in, cb200_px* out, const int64_t* indptr,
                              const int32_t* group_first, int32_t n_groups, int32_t round,
                              int32_t pairs_per_group, const cb200_remap* remap, int32_t* key_out,
                              int64_t* claim, void* stream);
/* The LAST merge round with the reduce-by-key of groupby(...).sum() (src/cooler/_reduce.py:616-620) fused
 * in: the two remaining lists of every coarse row are merged and adjacent equal coarse columns summed
 * (int64; *overflow = 1 if a sum leaves int32).  The reduced row g lands at out + indptr[group_first[g]]
 * (its input offset) and its length in row_count[g]; an exclusive scan of row_count is the new row
 * pointer and cb200_coarsen_compact moves the rows there (key_out, if not NULL: coarse row per element). */
int cb200_coarsen_merge_reduce(const cb200_px* in, cb200_px* out, const int64_t* indptr,
                               const int32_t* group_first, int32_t n_groups, int32_t round,
                               const cb200_remap* remap, int32_t* row_count, int32_t* overflow,
                               int64_t* claim, void* stream);
int cb200_coarsen_compact(const cb200_px* in, const int64_t* indptr, const int32_t* group_first,
                          const int64_t* new_indptr, int32_t n_groups, cb200_px* out, int32_t* key_out,
                          void* stream);
/* Reduce-by-key over a stream sorted so that equal (key, val.other) pairs are
 * adjacent.  Pass 1 counts run heads per tile of cb200_warp_tile() elements. */
int cb200_runs_count(const int32_t* keys, const cb200_px* vals, int64_t n,
                     int32_t* tile_count, void* stream);
/* Pass 2 writes one element per run at tile_off[t] + rank: key, {other, sum}
 * (swap != 0: key_out = other, val_out = {key, sum} -- the input of the second
 * sort).  Sums are accumulated in int64; *overflow is set to 1 if any sum does
 * not fit int32 (val_out.count then holds the low 32 bits; sum64_out, if not
 * NULL, the exact sums). */
int cb200_runs_write(const int32_t* keys, const cb200_px* vals, int64_t n,
                     const int64_t* tile_off, int32_t swap,
                     int32_t* key_out, cb200_px* val_out, int64_t* sum64_out,
                     int32_t* overflow, void* stream);
/* Generic form for any numeric value column and aggregator (the `agg` argument of
 * coarsen_cooler / CoolerCoarsener, src/cooler/_reduce.py:616-620): vals = {other, index into src},
 * stream sorted so that equal (key, other) are adjacent in pixel order.  src_dtype: 0 int32, 1 int64
 * (accumulated in int64 -> int_out), 2 float32, 3 float64 (accumulated in float64 -> float_out);
 * op: 0 sum, 1 min, 2 max, 3 first, 4 last, 5 mean (always -> float_out).  key_out / other_out
 * (NULL to skip) receive the run's (key, other).  tile_off as for cb200_runs_write. */
int cb200_runs_aggregate(const int32_t* keys, const cb200_px* vals, int64_t n, const int64_t* tile_off,
                         const void* src, int32_t src_dtype, int32_t op, int32_t* key_out,
                         int32_t* other_out, int64_t* int_out, double* float_out, void* stream);
/* (bin1', val={bin2', v}) sorted by (bin1', bin2') -> bin1_id/bin2_id int64 + count,
 * the ContactBinner chunk layout (src/cooler/create/_ingest.py:507-528). */
int cb200_unpack_pixels(const int32_t* bin1, const cb200_px* val, int64_t n,
                        int64_t* bin1_out, int64_t* bin2_out, int32_t* count_out, void* stream);

/* ---- reading .cool columns (replaces the H5Dread chunk pipeline of libhdf5 that the reference reaches
 *      through h5py in every worker, src/cooler/parallel.py:289-304 / core/_selectors.py) ------------- */
/* One stored chunk of a 1-D chunked dataset: `src_bytes` bytes at `src` (the mapped file); rows
 * [row0, row1) of the chunk go to destination rows dst_row ...; filter_mask = the chunk's HDF5 filter
 * mask (bit i set: filter i of the pipeline was not applied). */
typedef struct {
  const void* src; int64_t src_bytes; int64_t dst_row; int32_t row0; int32_t row1; uint32_t filter_mask;
} cb200_chunk;
/* Decodes a batch of chunks with n_threads host threads (0 = all cores) into dst (host memory, pinned or
 * not).  filters: bit 0 shuffle, bit 1 deflate, bit 2 fletcher32 -- present in that pipeline order, as
 * h5py writes them.  elem_bytes = stored element size; dst_elem_bytes <= elem_bytes: little-endian
 * integers are narrowed (rc -3 if a dropped byte is not zero).  rc -2: corrupt chunk.  Host-only: needs
 * no GPU. */
int cb200_decode_chunks(const cb200_chunk* chunks, int64_t n_chunks, int32_t chunk_rows,
                        int32_t elem_bytes, int32_t dst_elem_bytes, int32_t filters, void* dst,
                        int32_t n_threads);

/* ---- ingest binning (replaces create/_ingest.py:68-214 _sanitize_records for integer chromosome
 *      ids and fixed-size bins; the groupby(bin1, bin2).size() of aggregate_records, :452-504, is
 *      cb200_sort_pairs x2 + cb200_runs_count / cb200_runs_write as for coarsening) ------------- */
enum { CB200_TRIL_NONE = 0, CB200_TRIL_REFLECT = 1, CB200_TRIL_DROP = 2, CB200_TRIL_RAISE = 3 };
/* Record i -> key_out[i] = bin2, val_out[i] = {bin1, 1}; records with a chromosome id < 0, dropped
 * lower-triangle records and records failing a check get key = n_bins, val = {n_bins, 0} (they sort
 * behind every bin).  flags[0]: some anchor < 0; flags[1]: some anchor > chromosome length;
 * flags[2]: lower-triangle record seen with CB200_TRIL_RAISE; flags[3]: number of dropped records
 * (flags must be zero-initialised).  pos is 1-based when one_based != 0. */
int cb200_bin_pairs(const int32_t* chrom1, const int64_t* pos1, const int32_t* chrom2,
                    const int64_t* pos2, int64_t n, const int32_t* chrom_binoffset,
                    const int64_t* chromsizes, int64_t binsize, int32_t one_based,
                    int32_t tril_action, int32_t n_bins, int32_t* key_out, cb200_px* val_out,
                    int32_t* flags, void* stream);

/* ---- balanced matrix fetch (replaces cooler.api.matrix: the CSR range query of
 *      core/_rangequery.py:153-368 -- CSRReader, DirectRangeQuery2D, FillLowerRangeQuery2D --
 *      and the weight application of src/cooler/api.py:742-798) ----------------------------- */
/* Window [i0,i1) x [j0,j1) of the matrix stored in `csr` (the UNFILTERED bin1-major copy,
 * n_rows bins).  fill_lower != 0 (symmetric-upper storage): cells below the diagonal are
 * filled from their mirror images.  Query rows: (i1-i0) direct rows, then (j1-j0) mirrored
 * rows when fill_lower.
 * count: row_count[q] = pixels produced by query row q (scan it for cb200_fetch_coo).
 * dense: dense_out[(i1-i0)*(j1-j0)] int32, fully written (zeros where nothing is stored).
 * coo:   window-relative (row, col, value) triples of query row q at row_off[q] ... */
int cb200_fetch_count(const cb200_copy* csr, int32_t n_rows, int32_t i0, int32_t i1, int32_t j0,
                      int32_t j1, int32_t fill_lower, int32_t* row_count, void* stream);
int cb200_fetch_dense(const cb200_copy* csr, int32_t n_rows, int32_t i0, int32_t i1, int32_t j0,
                      int32_t j1, int32_t fill_lower, int32_t* dense_out, void* stream);
int cb200_fetch_coo(const cb200_copy* csr, int32_t n_rows, int32_t i0, int32_t i1, int32_t j0,
                    int32_t j1, int32_t fill_lower, const int64_t* row_off, int32_t* row_out,
                    int32_t* col_out, int32_t* val_out, void* stream);
/* out = dense * outer(w1, w2)  (w1 = weights[i0:i1], w2 = weights[j0:j1]; divisive != 0:
 * reciprocal weights), evaluated as x * (w1_i * w2_j) like `arr * np.outer(bias1, bias2)`. */
int cb200_fetch_balance_dense(const int32_t* dense, int64_t h, int64_t w, const double* w1,
                              const double* w2, int32_t divisive, double* out, void* stream);
/* out[k] = w1[row[k]] * w2[col[k]] * val[k]  (`bias1[mat.row] * bias2[mat.col] * mat.data`). */
int cb200_fetch_balance_coo(const int32_t* row, const int32_t* col, const int32_t* val, int64_t n,
                            const double* w1, const double* w2, int32_t divisive, double* out,
                            void* stream);
/* The same queries filled from ANOTHER value column of the pixel table (`field=` of Cooler.matrix,
 * src/cooler/api.py:326-329, 724-726: a floating-point count column, or any extra numeric column):
 * values[p] (float64) is the value of stored pixel p -- the bin1-major copy is in file order, so record
 * p of `csr` is pixel p.  Windows and COO values are float64. */
int cb200_fetch_dense_f64(const cb200_copy* csr, const double* values, int32_t n_rows, int32_t i0,
                          int32_t i1, int32_t j0, int32_t j1, int32_t fill_lower, double* dense_out,
                          void* stream);
int cb200_fetch_coo_f64(const cb200_copy* csr, const double* values, int32_t n_rows, int32_t i0,
                        int32_t i1, int32_t j0, int32_t j1, int32_t fill_lower, const int64_t* row_off,
                        int32_t* row_out, int32_t* col_out, double* val_out, void* stream);
int cb200_fetch_balance_dense_f64(const double* dense, int64_t h, int64_t w, const double* w1,
                                  const double* w2, int32_t divisive, double* out, void* stream);
int cb200_fetch_balance_coo_f64(const int32_t* row, const int32_t* col, const double* val, int64_t n,
                                const double* w1, const double* w2, int32_t divisive, double* out,
                                void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COOLER_B200_H */
