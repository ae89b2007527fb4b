// Balanced matrix fetch: 2-D range query on the resident bin1-major copy, lower-triangle
// fill for symmetric-upper storage, and w_i * w_j * x.
//
// Reference: open2c/cooler src/cooler/api.py:665-800 (matrix), core/_rangequery.py:153-368
// (CSRReader, DirectRangeQuery2D, FillLowerRangeQuery2D): the window [i0,i1) x [j0,j1) of the
// full matrix holds the stored pixels (r, c) with r in [i0,i1), c in [j0,j1), and -- when the
// storage is symmetric-upper -- the mirror images (c, r) of the stored pixels with r in
// [j0,j1), c in [i0,i1), c > r.  Every window cell is produced by at most one stored pixel, so
// the scatter needs no atomics.  Balancing multiplies by outer(w[i0:i1], w[j0:j1]) exactly as
// `arr * np.outer(bias1, bias2)` does (one rounding for w_i*w_j, one for the product with x;
// empty cells become 0 * (w_i*w_j), i.e. NaN where a weight is NaN).
#include "common.cuh"
#include "../../include/cooler_b200.h"

namespace cb200 {

// first position in [a, b) whose other id is >= key
__device__ __forceinline__ int64_t lower_bound_other(const int2* __restrict__ px, int64_t a, int64_t b, int key) {
  while (a < b) {
    const int64_t m = a + ((b - a) >> 1);
    if (px[m].x < key) a = m + 1; else b = m;
  }
  return a;
}

// One warp per query row.  Rows [0, h) are the direct part (matrix row i0 + k, columns
// [j0, j1)); rows [h, h + wm) the mirrored part (stored row j0 + k, columns [i0, i1), c > r).
// mode 0: count pixels per query row; 1: scatter into the dense window; 2: write COO triples.
// V = int32_t: the value is the record's count.  V = double: the value of stored pixel p is values[p]
// (any other value column of the pixel table, `field=` of Cooler.matrix, api.py:326-329; the bin1-major
// copy is in file order, so record p IS pixel p).
template <int kMode, typename V>
__global__ void __launch_bounds__(BLOCK_THREADS)
fetch_kernel(const int2* __restrict__ px, const int64_t* __restrict__ indptr, const V* __restrict__ values,
             int i0, int i1, int j0, int j1, int mirror_rows, V* __restrict__ dense,
             int32_t* __restrict__ row_count, const int64_t* __restrict__ row_off,
             int32_t* __restrict__ coo_row, int32_t* __restrict__ coo_col, V* __restrict__ coo_val) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int h = i1 - i0, w = j1 - j0;
  if (q >= h + mirror_rows) return;
  const bool mirrored = q >= h;
  const int r = mirrored ? j0 + (q - h) : i0 + q;          // stored row
  int c_lo = mirrored ? i0 : j0, c_hi = mirrored ? i1 : j1;
  if (mirrored && c_lo <= r) c_lo = r + 1;                 // strictly below the diagonal of the window
  int64_t a = indptr[r], b = indptr[r + 1];
  if (c_lo >= c_hi) { a = b; }
  else {
    a = lower_bound_other(px, a, b, c_lo);
    b = lower_bound_other(px, a, b, c_hi);
  }
  if (kMode == 0) {
    if (lane == 0) row_count[q] = (int32_t)(b - a);
    return;
  }
  const int64_t o = (kMode == 2) ? row_off[q] : 0;
  for (int64_t p = a + lane; p < b; p += 32) {
    const int2 e = px[p];
    const int wi = mirrored ? e.x - i0 : r - i0;           // window coordinates
    const int wj = mirrored ? r - j0 : e.x - j0;
    const V v = values ? values[p] : (V)e.y;
    if (kMode == 1) dense[(int64_t)wi * w + wj] = v;
    else {
      coo_row[o + (p - a)] = wi;
      coo_col[o + (p - a)] = wj;
      coo_val[o + (p - a)] = v;
    }
  }
}

template <typename V>
__global__ void fetch_scale_dense_kernel(const V* __restrict__ dense, int64_t h, int64_t w,
                                         const double* __restrict__ w1, const double* __restrict__ w2,
                                         int divisive, double* __restrict__ out) {
  const int64_t n = h * w;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    const int64_t i = k / w, j = k - i * w;
    double a = w1[i], b = w2[j];
    if (divisive) { a = 1.0 / a; b = 1.0 / b; }
    out[k] = (double)dense[k] * (a * b);                   // arr * np.outer(bias1, bias2)
  }
}

template <typename V>
__global__ void fetch_scale_coo_kernel(const int32_t* __restrict__ row, const int32_t* __restrict__ col,
                                       const V* __restrict__ val, int64_t n,
                                       const double* __restrict__ w1, const double* __restrict__ w2,
                                       int divisive, double* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    double a = w1[row[k]], b = w2[col[k]];
    if (divisive) { a = 1.0 / a; b = 1.0 / b; }
    out[k] = a * b * (double)val[k];                       // bias1[row] * bias2[col] * data
  }
}

}  // namespace cb200

using namespace cb200;

static int fetch_args_ok(const cb200_copy* csr, int32_t n_rows, int32_t i0, int32_t i1, int32_t j0, int32_t j1) {
  if (!csr || i0 < 0 || j0 < 0 || i1 < i0 || j1 < j0 || i1 > n_rows || j1 > n_rows) {
    set_error_msg("matrix_fetch: bad window");
    return 0;
  }
  return 1;
}

template <typename V>
static int fetch_dense_impl(const cb200_copy* csr, const V* values, int32_t n_rows, int32_t i0, int32_t i1,
                            int32_t j0, int32_t j1, int32_t fill_lower, V* dense_out, void* stream) {
  if (!fetch_args_ok(csr, n_rows, i0, i1, j0, j1)) return -1;
  const int64_t cells = (int64_t)(i1 - i0) * (j1 - j0);
  if (cells <= 0) return 0;
  CB_CHECK(cudaMemsetAsync(dense_out, 0, (size_t)cells * sizeof(V), (cudaStream_t)stream));   // 0.0 is all-zero bits
  const int q = (i1 - i0) + (fill_lower ? (j1 - j0) : 0);
  fetch_kernel<1, V><<<(unsigned)ceil_div64(q, WARPS_PER_BLOCK), BLOCK_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const int2*>(csr->px), csr->indptr, values, i0, i1, j0, j1, fill_lower ? (j1 - j0) : 0,
      dense_out, nullptr, nullptr, nullptr, nullptr, nullptr);
  CB_LAUNCH_CHECK("fetch_dense");
  return 0;
}

template <typename V>
static int fetch_coo_impl(const cb200_copy* csr, const V* values, int32_t n_rows, int32_t i0, int32_t i1,
                          int32_t j0, int32_t j1, int32_t fill_lower, const int64_t* row_off, int32_t* row_out,
                          int32_t* col_out, V* val_out, void* stream) {
  if (!fetch_args_ok(csr, n_rows, i0, i1, j0, j1)) return -1;
  const int q = (i1 - i0) + (fill_lower ? (j1 - j0) : 0);
  if (q <= 0) return 0;
  fetch_kernel<2, V><<<(unsigned)ceil_div64(q, WARPS_PER_BLOCK), BLOCK_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const int2*>(csr->px), csr->indptr, values, i0, i1, j0, j1, fill_lower ? (j1 - j0) : 0,
      nullptr, nullptr, row_off, row_out, col_out, val_out);
  CB_LAUNCH_CHECK("fetch_coo");
  return 0;
}

template <typename V>
static int fetch_balance_dense_impl(const V* dense, int64_t h, int64_t w, const double* w1, const double* w2,
                                    int32_t divisive, double* out, void* stream) {
  const int64_t n = h * w;
  if (n <= 0) return 0;
  int64_t blocks = ceil_div64(n, 256 * 4);
  if (blocks > 148 * 32) blocks = 148 * 32;
  fetch_scale_dense_kernel<V><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dense, h, w, w1, w2, divisive, out);
  CB_LAUNCH_CHECK("fetch_balance_dense");
  return 0;
}

template <typename V>
static int fetch_balance_coo_impl(const int32_t* row, const int32_t* col, const V* val, int64_t n,
                                  const double* w1, const double* w2, int32_t divisive, double* out,
                                  void* stream) {
  if (n <= 0) return 0;
  int64_t blocks = ceil_div64(n, 256 * 4);
  if (blocks > 148 * 32) blocks = 148 * 32;
  fetch_scale_coo_kernel<V><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(row, col, val, n, w1, w2, divisive, out);
  CB_LAUNCH_CHECK("fetch_balance_coo");
  return 0;
}

extern "C" {

int cb200_fetch_count(const cb200_copy* csr, int32_t n_rows, int32_t i0, int32_t i1, int32_t j0,
                      int32_t j1, int32_t fill_lower, int32_t* row_count, void* stream) {
  if (!fetch_args_ok(csr, n_rows, i0, i1, j0, j1)) return -1;
  const int q = (i1 - i0) + (fill_lower ? (j1 - j0) : 0);
  if (q <= 0) return 0;
  fetch_kernel<0, int32_t><<<(unsigned)ceil_div64(q, WARPS_PER_BLOCK), BLOCK_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const int2*>(csr->px), csr->indptr, nullptr, i0, i1, j0, j1, fill_lower ? (j1 - j0) : 0,
      nullptr, row_count, nullptr, nullptr, nullptr, nullptr);
  CB_LAUNCH_CHECK("fetch_count");
  return 0;
}

int cb200_fetch_dense(const cb200_copy* csr, int32_t n_rows, int32_t i0, int32_t i1, int32_t j0,
                      int32_t j1, int32_t fill_lower, int32_t* dense_out, void* stream) {
  return fetch_dense_impl<int32_t>(csr, nullptr, n_rows, i0, i1, j0, j1, fill_lower, dense_out, stream);
}

int cb200_fetch_coo(const cb200_copy* csr, int32_t n_rows, int32_t i0, int32_t i1, int32_t j0,
                    int32_t j1, int32_t fill_lower, const int64_t* row_off, int32_t* row_out,
                    int32_t* col_out, int32_t* val_out, void* stream) {
  return fetch_coo_impl<int32_t>(csr, nullptr, n_rows, i0, i1, j0, j1, fill_lower, row_off, row_out, col_out,
                                 val_out, stream);
}

int cb200_fetch_balance_dense(const int32_t* dense, int64_t h, int64_t w, const double* w1,
                              const double* w2, int32_t divisive, double* out, void* stream) {
  return fetch_balance_dense_impl<int32_t>(dense, h, w, w1, w2, divisive, out, stream);
}

int cb200_fetch_balance_coo(const int32_t* row, const int32_t* col, const int32_t* val, int64_t n,
                            const double* w1, const double* w2, int32_t divisive, double* out,
                            void* stream) {
  return fetch_balance_coo_impl<int32_t>(row, col, val, n, w1, w2, divisive, out, stream);
}

// the same queries on another value column of the pixel table: values[p] (float64) of stored pixel p
int cb200_fetch_dense_f64(const cb200_copy* csr, const double* values, int32_t n_rows, int32_t i0, int32_t i1,
                          int32_t j0, int32_t j1, int32_t fill_lower, double* dense_out, void* stream) {
  if (!values) { set_error_msg("matrix_fetch: values is NULL"); return -1; }
  return fetch_dense_impl<double>(csr, values, n_rows, i0, i1, j0, j1, fill_lower, dense_out, stream);
}

int cb200_fetch_coo_f64(const cb200_copy* csr, const double* values, int32_t n_rows, int32_t i0, int32_t i1,
                        int32_t j0, int32_t j1, int32_t fill_lower, const int64_t* row_off, int32_t* row_out,
                        int32_t* col_out, double* val_out, void* stream) {
  if (!values) { set_error_msg("matrix_fetch: values is NULL"); return -1; }
  return fetch_coo_impl<double>(csr, values, n_rows, i0, i1, j0, j1, fill_lower, row_off, row_out, col_out,
                                val_out, stream);
}

int cb200_fetch_balance_dense_f64(const double* dense, int64_t h, int64_t w, const double* w1,
                                  const double* w2, int32_t divisive, double* out, void* stream) {
  return fetch_balance_dense_impl<double>(dense, h, w, w1, w2, divisive, out, stream);
}

int cb200_fetch_balance_coo_f64(const int32_t* row, const int32_t* col, const double* val, int64_t n,
                                const double* w1, const double* w2, int32_t divisive, double* out,
                                void* stream) {
  return fetch_balance_coo_impl<double>(row, col, val, n, w1, w2, divisive, out, stream);
}

}  // extern "C"
