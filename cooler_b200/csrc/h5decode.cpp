// Host-side chunk decoder for .cool columns: the filter pipeline libhdf5 runs inside H5Dread
// (fletcher32 -> inflate -> un-shuffle) for 1-D chunked datasets as h5py / cooler write them
// (src/cooler/create/_create.py: shuffle + gzip), straight into a destination column -- which may be a
// pinned staging buffer -- with optional narrowing of little-endian integers (bin ids stored as int64,
// shipped to the GPU as int32: only the low byte planes of a shuffled chunk are assembled).
// A batch of chunks is decoded by a few host threads inside ONE call, so the Python caller holds no
// interpreter lock while the file is inflated.
#include <zlib.h>

#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/cooler_b200.h"
#include "common.cuh"

namespace {

enum { F_SHUFFLE = 1, F_DEFLATE = 2, F_FLETCHER = 4 };

// rows [r0, r1) of a shuffled chunk (byte plane k of element r at planes[k * n + r]) -> dst, `ws` low bytes
// per element; returns false if a dropped high byte is not zero.
bool unshuffle_rows(const uint8_t* planes, int64_t n, int es, int ws, int64_t r0, int64_t r1, uint8_t* dst) {
  if (ws == 4) {
    const uint8_t *p0 = planes, *p1 = planes + n, *p2 = planes + 2 * n, *p3 = planes + 3 * n;
    uint32_t* out = reinterpret_cast<uint32_t*>(dst);
    for (int64_t r = r0; r < r1; ++r)
      out[r - r0] = (uint32_t)p0[r] | ((uint32_t)p1[r] << 8) | ((uint32_t)p2[r] << 16) | ((uint32_t)p3[r] << 24);
  } else if (ws == 8) {
    uint64_t* out = reinterpret_cast<uint64_t*>(dst);
    for (int64_t r = r0; r < r1; ++r) {
      uint64_t v = 0;
      for (int k = 0; k < 8; ++k) v |= (uint64_t)planes[k * n + r] << (8 * k);
      out[r - r0] = v;
    }
  } else {
    for (int64_t r = r0; r < r1; ++r)
      for (int k = 0; k < ws; ++k) dst[(r - r0) * ws + k] = planes[k * n + r];
  }
  uint8_t high = 0;
  for (int k = ws; k < es; ++k) {
    const uint8_t* p = planes + k * n;
    for (int64_t r = r0; r < r1; ++r) high |= p[r];
  }
  return high == 0;
}

}  // namespace

extern "C" int cb200_decode_chunks(const cb200_chunk* chunks, int64_t n_chunks, int32_t chunk_rows,
                                   int32_t elem_bytes, int32_t dst_elem_bytes, int32_t filters, void* dst,
                                   int32_t n_threads) {
  if (n_chunks <= 0) return 0;
  if (chunk_rows <= 0 || elem_bytes <= 0 || dst_elem_bytes <= 0 || dst_elem_bytes > elem_bytes || dst == nullptr ||
      (filters & ~(F_SHUFFLE | F_DEFLATE | F_FLETCHER))) {
    cb200::set_error_msg("decode_chunks: bad arguments");
    return -1;
  }
  const int es = elem_bytes, ws = dst_elem_bytes;
  const int64_t full = (int64_t)chunk_rows * es;
  std::atomic<int64_t> next{0};
  std::atomic<int> status{0};                       // 1: corrupt chunk, 2: value does not fit the narrower type
  auto work = [&]() {
    std::vector<uint8_t> scratch((size_t)full + 16), tmp;
    while (status.load(std::memory_order_relaxed) == 0) {
      const int64_t i = next.fetch_add(1);
      if (i >= n_chunks) break;
      const cb200_chunk& c = chunks[i];
      const uint8_t* data = static_cast<const uint8_t*>(c.src);
      int64_t len = c.src_bytes;
      // filter i of the pipeline (in the order written: shuffle, deflate, fletcher32) is skipped for this
      // chunk when bit i of its mask is set
      int idx = 0;
      const int i_shuffle = (filters & F_SHUFFLE) ? idx++ : -1;
      const int i_deflate = (filters & F_DEFLATE) ? idx++ : -1;
      const int i_fletcher = (filters & F_FLETCHER) ? idx++ : -1;
      auto on = [&](int k) { return k >= 0 && !((c.filter_mask >> k) & 1u); };
      if (on(i_fletcher)) len -= 4;
      if (len < 0 || c.row0 < 0 || c.row1 > chunk_rows || c.row0 > c.row1) { status = 1; break; }
      if (on(i_deflate)) {
        uLongf out_len = (uLongf)full;
        if (uncompress(scratch.data(), &out_len, data, (uLong)len) != Z_OK) { status = 1; break; }
        data = scratch.data();
        len = (int64_t)out_len;
      }
      const int64_t n = len / es;                   // elements stored in this chunk
      if (c.row1 > n) { status = 1; break; }
      uint8_t* out = static_cast<uint8_t*>(dst) + c.dst_row * ws;
      if (on(i_shuffle) && es > 1) {
        if (!unshuffle_rows(data, n, es, ws, c.row0, c.row1, out)) { status = 2; break; }
      } else if (ws == es) {
        std::memcpy(out, data + (int64_t)c.row0 * es, (size_t)(c.row1 - c.row0) * es);
      } else {                                      // little-endian narrowing of plain elements
        uint8_t high = 0;
        for (int64_t r = c.row0; r < c.row1; ++r) {
          const uint8_t* e = data + r * es;
          for (int k = 0; k < ws; ++k) out[(r - c.row0) * ws + k] = e[k];
          for (int k = ws; k < es; ++k) high |= e[k];
        }
        if (high) { status = 2; break; }
      }
    }
  };
  int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > 64) nt = 64;
  if ((int64_t)nt > n_chunks) nt = (int)n_chunks;
  if (nt == 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    pool.reserve(nt);
    for (int t = 0; t < nt; ++t) pool.emplace_back(work);
    for (auto& t : pool) t.join();
  }
  if (status == 1) {
    cb200::set_error_msg("decode_chunks: corrupt chunk (inflate failed or sizes inconsistent)");
    return -2;
  }
  if (status == 2) {
    cb200::set_error_msg("decode_chunks: a value does not fit the narrower destination type");
    return -3;
  }
  return 0;
}
