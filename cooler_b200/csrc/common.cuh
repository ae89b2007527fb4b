// Shared device/host helpers for the cooler_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cb200 {

// Warp-tile: the unit of work of every row-walking kernel.  A warp owns
// WARP_TILE consecutive stored pixels of one copy of the pixel table, whatever
// rows they belong to (nnz-balanced split, deterministic summation tree).
constexpr int WARP_TILE = 2048;          // pixels per warp tile (16 KB of int2)
constexpr int CHUNK = 64;                // pixels per warp load (32 lanes x int4)
constexpr int WARPS_PER_BLOCK = 8;
constexpr int BLOCK_THREADS = WARPS_PER_BLOCK * 32;

void set_error(cudaError_t e, const char* where);
void set_error_msg(const char* msg);

#define CB_CHECK(expr)                                   \
  do {                                                   \
    cudaError_t _e = (expr);                             \
    if (_e != cudaSuccess) {                             \
      cb200::set_error(_e, #expr);                       \
      return (int)_e;                                    \
    }                                                    \
  } while (0)

#define CB_LAUNCH_CHECK(name)                            \
  do {                                                   \
    cudaError_t _e = cudaGetLastError();                 \
    if (_e != cudaSuccess) {                             \
      cb200::set_error(_e, name);                        \
      return (int)_e;                                    \
    }                                                    \
  } while (0)

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
// Streaming 128-bit load that does not allocate in L1 (the pixel stream is read
// once per pass; L1 is kept for the gathered bias vector).
__device__ __forceinline__ int4 ld_stream_int4(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ int2 ld_stream_int2(const int2* p) {
  int2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];"
               : "=r"(r.x), "=r"(r.y)
               : "l"(p));
  return r;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif

}  // namespace cb200
