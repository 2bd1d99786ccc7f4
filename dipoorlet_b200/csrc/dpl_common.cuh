// Shared device/host helpers for libdpl_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "dpl_b200.h"

namespace dpl {

// ---- host side -------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_status(cudaError_t e, const char* what);
int sm_count();  // SMs of the current device (148 on B200), cached per device

#define DPL_REQUIRE(cond, msg)                 \
  do {                                         \
    if (!(cond)) {                             \
      ::dpl::set_error("%s: %s", __func__, msg); \
      return DPL_E_BADARG;                     \
    }                                          \
  } while (0)

#define DPL_LAUNCH_CHECK(what)                                   \
  do {                                                           \
    int _st = ::dpl::cuda_status(cudaGetLastError(), what);      \
    if (_st) return _st;                                         \
  } while (0)

// ---- device side -----------------------------------------------------------
// Streaming 128-bit load: read-only path, do not allocate in L1 (each byte of a
// blob is read exactly once per pass).
__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_stream1(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Float atomic min/max through the integer ordering of IEEE-754 bit patterns
// (inputs are NaN-free by contract).
__device__ __forceinline__ void atomic_min_f32(float* addr, float v) {
  if (v >= 0.f)
    atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f32(float* addr, float v) {
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// Last blob whose `begin` field (selected by FIELD: 4 = seg_tile_begin,
// 5 = flat_tile_begin, as uint64 index into dpl_blob) is <= tile. Blobs with no
// tiles share their begin with the next blob, so "last" skips them.
template <int FIELD>
__device__ __forceinline__ int find_blob(const dpl_blob* blobs, int n_blobs, uint64_t tile) {
  int lo = 0, hi = n_blobs - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    uint64_t begin = reinterpret_cast<const uint64_t*>(blobs + mid)[FIELD];
    if (begin <= tile)
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}

}  // namespace dpl
