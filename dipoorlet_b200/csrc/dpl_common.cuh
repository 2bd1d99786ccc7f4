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

// Streaming 128-bit store: written once, read by a later kernel through L2.
__device__ __forceinline__ void stg_stream4(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// Counter-based uniform in [0, 1) keyed by (seed, element index): a 32-bit avalanche hash (two
// multiply-xorshift rounds) of the PAIR index i / 2 mixed with both halves of the seed; element i takes
// the low or the high 16 bits. Not torch's Philox stream - QDrop only needs an i.i.d. Bernoulli mask
// (brecq.py:169-170, ada_quant_layer.py:28-36) that the backward pass can regenerate from the same
// (seed, index); 16 bits resolve the drop probability to 1.5e-5 (0.5 is exact). One hash serves two
// elements: ~5 integer instructions per element (the 64-bit splitmix it replaces cost ~25).
__device__ __forceinline__ uint32_t hash_pair(uint64_t seed, uint64_t j) {
  uint32_t x = (uint32_t)j ^ (uint32_t)seed;
  x += ((uint32_t)(j >> 32) ^ (uint32_t)(seed >> 32)) * 0x9E3779B9u;
  x ^= x >> 16;
  x *= 0x7FEB352Du;
  x ^= x >> 15;
  x *= 0x846CA68Bu;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ float hash_u01(uint64_t seed, uint64_t i) {
  const uint32_t h = hash_pair(seed, i >> 1);
  return (float)((i & 1) ? (h >> 16) : (h & 0xFFFFu)) * (1.0f / 65536.0f);
}
// the four uniforms of elements e .. e + 3 (e a multiple of 4): two hashes
__device__ __forceinline__ float4 hash_u01x4(uint64_t seed, uint64_t e) {
  const uint32_t h0 = hash_pair(seed, e >> 1), h1 = hash_pair(seed, (e >> 1) + 1);
  const float c = 1.0f / 65536.0f;
  return make_float4((float)(h0 & 0xFFFFu) * c, (float)(h0 >> 16) * c, (float)(h1 & 0xFFFFu) * c,
                     (float)(h1 >> 16) * c);
}

// rint(x / s) with the IEEE-correct quotient, without paying for a full division per element.
// r = RN(1 / s) (one __frcp_rn per tensor / channel row); q1 = x r refined by one FMA step is within
// an ulp of x / s, so rint(q1) == rint(x / s) unless q1 sits within a few ulps of a rounding tie -
// only then (about 1 element in 10^5) the exact __fdiv_rn is evaluated. Non-finite products keep
// the plain quotient's behaviour (inf -> saturates in the caller's clamp, NaN propagates).
__device__ __forceinline__ float rint_div(float x, float s, float r) {
  const float q0 = __fmul_rn(x, r);
  const float q1 = __fmaf_rn(__fmaf_rn(-q0, s, x), r, q0);
  float t = rintf(q1);
  // distance of q1 to the nearest tie against ~8 ulp of q1; written so that a NaN (inf * 0 in the
  // correction step of a non-finite x) also takes the exact path
  const float d = 0.5f - fabsf(q1 - t);
  if (!(d > __fmaf_rn(fabsf(q1), 1e-6f, 1e-6f))) t = rintf(__fdiv_rn(x, s));
  return t;
}
// Four quotients at once: the fast evaluation is branch free, ONE (rarely taken) branch covers the
// exact re-evaluation of the whole vector, so the four chains interleave.
__device__ __forceinline__ float4 rint_div4(float4 x, float s, float r, bool fast) {
  float4 t;
  bool bad = !fast;
#define DPL_RD1(c)                                                    \
  {                                                                   \
    const float q0 = __fmul_rn(x.c, r);                               \
    const float q1 = __fmaf_rn(__fmaf_rn(-q0, s, x.c), r, q0);        \
    t.c = rintf(q1);                                                  \
    const float d = 0.5f - fabsf(q1 - t.c);                           \
    bad |= !(d > __fmaf_rn(fabsf(q1), 1e-6f, 1e-6f));                 \
  }
  DPL_RD1(x) DPL_RD1(y) DPL_RD1(z) DPL_RD1(w)
#undef DPL_RD1
  if (bad) {
    t.x = rintf(__fdiv_rn(x.x, s));
    t.y = rintf(__fdiv_rn(x.y, s));
    t.z = rintf(__fdiv_rn(x.z, s));
    t.w = rintf(__fdiv_rn(x.w, s));
  }
  return t;
}

// True when rint_div may be used for this scale (normal, finite reciprocal); else use __fdiv_rn.
__device__ __forceinline__ bool rint_div_ok(float s, float r) {
  return fabsf(s) >= 1.17549435e-38f && fabsf(s) < 1.0e37f && fabsf(r) >= 1.17549435e-38f && fabsf(r) < 1.0e37f;
}

// Grid of the streaming elementwise kernels: 8 CTAs of 256 threads per SM, fewer for small inputs.
inline unsigned stream_grid(uint64_t work_items) {
  const uint64_t cap = (uint64_t)sm_count() * 8;
  uint64_t g = (work_items + 255) / 256;
  if (g < 1) g = 1;
  return (unsigned)(g < cap ? g : cap);
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
