// Non-GEMM operators of the calibration forward: Relu / Clip, Add (+ the Relu that follows it),
// MaxPool, GlobalAveragePool — the nodes onnxruntime executes one image at a time in the reference
// (dipoorlet/forward_net.py:200-216). Each is a single streaming pass over a whole batch of blobs,
// HBM bound: 16-byte loads with several in flight per thread, grid = SMs x 8 CTAs.
// Every node output is a calibration blob, so outputs are always materialised; the fused
// Add + Relu kernel writes BOTH blobs from one read of its operands.

#include <math.h>

#include "dpl_common.cuh"

namespace dpl {
namespace {

constexpr int kEltThreads = 256;
constexpr int kEltCtasPerSm = 8;

__device__ __forceinline__ float clip1(float x, float lo, float hi) {
  // torch.clamp / ONNX Clip semantics incl. NaN propagation (a NaN fails both comparisons)
  float v = x < lo ? lo : x;
  return v > hi ? hi : v;
}

// Fused range statistics: every kernel here can fold the values it WRITES into the running per-blob
// extrema (the blob-level result of K1, dpl_segstats_f32) - the producer already has the data in
// registers, so pass 1 of a minmax / hist calibration needs no second read of these blobs.
// NaN-free by contract (fminf / fmaxf skip NaN); -0 is canonicalised for the integer-ordered atomics.
struct RangeAcc {
  float lo, hi;
  __device__ __forceinline__ RangeAcc() : lo(INFINITY), hi(-INFINITY) {}
  __device__ __forceinline__ void add(float v) {
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  __device__ __forceinline__ void add4(float4 v) {
    lo = fminf(fminf(lo, v.x), fminf(fminf(v.y, v.z), v.w));
    hi = fmaxf(fmaxf(hi, v.x), fmaxf(fmaxf(v.y, v.z), v.w));
  }
};

// all threads of the CTA must call it (one barrier); s_lo / s_hi: shared float[32]
__device__ __forceinline__ void range_flush(const RangeAcc& r, float* bmin, float* bmax, float* s_lo, float* s_hi) {
  if (!bmin && !bmax) return;
  const float lo = warp_min(r.lo), hi = warp_max(r.hi);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) {
    s_lo[warp] = lo;
    s_hi[warp] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float l = s_lo[0], h = s_hi[0];
    for (int w = 1; w < nw; ++w) {
      l = fminf(l, s_lo[w]);
      h = fmaxf(h, s_hi[w]);
    }
    if (bmin && l <= h) atomic_min_f32(bmin, l + 0.f);
    if (bmax && l <= h) atomic_max_f32(bmax, h + 0.f);
  }
}

__device__ __forceinline__ void st_stream4(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// Vector loops walk CTA-contiguous tiles: a CTA reads kEltThreads x R consecutive float4 (16 KB for
// R = 4) per step, so its loads fall into a few DRAM pages instead of R pages a grid-stride apart
// (grid-stride float4 loops measured 75 - 84 % of the HBM peak, tiles: see profiles/).
__global__ void __launch_bounds__(kEltThreads, kEltCtasPerSm)
clip_kernel(const float* __restrict__ x, float* __restrict__ y, uint64_t n, float lo, float hi, int vec,
            float* __restrict__ bmin, float* __restrict__ bmax) {
  __shared__ float s_lo[32], s_hi[32];
  RangeAcc acc;
  uint64_t done = 0;
  if (vec) {
    constexpr int R = 4;
    const uint64_t n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y);
    const uint64_t tiles = (n4 + kEltThreads * R - 1) / (kEltThreads * R);
    for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const uint64_t i0 = tile * (kEltThreads * R) + threadIdx.x;
      float4 v[R];
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (i0 + r * kEltThreads < n4) v[r] = ldg_stream4(x4 + i0 + r * kEltThreads);
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (i0 + r * kEltThreads < n4) {
          const float4 o = make_float4(clip1(v[r].x, lo, hi), clip1(v[r].y, lo, hi), clip1(v[r].z, lo, hi),
                                       clip1(v[r].w, lo, hi));
          st_stream4(y4 + i0 + r * kEltThreads, o);
          acc.add4(o);
        }
    }
    done = n4 << 2;
  }
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = done + tid; i < n; i += stride) {
    const float o = clip1(x[i], lo, hi);
    y[i] = o;
    acc.add(o);
  }
  range_flush(acc, bmin, bmax, s_lo, s_hi);
}

template <bool RELU>
__global__ void __launch_bounds__(kEltThreads, 4)   // 64 registers: four vectors in flight, no spills
add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y,
           float* __restrict__ yr, uint64_t n, int vec, float* __restrict__ bmin, float* __restrict__ bmax,
           float* __restrict__ rmin, float* __restrict__ rmax) {
  __shared__ float s_lo[32], s_hi[32];
  RangeAcc acc;   // of y; the Relu output's range follows from it: [max(lo, 0), max(hi, 0)]
  uint64_t done = 0;
  if (vec) {
    constexpr int R = 2;
    const uint64_t n4 = n >> 2;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    float4* y4 = reinterpret_cast<float4*>(y);
    float4* r4 = reinterpret_cast<float4*>(yr);
    const uint64_t tiles = (n4 + kEltThreads * R - 1) / (kEltThreads * R);
    for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const uint64_t i0 = tile * (kEltThreads * R) + threadIdx.x;
      float4 u[R], w[R];
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (i0 + r * kEltThreads < n4) {
          u[r] = ldg_stream4(a4 + i0 + r * kEltThreads);
          w[r] = ldg_stream4(b4 + i0 + r * kEltThreads);
        }
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (i0 + r * kEltThreads < n4) {
          const float4 s = make_float4(u[r].x + w[r].x, u[r].y + w[r].y, u[r].z + w[r].z, u[r].w + w[r].w);
          st_stream4(y4 + i0 + r * kEltThreads, s);
          acc.add4(s);
          if (RELU)
            st_stream4(r4 + i0 + r * kEltThreads,
                       make_float4(clip1(s.x, 0.f, INFINITY), clip1(s.y, 0.f, INFINITY), clip1(s.z, 0.f, INFINITY),
                                   clip1(s.w, 0.f, INFINITY)));
        }
    }
    done = n4 << 2;
  }
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = done + tid; i < n; i += stride) {
    const float s = a[i] + b[i];
    y[i] = s;
    acc.add(s);
    if (RELU) yr[i] = clip1(s, 0.f, INFINITY);
  }
  range_flush(acc, bmin, bmax, s_lo, s_hi);
  if (RELU && (rmin || rmax)) {
    RangeAcc racc;
    if (acc.lo <= acc.hi) {
      racc.lo = fmaxf(acc.lo, 0.f);
      racc.hi = fmaxf(acc.hi, 0.f);
    }
    __syncthreads();
    range_flush(racc, rmin, rmax, s_lo, s_hi);
  }
}

// One thread per output element, consecutive threads along Wo (coalesced stores; the k x k window
// reads overlap between neighbours and are served by L1). A CTA works inside one plane, so the index
// arithmetic is two 32-bit divisions per output; the window is a template parameter (0 = runtime size)
// and windows that lie inside the image skip the bounds checks. Padding never wins (-inf), as in ONNX.
template <int KH, int KW>
__global__ void __launch_bounds__(256)
maxpool2d_kernel(const float* __restrict__ x, float* __restrict__ y, uint32_t tiles_per_plane, int H, int W,
                 int kh_rt, int kw_rt, int sh, int sw, int pt, int pl, int Ho, int Wo, float* __restrict__ bmin,
                 float* __restrict__ bmax) {
  __shared__ float s_lo[32], s_hi[32];
  RangeAcc acc;
  const int kh = KH ? KH : kh_rt, kw = KW ? KW : kw_rt;
  const uint32_t plane = blockIdx.x / tiles_per_plane;
  const uint32_t tile = blockIdx.x - plane * tiles_per_plane;
  const uint32_t idx = tile * 256u + threadIdx.x;
  // no early exit: every thread of the CTA takes part in range_flush's shuffles and barrier
  if (idx < (uint32_t)(Ho * Wo)) {
    const int ho = (int)(idx / (uint32_t)Wo), wo = (int)(idx - (uint32_t)ho * (uint32_t)Wo);
    const float* xp = x + (uint64_t)plane * (uint32_t)(H * W);
    const int h0 = ho * sh - pt, w0 = wo * sw - pl;
    float m = -INFINITY;
    bool nan = false;
    if (h0 >= 0 && w0 >= 0 && h0 + kh <= H && w0 + kw <= W) {
      const float* p = xp + h0 * W + w0;
#pragma unroll
      for (int a = 0; a < kh; ++a) {
#pragma unroll
        for (int c = 0; c < kw; ++c) {
          const float v = __ldg(p + a * W + c);
          nan |= (v != v);
          m = fmaxf(m, v);
        }
      }
    } else {
      for (int a = 0; a < kh; ++a) {
        const int h = h0 + a;
        if (h < 0 || h >= H) continue;
        const float* row = xp + h * W;
        for (int c = 0; c < kw; ++c) {
          const int w = w0 + c;
          if (w < 0 || w >= W) continue;
          const float v = __ldg(row + w);
          nan |= (v != v);
          m = fmaxf(m, v);
        }
      }
    }
    y[(uint64_t)plane * (uint32_t)(Ho * Wo) + idx] = nan ? NAN : m;
    acc.add(m);
  }
  range_flush(acc, bmin, bmax, s_lo, s_hi);
}

// One warp per (image, channel) plane: fp32 lane partials, shuffle tree, one division.
__global__ void __launch_bounds__(256)
global_avgpool_kernel(const float* __restrict__ x, float* __restrict__ y, uint64_t planes, uint64_t hw,
                      float* __restrict__ bmin, float* __restrict__ bmax) {
  __shared__ float s_lo[32], s_hi[32];
  RangeAcc racc;
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t p = warp; p < planes; p += nwarps) {
    const float* xp = x + p * hw;
    float acc = 0.f;
    for (uint64_t i = lane; i < hw; i += 32) acc += __ldg(xp + i);
    acc = warp_sum(acc);
    const float mean = acc / (float)hw;
    if (lane == 0) y[p] = mean;
    racc.add(mean);
  }
  range_flush(racc, bmin, bmax, s_lo, s_hi);
}

inline unsigned elt_grid(uint64_t work_items) {
  const uint64_t cap = (uint64_t)sm_count() * kEltCtasPerSm;
  uint64_t g = (work_items + kEltThreads - 1) / kEltThreads;
  if (g < 1) g = 1;
  return (unsigned)(g < cap ? g : cap);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace dpl

using namespace dpl;

extern "C" int dpl_clip_f32(const float* d_x, float* d_y, uint64_t n, float lo, float hi, float* d_blob_min,
                            float* d_blob_max, void* stream) {
  DPL_REQUIRE(d_x && d_y, "null pointer");
  if (n == 0) return 0;
  const int vec = aligned16(d_x) && aligned16(d_y);
  clip_kernel<<<elt_grid(vec ? (n + 3) / 4 : n), kEltThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      d_x, d_y, n, lo, hi, vec, d_blob_min, d_blob_max);
  DPL_LAUNCH_CHECK("clip_kernel");
  return 0;
}

extern "C" int dpl_add_f32(const float* d_a, const float* d_b, float* d_y, float* d_y_relu, uint64_t n,
                           float* d_blob_min, float* d_blob_max, float* d_relu_min, float* d_relu_max,
                           void* stream) {
  DPL_REQUIRE(d_a && d_b && d_y, "null pointer");
  if (n == 0) return 0;
  const int vec = aligned16(d_a) && aligned16(d_b) && aligned16(d_y) && (!d_y_relu || aligned16(d_y_relu));
  const unsigned grid = elt_grid(vec ? (n + 3) / 4 : n);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d_y_relu)
    add_kernel<true><<<grid, kEltThreads, 0, st>>>(d_a, d_b, d_y, d_y_relu, n, vec, d_blob_min, d_blob_max,
                                                   d_relu_min, d_relu_max);
  else
    add_kernel<false><<<grid, kEltThreads, 0, st>>>(d_a, d_b, d_y, nullptr, n, vec, d_blob_min, d_blob_max,
                                                    nullptr, nullptr);
  DPL_LAUNCH_CHECK("add_kernel");
  return 0;
}

// 3 x 3 / stride 2 / pad 1 (top, left) max pooling with W a multiple of 8 - ResNet's stem pool, the only pooling
// layer of the measured path. A thread produces FOUR consecutive outputs: per input row two 16-byte loads (columns
// 8q .. 8q+7) plus the one column to their left, i.e. every input byte is requested once per output row it feeds
// and the stores are 16 bytes; the one-output-per-thread kernel issued nine scalar loads per output and reached
// 38 % of the HBM peak. Grid-stride over (plane, output row, column quad).
__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const float* __restrict__ x, float* __restrict__ y, uint64_t total_quads, int H, int W, int Ho,
                    int Wo, float* __restrict__ bmin, float* __restrict__ bmax) {
  __shared__ float s_lo[32], s_hi[32];
  RangeAcc acc;
  const int qpr = Wo >> 2;                       // quads per output row
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total_quads; e += stride) {
    const int q = (int)(e % (uint64_t)qpr);
    const uint64_t row = e / (uint64_t)qpr;      // plane * Ho + ho
    const int ho = (int)(row % (uint64_t)Ho);
    const uint64_t plane = row / (uint64_t)Ho;
    const float* xp = x + plane * (uint64_t)H * (uint64_t)W;
    float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
    bool n0 = false, n1 = false, n2 = false, n3 = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int h = 2 * ho - 1 + a;
      if (h < 0 || h >= H) continue;
      const float* r = xp + (uint64_t)h * (uint64_t)W + 8 * q;
      const float4 u = ldg_stream4(reinterpret_cast<const float4*>(r));
      const float4 v = ldg_stream4(reinterpret_cast<const float4*>(r) + 1);
      const float l = q > 0 ? __ldg(r - 1) : -INFINITY;     // column 8q - 1 (the left padding for q = 0)
      // output j covers columns 8q + 2j - 1 .. 8q + 2j + 1
      m0 = fmaxf(m0, fmaxf(l, fmaxf(u.x, u.y)));
      m1 = fmaxf(m1, fmaxf(u.y, fmaxf(u.z, u.w)));
      m2 = fmaxf(m2, fmaxf(u.w, fmaxf(v.x, v.y)));
      m3 = fmaxf(m3, fmaxf(v.y, fmaxf(v.z, v.w)));
      n0 |= (l != l) | (u.x != u.x) | (u.y != u.y);
      n1 |= (u.y != u.y) | (u.z != u.z) | (u.w != u.w);
      n2 |= (u.w != u.w) | (v.x != v.x) | (v.y != v.y);
      n3 |= (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
    }
    const float4 o = make_float4(n0 ? NAN : m0, n1 ? NAN : m1, n2 ? NAN : m2, n3 ? NAN : m3);
    *reinterpret_cast<float4*>(y + row * (uint64_t)Wo + 4 * q) = o;
    acc.add(m0);
    acc.add(m1);
    acc.add(m2);
    acc.add(m3);
  }
  range_flush(acc, bmin, bmax, s_lo, s_hi);
}

extern "C" int dpl_maxpool2d_f32(const float* d_x, float* d_y, uint64_t planes, int H, int W, int kh,
                                 int kw, int sh, int sw, int pad_top, int pad_left, int Ho, int Wo,
                                 float* d_blob_min, float* d_blob_max, void* stream) {
  DPL_REQUIRE(d_x && d_y, "null pointer");
  DPL_REQUIRE(H > 0 && W > 0 && kh > 0 && kw > 0 && sh > 0 && sw > 0 && Ho > 0 && Wo > 0, "bad geometry");
  if (planes == 0) return 0;
  DPL_REQUIRE((long long)H * W < (1ll << 31) && (long long)Ho * Wo < (1ll << 31), "plane too large");
  if (kh == 3 && kw == 3 && sh == 2 && sw == 2 && pad_top == 1 && pad_left == 1 && (W & 7) == 0 && Wo == W / 2 &&
      Ho == (H + 1) / 2 && (reinterpret_cast<uintptr_t>(d_x) & 15u) == 0 && (reinterpret_cast<uintptr_t>(d_y) & 15u) == 0) {
    const uint64_t quads = planes * (uint64_t)Ho * (uint64_t)(Wo / 4);
    maxpool3x3s2_kernel<<<stream_grid(quads), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_x, d_y, quads, H, W, Ho, Wo,
                                                                                          d_blob_min, d_blob_max);
    DPL_LAUNCH_CHECK("maxpool3x3s2_kernel");
    return 0;
  }
  const uint64_t tiles_per_plane = ((uint64_t)Ho * Wo + 255) / 256;
  DPL_REQUIRE(planes * tiles_per_plane < (1ull << 31), "too many tiles");
  auto kern = (kh == 3 && kw == 3) ? maxpool2d_kernel<3, 3> : ((kh == 2 && kw == 2) ? maxpool2d_kernel<2, 2>
                                                                                      : maxpool2d_kernel<0, 0>);
  kern<<<(unsigned)(planes * tiles_per_plane), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_x, d_y, (uint32_t)tiles_per_plane, H, W, kh, kw, sh, sw, pad_top, pad_left, Ho, Wo, d_blob_min, d_blob_max);
  DPL_LAUNCH_CHECK("maxpool2d_kernel");
  return 0;
}

extern "C" int dpl_global_avgpool_f32(const float* d_x, float* d_y, uint64_t planes, uint64_t hw,
                                      float* d_blob_min, float* d_blob_max, void* stream) {
  DPL_REQUIRE(d_x && d_y, "null pointer");
  DPL_REQUIRE(hw > 0, "empty plane");
  if (planes == 0) return 0;
  uint64_t grid = (planes * 32 + 255) / 256;
  const uint64_t cap = (uint64_t)sm_count() * 8;
  if (grid > cap) grid = cap;
  global_avgpool_kernel<<<(unsigned)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_x, d_y, planes, hw,
                                                                                       d_blob_min, d_blob_max);
  DPL_LAUNCH_CHECK("global_avgpool_kernel");
  return 0;
}
