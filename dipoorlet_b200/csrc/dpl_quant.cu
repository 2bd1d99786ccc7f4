// Elementwise / reduction kernels around the quantised graph (sm_100a, HBM bound):
//   K5  dpl_fakequant_f32        QuantizeLinear o DequantizeLinear, optional QDrop
//   K7a dpl_channel_sumdiff_f32  bias-correction reduction
//   K7b dpl_cosine3_f32          the three sums of cos_similarity
//   K6* dpl_adaround_*           soft rounding, hard rounding, fused grad + Adam step
// Reference semantics: dipoorlet/quantize.py:197-239 (ONNX opset-13 Q/DQ executed by
// onnxruntime), weight_transform/ada_quant_layer.py:28-50,96-130, bias_correction.py:10-13,
// utils.py:273-278, adaround.py:119-144 (torch.optim.Adam defaults).

#include <math.h>

#include "dpl_common.cuh"

namespace dpl {
namespace {

// ---- K5 ---------------------------------------------------------------------
// Counter-based uniform in [0, 1): splitmix64 of (seed, element index). Not torch's
// Philox stream — QDrop only needs an i.i.d. Bernoulli mask (brecq.py:169-170).
__device__ __forceinline__ float uniform01(uint64_t seed, uint64_t i) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (i + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ float fq1(float x, float s, float zp, float qlo, float qhi) {
  // round-half-even(x / s) + zp, saturate, dequantise (IEEE division, no reciprocal)
  float q = rintf(__fdiv_rn(x, s)) + zp;
  q = fminf(fmaxf(q, qlo), qhi);
  return __fmul_rn(q - zp, s);
}

__global__ void __launch_bounds__(256)
fakequant_kernel(const float* __restrict__ x, float* __restrict__ y, uint64_t n,
                 const float* __restrict__ scale, const int32_t* __restrict__ zero_point,
                 int n_channels, uint64_t inner, float qlo, float qhi, float drop_prob,
                 uint64_t seed) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15u) == 0 &&
                   (n_channels == 1 || (inner & 3u) == 0);
  const bool drop = drop_prob < 1.0f;
  uint64_t done = 0;
  if (vec) {
    const uint64_t n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y);
    for (uint64_t i = t; i < n4; i += stride) {
      const int c = n_channels == 1 ? 0 : (int)(((i << 2) / inner) % (uint64_t)n_channels);
      const float s = scale[c];
      const float zp = zero_point ? (float)zero_point[c] : 0.f;
      const float4 v = ldg_stream4(x4 + i);
      float4 o;
      o.x = fq1(v.x, s, zp, qlo, qhi);
      o.y = fq1(v.y, s, zp, qlo, qhi);
      o.z = fq1(v.z, s, zp, qlo, qhi);
      o.w = fq1(v.w, s, zp, qlo, qhi);
      if (drop) {
        const uint64_t e = i << 2;
        if (!(uniform01(seed, e) < drop_prob)) o.x = v.x;
        if (!(uniform01(seed, e + 1) < drop_prob)) o.y = v.y;
        if (!(uniform01(seed, e + 2) < drop_prob)) o.z = v.z;
        if (!(uniform01(seed, e + 3) < drop_prob)) o.w = v.w;
      }
      y4[i] = o;
    }
    done = n4 << 2;
  }
  for (uint64_t i = done + t; i < n; i += stride) {
    const int c = n_channels == 1 ? 0 : (int)((i / inner) % (uint64_t)n_channels);
    const float s = scale[c];
    const float zp = zero_point ? (float)zero_point[c] : 0.f;
    const float v = x[i];
    float o = fq1(v, s, zp, qlo, qhi);
    if (drop && !(uniform01(seed, i) < drop_prob)) o = v;
    y[i] = o;
  }
}

// ---- K7a --------------------------------------------------------------------
// One warp per (image, channel) row of `inner` elements; double atomics per channel.
__global__ void __launch_bounds__(256)
channel_sumdiff_kernel(const float* __restrict__ a, const float* __restrict__ b, uint64_t n_rows,
                       uint64_t channels, uint64_t inner, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t row = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); row < n_rows;
       row += warps) {
    const float* pa = a + row * inner;
    const float* pb = b + row * inner;
    double acc = 0.0;
    float blk = 0.f;
    int k = 0;
    for (uint64_t i = lane; i < inner; i += 32) {
      blk += pa[i] - pb[i];
      if (++k == 16) {
        acc += (double)blk;
        blk = 0.f;
        k = 0;
      }
    }
    acc += (double)blk;
    acc = warp_sum(acc);
    if (lane == 0) atomicAdd(out + (row % channels), acc);
  }
}

// ---- K7b --------------------------------------------------------------------
constexpr uint32_t kCosChunk = 16384;
__global__ void __launch_bounds__(256)
cosine3_kernel(const float* __restrict__ a, const float* __restrict__ b, uint64_t seg_len,
               uint64_t chunks_per_seg, double* __restrict__ out) {
  __shared__ double s_red[3][8];
  const uint64_t seg = blockIdx.x / chunks_per_seg;
  const uint64_t ch = blockIdx.x % chunks_per_seg;
  const uint64_t e0 = ch * kCosChunk, e1 = min(seg_len, e0 + kCosChunk);
  const float* pa = a + seg * seg_len;
  const float* pb = b + seg * seg_len;
  float ab = 0.f, aa = 0.f, bb = 0.f;
  for (uint64_t i = e0 + threadIdx.x; i < e1; i += 256) {  // <= 64 elements per thread
    const float u = pa[i], v = pb[i];
    ab = fmaf(u, v, ab);
    aa = fmaf(u, u, aa);
    bb = fmaf(v, v, bb);
  }
  double d0 = warp_sum((double)ab), d1 = warp_sum((double)aa), d2 = warp_sum((double)bb);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    s_red[0][warp] = d0;
    s_red[1][warp] = d1;
    s_red[2][warp] = d2;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_red[threadIdx.x][w];
    atomicAdd(out + seg * 3 + threadIdx.x, t);
  }
}

// ---- K6 elementwise -----------------------------------------------------------
constexpr float kZeta = 1.1f, kGamma = -0.1f;

__device__ __forceinline__ float sigmoidf_(float a) { return __fdiv_rn(1.0f, 1.0f + expf(-a)); }
// h(alpha) = clamp((zeta - gamma) * sigmoid(alpha) + gamma, 0, 1)
// torch evaluates the product and the sum separately (no FMA); at the clamp boundary
// (rest == 0, i.e. the largest weight of every channel, w/s = +-127) a fused evaluation can
// land on the other side of 0 and flip the clamp's gradient mask.
__device__ __forceinline__ float rect_sigmoid_raw(float sg) {
  return __fadd_rn(__fmul_rn(kZeta - kGamma, sg), kGamma);
}
__device__ __forceinline__ float rect_sigmoid(float a) {
  return fminf(fmaxf(rect_sigmoid_raw(sigmoidf_(a)), 0.f), 1.f);
}

__global__ void __launch_bounds__(256)
adaround_init_kernel(const float* __restrict__ w, const float* __restrict__ scale, int n_channels,
                     uint64_t inner, uint64_t n, float* __restrict__ alpha,
                     float* __restrict__ wfloor) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float s = scale[n_channels == 1 ? 0 : (i / inner) % (uint64_t)n_channels];
    const float q = __fdiv_rn(w[i], s);
    const float fl = floorf(q);
    const float rest = q - fl;
    // -log((zeta - gamma) / (rest - gamma) - 1)
    alpha[i] = -logf(__fdiv_rn(kZeta - kGamma, rest - kGamma) - 1.0f);
    wfloor[i] = fl;
  }
}

__global__ void __launch_bounds__(256)
adaround_weight_kernel(const float* __restrict__ wfloor, const float* __restrict__ alpha,
                       const float* __restrict__ scale, int n_channels, uint64_t inner, uint64_t n,
                       float qmin, float qmax, int soft, float* __restrict__ wq) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float s = scale[n_channels == 1 ? 0 : (i / inner) % (uint64_t)n_channels];
    const float a = alpha[i];
    const float r = soft ? rect_sigmoid(a) : (a >= 0.f ? 1.f : 0.f);
    float q = wfloor[i] + r;
    q = fminf(fmaxf(q, qmin), qmax);
    wq[i] = __fmul_rn(q, s);
  }
}

__global__ void __launch_bounds__(256)
adaround_step_kernel(const float* __restrict__ grad_w, const float* __restrict__ wfloor,
                     const float* __restrict__ scale, int n_channels, uint64_t inner, uint64_t n,
                     float qmin, float qmax, float beta, float reg_alpha, float lr, float b1,
                     float b2, float eps, float bc1, float bc2_sqrt, float grad_scale,
                     float* __restrict__ alpha, float* __restrict__ m, float* __restrict__ v,
                     double* __restrict__ reg_out, const float* __restrict__ sched) {
  if (sched) {   // per-iteration scalars from device memory (CUDA-graph replay)
    beta = sched[0];
    bc1 = sched[1];
    bc2_sqrt = sched[2];
  }
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  double reg_acc = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float s = scale[n_channels == 1 ? 0 : (i / inner) % (uint64_t)n_channels];
    const float a = alpha[i];
    const float sg = sigmoidf_(a);
    const float hraw = rect_sigmoid_raw(sg);
    const float h = fminf(fmaxf(hraw, 0.f), 1.f);
    // clamp(0,1) passes the gradient on the closed interval (torch.clamp backward)
    const float dh = (hraw >= 0.f && hraw <= 1.f) ? (kZeta - kGamma) * sg * (1.f - sg) : 0.f;
    // max(., qmin) / min(., qmax): pass where strictly inside, half on an exact tie
    const float q = wfloor[i] + h;
    float pass = 1.f;
    if (q < qmin || q > qmax) pass = 0.f;
    else if (q == qmin || q == qmax) pass = 0.5f;
    float g = grad_scale * grad_w[i] * s * pass * dh;
    // regulariser: reg_alpha * sum(1 - |2h - 1|^beta)
    if (beta > 0.f) {
      const float u = fabsf(h - 0.5f) * 2.f;
      const float sgn = (h > 0.5f) ? 1.f : ((h < 0.5f) ? -1.f : 0.f);
      const float pw1 = powf(u, beta - 1.f);
      g += -reg_alpha * beta * pw1 * 2.f * sgn * dh;
      reg_acc += (double)(reg_alpha * (1.f - pw1 * u));
    }
    // torch.optim.Adam (single-tensor path): lerp, addcmul, addcdiv
    float mi = m[i], vi = v[i];
    mi = mi + (g - mi) * (1.f - b1);
    vi = vi * b2 + (1.f - b2) * g * g;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    alpha[i] = a - (lr / bc1) * (mi / denom);
    m[i] = mi;
    v[i] = vi;
  }
  if (reg_out) {
    reg_acc = warp_sum(reg_acc);
    if ((threadIdx.x & 31) == 0 && reg_acc != 0.0) atomicAdd(reg_out, reg_acc);
  }
}

inline unsigned ew_grid(uint64_t n, int per_thread) {
  uint64_t blocks = (n + 256ull * per_thread - 1) / (256ull * per_thread);
  const uint64_t cap = (uint64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace
}  // namespace dpl

using namespace dpl;

extern "C" int dpl_fakequant_f32(const float* d_x, float* d_y, uint64_t n, const float* d_scale,
                                 const int32_t* d_zero_point, int n_channels, uint64_t inner,
                                 int qlo, int qhi, float drop_prob, uint64_t seed, void* stream) {
  DPL_REQUIRE(d_x && d_y && d_scale, "null pointer");
  DPL_REQUIRE(n_channels >= 1 && inner >= 1, "bad channel layout");
  if (n == 0) return 0;
  fakequant_kernel<<<ew_grid(n, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_x, d_y, n, d_scale, d_zero_point, n_channels, inner, (float)qlo, (float)qhi, drop_prob,
      seed);
  DPL_LAUNCH_CHECK("fakequant_kernel");
  return 0;
}

extern "C" int dpl_channel_sumdiff_f32(const float* d_a, const float* d_b, uint64_t n_img,
                                       uint64_t channels, uint64_t inner, double* d_sum,
                                       void* stream) {
  DPL_REQUIRE(d_a && d_b && d_sum, "null pointer");
  DPL_REQUIRE(channels >= 1 && inner >= 1, "bad layout");
  const uint64_t rows = n_img * channels;
  if (rows == 0) return 0;
  uint64_t blocks = (rows + 7) / 8;
  const uint64_t cap = (uint64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  channel_sumdiff_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_a, d_b, rows, channels, inner, d_sum);
  DPL_LAUNCH_CHECK("channel_sumdiff_kernel");
  return 0;
}

extern "C" int dpl_cosine3_f32(const float* d_a, const float* d_b, uint64_t n_seg,
                               uint64_t seg_len, double* d_out, void* stream) {
  DPL_REQUIRE(d_a && d_b && d_out, "null pointer");
  if (n_seg == 0 || seg_len == 0) return 0;
  const uint64_t cps = (seg_len + kCosChunk - 1) / kCosChunk;
  DPL_REQUIRE(n_seg * cps < (1ull << 31), "too many chunks");
  cosine3_kernel<<<(unsigned)(n_seg * cps), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_a, d_b, seg_len, cps, d_out);
  DPL_LAUNCH_CHECK("cosine3_kernel");
  return 0;
}

extern "C" int dpl_adaround_init_f32(const float* d_w, const float* d_scale, int n_channels,
                                     uint64_t inner, float* d_alpha, float* d_wfloor,
                                     void* stream) {
  DPL_REQUIRE(d_w && d_scale && d_alpha && d_wfloor, "null pointer");
  DPL_REQUIRE(n_channels >= 1 && inner >= 1, "bad channel layout");
  const uint64_t n = (uint64_t)n_channels * inner;
  adaround_init_kernel<<<ew_grid(n, 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_w, d_scale, n_channels, inner, n, d_alpha, d_wfloor);
  DPL_LAUNCH_CHECK("adaround_init_kernel");
  return 0;
}

extern "C" int dpl_adaround_weight_f32(const float* d_wfloor, const float* d_alpha,
                                       const float* d_scale, int n_channels, uint64_t inner,
                                       float qmin, float qmax, int soft, float* d_wq,
                                       void* stream) {
  DPL_REQUIRE(d_wfloor && d_alpha && d_scale && d_wq, "null pointer");
  DPL_REQUIRE(n_channels >= 1 && inner >= 1, "bad channel layout");
  const uint64_t n = (uint64_t)n_channels * inner;
  adaround_weight_kernel<<<ew_grid(n, 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_wfloor, d_alpha, d_scale, n_channels, inner, n, qmin, qmax, soft, d_wq);
  DPL_LAUNCH_CHECK("adaround_weight_kernel");
  return 0;
}

extern "C" int dpl_adaround_step_f32(const float* d_grad_w, const float* d_wfloor,
                                     const float* d_scale, int n_channels, uint64_t inner,
                                     float qmin, float qmax, float beta, float reg_alpha, float lr,
                                     float b1, float b2, float eps, int step, float grad_scale,
                                     float* d_alpha, float* d_m, float* d_v, double* d_reg,
                                     const float* d_sched, void* stream) {
  DPL_REQUIRE(d_grad_w && d_wfloor && d_scale && d_alpha && d_m && d_v, "null pointer");
  DPL_REQUIRE(n_channels >= 1 && inner >= 1 && step >= 1, "bad arguments");
  const uint64_t n = (uint64_t)n_channels * inner;
  // bias corrections in double like torch (Python floats), then float for the kernel
  const double bc1 = 1.0 - pow((double)b1, (double)step);
  const double bc2 = 1.0 - pow((double)b2, (double)step);
  adaround_step_kernel<<<ew_grid(n, 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_grad_w, d_wfloor, d_scale, n_channels, inner, n, qmin, qmax, beta, reg_alpha, lr, b1, b2,
      eps, (float)bc1, (float)sqrt(bc2), grad_scale, d_alpha, d_m, d_v, d_reg, d_sched);
  DPL_LAUNCH_CHECK("adaround_step_kernel");
  return 0;
}
