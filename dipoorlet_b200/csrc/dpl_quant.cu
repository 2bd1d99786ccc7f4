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
__device__ __forceinline__ float uniform01(uint64_t seed, uint64_t i) { return hash_u01(seed, i); }

__device__ __forceinline__ float fq1(float x, float s, float zp, float qlo, float qhi) {
  // round-half-even(x / s) + zp, saturate, dequantise (IEEE division, no reciprocal)
  float q = rintf(__fdiv_rn(x, s)) + zp;
  q = fminf(fmaxf(q, qlo), qhi);
  return __fmul_rn(q - zp, s);
}
// same result through rint_div (r = RN(1 / s), fast == rint_div_ok(s, r))
__device__ __forceinline__ float fq1r(float x, float s, float r, bool fast, float zp, float qlo, float qhi) {
  float q = (fast ? rint_div(x, s, r) : rintf(__fdiv_rn(x, s))) + zp;
  q = fminf(fmaxf(q, qlo), qhi);
  return __fmul_rn(q - zp, s);
}

// Channel of element e for a per-channel tensor: (e / inner) % n_channels, in 32-bit arithmetic
// when the tensor has fewer than 2^32 elements (a 64-bit divide costs ~10x a 32-bit one).
__device__ __forceinline__ int channel_of(uint64_t e, uint64_t inner, int n_channels, bool small) {
  if (small) return (int)(((uint32_t)e / (uint32_t)inner) % (uint32_t)n_channels);
  return (int)((e / inner) % (uint64_t)n_channels);
}

__device__ __forceinline__ float fq_tail(float t, float s, float zp, float qlo, float qhi) {
  float q = t + zp;
  q = fminf(fmaxf(q, qlo), qhi);
  return __fmul_rn(q - zp, s);
}

__device__ __forceinline__ float4 fq4(float4 v, float s, float r, bool fast, float zp, float qlo, float qhi,
                                      bool drop, float drop_prob, uint64_t seed, uint64_t e) {
  const float4 t = rint_div4(v, s, r, fast);
  float4 o = make_float4(fq_tail(t.x, s, zp, qlo, qhi), fq_tail(t.y, s, zp, qlo, qhi), fq_tail(t.z, s, zp, qlo, qhi),
                         fq_tail(t.w, s, zp, qlo, qhi));
  if (drop) {
    const float4 u = hash_u01x4(seed, e);
    o.x = (u.x < drop_prob) ? o.x : v.x;
    o.y = (u.y < drop_prob) ? o.y : v.y;
    o.z = (u.z < drop_prob) ? o.z : v.z;
    o.w = (u.w < drop_prob) ? o.w : v.w;
  }
  return o;
}

__global__ void __launch_bounds__(256, 4)
fakequant_kernel(const float* __restrict__ x, float* __restrict__ y, uint64_t n,
                 const float* __restrict__ scale, const int32_t* __restrict__ zero_point,
                 int n_channels, uint64_t inner, float qlo, float qhi, float drop_prob,
                 uint64_t seed) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15u) == 0 &&
                   (n_channels == 1 || (inner & 3u) == 0);
  const bool drop = drop_prob < 1.0f;
  const bool small = n < (1ull << 32);
  const bool per_tensor = n_channels == 1;
  // per-tensor: scale, its reciprocal and the zero point are loop invariants
  float s0 = scale[0], r0 = __frcp_rn(s0), z0 = zero_point ? (float)zero_point[0] : 0.f;
  bool f0 = rint_div_ok(s0, r0);
  uint64_t done = 0;
  if (vec) {
    // CTA-contiguous tiles of 256 x R float4 (see dpl_eltwise.cu), R loads in flight per thread
    constexpr int R = 4;
    const uint64_t n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y);
    const uint64_t tiles = (n4 + 256 * R - 1) / (256 * R);
    for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const uint64_t i0 = tile * (256 * R) + threadIdx.x;
      float4 v[R];
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (i0 + r * 256 < n4) v[r] = ldg_stream4(x4 + i0 + r * 256);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const uint64_t i = i0 + r * 256;
        if (i < n4) {
          if (!per_tensor) {
            const int c0 = channel_of(i << 2, inner, n_channels, small);
            s0 = scale[c0];
            z0 = zero_point ? (float)zero_point[c0] : 0.f;
            r0 = __frcp_rn(s0);
            f0 = rint_div_ok(s0, r0);
          }
          stg_stream4(y4 + i, fq4(v[r], s0, r0, f0, z0, qlo, qhi, drop, drop_prob, seed, i << 2));
        }
      }
    }
    done = n4 << 2;
  }
  for (uint64_t i = done + t; i < n; i += stride) {
    const int c = n_channels == 1 ? 0 : channel_of(i, inner, n_channels, small);
    const float s = scale[c];
    const float zp = zero_point ? (float)zero_point[c] : 0.f;
    const float v = x[i];
    float o = fq1(v, s, zp, qlo, qhi);
    if (drop && !(uniform01(seed, i) < drop_prob)) o = v;
    y[i] = o;
  }
}

// ---- K7a --------------------------------------------------------------------
// One warp per (image, channel) row of `inner` elements (16-byte loads, four rows of lanes in
// flight when the rows are aligned); double atomics per channel.
__global__ void __launch_bounds__(256)
channel_sumdiff_kernel(const float* __restrict__ a, const float* __restrict__ b, uint64_t n_rows,
                       uint64_t channels, uint64_t inner, double* __restrict__ out, int vec) {
  const int lane = threadIdx.x & 31;
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t row = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); row < n_rows;
       row += warps) {
    const float* pa = a + row * inner;
    const float* pb = b + row * inner;
    double acc = 0.0;
    uint64_t i0 = 0;
    if (vec) {
      const float4* a4 = reinterpret_cast<const float4*>(pa);
      const float4* b4 = reinterpret_cast<const float4*>(pb);
      const uint64_t n4 = inner >> 2;
      for (uint64_t i = lane; i < n4; i += 128) {
        float blk = 0.f;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const uint64_t j = i + r * 32;
          if (j < n4) {
            const float4 u = ldg_stream4(a4 + j), v = ldg_stream4(b4 + j);
            blk += (u.x - v.x) + (u.y - v.y) + (u.z - v.z) + (u.w - v.w);
          }
        }
        acc += (double)blk;
      }
      i0 = n4 << 2;
    }
    float blk = 0.f;
    int k = 0;
    for (uint64_t i = i0 + lane; i < inner; i += 32) {
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
               uint64_t chunks_per_seg, double* __restrict__ out, int vec) {
  __shared__ double s_red[3][8];
  const uint64_t seg = blockIdx.x / chunks_per_seg;
  const uint64_t ch = blockIdx.x % chunks_per_seg;
  const uint64_t e0 = ch * kCosChunk, e1 = min(seg_len, e0 + kCosChunk);
  const float* pa = a + seg * seg_len;
  const float* pb = b + seg * seg_len;
  float ab = 0.f, aa = 0.f, bb = 0.f;
  if (vec) {   // 16 float4 per thread and operand, all issued before use
    const float4* a4 = reinterpret_cast<const float4*>(pa + e0);
    const float4* b4 = reinterpret_cast<const float4*>(pb + e0);
    const uint32_t n4 = (uint32_t)((e1 - e0) >> 2);
#pragma unroll 4
    for (uint32_t i = threadIdx.x; i < n4; i += 256) {
      const float4 u = ldg_stream4(a4 + i), v = ldg_stream4(b4 + i);
      ab = fmaf(u.x, v.x, ab); aa = fmaf(u.x, u.x, aa); bb = fmaf(v.x, v.x, bb);
      ab = fmaf(u.y, v.y, ab); aa = fmaf(u.y, u.y, aa); bb = fmaf(v.y, v.y, bb);
      ab = fmaf(u.z, v.z, ab); aa = fmaf(u.z, u.z, aa); bb = fmaf(v.z, v.z, bb);
      ab = fmaf(u.w, v.w, ab); aa = fmaf(u.w, u.w, aa); bb = fmaf(v.w, v.w, bb);
    }
    for (uint64_t i = e0 + ((uint64_t)n4 << 2) + threadIdx.x; i < e1; i += 256) {
      const float u = pa[i], v = pb[i];
      ab = fmaf(u, v, ab); aa = fmaf(u, u, aa); bb = fmaf(v, v, bb);
    }
  } else {
    for (uint64_t i = e0 + threadIdx.x; i < e1; i += 256) {  // <= 64 elements per thread
      const float u = pa[i], v = pb[i];
      ab = fmaf(u, v, ab);
      aa = fmaf(u, u, aa);
      bb = fmaf(v, v, bb);
    }
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
    const float s = scale[n_channels == 1 ? 0 : channel_of(i, inner, n_channels, n < (1ull << 32))];
    const float q = __fdiv_rn(w[i], s);
    const float fl = floorf(q);
    const float rest = q - fl;
    // -log((zeta - gamma) / (rest - gamma) - 1)
    alpha[i] = -logf(__fdiv_rn(kZeta - kGamma, rest - kGamma) - 1.0f);
    wfloor[i] = fl;
  }
}

__device__ __forceinline__ float soft_w1(float wfl, float a, float s, float qmin, float qmax, int soft) {
  const float r = soft ? rect_sigmoid(a) : (a >= 0.f ? 1.f : 0.f);
  float q = wfl + r;
  q = fminf(fmaxf(q, qmin), qmax);
  return __fmul_rn(q, s);
}

// Weight-shaped tensors [n_channels][inner]: 16-byte accesses when a channel row is a multiple of four
// elements (then the four lanes of a vector share their scale), channel index in 32-bit arithmetic.
__global__ void __launch_bounds__(256, 4)
adaround_weight_kernel(const float* __restrict__ wfloor, const float* __restrict__ alpha,
                       const float* __restrict__ scale, int n_channels, uint64_t inner, uint64_t n,
                       float qmin, float qmax, int soft, float* __restrict__ wq, int vec) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool small = n < (1ull << 32);
  if (vec) {
    const uint64_t n4 = n >> 2;
    const float4* f4 = reinterpret_cast<const float4*>(wfloor);
    const float4* a4 = reinterpret_cast<const float4*>(alpha);
    float4* q4 = reinterpret_cast<float4*>(wq);
    for (uint64_t i = tid; i < n4; i += stride) {
      const float4 f = ldg_stream4(f4 + i), a = ldg_stream4(a4 + i);
      const float s = scale[n_channels == 1 ? 0 : channel_of(i << 2, inner, n_channels, small)];
      stg_stream4(q4 + i, make_float4(soft_w1(f.x, a.x, s, qmin, qmax, soft), soft_w1(f.y, a.y, s, qmin, qmax, soft),
                                      soft_w1(f.z, a.z, s, qmin, qmax, soft), soft_w1(f.w, a.w, s, qmin, qmax, soft)));
    }
    return;   // vec implies n % 4 == 0
  }
  for (uint64_t i = tid; i < n; i += stride) {
    const float s = scale[n_channels == 1 ? 0 : channel_of(i, inner, n_channels, small)];
    wq[i] = soft_w1(wfloor[i], alpha[i], s, qmin, qmax, soft);
  }
}

struct StepCfg {
  float qmin, qmax, beta, reg_alpha, lr, b1, b2, eps, bc1, bc2_sqrt, grad_scale;
};

// One element of the fused step: dL/d-alpha through the rectified sigmoid and the clamp, the
// regulariser gradient, torch.optim.Adam's single-tensor update. Returns the regulariser term.
__device__ __forceinline__ float step1(float gw, float wfl, float s, const StepCfg& c, float& a, float& mi,
                                       float& vi) {
  const float sg = sigmoidf_(a);
  const float hraw = rect_sigmoid_raw(sg);
  const float h = fminf(fmaxf(hraw, 0.f), 1.f);
  // clamp(0,1) passes the gradient on the closed interval (torch.clamp backward)
  const float dh = (hraw >= 0.f && hraw <= 1.f) ? (kZeta - kGamma) * sg * (1.f - sg) : 0.f;
  // max(., qmin) / min(., qmax): pass where strictly inside, half on an exact tie
  const float q = wfl + h;
  float pass = 1.f;
  if (q < c.qmin || q > c.qmax) pass = 0.f;
  else if (q == c.qmin || q == c.qmax) pass = 0.5f;
  float g = c.grad_scale * gw * s * pass * dh;
  float reg = 0.f;
  // regulariser: reg_alpha * sum(1 - |2h - 1|^beta)
  if (c.beta > 0.f) {
    const float u = fabsf(h - 0.5f) * 2.f;
    const float sgn = (h > 0.5f) ? 1.f : ((h < 0.5f) ? -1.f : 0.f);
    const float pw1 = powf(u, c.beta - 1.f);
    g += -c.reg_alpha * c.beta * pw1 * 2.f * sgn * dh;
    reg = c.reg_alpha * (1.f - pw1 * u);
  }
  // torch.optim.Adam (single-tensor path): lerp, addcmul, addcdiv
  mi = mi + (g - mi) * (1.f - c.b1);
  vi = vi * c.b2 + (1.f - c.b2) * g * g;
  const float denom = sqrtf(vi) / c.bc2_sqrt + c.eps;
  a = a - (c.lr / c.bc1) * (mi / denom);
  return reg;
}

__global__ void __launch_bounds__(256, 4)
adaround_step_kernel(const float* __restrict__ grad_w, const float* __restrict__ wfloor,
                     const float* __restrict__ scale, int n_channels, uint64_t inner, uint64_t n,
                     StepCfg c, float* __restrict__ alpha, float* __restrict__ m, float* __restrict__ v,
                     double* __restrict__ reg_out, const float* __restrict__ sched, int vec) {
  if (sched) {   // per-iteration scalars from device memory (CUDA-graph replay)
    c.beta = sched[0];
    c.bc1 = sched[1];
    c.bc2_sqrt = sched[2];
  }
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool small = n < (1ull << 32);
  double reg_acc = 0.0;
  if (vec) {
    const uint64_t n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(grad_w);
    const float4* f4 = reinterpret_cast<const float4*>(wfloor);
    float4* a4 = reinterpret_cast<float4*>(alpha);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (uint64_t i = tid; i < n4; i += stride) {
      const float4 g = ldg_stream4(g4 + i), f = ldg_stream4(f4 + i);
      float4 a = a4[i], mm = m4[i], vv = v4[i];
      const float s = scale[n_channels == 1 ? 0 : channel_of(i << 2, inner, n_channels, small)];
      float reg = step1(g.x, f.x, s, c, a.x, mm.x, vv.x);
      reg += step1(g.y, f.y, s, c, a.y, mm.y, vv.y);
      reg += step1(g.z, f.z, s, c, a.z, mm.z, vv.z);
      reg += step1(g.w, f.w, s, c, a.w, mm.w, vv.w);
      a4[i] = a;
      m4[i] = mm;
      v4[i] = vv;
      reg_acc += (double)reg;
    }
  } else {
    for (uint64_t i = tid; i < n; i += stride) {
      const float s = scale[n_channels == 1 ? 0 : channel_of(i, inner, n_channels, small)];
      float a = alpha[i], mi = m[i], vi = v[i];
      reg_acc += (double)step1(grad_w[i], wfloor[i], s, c, a, mi, vi);
      alpha[i] = a;
      m[i] = mi;
      v[i] = vi;
    }
  }
  if (reg_out) {
    reg_acc = warp_sum(reg_acc);
    if ((threadIdx.x & 31) == 0 && reg_acc != 0.0) atomicAdd(reg_out, reg_acc);
  }
}

// ---- K6 step with the gradient all-reduce inside (opt-in, DPL_PEER_ALLREDUCE=1) ----
// The reference averages dL/dW over ranks with DistributedDataParallel's bucketed NCCL all-reduce before
// Adam runs (adaround.py:121, brecq.py:163). Here every rank keeps its dL/dW in a buffer that all peers have
// mapped over NVLink (CUDA IPC through torch's symmetric memory) and the step kernel itself reads the W
// copies, adds them in rank order — so every replica computes bit-identical sums and stays identical — and
// applies the update: one launch and W reads of the gradient instead of an all-reduce (2 (W - 1) / W
// transfers plus a launch-latency-bound collective) followed by the step.
// Arrival protocol, per layer: each rank owns W 32-bit words, word r = the last epoch at which rank r
// announced "my gradient slot for this epoch is written". Every CTA stores its own rank's word on all peers
// (release, system scope; the same value from every CTA, so the store is idempotent and no CTA depends on
// another one being scheduled) and waits (acquire, bounded) until its own words show this epoch for all
// peers. Epochs only grow, so no reset and no second barrier is needed; the gradient lives in two slots
// used alternately: a rank can only overwrite slot e % 2 in iteration e + 2, after the arrival of iteration
// e + 1, by which time every peer has finished the step kernel of iteration e that read it.
constexpr int kMaxPeers = 8;
struct PeerCfg {
  const float* grads[kMaxPeers];   // this epoch's gradient slot on every rank (rank order; [rank] is local)
  uint32_t* words[kMaxPeers];      // the arrival words of every rank
  int world, rank;
  uint32_t epoch;                  // >= 1, identical on all ranks, +1 per iteration
};

__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// peer memory: coherent at system scope, never through the non-coherent path
__device__ __forceinline__ float4 ld_sys4(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ float ld_sys1(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256, 4)
adaround_step_peer_kernel(const PeerCfg pc, const float* __restrict__ wfloor, const float* __restrict__ scale,
                          int n_channels, uint64_t inner, uint64_t n, StepCfg c, float* __restrict__ alpha,
                          float* __restrict__ m, float* __restrict__ v, double* __restrict__ reg_out,
                          const float* __restrict__ sched, int vec, int* __restrict__ error_flag) {
  int late = 0;
  if (threadIdx.x < pc.world) {
    __threadfence_system();
    st_release_sys_u32(pc.words[threadIdx.x] + pc.rank, pc.epoch);
    const uint32_t* mine = pc.words[pc.rank] + threadIdx.x;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys_u32(mine) - pc.epoch) < 0) {
      if (clock64() - t0 > 4000000000ll) {   // ~2 s: a peer is gone; report instead of hanging the GPU
        late = 1;
        break;
      }
    }
  }
  if (__syncthreads_or(late)) {              // leave alpha / m / v untouched
    if (threadIdx.x == 0 && error_flag) atomicExch(error_flag, 1);
    return;
  }
  if (sched) {
    c.beta = sched[0];
    c.bc1 = sched[1];
    c.bc2_sqrt = sched[2];
  }
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool small = n < (1ull << 32);
  double reg_acc = 0.0;
  if (vec) {
    const uint64_t n4 = n >> 2;
    const float4* f4 = reinterpret_cast<const float4*>(wfloor);
    float4* a4 = reinterpret_cast<float4*>(alpha);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (uint64_t i = tid; i < n4; i += stride) {
      float4 g = ld_sys4(reinterpret_cast<const float4*>(pc.grads[0]) + i);
      for (int r = 1; r < pc.world; ++r) {
        const float4 t = ld_sys4(reinterpret_cast<const float4*>(pc.grads[r]) + i);
        g.x += t.x;
        g.y += t.y;
        g.z += t.z;
        g.w += t.w;
      }
      const float4 f = ldg_stream4(f4 + i);
      float4 a = a4[i], mm = m4[i], vv = v4[i];
      const float s = scale[n_channels == 1 ? 0 : channel_of(i << 2, inner, n_channels, small)];
      float reg = step1(g.x, f.x, s, c, a.x, mm.x, vv.x);
      reg += step1(g.y, f.y, s, c, a.y, mm.y, vv.y);
      reg += step1(g.z, f.z, s, c, a.z, mm.z, vv.z);
      reg += step1(g.w, f.w, s, c, a.w, mm.w, vv.w);
      a4[i] = a;
      m4[i] = mm;
      v4[i] = vv;
      reg_acc += (double)reg;
    }
  } else {
    for (uint64_t i = tid; i < n; i += stride) {
      float g = ld_sys1(pc.grads[0] + i);
      for (int r = 1; r < pc.world; ++r) g += ld_sys1(pc.grads[r] + i);
      const float s = scale[n_channels == 1 ? 0 : channel_of(i, inner, n_channels, small)];
      float a = alpha[i], mi = m[i], vi = v[i];
      reg_acc += (double)step1(g, wfloor[i], s, c, a, mi, vi);
      alpha[i] = a;
      m[i] = mi;
      v[i] = vi;
    }
  }
  if (reg_out) {
    reg_acc = warp_sum(reg_acc);
    if ((threadIdx.x & 31) == 0 && reg_acc != 0.0) atomicAdd(reg_out, reg_acc);
  }
}

inline bool al16q(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

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
  fakequant_kernel<<<stream_grid((n + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
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
  const int vec = ((reinterpret_cast<uintptr_t>(d_a) | reinterpret_cast<uintptr_t>(d_b)) & 15u) == 0 &&
                  (inner & 3u) == 0;
  channel_sumdiff_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_a, d_b, rows, channels, inner, d_sum, vec);
  DPL_LAUNCH_CHECK("channel_sumdiff_kernel");
  return 0;
}

extern "C" int dpl_cosine3_f32(const float* d_a, const float* d_b, uint64_t n_seg,
                               uint64_t seg_len, double* d_out, void* stream) {
  DPL_REQUIRE(d_a && d_b && d_out, "null pointer");
  if (n_seg == 0 || seg_len == 0) return 0;
  const uint64_t cps = (seg_len + kCosChunk - 1) / kCosChunk;
  DPL_REQUIRE(n_seg * cps < (1ull << 31), "too many chunks");
  // chunk starts are multiples of 16384 elements: rows are 16-byte aligned iff the bases and seg_len are
  const int vec = ((reinterpret_cast<uintptr_t>(d_a) | reinterpret_cast<uintptr_t>(d_b)) & 15u) == 0 &&
                  (seg_len & 3u) == 0;
  cosine3_kernel<<<(unsigned)(n_seg * cps), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_a, d_b, seg_len, cps, d_out, vec);
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
  const int vec = al16q(d_wfloor) && al16q(d_alpha) && al16q(d_wq) && (inner & 3u) == 0;
  adaround_weight_kernel<<<stream_grid(vec ? n / 4 : n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_wfloor, d_alpha, d_scale, n_channels, inner, n, qmin, qmax, soft, d_wq, vec);
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
  StepCfg c = {qmin, qmax, beta, reg_alpha, lr, b1, b2, eps, (float)bc1, (float)sqrt(bc2), grad_scale};
  const int vec = al16q(d_grad_w) && al16q(d_wfloor) && al16q(d_alpha) && al16q(d_m) && al16q(d_v) &&
                  (inner & 3u) == 0;
  adaround_step_kernel<<<stream_grid(vec ? n / 4 : n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_grad_w, d_wfloor, d_scale, n_channels, inner, n, c, d_alpha, d_m, d_v, d_reg, d_sched, vec);
  DPL_LAUNCH_CHECK("adaround_step_kernel");
  return 0;
}

extern "C" int dpl_adaround_step_peer_f32(const void* const* peer_grads, void* const* peer_words, int world,
                                          int rank, uint32_t epoch, const float* d_wfloor,
                                          const float* d_scale, int n_channels, uint64_t inner, float qmin,
                                          float qmax, float beta, float reg_alpha, float lr, float b1, float b2,
                                          float eps, int step, float* d_alpha, float* d_m, float* d_v,
                                          double* d_reg, const float* d_sched, int* d_error, void* stream) {
  DPL_REQUIRE(peer_grads && peer_words && d_wfloor && d_scale && d_alpha && d_m && d_v, "null pointer");
  DPL_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "bad world / rank");
  DPL_REQUIRE(n_channels >= 1 && inner >= 1 && step >= 1 && epoch >= 1, "bad arguments");
  const uint64_t n = (uint64_t)n_channels * inner;
  PeerCfg pc;
  int vec = al16q(d_wfloor) && al16q(d_alpha) && al16q(d_m) && al16q(d_v) && (inner & 3u) == 0;
  for (int r = 0; r < kMaxPeers; ++r) {
    pc.grads[r] = r < world ? static_cast<const float*>(peer_grads[r]) : nullptr;
    pc.words[r] = r < world ? static_cast<uint32_t*>(peer_words[r]) : nullptr;
    if (r < world) {
      DPL_REQUIRE(pc.grads[r] && pc.words[r], "null peer pointer");
      vec = vec && al16q(pc.grads[r]);
    }
  }
  pc.world = world;
  pc.rank = rank;
  pc.epoch = epoch;
  const double bc1 = 1.0 - pow((double)b1, (double)step);
  const double bc2 = 1.0 - pow((double)b2, (double)step);
  // the mean over ranks (DistributedDataParallel) is folded into the step
  StepCfg c = {qmin, qmax, beta, reg_alpha, lr, b1, b2, eps, (float)bc1, (float)sqrt(bc2), 1.0f / (float)world};
  // at most 2 CTAs per SM: every CTA announces to every peer, and all of them can be resident at once
  uint64_t g = ((vec ? n / 4 : n) + 255) / 256;
  const uint64_t cap = (uint64_t)sm_count() * 2;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  adaround_step_peer_kernel<<<(unsigned)g, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      pc, d_wfloor, d_scale, n_channels, inner, n, c, d_alpha, d_m, d_v, d_reg, d_sched, vec, d_error);
  DPL_LAUNCH_CHECK("adaround_step_peer_kernel");
  return 0;
}
