// Elementwise kernels of the AdaRound / BRECQ / QDrop reconstruction loop around the
// conv/GEMM re-evaluation (sm_100a, HBM bound). They replace the chains of small torch
// kernels + autograd bookkeeping of weight_transform/ada_quant_layer.py:28-36,113-114,
// 224-252 and brecq.py:167-172.
//   dpl_recon_act_f32      y = [drop-]fakequant(relu(o))            layer epilogue, forward
//   dpl_recon_act_bwd_f32  go = gy * [o > 0] * [element not quantised]   its backward
//   dpl_recon_loss_f32     L2 loss of the block output + dL/do in one pass
//   dpl_mix_drop_f32       QDrop block input: where(u < p, q_in, fp_in)
// The Bernoulli masks come from a counter-based generator keyed by (seed, element index),
// so the backward pass regenerates the forward's mask instead of storing it. round() has
// zero gradient in the reference (no straight-through estimator): gradients flow only
// through the elements that were NOT replaced by their quantised value.

#include <math.h>

#include "dpl_common.cuh"

namespace dpl {
namespace {

__device__ __forceinline__ float u01(uint64_t seed, uint64_t i) { return hash_u01(seed, i); }

// Per-iteration scalars live in device memory so that one captured CUDA graph can be replayed
// for every iteration of the optimisation loop:
//   sched[0] = beta(t), sched[1] = 1 - b1^(t+1), sched[2] = sqrt(1 - b2^(t+1)), seeds[l] per layer
struct ActCfg {
  int relu;        // apply max(o, 0)
  int quant;       // apply fake-quant (per-tensor scale, symmetric [qmin, qmax])
  float scale, qmin, qmax;
  float rscale;    // RN(1 / scale) for rint_div; 0 = use the plain IEEE division
  float prob;      // probability of taking the quantised value (1 = always, QDrop uses 0.5)
  uint64_t seed;
};

// returns the activation and whether the gradient passes through this element
__device__ __forceinline__ float act_fwd(float o, uint64_t i, const ActCfg& c, bool& pass) {
  pass = true;
  float y = o;
  if (c.relu) {
    pass = o > 0.f;
    y = fmaxf(o, 0.f);
  }
  if (c.quant) {
    // branch-free: the quantised value is always formed, the mask selects (a divergent branch
    // around the division cost more than the division)
    const bool take_q = (c.prob >= 1.0f) || (u01(c.seed, i) < c.prob);
    float q = c.rscale != 0.f ? rint_div(y, c.scale, c.rscale) : rintf(__fdiv_rn(y, c.scale));
    q = fminf(fmaxf(q, c.qmin), c.qmax);
    const float yq = __fmul_rn(q, c.scale);
    y = take_q ? yq : y;
    pass = pass && !take_q;
  }
  return y;
}

// All of them walk the tensor with 16-byte streaming loads / stores, two per operand in flight per
// thread (the scalar versions with a 64-bit splitmix mask reached 31 - 63 % of the HBM peak on a
// 205 MB blob; see profiles/ for these).
// Four elements at once (the quotients through rint_div4): y and the gradient-pass flags.
__device__ __forceinline__ float4 act_fwd4(float4 o, uint64_t e, const ActCfg& c, bool (&pass)[4]) {
  float4 y = o;
  pass[0] = pass[1] = pass[2] = pass[3] = true;
  if (c.relu) {
    pass[0] = o.x > 0.f; pass[1] = o.y > 0.f; pass[2] = o.z > 0.f; pass[3] = o.w > 0.f;
    y = make_float4(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f), fmaxf(o.z, 0.f), fmaxf(o.w, 0.f));
  }
  if (c.quant) {
    const float4 t = rint_div4(y, c.scale, c.rscale, c.rscale != 0.f);
    const bool all_q = c.prob >= 1.0f;
    const float4 u = all_q ? make_float4(0.f, 0.f, 0.f, 0.f) : hash_u01x4(c.seed, e);
    const bool q0 = all_q || (u.x < c.prob), q1 = all_q || (u.y < c.prob);
    const bool q2 = all_q || (u.z < c.prob), q3 = all_q || (u.w < c.prob);
    y.x = q0 ? __fmul_rn(fminf(fmaxf(t.x, c.qmin), c.qmax), c.scale) : y.x;
    y.y = q1 ? __fmul_rn(fminf(fmaxf(t.y, c.qmin), c.qmax), c.scale) : y.y;
    y.z = q2 ? __fmul_rn(fminf(fmaxf(t.z, c.qmin), c.qmax), c.scale) : y.z;
    y.w = q3 ? __fmul_rn(fminf(fmaxf(t.w, c.qmin), c.qmax), c.scale) : y.w;
    pass[0] = pass[0] && !q0; pass[1] = pass[1] && !q1; pass[2] = pass[2] && !q2; pass[3] = pass[3] && !q3;
  }
  return y;
}

__global__ void __launch_bounds__(256, 4)
recon_act_kernel(const float* __restrict__ o, float* __restrict__ y, uint64_t n, ActCfg c,
                 const unsigned long long* __restrict__ seed_ptr, int vec) {
  if (seed_ptr) c.seed = *seed_ptr;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t done = 0;
  if (vec) {
    // CTA-contiguous tiles of 256 x R float4 (see dpl_eltwise.cu), R loads in flight per thread
    constexpr int R = 4;
    const uint64_t n4 = n >> 2;
    const float4* o4 = reinterpret_cast<const float4*>(o);
    float4* y4 = reinterpret_cast<float4*>(y);
    const uint64_t tiles = (n4 + 256 * R - 1) / (256 * R);
    for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const uint64_t i0 = tile * (256 * R) + threadIdx.x;
      float4 v[R];
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (i0 + r * 256 < n4) v[r] = ldg_stream4(o4 + i0 + r * 256);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const uint64_t i = i0 + r * 256;
        bool pass[4];
        if (i < n4) stg_stream4(y4 + i, act_fwd4(v[r], i << 2, c, pass));
      }
    }
    done = n4 << 2;
  }
  for (uint64_t i = done + tid; i < n; i += stride) {
    bool pass;
    y[i] = act_fwd(o[i], i, c, pass);
  }
}

__device__ __forceinline__ float act_bwd1(float o, float gy, uint64_t i, const ActCfg& c) {
  bool pass;
  (void)act_fwd(o, i, c, pass);
  return pass ? gy : 0.f;
}

__global__ void __launch_bounds__(256, 4)
recon_act_bwd_kernel(const float* __restrict__ o, const float* __restrict__ gy,
                     float* __restrict__ go, uint64_t n, ActCfg c,
                     const unsigned long long* __restrict__ seed_ptr, int vec) {
  if (seed_ptr) c.seed = *seed_ptr;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t done = 0;
  if (vec) {
    const uint64_t n4 = n >> 2;
    const float4* o4 = reinterpret_cast<const float4*>(o);
    const float4* g4 = reinterpret_cast<const float4*>(gy);
    float4* r4 = reinterpret_cast<float4*>(go);
    for (uint64_t i = tid; i < n4; i += 2 * stride) {
      const bool two = i + stride < n4;
      const float4 a = ldg_stream4(o4 + i), ga = ldg_stream4(g4 + i);
      const float4 b = two ? ldg_stream4(o4 + i + stride) : a;
      const float4 gb = two ? ldg_stream4(g4 + i + stride) : ga;
      const uint64_t e = i << 2;
      stg_stream4(r4 + i, make_float4(act_bwd1(a.x, ga.x, e, c), act_bwd1(a.y, ga.y, e + 1, c),
                                      act_bwd1(a.z, ga.z, e + 2, c), act_bwd1(a.w, ga.w, e + 3, c)));
      if (two) {
        const uint64_t f = (i + stride) << 2;
        stg_stream4(r4 + i + stride, make_float4(act_bwd1(b.x, gb.x, f, c), act_bwd1(b.y, gb.y, f + 1, c),
                                                 act_bwd1(b.z, gb.z, f + 2, c), act_bwd1(b.w, gb.w, f + 3, c)));
      }
    }
    done = n4 << 2;
  }
  for (uint64_t i = done + tid; i < n; i += stride) go[i] = act_bwd1(o[i], gy[i], i, c);
}

// loss = sum((act(o) - tgt)^2) * inv_count ; go = 2 * (act(o) - tgt) * inv_count * pass
__device__ __forceinline__ float loss1(float o, float t, uint64_t i, const ActCfg& c, float inv_count,
                                       float& blk) {
  bool pass;
  const float y = act_fwd(o, i, c, pass);
  const float d = y - t;
  blk = fmaf(d, d, blk);
  return pass ? 2.f * d * inv_count : 0.f;
}

__global__ void __launch_bounds__(256, 4)
recon_loss_kernel(const float* __restrict__ o, const float* __restrict__ tgt,
                  float* __restrict__ go, uint64_t n, ActCfg c, float inv_count,
                  double* __restrict__ loss, const unsigned long long* __restrict__ seed_ptr, int vec) {
  if (seed_ptr) c.seed = *seed_ptr;
  __shared__ double s_red[8];
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0.0;
  uint64_t done = 0;
  if (vec) {
    const uint64_t n4 = n >> 2;
    const float4* o4 = reinterpret_cast<const float4*>(o);
    const float4* t4 = reinterpret_cast<const float4*>(tgt);
    float4* r4 = reinterpret_cast<float4*>(go);
    int k = 0;
    float blk = 0.f;
    for (uint64_t i = tid; i < n4; i += 2 * stride) {
      const bool two = i + stride < n4;
      const float4 a = ldg_stream4(o4 + i), ta = ldg_stream4(t4 + i);
      const float4 b = two ? ldg_stream4(o4 + i + stride) : a;
      const float4 tb = two ? ldg_stream4(t4 + i + stride) : ta;
      const uint64_t e = i << 2;
      float4 r;
      r.x = loss1(a.x, ta.x, e, c, inv_count, blk);
      r.y = loss1(a.y, ta.y, e + 1, c, inv_count, blk);
      r.z = loss1(a.z, ta.z, e + 2, c, inv_count, blk);
      r.w = loss1(a.w, ta.w, e + 3, c, inv_count, blk);
      stg_stream4(r4 + i, r);
      if (two) {
        const uint64_t f = (i + stride) << 2;
        r.x = loss1(b.x, tb.x, f, c, inv_count, blk);
        r.y = loss1(b.y, tb.y, f + 1, c, inv_count, blk);
        r.z = loss1(b.z, tb.z, f + 2, c, inv_count, blk);
        r.w = loss1(b.w, tb.w, f + 3, c, inv_count, blk);
        stg_stream4(r4 + i + stride, r);
      }
      if (++k == 4) {   // fold the float partial (<= 32 squares) into the double sum
        acc += (double)blk;
        blk = 0.f;
        k = 0;
      }
    }
    acc += (double)blk;
    done = n4 << 2;
  }
  {
    float blk = 0.f;
    for (uint64_t i = done + tid; i < n; i += stride) go[i] = loss1(o[i], tgt[i], i, c, inv_count, blk);
    acc += (double)blk;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && loss) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_red[w];
    atomicAdd(loss, t * (double)inv_count);
  }
}

// One thread: the schedule of iteration t = *iter (TempDecay, ada_quant_layer.py:117-130; Adam
// bias corrections; per-layer mask seeds), then t += 1.
__global__ void recon_schedule_kernel(int* __restrict__ iter, float* __restrict__ sched,
                                      unsigned long long* __restrict__ seeds, int n_seeds, double t_max,
                                      double rel_start, double start_b, double end_b, double b1, double b2,
                                      unsigned long long seed_base) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int t = *iter;
  const double start_decay = rel_start * t_max;
  double beta = 0.0;
  if (!((double)t < start_decay)) {
    const double rel_t = ((double)t - start_decay) / (t_max - start_decay);
    beta = end_b + 0.5 * (start_b - end_b) * (1.0 + cos(rel_t * 3.141592653589793));
  }
  sched[0] = (float)beta;
  sched[1] = (float)(1.0 - pow(b1, (double)(t + 1)));
  sched[2] = (float)sqrt(1.0 - pow(b2, (double)(t + 1)));
  for (int l = 0; l < n_seeds; ++l)
    seeds[l] = (seed_base * 1000003ull + (unsigned long long)t * 131ull + (unsigned long long)l * 7ull + 12345ull) &
               0x7FFFFFFFFFFFFFFFull;
  *iter = t + 1;
}

__global__ void __launch_bounds__(256, 4)
mix_drop_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y,
                uint64_t n, float prob, uint64_t seed, int vec) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t done = 0;
  if (vec) {
    const uint64_t n4 = n >> 2;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    float4* y4 = reinterpret_cast<float4*>(y);
    for (uint64_t i = tid; i < n4; i += 2 * stride) {
      const bool two = i + stride < n4;
      const float4 u0 = ldg_stream4(a4 + i), v0 = ldg_stream4(b4 + i);
      const float4 u1 = two ? ldg_stream4(a4 + i + stride) : u0;
      const float4 v1 = two ? ldg_stream4(b4 + i + stride) : v0;
      const float4 r0 = hash_u01x4(seed, i << 2);
      stg_stream4(y4 + i, make_float4(r0.x < prob ? u0.x : v0.x, r0.y < prob ? u0.y : v0.y,
                                      r0.z < prob ? u0.z : v0.z, r0.w < prob ? u0.w : v0.w));
      if (two) {
        const float4 r1 = hash_u01x4(seed, (i + stride) << 2);
        stg_stream4(y4 + i + stride, make_float4(r1.x < prob ? u1.x : v1.x, r1.y < prob ? u1.y : v1.y,
                                                 r1.z < prob ? u1.z : v1.z, r1.w < prob ? u1.w : v1.w));
      }
    }
    done = n4 << 2;
  }
  for (uint64_t i = done + tid; i < n; i += stride) y[i] = (u01(seed, i) < prob) ? a[i] : b[i];
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline unsigned grid_for(uint64_t n) { return stream_grid((n + 7) / 8); }

inline ActCfg make_cfg(int relu, int quant, float scale, float qmin, float qmax, float prob,
                       uint64_t seed) {
  ActCfg c;
  c.relu = relu;
  c.quant = quant;
  c.scale = scale;
  {
    const float r = 1.0f / scale;   // host division: correctly rounded
    const float as = fabsf(scale), ar = fabsf(r);
    c.rscale = (as >= 1.17549435e-38f && as < 1.0e37f && ar >= 1.17549435e-38f && ar < 1.0e37f) ? r : 0.f;
  }
  c.qmin = qmin;
  c.qmax = qmax;
  c.prob = prob;
  c.seed = seed;
  return c;
}

}  // namespace
}  // namespace dpl

using namespace dpl;

extern "C" int dpl_recon_act_f32(const float* d_o, float* d_y, uint64_t n, int relu, int quant,
                                 float scale, float qmin, float qmax, float prob, uint64_t seed,
                                 const unsigned long long* d_seed, void* stream) {
  DPL_REQUIRE(d_o && d_y, "null pointer");
  if (n == 0) return 0;
  recon_act_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_o, d_y, n, make_cfg(relu, quant, scale, qmin, qmax, prob, seed), d_seed, al16(d_o) && al16(d_y));
  DPL_LAUNCH_CHECK("recon_act_kernel");
  return 0;
}

extern "C" int dpl_recon_act_bwd_f32(const float* d_o, const float* d_gy, float* d_go, uint64_t n,
                                     int relu, int quant, float scale, float qmin, float qmax,
                                     float prob, uint64_t seed, const unsigned long long* d_seed,
                                     void* stream) {
  DPL_REQUIRE(d_o && d_gy && d_go, "null pointer");
  if (n == 0) return 0;
  recon_act_bwd_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_o, d_gy, d_go, n, make_cfg(relu, quant, scale, qmin, qmax, prob, seed), d_seed,
      al16(d_o) && al16(d_gy) && al16(d_go));
  DPL_LAUNCH_CHECK("recon_act_bwd_kernel");
  return 0;
}

extern "C" int dpl_recon_loss_f32(const float* d_o, const float* d_tgt, float* d_go, uint64_t n,
                                  int relu, int quant, float scale, float qmin, float qmax,
                                  float prob, uint64_t seed, float inv_count, double* d_loss,
                                  const unsigned long long* d_seed, void* stream) {
  DPL_REQUIRE(d_o && d_tgt && d_go, "null pointer");
  if (n == 0) return 0;
  recon_loss_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_o, d_tgt, d_go, n, make_cfg(relu, quant, scale, qmin, qmax, prob, seed), inv_count, d_loss,
      d_seed, al16(d_o) && al16(d_tgt) && al16(d_go));
  DPL_LAUNCH_CHECK("recon_loss_kernel");
  return 0;
}

extern "C" int dpl_mix_drop_f32(const float* d_a, const float* d_b, float* d_y, uint64_t n,
                                float prob, uint64_t seed, void* stream) {
  DPL_REQUIRE(d_a && d_b && d_y, "null pointer");
  if (n == 0) return 0;
  mix_drop_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_a, d_b, d_y, n, prob, seed, al16(d_a) && al16(d_b) && al16(d_y));
  DPL_LAUNCH_CHECK("mix_drop_kernel");
  return 0;
}

extern "C" int dpl_recon_schedule(int* d_iter, float* d_sched, unsigned long long* d_seeds, int n_seeds,
                                  double t_max, double rel_start, double start_b, double end_b, double b1,
                                  double b2, uint64_t seed_base, void* stream) {
  DPL_REQUIRE(d_iter && d_sched, "null pointer");
  DPL_REQUIRE(n_seeds == 0 || d_seeds, "null seeds");
  recon_schedule_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
      d_iter, d_sched, d_seeds, n_seeds, t_max, rel_start, start_b, end_b, b1, b2, seed_base);
  DPL_LAUNCH_CHECK("recon_schedule_kernel");
  return 0;
}
