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

__device__ __forceinline__ float u01(uint64_t seed, uint64_t i) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (i + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}

// Per-iteration scalars live in device memory so that one captured CUDA graph can be replayed
// for every iteration of the optimisation loop:
//   sched[0] = beta(t), sched[1] = 1 - b1^(t+1), sched[2] = sqrt(1 - b2^(t+1)), seeds[l] per layer
struct ActCfg {
  int relu;        // apply max(o, 0)
  int quant;       // apply fake-quant (per-tensor scale, symmetric [qmin, qmax])
  float scale, qmin, qmax;
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
    const bool take_q = (c.prob >= 1.0f) || (u01(c.seed, i) < c.prob);
    if (take_q) {
      float q = rintf(__fdiv_rn(y, c.scale));
      q = fminf(fmaxf(q, c.qmin), c.qmax);
      y = __fmul_rn(q, c.scale);
      pass = false;
    }
  }
  return y;
}

__global__ void __launch_bounds__(256)
recon_act_kernel(const float* __restrict__ o, float* __restrict__ y, uint64_t n, ActCfg c,
                 const unsigned long long* __restrict__ seed_ptr) {
  if (seed_ptr) c.seed = *seed_ptr;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    bool pass;
    y[i] = act_fwd(o[i], i, c, pass);
  }
}

__global__ void __launch_bounds__(256)
recon_act_bwd_kernel(const float* __restrict__ o, const float* __restrict__ gy,
                     float* __restrict__ go, uint64_t n, ActCfg c,
                     const unsigned long long* __restrict__ seed_ptr) {
  if (seed_ptr) c.seed = *seed_ptr;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    bool pass;
    (void)act_fwd(o[i], i, c, pass);
    go[i] = pass ? gy[i] : 0.f;
  }
}

// loss = sum((act(o) - tgt)^2) * inv_count ; go = 2 * (act(o) - tgt) * inv_count * pass
__global__ void __launch_bounds__(256)
recon_loss_kernel(const float* __restrict__ o, const float* __restrict__ tgt,
                  float* __restrict__ go, uint64_t n, ActCfg c, float inv_count,
                  double* __restrict__ loss, const unsigned long long* __restrict__ seed_ptr) {
  if (seed_ptr) c.seed = *seed_ptr;
  __shared__ double s_red[8];
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  double acc = 0.0;
  float blk = 0.f;
  int k = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    bool pass;
    const float y = act_fwd(o[i], i, c, pass);
    const float d = y - tgt[i];
    blk = fmaf(d, d, blk);
    if (++k == 32) {
      acc += (double)blk;
      blk = 0.f;
      k = 0;
    }
    go[i] = pass ? 2.f * d * inv_count : 0.f;
  }
  acc += (double)blk;
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

__global__ void __launch_bounds__(256)
mix_drop_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y,
                uint64_t n, float prob, uint64_t seed) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = (u01(seed, i) < prob) ? a[i] : b[i];
}

inline unsigned grid_for(uint64_t n) {
  uint64_t blocks = (n + 256ull * 4 - 1) / (256ull * 4);
  const uint64_t cap = (uint64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

inline ActCfg make_cfg(int relu, int quant, float scale, float qmin, float qmax, float prob,
                       uint64_t seed) {
  ActCfg c;
  c.relu = relu;
  c.quant = quant;
  c.scale = scale;
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
      d_o, d_y, n, make_cfg(relu, quant, scale, qmin, qmax, prob, seed), d_seed);
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
      d_o, d_gy, d_go, n, make_cfg(relu, quant, scale, qmin, qmax, prob, seed), d_seed);
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
      d_seed);
  DPL_LAUNCH_CHECK("recon_loss_kernel");
  return 0;
}

extern "C" int dpl_mix_drop_f32(const float* d_a, const float* d_b, float* d_y, uint64_t n,
                                float prob, uint64_t seed, void* stream) {
  DPL_REQUIRE(d_a && d_b && d_y, "null pointer");
  if (n == 0) return 0;
  mix_drop_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_a, d_b, d_y, n, prob,
                                                                             seed);
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
