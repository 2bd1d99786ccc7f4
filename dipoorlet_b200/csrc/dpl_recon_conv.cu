// K6 dense contraction for the k x k / strided convolutions of the AdaRound / BRECQ reconstruction loop
// (dipoorlet/weight_transform/ada_quant_layer.py:224-244: F.conv2d and, through autograd, its weight and
// data gradients, which torch runs on cuDNN with TF32 allowed). Single-pass TF32 on the 5th-generation
// tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM, operands staged by TMA) - the reference's
// own numerics for this step; the calibration forward keeps the 3xTF32 kernels of dpl_gemm.cu.
//
// All three contractions run on the channel-last, zero-bordered staging copies that dpl_pad_plane_f32
// writes (Xp of the layer input, Gp of the output gradient): in that layout a filter tap is a ROW SHIFT
// of the same matrix, so
//   forward  Y[q][co]      = sum_tap sum_ci Xp[q + shift(tap)][ci] * Wf[tap][co][ci]        (tap_conv)
//   dgrad    dX[q'][ci]    = sum_tap sum_co Gp[q' + shift'(tap)][co] * Wd[tap][ci][co]      (tap_conv, per
//                            output-parity class for stride 2, written with an output stride)
//   wgrad    dW[co][ci][t] = sum_q Gp[q][co] * Xp[q + shift(t)][ci]                         (tap_wgrad:
//                            both operands MN-major, the whole batch is the K dimension, split over CTAs)
// Border rows of Gp are zero, so the junk the shifted windows read at plane borders never contributes.
// Depthwise convolutions (MobileNetV2) do not fit the tensor cores: exact-fp32 SIMT weight / data
// gradients below (forward: dpl_dwconv2d_f32).

#include <cuda.h>
#include <math.h>

#include "dpl_common.cuh"
#include "dpl_tc.cuh"

namespace dpl {
// dpl_gemm.cu: the persistent kernel of dpl_x3p.cuh in single-pass mode (overlapped epilogue on eight warps)
int tap_conv_tf32_persistent(const float* d_xp, long long total_rows, const float* d_w_taps, int n_w_taps, float* d_y,
                             int n_img, int ck, int cn, int H, int W, int Hp, int Wp, int origin, int out_stride,
                             int out_a, int out_b, int n_taps, const int* tap_shift, const int* tap_w,
                             const float* d_bias, int* d_error_flag, cudaStream_t stream);
namespace {

constexpr int kRcStages = 3;                    // 3 x 32 KB: two CTAs per SM (one's epilogue under the other's loop)
constexpr int kRcStageBytes = 2 * kTileBytes;
constexpr int kRcThreads = 128;
constexpr int kRcTmemCols = 128;
constexpr int kMaxTaps = 9;

struct TapConvParams {
  int ck, cn;               // reduction channels (columns of Xp) / output channels
  int Wp, plane, origin;    // padded plane geometry of the q index
  int n_taps;
  int tap_shift[kMaxTaps];  // row offset of every tap in Xp
  int tap_w[kMaxTaps];      // its slice of Wt [T][cn][ck]
  int bn;                   // output channels per CTA (64 or 128)
  long long q_total;        // n_img * plane
  int H, W;                 // output image size
  int os, oa, ob;           // plane point (hq, wq) is output pixel (hq * os + oa, wq * os + ob)
  float* Y;                 // [n_img][cn][H][W]
  const float* bias;
  int* error_flag;
};

__global__ void __launch_bounds__(kRcThreads, 2)
tap_conv_tf32_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                     const TapConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_full[kRcStages], s_empty[kRcStages], s_tmem_full;
  __shared__ uint32_t s_tmem_base;
  __shared__ int s_fail;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tiles = (smem_addr(smem_raw) + 1023u) & ~1023u;
  const long long q0 = (long long)blockIdx.x * kBM;
  const int co0 = blockIdx.y * p.bn;
  const int num_kb = (p.ck + kBK - 1) / kBK;
  const int total_iters = p.n_taps * num_kb;
  const uint32_t w_tile_bytes = (uint32_t)p.bn * 128u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kRcStages; ++s) {
      bar_init(smem_addr(&s_full[s]), 1);
      bar_init(smem_addr(&s_empty[s]), 1);
    }
    bar_init(smem_addr(&s_tmem_full), 1);
    s_fail = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_addr(&s_tmem_base)),
                 "r"(kRcTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_acc = s_tmem_base;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer: X tile (shifted rows of the staging copy) + the tap's W tile, both K-major =====
    for (int it = 0; it < total_iters; ++it) {
      const int s = it % kRcStages;
      const uint32_t ph = (it / kRcStages) & 1;
      if (!bar_wait(smem_addr(&s_empty[s]), ph ^ 1)) {
        s_fail = 1;
        break;
      }
      const uint32_t full = smem_addr(&s_full[s]);
      bar_expect_tx(full, kTileBytes + w_tile_bytes);
      const int kb = it / p.n_taps, tap = it - kb * p.n_taps;
      const int k0 = kb * kBK;
      const uint32_t x_tile = tiles + s * kRcStageBytes, w_tile = x_tile + kTileBytes;
      tma_load_3d(x_tile, &tmX, k0, (int)q0 + p.tap_shift[tap], 0, full);
      tma_load_3d(w_tile, &tmW, k0, co0, p.tap_w[tap], full);
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.bn >> 3) << 17) |
                           ((uint32_t)(kBM >> 4) << 24);      // D f32, A = B = tf32, both K-major
    bool failed = false;
    for (int it = 0; it < total_iters; ++it) {
      const int s = it % kRcStages;
      const uint32_t ph = (it / kRcStages) & 1;
      if (!bar_wait(smem_addr(&s_full[s]), ph)) {
        s_fail = 1;
        failed = true;
        break;
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t x_tile = tiles + s * kRcStageBytes, w_tile = x_tile + kTileBytes;
#pragma unroll
      for (int j = 0; j < kBK / kUmmaK; ++j) {
        const uint64_t da = desc_k_major(x_tile, j), db = desc_k_major(w_tile, j);
        const uint32_t accumulate = (it > 0 || j > 0) ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_acc), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
            : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_addr(&s_empty[s]))
                   : "memory");
    }
    if (!failed)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_addr(&s_tmem_full))
                   : "memory");
  }
  __syncwarp();

  // ===== epilogue: TMEM lane = plane point q (consecutive lanes = consecutive pixels), columns = channels =====
  bool ok = bar_wait(smem_addr(&s_tmem_full), 0);
  ok = __all_sync(0xffffffffu, ok);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (ok) {
    const long long q = q0 + warp * 32 + lane;
    bool valid = q < p.q_total;
    long long out_base = 0;
    if (valid) {
      const int img = (int)(q / p.plane);
      const int r = (int)(q - (long long)img * p.plane);
      const int hp = r / p.Wp, wp = r - hp * p.Wp;
      const int hq = hp - p.origin, wq = wp - p.origin;
      const int ho = hq * p.os + p.oa, wo = wq * p.os + p.ob;
      valid = hq >= 0 && wq >= 0 && ho < p.H && wo < p.W;
      out_base = (((long long)img * p.cn) * p.H + ho) * p.W + wo;
    }
    const long long ch_stride = (long long)p.H * p.W;
#pragma unroll 1
    for (int c = 0; c < p.bn / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (valid) {
        const int cb = co0 + c * 32;
        float* dst = p.Y + out_base + (long long)cb * ch_stride;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (cb + j < p.cn) {
            float v = __uint_as_float(r[j]);
            if (p.bias) v += __ldg(p.bias + cb + j);
            dst[(long long)j * ch_stride] = v;
          }
        }
      }
    }
  }
  if (!ok || s_fail) {
    if ((threadIdx.x & 31) == 0 && p.error_flag) atomicExch(p.error_flag, 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(kRcTmemCols)
                 : "memory");
  }
}

// ---- weight gradient: dW[co][ci][t] = sum_q Gp[q][co] * Xp[q + shift(t)][ci] ------------------------------
struct TapWgradParams {
  int M, N;                 // c_out, c_in
  int n_taps;
  int tap_shift[kMaxTaps];  // row offset of the tap in Xp (plane base + window shift)
  int tap_col[kMaxTaps];    // its position in the filter's [kh * kw] axis
  int t_full;               // kh * kw of the filter (stride of ci in dW)
  int kb_total, kb_per_cta; // K blocks of 32 plane points, and how many one CTA reduces
  int bn;                   // input channels per CTA (64 or 128)
  float* DW;                // [c_out][c_in][t_full]
  int atomic_out;           // several CTAs along K: red.add into the zeroed output
  int* error_flag;
};

__global__ void __launch_bounds__(kRcThreads, 2)
tap_wgrad_tf32_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX,
                      const TapWgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_full[kRcStages], s_empty[kRcStages], s_tmem_full;
  __shared__ uint32_t s_tmem_base;
  __shared__ int s_fail;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tiles = (smem_addr(smem_raw) + 1023u) & ~1023u;
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * p.bn;
  const int tap = blockIdx.z % p.n_taps, split = blockIdx.z / p.n_taps;
  const int kb_begin = split * p.kb_per_cta;
  const int total_iters = min(p.kb_per_cta, p.kb_total - kb_begin);
  const int a_blocks = min(kBM / 32, (p.M - m0 + 31) / 32);     // 32-channel column blocks that exist
  const int b_blocks = min(p.bn / 32, (p.N - n0 + 31) / 32);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kRcStages; ++s) {
      bar_init(smem_addr(&s_full[s]), 1);
      bar_init(smem_addr(&s_empty[s]), 1);
    }
    bar_init(smem_addr(&s_tmem_full), 1);
    s_fail = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_addr(&s_tmem_base)),
                 "r"(kRcTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_acc = s_tmem_base;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer: [32 plane points][32 channels] boxes of Gp and of the shifted Xp =====
    // Column blocks past the channel count are not loaded: the rows / columns of D they feed are never
    // stored, and a shared-memory row of A (B) only reaches its own row (column) of D.
    const int shift = p.tap_shift[tap];
    for (int it = 0; it < total_iters; ++it) {
      const int s = it % kRcStages;
      const uint32_t ph = (it / kRcStages) & 1;
      if (!bar_wait(smem_addr(&s_empty[s]), ph ^ 1)) {
        s_fail = 1;
        break;
      }
      const uint32_t full = smem_addr(&s_full[s]);
      bar_expect_tx(full, (uint32_t)(a_blocks + b_blocks) * (kBK * 128));
      const int k0 = (kb_begin + it) * kBK;
      const uint32_t a_tile = tiles + s * kRcStageBytes, b_tile = a_tile + kTileBytes;
      for (int j = 0; j < a_blocks; ++j) tma_load_3d(a_tile + j * (kBK * 128), &tmG, m0 + 32 * j, k0, 0, full);
      for (int j = 0; j < b_blocks; ++j)
        tma_load_3d(b_tile + j * (kBK * 128), &tmX, n0 + 32 * j, k0 + shift, 0, full);
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer: A and B both MN-major =====
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                           ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
    bool failed = false;
    for (int it = 0; it < total_iters; ++it) {
      const int s = it % kRcStages;
      const uint32_t ph = (it / kRcStages) & 1;
      if (!bar_wait(smem_addr(&s_full[s]), ph)) {
        s_fail = 1;
        failed = true;
        break;
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_tile = tiles + s * kRcStageBytes, b_tile = a_tile + kTileBytes;
#pragma unroll
      for (int j = 0; j < kBK / kUmmaK; ++j) {
        const uint64_t da = desc_mn_major(a_tile, j), db = desc_mn_major(b_tile, j);
        const uint32_t accumulate = (it > 0 || j > 0) ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_acc), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
            : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_addr(&s_empty[s]))
                   : "memory");
    }
    if (!failed)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_addr(&s_tmem_full))
                   : "memory");
  }
  __syncwarp();

  // ===== epilogue: TMEM lane = output channel, columns = input channels; dW is [co][ci][t_full] =====
  bool ok = true;
  if (total_iters > 0) ok = bar_wait(smem_addr(&s_tmem_full), 0);
  ok = __all_sync(0xffffffffu, ok);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (ok && total_iters > 0) {
    const int m = m0 + warp * 32 + lane;
    float* drow = p.DW + (long long)m * p.N * p.t_full + p.tap_col[tap];
#pragma unroll 1
    for (int c = 0; c < p.bn / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (m < p.M) {
        const int nc = n0 + c * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (nc + j < p.N) {
            float* dst = drow + (long long)(nc + j) * p.t_full;
            if (p.atomic_out)
              atomicAdd(dst, __uint_as_float(r[j]));
            else
              *dst = __uint_as_float(r[j]);
          }
        }
      }
    }
  }
  if (!ok || s_fail) {
    if ((threadIdx.x & 31) == 0 && p.error_flag) atomicExch(p.error_flag, 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(kRcTmemCols)
                 : "memory");
  }
}

// w [co][ci][T] -> wf [T][co][ci] (forward taps) and wd [T][ci][co] (data-gradient taps); either may be null.
__global__ void __launch_bounds__(256)
taps_layout_kernel(const float* __restrict__ w, float* __restrict__ wf, float* __restrict__ wd, int co, int ci,
                   int T) {
  const long long total = (long long)co * ci * T;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int t = (int)(e % T);
    const long long oi = e / T;
    const int i = (int)(oi % ci), o = (int)(oi / ci);
    const float v = w[e];
    if (wf) wf[((long long)t * co + o) * ci + i] = v;
    if (wd) wd[((long long)t * ci + i) * co + o] = v;
  }
}

// ---- depthwise convolution gradients (exact fp32, SIMT) ----------------------------------------------------
// dW[c][kh][kw] = sum_{img, ho, wo} dY[img][c][ho][wo] * X[img][c][ho s + kh - p][wo s + kw - p]
// grid (C, splits over images); a thread keeps the KS x KS partial sums of its pixels in registers.
template <int KS>
__global__ void __launch_bounds__(256)
dw_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gw, int n_img, int C,
                int H, int W, int Ho, int Wo, int stride, int pad, int img_per_cta) {
  const int c = blockIdx.x;
  const int img0 = blockIdx.y * img_per_cta, img1 = min(n_img, img0 + img_per_cta);
  float acc[KS * KS];
#pragma unroll
  for (int k = 0; k < KS * KS; ++k) acc[k] = 0.f;
  const int px = Ho * Wo;
  for (int img = img0; img < img1; ++img) {
    const float* xi = x + ((long long)img * C + c) * H * W;
    const float* gi = gy + ((long long)img * C + c) * px;
    for (int q = threadIdx.x; q < px; q += blockDim.x) {
      const int ho = q / Wo, wo = q - ho * Wo;
      const float g = ldg_stream1(gi + q);
      const int h0 = ho * stride - pad, w0 = wo * stride - pad;
#pragma unroll
      for (int a = 0; a < KS; ++a) {
        const int h = h0 + a;
        if (h < 0 || h >= H) continue;
#pragma unroll
        for (int b = 0; b < KS; ++b) {
          const int w = w0 + b;
          if (w >= 0 && w < W) acc[a * KS + b] = fmaf(g, __ldg(xi + h * W + w), acc[a * KS + b]);
        }
      }
    }
  }
  __shared__ float red[8][KS * KS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < KS * KS; ++k) {
    const float v = warp_sum(acc[k]);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < KS * KS) {
    float v = 0.f;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) v += red[wv][threadIdx.x];
    atomicAdd(gw + (long long)c * KS * KS + threadIdx.x, v);
  }
}

// dX[img][c][h][w] = sum_{kh, kw : (h + p - kh) % s == 0, ...} dY[img][c][(h + p - kh) / s][(w + p - kw) / s] * Wt[c][kh][kw]
template <int KS>
__global__ void __launch_bounds__(256)
dw_dgrad_kernel(const float* __restrict__ gy, const float* __restrict__ wt, float* __restrict__ gx, long long total,
                int C, int H, int W, int Ho, int Wo, int stride, int pad) {
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {
    const int w = (int)(e % W);
    const int h = (int)((e / W) % H);
    const long long ic = e / ((long long)W * H);     // img * C + c
    const int c = (int)(ic % C);
    const float* gi = gy + ic * Ho * Wo;
    const float* wc = wt + (long long)c * KS * KS;
    float v = 0.f;
#pragma unroll
    for (int a = 0; a < KS; ++a) {
      const int hn = h + pad - a;
      if (hn < 0 || hn % stride) continue;
      const int ho = hn / stride;
      if (ho >= Ho) continue;
#pragma unroll
      for (int b = 0; b < KS; ++b) {
        const int wn = w + pad - b;
        if (wn < 0 || wn % stride) continue;
        const int wo = wn / stride;
        if (wo < Wo) v = fmaf(__ldg(gi + ho * Wo + wo), __ldg(wc + a * KS + b), v);
      }
    }
    gx[e] = v;
  }
}

}  // namespace
}  // namespace dpl

using namespace dpl;

// Tap-table convolution in single-pass TF32 (forward and data gradient of the reconstruction loop):
//   Y[img][co][hq os + oa][wq os + ob] = bias[co] + sum_tap sum_ck Wt[tap_w[tap]][co][ck] * Xp[q + tap_shift[tap]][ck],
//   q = img * Hp * Wp + (hq + origin) * Wp + (wq + origin); points whose pixel falls outside H x W are dropped.
//   d_xp       channel-last staging copy from dpl_pad_plane_f32, [total_rows][ck]
//   d_w_taps   [n_w_taps][cn][ck]
extern "C" int dpl_tap_conv_tf32(const float* d_xp, long long total_rows, const float* d_w_taps, int n_w_taps,
                                 float* d_y, int n_img, int ck, int cn, int H, int W, int Hp, int Wp, int origin,
                                 int out_stride, int out_a, int out_b, int n_taps, const int* tap_shift,
                                 const int* tap_w, const float* d_bias, int* d_error_flag, void* stream) {
  DPL_REQUIRE(d_xp && d_w_taps && d_y && tap_shift && tap_w, "null pointer");
  DPL_REQUIRE(n_img > 0 && ck > 0 && cn > 0 && H > 0 && W > 0 && Hp > 0 && Wp > 0, "empty problem");
  DPL_REQUIRE(n_taps >= 1 && n_taps <= kMaxTaps && n_w_taps >= 1, "1 <= n_taps <= 9");
  DPL_REQUIRE(origin == 0 || origin == 1, "origin must be 0 or 1");
  DPL_REQUIRE(out_stride >= 1 && out_a >= 0 && out_b >= 0 && out_a < out_stride && out_b < out_stride,
              "output stride / phase");
  const long long plane = (long long)Hp * Wp;
  const long long q_total = (long long)n_img * plane;
  DPL_REQUIRE(total_rows >= q_total && total_rows < (1ll << 31) - 4096, "total_rows out of range");
  if (ck & 3) {
    set_error("dpl_tap_conv_tf32: the reduction channel count must be a multiple of 4 (TMA stride alignment)");
    return DPL_E_UNSUPPORTED;
  }
  for (int t = 0; t < n_taps; ++t) DPL_REQUIRE(tap_w[t] >= 0 && tap_w[t] < n_w_taps, "tap_w out of range");
  {
    const int st_p = tap_conv_tf32_persistent(d_xp, total_rows, d_w_taps, n_w_taps, d_y, n_img, ck, cn, H, W, Hp, Wp,
                                              origin, out_stride, out_a, out_b, n_taps, tap_shift, tap_w, d_bias,
                                              d_error_flag, static_cast<cudaStream_t>(stream));
    if (st_p >= 0) return st_p;     // < 0: persistent path switched off, take the one-tile-per-CTA kernel below
  }
  CUtensorMap tmX, tmW;
  int st = make_map(&tmX, d_xp, (uint64_t)ck, (uint64_t)total_rows, 1, (uint64_t)ck, 0, kBM, false);
  if (st) return st;
  const int bn = cn <= 64 ? 64 : 128;
  st = make_map(&tmW, d_w_taps, (uint64_t)ck, (uint64_t)cn, (uint64_t)n_w_taps, (uint64_t)ck, (uint64_t)cn * ck,
                (uint32_t)bn, false);
  if (st) return st;
  TapConvParams p;
  p.ck = ck;
  p.cn = cn;
  p.Wp = Wp;
  p.plane = (int)plane;
  p.origin = origin;
  p.n_taps = n_taps;
  for (int t = 0; t < kMaxTaps; ++t) {
    p.tap_shift[t] = t < n_taps ? tap_shift[t] : 0;
    p.tap_w[t] = t < n_taps ? tap_w[t] : 0;
  }
  p.bn = bn;
  p.q_total = q_total;
  p.H = H;
  p.W = W;
  p.os = out_stride;
  p.oa = out_a;
  p.ob = out_b;
  p.Y = d_y;
  p.bias = d_bias;
  p.error_flag = d_error_flag;
  dim3 grid((unsigned)((q_total + kBM - 1) / kBM), (unsigned)((cn + bn - 1) / bn), 1);
  const size_t smem = (size_t)kRcStages * kRcStageBytes + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    int e = cuda_status(cudaFuncSetAttribute(tap_conv_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem),
                        "cudaFuncSetAttribute(tap_conv_tf32_kernel)");
    if (e) return e;
    attr_done = true;
  }
  tap_conv_tf32_kernel<<<grid, kRcThreads, smem, static_cast<cudaStream_t>(stream)>>>(tmX, tmW, p);
  DPL_LAUNCH_CHECK("tap_conv_tf32_kernel");
  return 0;
}

// Weight gradient of a tap-table convolution (all taps in one launch, the batch folded into K):
//   dW[co][ci][tap_col[t]] = sum_{q < q_total} Gp[q][co] * Xp[q + tap_shift[t]][ci]
//   d_gp  [q_total][c_out]  channel-last staging copy of dY with ZERO border rows (dpl_pad_plane_f32)
//   d_xp  [x_rows][c_in]    staging copy of the layer input (the forward's)
//   d_dw  [c_out][c_in][t_full], fully overwritten when n_taps == t_full (zeroed here first)
extern "C" int dpl_tap_wgrad_tf32(const float* d_gp, long long q_total, const float* d_xp, long long x_rows,
                                  float* d_dw, int c_out, int c_in, int t_full, int n_taps, const int* tap_shift,
                                  const int* tap_col, int* d_error_flag, void* stream) {
  DPL_REQUIRE(d_gp && d_xp && d_dw && tap_shift && tap_col, "null pointer");
  DPL_REQUIRE(c_out > 0 && c_in > 0 && q_total > 0 && x_rows > 0, "empty problem");
  DPL_REQUIRE(n_taps >= 1 && n_taps <= kMaxTaps && t_full >= n_taps, "1 <= n_taps <= 9");
  DPL_REQUIRE(q_total < (1ll << 31) - 4096 && x_rows < (1ll << 31) - 4096, "row count out of range");
  if ((c_out & 3) || (c_in & 3)) {
    set_error("dpl_tap_wgrad_tf32: channel counts must be multiples of 4 (TMA stride alignment)");
    return DPL_E_UNSUPPORTED;
  }
  for (int t = 0; t < n_taps; ++t) DPL_REQUIRE(tap_col[t] >= 0 && tap_col[t] < t_full, "tap_col out of range");
  CUtensorMap tmG, tmX;
  int st = make_map(&tmG, d_gp, (uint64_t)c_out, (uint64_t)q_total, 1, (uint64_t)c_out, 0, kBK, true);
  if (!st) st = make_map(&tmX, d_xp, (uint64_t)c_in, (uint64_t)x_rows, 1, (uint64_t)c_in, 0, kBK, true);
  if (st) return st;
  TapWgradParams p;
  p.M = c_out;
  p.N = c_in;
  p.n_taps = n_taps;
  for (int t = 0; t < kMaxTaps; ++t) {
    p.tap_shift[t] = t < n_taps ? tap_shift[t] : 0;
    p.tap_col[t] = t < n_taps ? tap_col[t] : 0;
  }
  p.t_full = t_full;
  p.bn = c_in <= 64 ? 64 : 128;
  p.kb_total = (int)((q_total + kBK - 1) / kBK);
  const int tiles = ((c_out + kBM - 1) / kBM) * ((c_in + p.bn - 1) / p.bn) * n_taps;
  // two CTAs per SM; a CTA should still reduce >= 16 K blocks so that the pipeline fills
  int splits = (2 * sm_count() + tiles - 1) / tiles;
  const int max_splits = (p.kb_total + 15) / 16;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.kb_per_cta = (p.kb_total + splits - 1) / splits;
  splits = (p.kb_total + p.kb_per_cta - 1) / p.kb_per_cta;
  DPL_REQUIRE((long long)splits * n_taps <= 65535, "grid limit");
  p.atomic_out = splits > 1 ? 1 : 0;
  p.DW = d_dw;
  p.error_flag = d_error_flag;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (p.atomic_out || n_taps < t_full) {
    int e = cuda_status(cudaMemsetAsync(d_dw, 0, (size_t)c_out * c_in * t_full * sizeof(float), s),
                        "cudaMemsetAsync(dW)");
    if (e) return e;
  }
  dim3 grid((unsigned)((c_out + kBM - 1) / kBM), (unsigned)((c_in + p.bn - 1) / p.bn), (unsigned)(splits * n_taps));
  const size_t smem = (size_t)kRcStages * kRcStageBytes + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    int e = cuda_status(cudaFuncSetAttribute(tap_wgrad_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem),
                        "cudaFuncSetAttribute(tap_wgrad_tf32_kernel)");
    if (e) return e;
    attr_done = true;
  }
  tap_wgrad_tf32_kernel<<<grid, kRcThreads, smem, s>>>(tmG, tmX, p);
  DPL_LAUNCH_CHECK("tap_wgrad_tf32_kernel");
  return 0;
}

// Filter re-layout for the tap-table kernels: w [c_out][c_in][T] -> d_wf [T][c_out][c_in] and / or
// d_wd [T][c_in][c_out] (the data gradient contracts over c_out).
extern "C" int dpl_taps_layout_f32(const float* d_w, float* d_wf, float* d_wd, int c_out, int c_in, int T,
                                   void* stream) {
  DPL_REQUIRE(d_w && (d_wf || d_wd), "null pointer");
  DPL_REQUIRE(c_out > 0 && c_in > 0 && T > 0, "empty filter");
  const uint64_t total = (uint64_t)c_out * c_in * T;
  taps_layout_kernel<<<stream_grid(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_w, d_wf, d_wd, c_out,
                                                                                        c_in, T);
  DPL_LAUNCH_CHECK("taps_layout_kernel");
  return 0;
}

// Depthwise convolution weight gradient, exact fp32: d_gw [C][k][k] (zeroed here, then accumulated).
extern "C" int dpl_dwconv2d_wgrad_f32(const float* d_x, const float* d_gy, float* d_gw, int n_img, int C, int H, int W,
                                      int k, int stride, int pad, int Ho, int Wo, void* stream) {
  DPL_REQUIRE(d_x && d_gy && d_gw, "null pointer");
  DPL_REQUIRE(n_img > 0 && C > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0 && stride > 0 && pad >= 0, "bad geometry");
  if (k != 3 && k != 5) {
    set_error("dpl_dwconv2d_wgrad_f32: kernel size %d not supported (3 or 5)", k);
    return DPL_E_UNSUPPORTED;
  }
  DPL_REQUIRE(C <= 65535 * 32, "too many channels");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int e = cuda_status(cudaMemsetAsync(d_gw, 0, (size_t)C * k * k * sizeof(float), s), "cudaMemsetAsync(gw)");
  if (e) return e;
  // enough CTAs to fill the machine: C x splits >= 4 per SM when the batch allows it
  int splits = (4 * sm_count() + C - 1) / C;
  if (splits > n_img) splits = n_img;
  if (splits < 1) splits = 1;
  const int per = (n_img + splits - 1) / splits;
  splits = (n_img + per - 1) / per;
  dim3 grid((unsigned)C, (unsigned)splits, 1);
  if (k == 3)
    dw_wgrad_kernel<3><<<grid, 256, 0, s>>>(d_x, d_gy, d_gw, n_img, C, H, W, Ho, Wo, stride, pad, per);
  else
    dw_wgrad_kernel<5><<<grid, 256, 0, s>>>(d_x, d_gy, d_gw, n_img, C, H, W, Ho, Wo, stride, pad, per);
  DPL_LAUNCH_CHECK("dw_wgrad_kernel");
  return 0;
}

// Depthwise convolution data gradient, exact fp32: d_gx [n_img][C][H][W] fully written.
extern "C" int dpl_dwconv2d_dgrad_f32(const float* d_gy, const float* d_w, float* d_gx, int n_img, int C, int H, int W,
                                      int k, int stride, int pad, int Ho, int Wo, void* stream) {
  DPL_REQUIRE(d_gy && d_w && d_gx, "null pointer");
  DPL_REQUIRE(n_img > 0 && C > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0 && stride > 0 && pad >= 0, "bad geometry");
  if (k != 3 && k != 5) {
    set_error("dpl_dwconv2d_dgrad_f32: kernel size %d not supported (3 or 5)", k);
    return DPL_E_UNSUPPORTED;
  }
  const long long total = (long long)n_img * C * H * W;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (k == 3)
    dw_dgrad_kernel<3><<<stream_grid((uint64_t)total), 256, 0, s>>>(d_gy, d_w, d_gx, total, C, H, W, Ho, Wo, stride,
                                                                    pad);
  else
    dw_dgrad_kernel<5><<<stream_grid((uint64_t)total), 256, 0, s>>>(d_gy, d_w, d_gx, total, C, H, W, Ho, Wo, stride,
                                                                    pad);
  DPL_LAUNCH_CHECK("dw_dgrad_kernel");
  return 0;
}
