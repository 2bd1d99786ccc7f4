// Direct fp32 convolution for layers with very few input channels — ResNet's 7x7 / stride 2 stem and
// MobileNetV2's 3x3 / stride 2 stem on the 3 image channels (the Conv node ORT executes in
// dipoorlet/forward_net.py:200-216). With C_in = 3 the reduction is only 27 .. 147 long and the
// tensor-core tiles would spend their time staging an im2col copy (measured: 475 MB per 64 images,
// slower than cuDNN), so this one runs on the FMA pipe, exact fp32 like the reference's CPU path:
//   * a CTA computes an 8 x 32 pixel tile of up to 64 output channels; the input patch (zero filled
//     outside the image = the padding) and the [k][co] transposed filter live in shared memory;
//   * a thread owns 4 consecutive pixels x 16 channels (64 accumulators). Per (channel, filter row) it
//     loads its input row segment ONCE with 16-byte shared loads and reuses it for every filter
//     column; the 16 filter values of a tap are the same address for the whole warp (broadcast);
//     4 + 4 * KW vector loads feed 64 * KW FMAs, so the kernel is FMA-issue bound;
//   * epilogue: + bias, 16-byte stores of 4 pixels per channel (a warp writes 128-byte row segments),
//     optional second output max(y, 0) and the fused range statistics of dpl_clip_f32.

#include <math.h>

#include "dpl_common.cuh"

namespace dpl {
namespace {

constexpr int kDcTH = 8, kDcTW = 32;       // output tile (rows x columns)
constexpr int kDcPX = 4, kDcCH = 16;       // per thread: consecutive pixels x channels
constexpr int kDcCoTile = 64;              // output channels per CTA
constexpr int kDcThreads = (kDcTH * kDcTW / kDcPX) * (kDcCoTile / kDcCH);   // 64 pixel groups x 4 = 256

template <int KW, int S>
struct DcGeom {
  static constexpr int kXin = ((kDcPX - 1) * S + KW + 3) / 4 * 4;          // input floats a thread reads per row
  static constexpr int kPatchW = (kDcTW / kDcPX - 1) * kDcPX * S + kXin;   // padded patch width (multiple of 4)
};

struct DcParams {
  const float* x;
  const float* w;       // [c_out][C][KH][KW]
  const float* bias;    // or null
  float* y;
  float* y2;            // optional max(y, 0)
  float *bmin, *bmax, *rmin, *rmax;
  int n_img, C, H, W, c_out, pad, Ho, Wo, co_tiles;
};

template <int KH, int KW, int S>
__global__ void __launch_bounds__(kDcThreads, 2)
conv_direct_kernel(const DcParams p) {
  using G = DcGeom<KW, S>;
  constexpr int PR = (kDcTH - 1) * S + KH;      // patch rows
  constexpr int PW = G::kPatchW;
  extern __shared__ __align__(16) float smem[];
  float* x_s = smem;                             // [C][PR][PW]
  float* w_s = smem + p.C * PR * PW;             // [C * KH * KW][64]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int img = blockIdx.z / p.co_tiles, cot = blockIdx.z - img * p.co_tiles;
  const int co0 = cot * kDcCoTile;
  const int ho0 = blockIdx.y * kDcTH, wo0 = blockIdx.x * kDcTW;
  const int K = p.C * KH * KW;

  // ---- stage the filter ([k][co], zero for channels past c_out) and the input patch ----
  for (int i = tid; i < K * kDcCoTile; i += kDcThreads) {
    const int k = i / kDcCoTile, co = i - k * kDcCoTile;
    w_s[i] = (co0 + co < p.c_out) ? __ldg(p.w + (long long)(co0 + co) * K + k) : 0.f;
  }
  {
    const int h_base = ho0 * S - p.pad, w_base = wo0 * S - p.pad;
    const float* xi = p.x + (long long)img * p.C * p.H * p.W;
    for (int i = tid; i < p.C * PR * PW; i += kDcThreads) {
      const int col = i % PW, r = (i / PW) % PR, c = i / (PW * PR);
      const int h = h_base + r, w = w_base + col;
      x_s[i] = (h >= 0 && h < p.H && w >= 0 && w < p.W) ? __ldg(xi + ((long long)c * p.H + h) * p.W + w) : 0.f;
    }
  }
  __syncthreads();

  // ---- thread tile: warp = (channel group, spatial half); lane = (row, pixel group) ----
  const int cg = warp & 3, half = warp >> 2;
  const int row = half * 4 + (lane >> 3), pxg = lane & 7;
  float acc[kDcPX][kDcCH];
#pragma unroll
  for (int q = 0; q < kDcPX; ++q)
#pragma unroll
    for (int j = 0; j < kDcCH; ++j) acc[q][j] = 0.f;

  for (int c = 0; c < p.C; ++c) {
#pragma unroll 1
    for (int a = 0; a < KH; ++a) {
      float xin[G::kXin];
      const float4* xr = reinterpret_cast<const float4*>(x_s + (c * PR + row * S + a) * PW + pxg * kDcPX * S);
#pragma unroll
      for (int v = 0; v < G::kXin / 4; ++v) {
        const float4 t = xr[v];
        xin[4 * v] = t.x;
        xin[4 * v + 1] = t.y;
        xin[4 * v + 2] = t.z;
        xin[4 * v + 3] = t.w;
      }
      const float4* wr = reinterpret_cast<const float4*>(w_s + ((c * KH + a) * KW) * kDcCoTile + cg * kDcCH);
#pragma unroll
      for (int b = 0; b < KW; ++b) {
        float wv[kDcCH];
#pragma unroll
        for (int v = 0; v < kDcCH / 4; ++v) {
          const float4 t = wr[b * (kDcCoTile / 4) + v];
          wv[4 * v] = t.x;
          wv[4 * v + 1] = t.y;
          wv[4 * v + 2] = t.z;
          wv[4 * v + 3] = t.w;
        }
#pragma unroll
        for (int q = 0; q < kDcPX; ++q)
#pragma unroll
          for (int j = 0; j < kDcCH; ++j) acc[q][j] = fmaf(xin[q * S + b], wv[j], acc[q][j]);
      }
    }
  }

  // ---- epilogue ----
  const int ho = ho0 + row, wo = wo0 + pxg * kDcPX;
  float rlo = INFINITY, rhi = -INFINITY;
  if (ho < p.Ho && wo < p.Wo) {
    const long long plane = (long long)p.Ho * p.Wo;
    const bool vec = (wo + kDcPX <= p.Wo) && ((p.Wo & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(p.y) & 15u) == 0) &&
                     (!p.y2 || (reinterpret_cast<uintptr_t>(p.y2) & 15u) == 0);
#pragma unroll
    for (int j = 0; j < kDcCH; ++j) {
      const int co = co0 + cg * kDcCH + j;
      if (co < p.c_out) {
        const float bj = p.bias ? __ldg(p.bias + co) : 0.f;
        const long long off = ((long long)img * p.c_out + co) * plane + (long long)ho * p.Wo + wo;
        float v[kDcPX];
#pragma unroll
        for (int q = 0; q < kDcPX; ++q) v[q] = acc[q][j] + bj;
        if (vec) {
          *reinterpret_cast<float4*>(p.y + off) = make_float4(v[0], v[1], v[2], v[3]);
          if (p.y2)
            *reinterpret_cast<float4*>(p.y2 + off) =
                make_float4(v[0] < 0.f ? 0.f : v[0], v[1] < 0.f ? 0.f : v[1], v[2] < 0.f ? 0.f : v[2],
                            v[3] < 0.f ? 0.f : v[3]);
#pragma unroll
          for (int q = 0; q < kDcPX; ++q) {
            rlo = fminf(rlo, v[q]);
            rhi = fmaxf(rhi, v[q]);
          }
        } else {
#pragma unroll
          for (int q = 0; q < kDcPX; ++q)
            if (wo + q < p.Wo) {
              p.y[off + q] = v[q];
              if (p.y2) p.y2[off + q] = v[q] < 0.f ? 0.f : v[q];
              rlo = fminf(rlo, v[q]);
              rhi = fmaxf(rhi, v[q]);
            }
        }
      }
    }
  }
  // fused range statistics (every lane takes part in the shuffles)
  if (p.bmin || p.bmax || p.rmin || p.rmax) {
    rlo = warp_min(rlo);
    rhi = warp_max(rhi);
    if (lane == 0 && rlo <= rhi) {
      if (p.bmin) atomic_min_f32(p.bmin, rlo + 0.f);
      if (p.bmax) atomic_max_f32(p.bmax, rhi + 0.f);
      if (p.rmin) atomic_min_f32(p.rmin, fmaxf(rlo, 0.f));
      if (p.rmax) atomic_max_f32(p.rmax, fmaxf(rhi, 0.f));
    }
  }
}

template <int KH, int KW, int S>
int launch_direct(const DcParams& p, cudaStream_t st) {
  using G = DcGeom<KW, S>;
  constexpr int PR = (kDcTH - 1) * S + KH;
  const size_t smem = ((size_t)p.C * PR * G::kPatchW + (size_t)p.C * KH * KW * kDcCoTile) * sizeof(float);
  if (smem > 100 * 1024) {
    set_error("dpl_conv_direct_f32: %zu bytes of shared memory needed (too many input channels)", smem);
    return DPL_E_UNSUPPORTED;
  }
  static bool attr_done = false;
  if (!attr_done) {
    int e = cuda_status(cudaFuncSetAttribute(conv_direct_kernel<KH, KW, S>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             100 * 1024),
                        "cudaFuncSetAttribute(conv_direct_kernel)");
    if (e) return e;
    attr_done = true;
  }
  dim3 grid((unsigned)((p.Wo + kDcTW - 1) / kDcTW), (unsigned)((p.Ho + kDcTH - 1) / kDcTH),
            (unsigned)(p.n_img * p.co_tiles));
  conv_direct_kernel<KH, KW, S><<<grid, kDcThreads, smem, st>>>(p);
  return cuda_status(cudaGetLastError(), "conv_direct_kernel");
}


// ---- depthwise convolution (MobileNetV2's 3x3 group = C layers) -------------------------------
// One multiply-add per filter tap and output: purely HBM bound, nothing for the tensor cores. One
// thread per output, a CTA inside one (image, channel) plane (filter taps and bias are CTA-uniform:
// registers), windows inside the image skip the bounds checks; fp32 FMA chain in tap order.
template <int K>
__global__ void __launch_bounds__(256)
dwconv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
              float* __restrict__ y, uint32_t tiles_per_plane, int C, int H, int W, int stride, int pad, int Ho,
              int Wo, float* __restrict__ bmin, float* __restrict__ bmax) {
  const uint32_t plane = blockIdx.x / tiles_per_plane;
  const uint32_t tile = blockIdx.x - plane * tiles_per_plane;
  const int c = (int)(plane % (uint32_t)C);
  float wt[K * K];
#pragma unroll
  for (int i = 0; i < K * K; ++i) wt[i] = __ldg(w + (long long)c * K * K + i);
  const float b0 = bias ? __ldg(bias + c) : 0.f;
  const uint32_t idx = tile * 256u + threadIdx.x;
  float rlo = INFINITY, rhi = -INFINITY;
  if (idx < (uint32_t)(Ho * Wo)) {
    const int ho = (int)(idx / (uint32_t)Wo), wo = (int)(idx - (uint32_t)ho * (uint32_t)Wo);
    const float* xp = x + (uint64_t)plane * (uint32_t)(H * W);
    const int h0 = ho * stride - pad, w0 = wo * stride - pad;
    float acc = 0.f;
    if (h0 >= 0 && w0 >= 0 && h0 + K <= H && w0 + K <= W) {
      const float* q = xp + h0 * W + w0;
#pragma unroll
      for (int a = 0; a < K; ++a)
#pragma unroll
        for (int b = 0; b < K; ++b) acc = fmaf(__ldg(q + a * W + b), wt[a * K + b], acc);
    } else {
#pragma unroll
      for (int a = 0; a < K; ++a) {
        const int h = h0 + a;
#pragma unroll
        for (int b = 0; b < K; ++b) {
          const int ww = w0 + b;
          const float v = (h >= 0 && h < H && ww >= 0 && ww < W) ? __ldg(xp + h * W + ww) : 0.f;
          acc = fmaf(v, wt[a * K + b], acc);
        }
      }
    }
    acc += b0;
    y[(uint64_t)plane * (uint32_t)(Ho * Wo) + idx] = acc;
    rlo = rhi = acc;
  }
  if (bmin || bmax) {   // every lane takes part
    rlo = warp_min(rlo);
    rhi = warp_max(rhi);
    if ((threadIdx.x & 31) == 0 && rlo <= rhi) {
      if (bmin) atomic_min_f32(bmin, rlo + 0.f);
      if (bmax) atomic_max_f32(bmax, rhi + 0.f);
    }
  }
}

}  // namespace
}  // namespace dpl

using namespace dpl;

// Y[img][co][ho][wo] = bias[co] + sum_{c,a,b} W[co][c][a][b] * X[img][c][ho * stride - pad + a][wo * stride - pad + b]
// for few input channels; (kh, kw, stride) in {(7, 7, 2), (3, 3, 2), (3, 3, 1), (5, 5, 2), (5, 5, 1)}, square
// symmetric padding, else DPL_E_UNSUPPORTED. d_y_relu / d_blob_* / d_relu_* as in dpl_gemm_tf32x3.
extern "C" int dpl_conv_direct_f32(const float* d_x, const float* d_w, const float* d_bias, float* d_y, int n_img,
                                   int channels, int H, int W, int c_out, int kh, int kw, int stride, int pad, int Ho,
                                   int Wo, float* d_y_relu, float* d_blob_min, float* d_blob_max, float* d_relu_min,
                                   float* d_relu_max, void* stream) {
  DPL_REQUIRE(d_x && d_w && d_y, "null pointer");
  DPL_REQUIRE(n_img > 0 && channels > 0 && H > 0 && W > 0 && c_out > 0 && Ho > 0 && Wo > 0 && pad >= 0, "bad geometry");
  DcParams p;
  p.x = d_x;
  p.w = d_w;
  p.bias = d_bias;
  p.y = d_y;
  p.y2 = d_y_relu;
  p.bmin = d_blob_min;
  p.bmax = d_blob_max;
  p.rmin = d_relu_min;
  p.rmax = d_relu_max;
  p.n_img = n_img;
  p.C = channels;
  p.H = H;
  p.W = W;
  p.c_out = c_out;
  p.pad = pad;
  p.Ho = Ho;
  p.Wo = Wo;
  p.co_tiles = (c_out + kDcCoTile - 1) / kDcCoTile;
  DPL_REQUIRE((long long)n_img * p.co_tiles <= 65535 && (Ho + kDcTH - 1) / kDcTH <= 65535, "grid limit");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (kh == 7 && kw == 7 && stride == 2) return launch_direct<7, 7, 2>(p, st);
  if (kh == 3 && kw == 3 && stride == 2) return launch_direct<3, 3, 2>(p, st);
  if (kh == 3 && kw == 3 && stride == 1) return launch_direct<3, 3, 1>(p, st);
  if (kh == 5 && kw == 5 && stride == 2) return launch_direct<5, 5, 2>(p, st);
  if (kh == 5 && kw == 5 && stride == 1) return launch_direct<5, 5, 1>(p, st);
  set_error("dpl_conv_direct_f32: kernel %dx%d stride %d is not instantiated", kh, kw, stride);
  return DPL_E_UNSUPPORTED;
}

// Depthwise convolution, group = channels = c_out, depth multiplier 1 (MobileNetV2's 3x3 layers):
//   Y[img][c][ho][wo] = bias[c] + sum_{a,b} W[c][0][a][b] * X[img][c][ho * stride - pad + a][wo * stride - pad + b]
// k in {3, 5}, symmetric padding; d_blob_min / d_blob_max: fused range statistics as in dpl_clip_f32.
extern "C" int dpl_dwconv2d_f32(const float* d_x, const float* d_w, const float* d_bias, float* d_y, int n_img,
                                int channels, int H, int W, int k, int stride, int pad, int Ho, int Wo,
                                float* d_blob_min, float* d_blob_max, void* stream) {
  DPL_REQUIRE(d_x && d_w && d_y, "null pointer");
  DPL_REQUIRE(n_img > 0 && channels > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0 && stride > 0 && pad >= 0, "bad geometry");
  DPL_REQUIRE((long long)H * W < (1ll << 31) && (long long)Ho * Wo < (1ll << 31), "plane too large");
  const uint64_t planes = (uint64_t)n_img * channels;
  const uint64_t tiles_per_plane = ((uint64_t)Ho * Wo + 255) / 256;
  DPL_REQUIRE(planes * tiles_per_plane < (1ull << 31), "too many tiles");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (k == 3)
    dwconv_kernel<3><<<(unsigned)(planes * tiles_per_plane), 256, 0, st>>>(
        d_x, d_w, d_bias, d_y, (uint32_t)tiles_per_plane, channels, H, W, stride, pad, Ho, Wo, d_blob_min, d_blob_max);
  else if (k == 5)
    dwconv_kernel<5><<<(unsigned)(planes * tiles_per_plane), 256, 0, st>>>(
        d_x, d_w, d_bias, d_y, (uint32_t)tiles_per_plane, channels, H, W, stride, pad, Ho, Wo, d_blob_min, d_blob_max);
  else {
    set_error("dpl_dwconv2d_f32: kernel size %d is not instantiated", k);
    return DPL_E_UNSUPPORTED;
  }
  DPL_LAUNCH_CHECK("dwconv_kernel");
  return 0;
}
