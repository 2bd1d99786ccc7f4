// K6 dense tile: TF32 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in
// TMEM, operands staged in shared memory by TMA) for the Conv-1x1 / Gemm re-evaluation of the
// AdaRound / BRECQ reconstruction loop — the one dense contraction on the path.
//
//   D[z][m][n] (+)= sum_k A[za][m][k] * B[z][k][n]   (+ bias, relu)        fp32 in, TF32 MMA, fp32 out
//
// Operands are described as 3-D tensors (inner, outer, batch) and can each be K-major
// (k contiguous) or MN-major (m / n contiguous), which covers, for NCHW activations,
//   conv1x1 forward   O[img][co][hw] = W[co][ci]      (K-major)  x X[img][ci][hw]      (MN-major)
//   conv1x1 wgrad     dW[co][ci]     = dO[img][co][hw] (K-major)  x X[img][ci][hw]^T    (K-major), batch folded into K
//   conv1x1 dgrad     dX[img][ci][hw]= W[co][ci]^T     (MN-major) x dO[img][co][hw]     (MN-major)
//   fc forward        Y[n][out]      = X[n][k]         (K-major)  x W[out][k]^T         (K-major)
// kind::tf32 reads fp32 bit patterns and uses the top 19 bits — numerically in family with
// the reference, whose torch conv runs with cudnn.allow_tf32 = True.
//
// One CTA = one 128 x 128 output tile, 4 warps: warp 0 / lane 0 issues TMA, warp 1 / lane 0
// issues the MMAs, warp 2 owns the TMEM allocation, all four warps run the epilogue (TMEM lane
// quarter = warp id). 3-stage shared-memory ring (32 KB per stage, two CTAs per SM) with full/empty mbarriers;
// tcgen05.commit releases a stage when the MMAs that read it have retired. Every mbarrier wait
// is bounded in time: a protocol error surfaces as DPL_E_TIMEOUT instead of a hung GPU.

#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include "dpl_common.cuh"
#include "dpl_tc.cuh"

namespace dpl {
namespace {

constexpr int kStages = 3;                           // 96 KB: two CTAs per SM overlap epilogue and main loop
constexpr int kStageBytes = 2 * kTileBytes;
constexpr int kTmemCols = 128;
constexpr int kGemmThreads = 128;

struct GemmParams {
  int M, N, K;
  int batch;            // batch slices of B (and of A when a_batched)
  int a_batched;        // 0: A shared by all slices (weights)
  int fold_batch;       // 1: the batch is an extension of K (wgrad); output has no batch dim
  int z_per_cta;        // fold_batch: slices handled by one CTA (split-K over gridDim.z)
  float* D;
  long long ldd, d_batch_stride;
  const float* bias;
  int bias_mode;        // 0 none, 1 per row (m), 2 per column (n)
  int relu;
  int atomic_out;       // 1: red.add into D (split-K partial sums)
  int hi_alt;           // 3xTF32: alternate the leading term between two accumulators
  int* error_flag;
  float* D2;            // optional second output max(D, 0) with D's layout (the Relu blob behind a Conv)
  float* bmin;          // optional fused range statistics of D (one float each, see dpl_clip_f32)
  float* bmax;
  float* rmin;          // ... and of D2
  float* rmax;
};

// Warp-level fold of a thread's range of written values into the running per-blob extrema.
// All 32 lanes must call it. NaN-free by contract; -0 canonicalised for the integer-ordered atomics.
__device__ __forceinline__ void warp_range_flush(float lo, float hi, float* bmin, float* bmax, float* rmin,
                                                 float* rmax) {
  if (!bmin && !bmax && !rmin && !rmax) return;
  lo = warp_min(lo);
  hi = warp_max(hi);
  if ((threadIdx.x & 31) == 0 && lo <= hi) {
    if (bmin) atomic_min_f32(bmin, lo + 0.f);
    if (bmax) atomic_max_f32(bmax, hi + 0.f);
    if (rmin) atomic_min_f32(rmin, fmaxf(lo, 0.f));
    if (rmax) atomic_max_f32(rmax, fmaxf(hi, 0.f));
  }
}

__device__ __forceinline__ float relu_keep_nan(float v) { return v < 0.f ? 0.f : v; }


template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kGemmThreads, 2)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_full[kStages], s_empty[kStages], s_tmem_full;
  __shared__ uint32_t s_tmem_base;
  __shared__ int s_fail;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 1024-byte aligned tile storage (the swizzle pattern is anchored to 1024-byte atoms)
  const uint32_t tiles = (smem_addr(smem_raw) + 1023u) & ~1023u;
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * kBN;
  const int num_kb = (p.K + kBK - 1) / kBK;
  int z_begin, z_count;
  if (p.fold_batch) {
    z_begin = blockIdx.z * p.z_per_cta;
    z_count = min(p.z_per_cta, p.batch - z_begin);
  } else {
    z_begin = blockIdx.z;
    z_count = 1;
  }
  const int total_iters = z_count * num_kb;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
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
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_acc = s_tmem_base;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    for (int it = 0; it < total_iters; ++it) {
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      if (!bar_wait(smem_addr(&s_empty[s]), ph ^ 1)) {
        s_fail = 1;
        break;
      }
      const uint32_t full = smem_addr(&s_full[s]);
      bar_expect_tx(full, kStageBytes);
      const int z = z_begin + it / num_kb;
      const int k0 = (it % num_kb) * kBK;
      const uint32_t a_tile = tiles + s * kStageBytes, b_tile = a_tile + kTileBytes;
      const int za = p.a_batched ? z : 0;
      if (A_MN) {
#pragma unroll
        for (int j = 0; j < kBM / 32; ++j) tma_load_3d(a_tile + j * (kBK * 128), &tmA, m0 + 32 * j, k0, za, full);
      } else {
        tma_load_3d(a_tile, &tmA, k0, m0, za, full);
      }
      if (B_MN) {
#pragma unroll
        for (int j = 0; j < kBN / 32; ++j) tma_load_3d(b_tile + j * (kBK * 128), &tmB, n0 + 32 * j, k0, z, full);
      } else {
        tma_load_3d(b_tile, &tmB, k0, n0, z, full);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, majors, N >> 3, M >> 4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) |
                           ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(kBN >> 3) << 17) |
                           ((uint32_t)(kBM >> 4) << 24);
    bool failed = false;
    for (int it = 0; it < total_iters; ++it) {
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      if (!bar_wait(smem_addr(&s_full[s]), ph)) {
        s_fail = 1;
        failed = true;
        break;
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_tile = tiles + s * kStageBytes, b_tile = a_tile + kTileBytes;
#pragma unroll
      for (int j = 0; j < kBK / kUmmaK; ++j) {
        const uint64_t da = A_MN ? desc_mn_major(a_tile, j) : desc_k_major(a_tile, j);
        const uint64_t db = B_MN ? desc_mn_major(b_tile, j) : desc_k_major(b_tile, j);
        const uint32_t accumulate = (it > 0 || j > 0) ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_acc), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
            : "memory");
      }
      // frees the stage once the MMAs above have finished reading shared memory
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

  // ===== epilogue: TMEM -> registers -> global (all four warps; lane quarter = warp) =====
  bool ok = true;
  if (total_iters > 0) ok = bar_wait(smem_addr(&s_tmem_full), 0);
  ok = __all_sync(0xffffffffu, ok);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (ok && total_iters > 0) {
    const int m = m0 + warp * 32 + lane;
    float* drow = p.D + (p.fold_batch ? 0 : (long long)blockIdx.z * p.d_batch_stride) + (long long)m * p.ldd;
    const float bias_m = (p.bias_mode == 1 && m < p.M) ? p.bias[m] : 0.f;
#pragma unroll 1
    for (int c = 0; c < kBN / 32; ++c) {
      uint32_t r[32];
      const uint32_t taddr = tmem_acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
            "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
            "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
            "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (m < p.M) {
        const int nc = n0 + c * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = __uint_as_float(r[j]) + bias_m;
          if (p.bias_mode == 2 && nc + j < p.N) v[j] += p.bias[nc + j];
          if (p.relu) v[j] = fmaxf(v[j], 0.f);
        }
        float* dst = drow + nc;
        if (!p.atomic_out && nc + 32 <= p.N && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
          // the thread owns 128 contiguous bytes of its output row: 8 x 16-byte stores
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (nc + j < p.N) {
              if (p.atomic_out)
                atomicAdd(dst + j, v[j]);
              else
                dst[j] = v[j];
            }
          }
        }
      }
    }
  }
  if (!ok || s_fail) {
    if (threadIdx.x == 0 && p.error_flag) atomicExch(p.error_flag, 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(kTmemCols)
                 : "memory");
  }
}

// ---- 3xTF32: fp32-accurate products on the TF32 tensor cores ----------------------------
// a * b = (a_hi + a_lo)(b_hi + b_lo) ~= a_hi b_hi + a_lo b_hi + a_hi b_lo   (a_lo b_lo ~ 2^-22 dropped)
// kind::tf32 truncates its operands, so feeding the raw fp32 pattern IS the hi part; the lo
// parts (x - trunc(x)) come from the host for the weights (A_lo, a second tensor map) and are
// computed in shared memory for the activations: four "transform" warps turn each landed B tile
// into a sibling B_lo tile (elementwise, so the swizzled layout is preserved), make it visible to
// the async proxy and hand the stage to the MMA thread, which issues three MMAs per K step.
// Used by the calibration forward, where activations must stay within fp32 rounding of the
// reference's CPU path (clip values are compared at 1e-5).
constexpr int kStages3 = 3;
constexpr int kStageBytes3 = 4 * kTileBytes;        // A_hi | A_lo | B | B_lo
#ifndef DPL_XFORM_WARPS
#define DPL_XFORM_WARPS 4   // 8 measured: no change (the transform is not the critical path)
#endif
constexpr int kXformWarps = DPL_XFORM_WARPS;        // transform warps of the one-tile-per-CTA 3xTF32 kernels
constexpr int kXformThreads = 32 * kXformWarps;
constexpr int kGemm3Threads = 128 + kXformThreads;
constexpr int kTmemCols3 = 512;                     // three 128-column accumulators (power of two)


__device__ __forceinline__ float tf32_residual(float x) {
  return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kGemm3Threads, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
                   const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_full[kStages3], s_ready[kStages3], s_empty[kStages3], s_tmem_full;
  __shared__ uint32_t s_tmem_base;
  __shared__ int s_fail;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tiles = (smem_addr(smem_raw) + 1023u) & ~1023u;
  uint8_t* tiles_ptr = smem_raw + (tiles - smem_addr(smem_raw));
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * kBN;
  const int num_kb = (p.K + kBK - 1) / kBK;
  const int z = blockIdx.z;
  const int total_iters = num_kb;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages3; ++s) {
      bar_init(smem_addr(&s_full[s]), 1);
      bar_init(smem_addr(&s_ready[s]), kXformWarps);   // one arrival per transform warp
      bar_init(smem_addr(&s_empty[s]), 1);
    }
    bar_init(smem_addr(&s_tmem_full), 1);
    s_fail = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_addr(&s_tmem_base)),
                 "r"(kTmemCols3)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_acc = s_tmem_base;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer: A (hi pattern), A_lo, B =====
    for (int it = 0; it < total_iters; ++it) {
      const int s = it % kStages3;
      const uint32_t ph = (it / kStages3) & 1;
      if (!bar_wait(smem_addr(&s_empty[s]), ph ^ 1)) {
        s_fail = 1;
        break;
      }
      const uint32_t full = smem_addr(&s_full[s]);
      bar_expect_tx(full, 3 * kTileBytes);
      const int k0 = it * kBK;
      const uint32_t a_tile = tiles + s * kStageBytes3, alo_tile = a_tile + kTileBytes,
                     b_tile = a_tile + 2 * kTileBytes;
      const int za = p.a_batched ? z : 0;
      if (A_MN) {
#pragma unroll
        for (int j = 0; j < kBM / 32; ++j) {
          tma_load_3d(a_tile + j * (kBK * 128), &tmA, m0 + 32 * j, k0, za, full);
          tma_load_3d(alo_tile + j * (kBK * 128), &tmAlo, m0 + 32 * j, k0, za, full);
        }
      } else {
        tma_load_3d(a_tile, &tmA, k0, m0, za, full);
        tma_load_3d(alo_tile, &tmAlo, k0, m0, za, full);
      }
      if (B_MN) {
#pragma unroll
        for (int j = 0; j < kBN / 32; ++j) tma_load_3d(b_tile + j * (kBK * 128), &tmB, n0 + 32 * j, k0, z, full);
      } else {
        tma_load_3d(b_tile, &tmB, k0, n0, z, full);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer: three MMAs per K step =====
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) |
                           ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(kBN >> 3) << 17) |
                           ((uint32_t)(kBM >> 4) << 24);
    bool failed = false;
    for (int it = 0; it < total_iters; ++it) {
      const int s = it % kStages3;
      const uint32_t ph = (it / kStages3) & 1;
      if (!bar_wait(smem_addr(&s_ready[s]), ph)) {
        s_fail = 1;
        failed = true;
        break;
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_tile = tiles + s * kStageBytes3, alo_tile = a_tile + kTileBytes,
                     b_tile = a_tile + 2 * kTileBytes, blo_tile = a_tile + 3 * kTileBytes;
#pragma unroll
      for (int j = 0; j < kBK / kUmmaK; ++j) {
        const uint64_t da = A_MN ? desc_mn_major(a_tile, j) : desc_k_major(a_tile, j);
        const uint64_t dal = A_MN ? desc_mn_major(alo_tile, j) : desc_k_major(alo_tile, j);
        const uint64_t db = B_MN ? desc_mn_major(b_tile, j) : desc_k_major(b_tile, j);
        const uint64_t dbl = B_MN ? desc_mn_major(blo_tile, j) : desc_k_major(blo_tile, j);
        // Three TMEM accumulators: the two small cross terms are summed apart from the leading
        // term (they would lose their low bits when aligned against a 2^11 larger partial sum),
        // and the leading term alternates between two accumulators to halve its chain length;
        // the epilogue adds the three in fp32 (round to nearest).
        const uint32_t acc_lo_first = (it > 0 || j > 0) ? 1u : 0u;
        const uint32_t acc_hi_first = (it > (p.hi_alt ? 1 : 0) || j > 0) ? 1u : 0u;
        const uint32_t acc_hi = tmem_acc + (uint32_t)((p.hi_alt ? (it & 1) : 0) * kBN);
        const uint32_t acc_lo = tmem_acc + 2u * kBN;
        asm volatile(
            "{\n\t.reg .pred p, q, t;\n\t"
            "setp.ne.b32 p, %7, 0;\n\t"
            "setp.ne.b32 q, %8, 0;\n\t"
            "setp.eq.b32 t, %6, %6;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%1], %3, %4, %6, p;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%1], %2, %5, %6, t;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %2, %4, %6, q;\n\t}"
            ::"r"(acc_hi), "r"(acc_lo), "l"(da), "l"(dal), "l"(db), "l"(dbl), "r"(idesc), "r"(acc_lo_first),
              "r"(acc_hi_first)
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
  } else if (warp >= 4) {
    // ===== transform warps: B_lo = B - trunc_tf32(B), same offsets (layout preserved) =====
    const int t = threadIdx.x - 128;
    for (int it = 0; it < total_iters; ++it) {
      const int s = it % kStages3;
      const uint32_t ph = (it / kStages3) & 1;
      if (!bar_wait(smem_addr(&s_full[s]), ph)) {
        s_fail = 1;
        break;
      }
      const float4* src = reinterpret_cast<const float4*>(tiles_ptr + s * kStageBytes3 + 2 * kTileBytes);
      float4* dst = reinterpret_cast<float4*>(tiles_ptr + s * kStageBytes3 + 3 * kTileBytes);
#pragma unroll
      for (int j = 0; j < kTileBytes / 16 / kXformThreads; ++j) {
        const float4 v = src[t + j * kXformThreads];
        dst[t + j * kXformThreads] = make_float4(tf32_residual(v.x), tf32_residual(v.y), tf32_residual(v.z),
                                       tf32_residual(v.w));
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> async proxy (MMA)
      __syncwarp();
      if (lane == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&s_ready[s])) : "memory");
    }
  }
  __syncwarp();

  // ===== epilogue: warps 0-3 (TMEM lane quarter = warp) =====
  bool ok = true;
  if (warp < 4) {
    if (total_iters > 0) ok = bar_wait(smem_addr(&s_tmem_full), 0);
    ok = __all_sync(0xffffffffu, ok);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (ok && total_iters > 0) {
      const int m = m0 + warp * 32 + lane;
      float* drow = p.D + (long long)z * p.d_batch_stride + (long long)m * p.ldd;
      const float bias_m = (p.bias_mode == 1 && m < p.M) ? p.bias[m] : 0.f;
      float rlo = INFINITY, rhi = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < kBN / 32; ++c) {
        uint32_t r[32], r1[32], r2[32];
        const uint32_t taddr = tmem_acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32);
        tmem_ld32(taddr, r);
        tmem_ld32(taddr + 2u * kBN, r2);
        if (total_iters > 1 && p.hi_alt) tmem_ld32(taddr + kBN, r1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (m < p.M) {
          const int nc = n0 + c * 32;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float acc = __uint_as_float(r[j]);
            if (total_iters > 1 && p.hi_alt) acc += __uint_as_float(r1[j]);
            v[j] = (acc + __uint_as_float(r2[j])) + bias_m;
            if (p.bias_mode == 2 && nc + j < p.N) v[j] += p.bias[nc + j];
            if (p.relu) v[j] = fmaxf(v[j], 0.f);
            if (nc + j < p.N) {
              rlo = fminf(rlo, v[j]);
              rhi = fmaxf(rhi, v[j]);
            }
          }
          float* dst = drow + nc;
          float* dst2 = p.D2 ? p.D2 + (drow - p.D) + nc : nullptr;
          if (nc + 32 <= p.N && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) &&
              ((reinterpret_cast<uintptr_t>(dst2) & 15u) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            if (dst2) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(dst2 + j) = make_float4(relu_keep_nan(v[j]), relu_keep_nan(v[j + 1]),
                                                                   relu_keep_nan(v[j + 2]), relu_keep_nan(v[j + 3]));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nc + j < p.N) {
                dst[j] = v[j];
                if (dst2) dst2[j] = relu_keep_nan(v[j]);
              }
          }
        }
      }
      warp_range_flush(rlo, rhi, p.bmin, p.bmax, p.rmin, p.rmax);
    }
  }
  if (!ok || s_fail) {
    if ((threadIdx.x & 31) == 0 && p.error_flag) atomicExch(p.error_flag, 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(kTmemCols3)
                 : "memory");
  }
}

// ---- persistent 3xTF32 GEMM for the short-K 1x1 convolutions ------------------------------
// The expanding / reducing 1x1 convolutions of a bottleneck have K = C_in of 64 .. 512, i.e. 2 .. 16
// K blocks per 128 x 128 output tile: a one-tile-per-CTA kernel spends its time in the prologue
// (barrier init, TMEM allocation, first TMA round trip) and in the 64 KB epilogue. Here one CTA
// per SM walks a static tile list; the TMA ring runs across tile boundaries, the accumulators are
// double-buffered in TMEM (2 x {hi, lo} x 128 columns = all 512) and four dedicated warps drain
// tile t while the MMA thread accumulates tile t + 1.
// A = W (K-major, rows = output channels), B = X (MN-major, pixels contiguous), batch = image.
// Leading term in ONE accumulator: the tensor core accumulates with truncation, which costs
// ~3e-9 relative per K step (measured, tools/x3_accuracy.py) - below fp32 rounding for K <= 512.
constexpr int kGemmPThreads = 384;
constexpr int kStgPitch = 36;                                  // floats per staged row: 16-byte aligned, conflict free
constexpr int kStgBytes = 4 * 32 * kStgPitch * 4;              // one 32 x 32 staging tile per epilogue warp

__global__ void __launch_bounds__(kGemmPThreads, 1)
gemm_tf32x3_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
                              const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_full[kStages3], s_ready[kStages3], s_empty[kStages3], s_acc_full[2],
      s_acc_empty[2];
  __shared__ uint32_t s_tmem_base;
  __shared__ int s_fail;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tiles = (smem_addr(smem_raw) + 1023u) & ~1023u;
  uint8_t* tiles_ptr = smem_raw + (tiles - smem_addr(smem_raw));
  const int num_kb = (p.K + kBK - 1) / kBK;
  const int m_tiles = (p.M + kBM - 1) / kBM, n_tiles = (p.N + kBN - 1) / kBN;
  const int total_tiles = m_tiles * n_tiles * p.batch;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages3; ++s) {
      bar_init(smem_addr(&s_full[s]), 1);
      bar_init(smem_addr(&s_ready[s]), 4);
      bar_init(smem_addr(&s_empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      bar_init(smem_addr(&s_acc_full[b]), 1);
      bar_init(smem_addr(&s_acc_empty[b]), 4);   // one arrival per epilogue warp
    }
    s_fail = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_addr(&s_tmem_base)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_acc = s_tmem_base;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles && !s_fail; tile += gridDim.x) {
      const int mt = tile % m_tiles, nt = (tile / m_tiles) % n_tiles, z = tile / (m_tiles * n_tiles);
      const int m0 = mt * kBM, n0 = nt * kBN;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % kStages3;
        const uint32_t ph = (it / kStages3) & 1;
        if (!bar_wait(smem_addr(&s_empty[s]), ph ^ 1)) {
          s_fail = 1;
          break;
        }
        const uint32_t full = smem_addr(&s_full[s]);
        bar_expect_tx(full, 3 * kTileBytes);
        const int k0 = kb * kBK;
        const uint32_t a_tile = tiles + s * kStageBytes3, alo_tile = a_tile + kTileBytes,
                       b_tile = a_tile + 2 * kTileBytes;
        tma_load_3d(a_tile, &tmA, k0, m0, 0, full);
        tma_load_3d(alo_tile, &tmAlo, k0, m0, 0, full);
#pragma unroll
        for (int j = 0; j < kBN / 32; ++j) tma_load_3d(b_tile + j * (kBK * 128), &tmB, n0 + 32 * j, k0, z, full);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(kBN >> 3) << 17) |
                           ((uint32_t)(kBM >> 4) << 24);      // A K-major, B MN-major
    int it = 0, t = 0;
    for (int tile = blockIdx.x; tile < total_tiles && !s_fail; tile += gridDim.x, ++t) {
      const int buf = t & 1;
      if (!bar_wait(smem_addr(&s_acc_empty[buf]), ((t >> 1) & 1) ^ 1)) {
        s_fail = 1;
        break;
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t acc_hi = tmem_acc + (uint32_t)(buf * 2 * kBN), acc_lo = acc_hi + kBN;
      bool failed = false;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % kStages3;
        const uint32_t ph = (it / kStages3) & 1;
        if (!bar_wait(smem_addr(&s_ready[s]), ph)) {
          s_fail = 1;
          failed = true;
          break;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_tile = tiles + s * kStageBytes3, alo_tile = a_tile + kTileBytes,
                       b_tile = a_tile + 2 * kTileBytes, blo_tile = a_tile + 3 * kTileBytes;
#pragma unroll
        for (int j = 0; j < kBK / kUmmaK; ++j) {
          const uint64_t da = desc_k_major(a_tile, j), dal = desc_k_major(alo_tile, j);
          const uint64_t db = desc_mn_major(b_tile, j), dbl = desc_mn_major(blo_tile, j);
          const uint32_t accumulate = (kb > 0 || j > 0) ? 1u : 0u;
          asm volatile(
              "{\n\t.reg .pred p, t;\n\t"
              "setp.ne.b32 p, %7, 0;\n\t"
              "setp.eq.b32 t, %6, %6;\n\t"
              "tcgen05.mma.cta_group::1.kind::tf32 [%1], %3, %4, %6, p;\n\t"
              "tcgen05.mma.cta_group::1.kind::tf32 [%1], %2, %5, %6, t;\n\t"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], %2, %4, %6, p;\n\t}"
              ::"r"(acc_hi), "r"(acc_lo), "l"(da), "l"(dal), "l"(db), "l"(dbl), "r"(idesc), "r"(accumulate)
              : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         smem_addr(&s_empty[s]))
                     : "memory");
      }
      if (failed) break;
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_addr(&s_acc_full[buf]))
                   : "memory");
    }
  } else if (warp >= 4 && warp < 8) {
    // ===== transform warps: B_lo = B - trunc_tf32(B) =====
    const int tt = threadIdx.x - 128;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles && !s_fail; tile += gridDim.x) {
      bool failed = false;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % kStages3;
        const uint32_t ph = (it / kStages3) & 1;
        if (!bar_wait(smem_addr(&s_full[s]), ph)) {
          s_fail = 1;
          failed = true;
          break;
        }
        const float4* src = reinterpret_cast<const float4*>(tiles_ptr + s * kStageBytes3 + 2 * kTileBytes);
        float4* dst = reinterpret_cast<float4*>(tiles_ptr + s * kStageBytes3 + 3 * kTileBytes);
#pragma unroll
        for (int j = 0; j < kTileBytes / 16 / 128; ++j) {
          const float4 v = src[tt + j * 128];
          dst[tt + j * 128] = make_float4(tf32_residual(v.x), tf32_residual(v.y), tf32_residual(v.z),
                                          tf32_residual(v.w));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0)
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&s_ready[s])) : "memory");
      }
      if (failed) break;
    }
  } else if (warp >= 8) {
    // ===== epilogue warps (TMEM lane quarter = warp - 8) =====
    const int wq = warp - 8;
    int t = 0;
    float rlo = INFINITY, rhi = -INFINITY;   // range of everything this thread stores, flushed once
    float* stg = reinterpret_cast<float*>(tiles_ptr + kStages3 * kStageBytes3) + wq * (32 * kStgPitch);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++t) {
      const int mt = tile % m_tiles, nt = (tile / m_tiles) % n_tiles, z = tile / (m_tiles * n_tiles);
      const int m0 = mt * kBM, n0 = nt * kBN;
      const int buf = t & 1;
      bool ok = bar_wait(smem_addr(&s_acc_full[buf]), (t >> 1) & 1);
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) {
        s_fail = 1;
        break;
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int m = m0 + wq * 32 + lane;
      float* drow = p.D + (long long)z * p.d_batch_stride + (long long)m * p.ldd;
      const float bias_m = (p.bias_mode == 1 && m < p.M) ? p.bias[m] : 0.f;
#pragma unroll 1
      for (int c = 0; c < kBN / 32; ++c) {
        uint32_t r[32], r2[32];
        const uint32_t taddr = tmem_acc + ((uint32_t)(wq * 32) << 16) + (uint32_t)(buf * 2 * kBN + c * 32);
        tmem_ld32(taddr, r);
        tmem_ld32(taddr + kBN, r2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c == kBN / 32 - 1) {
          // all TMEM reads of this buffer are done: hand it back before the global stores
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0)
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&s_acc_empty[buf]))
                         : "memory");
        }
        const int nc = n0 + c * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = (__uint_as_float(r[j]) + __uint_as_float(r2[j])) + bias_m;
          if (p.relu) v[j] = fmaxf(v[j], 0.f);
          if (m < p.M && nc + j < p.N) {
            rlo = fminf(rlo, v[j]);
            rhi = fmaxf(rhi, v[j]);
          }
        }
        // A thread holds 32 consecutive pixels of ONE output row, so a direct 16-byte store touches 32
        // rows x 16 bytes per instruction (half sectors). For full chunks the 32 x 32 block goes through a
        // padded shared-memory tile and is written as 4 rows x 128 contiguous bytes per instruction.
        float* blk = p.D + (long long)z * p.d_batch_stride + (long long)(m0 + wq * 32) * p.ldd + nc;
        float* blk2 = p.D2 ? p.D2 + (blk - p.D) : nullptr;
        const bool full = (m0 + wq * 32 + 32 <= p.M) && (nc + 32 <= p.N) && ((p.ldd & 3) == 0) &&
                          ((reinterpret_cast<uintptr_t>(blk) & 15u) == 0) &&
                          ((reinterpret_cast<uintptr_t>(blk2) & 15u) == 0);
        if (full) {
          __syncwarp();   // the previous chunk's reads of the staging tile are done
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(stg + lane * kStgPitch + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          __syncwarp();
          const int rr = lane >> 3, cc = (lane & 7) * 4;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int row = 4 * q + rr;
            const float4 t = *reinterpret_cast<const float4*>(stg + row * kStgPitch + cc);
            *reinterpret_cast<float4*>(blk + (long long)row * p.ldd + cc) = t;
            if (blk2)
              *reinterpret_cast<float4*>(blk2 + (long long)row * p.ldd + cc) =
                  make_float4(relu_keep_nan(t.x), relu_keep_nan(t.y), relu_keep_nan(t.z), relu_keep_nan(t.w));
          }
        } else if (m < p.M) {
          float* dst = drow + nc;
          float* dst2 = p.D2 ? p.D2 + (drow - p.D) + nc : nullptr;
          if (nc + 32 <= p.N && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) &&
              ((reinterpret_cast<uintptr_t>(dst2) & 15u) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            if (dst2) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(dst2 + j) = make_float4(relu_keep_nan(v[j]), relu_keep_nan(v[j + 1]),
                                                                   relu_keep_nan(v[j + 2]), relu_keep_nan(v[j + 3]));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nc + j < p.N) {
                dst[j] = v[j];
                if (dst2) dst2[j] = relu_keep_nan(v[j]);
              }
          }
        }
      }
    }
    warp_range_flush(rlo, rhi, p.bmin, p.bmax, p.rmin, p.rmax);
  }
  __syncwarp();
  if (s_fail) {
    if ((threadIdx.x & 31) == 0 && p.error_flag) atomicExch(p.error_flag, 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(512) : "memory");
  }
}

__device__ __forceinline__ float tf32_rn_host_side(float x) {   // round to nearest (even) TF32
  uint32_t u = __float_as_uint(x);
  u += 0xFFFu + ((u >> 13) & 1u);
  return __uint_as_float(u & 0xFFFFE000u);
}

// residual of the truncated pattern (what kind::tf32 reads from raw fp32), rounded to TF32 so that the tensor
// core reads it exactly
__global__ void __launch_bounds__(256)
tf32_residual_kernel(const float* __restrict__ x, float* __restrict__ lo, uint64_t n) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    lo[i] = tf32_rn_host_side(tf32_residual(x[i]));
}

// unbiased split for weights: hi = RN_tf32(x), lo = RN_tf32(x - hi)
__global__ void __launch_bounds__(256)
tf32_split_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, uint64_t n) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = x[i], h = tf32_rn_host_side(v);
    hi[i] = h;
    lo[i] = tf32_rn_host_side(v - h);
  }
}

// ---- 3x3 convolution (stride 1, pad 1) as a shifted-window implicit GEMM, 3xTF32 ----------
// The calibration forward's k x k convolutions (forward_net.py:200-216 run them through ORT).
// The input is first copied into a zero-bordered, channel-last layout
//     Xp[q][ci],   q = img * Pp + (h + 1)(W + 2) + (w + 1),   Pp = (H + 2)(W + 2)
// in which tap (kh, kw) of the filter is the SAME matrix shifted by (kh - 1)(W + 2) + (kw - 1)
// ROWS. So the convolution is 9 * C_in / 32 accumulating GEMM steps
//     D[q][co] += Xp[q + shift(tap)][ci-block] (K-major A: one TMA box at a shifted row coordinate,
//                 rows outside the tensor zero-filled)  x  Wt[tap][co][ci-block]^T (K-major B)
// (channel-last because a swizzled TMA box may start at any row but only at 16-byte multiples
// along the contiguous dimension — measured, tools/probe/tma_probe.cu — and the taps shift by
// single pixels). Pixels sit on the TMEM lanes: the epilogue's lanes are consecutive pixels, so
// every store of an output channel is a coalesced row segment of the unpadded NCHW output;
// border positions of the padded plane are computed and dropped (7 % extra at 56 x 56).
// fp32 accuracy as gemm_tf32x3_kernel: A_lo formed in shared memory, B_lo = host residual.
// Generalised to a tap table so that the same kernel runs
//   3x3 stride 2  : the input is split into its four (row, column) parity planes, each with a
//                   one-pixel top/left zero border; tap (kh, kw) reads plane ((kh+1)&1, (kw+1)&1)
//                   at a shift of {-1, 0} rows / columns — no wasted outputs;
//   1x1 stride 2  : one tap, plane (0, 0) without border (the gather is the "padding" copy).
struct ConvParams {
  int n_img, c_in, c_out, H, W, Wp;   // H, W: OUTPUT size; Wp: padded plane width
  int plane;            // Pp = Hp * Wp
  int origin;           // 1: one-pixel top/left border in the padded plane, 0: none
  int n_taps;
  int tap_shift[9];     // row offset of every tap in Xp (plane base + window shift)
  int tap_w[9];         // x3p MODE 1 only: weight slice of every tap (identity for the forward kernels)
  int bn;               // output channels per CTA (64 or 128)
  int hi_alt;           // alternate the leading term between two accumulators
  long long q_total;    // n_img * Pp
  float* Y;             // [n_img][c_out][H][W]
  const float* bias;    // per output channel or null
  int relu;
  int* error_flag;
  float* Y2;            // optional second output max(Y, 0) (the Relu blob behind the Conv)
  float *bmin, *bmax, *rmin, *rmax;   // optional fused range statistics of Y / Y2
  int os, oa, ob;       // x3p MODE 1 only: plane point (hq, wq) is output pixel (hq * os + oa, wq * os + ob); 1, 0, 0 = dense
};

__global__ void __launch_bounds__(kGemm3Threads, 1)
conv_taps_tf32x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                      const __grid_constant__ CUtensorMap tmWlo, const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_full[kStages3], s_ready[kStages3], s_empty[kStages3], s_tmem_full;
  __shared__ uint32_t s_tmem_base;
  __shared__ int s_fail;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tiles = (smem_addr(smem_raw) + 1023u) & ~1023u;
  uint8_t* tiles_ptr = smem_raw + (tiles - smem_addr(smem_raw));
  const long long q0 = (long long)blockIdx.x * kBM;
  const int co0 = blockIdx.y * p.bn;
  const int num_kb = (p.c_in + kBK - 1) / kBK;
  const int total_iters = p.n_taps * num_kb;
  const uint32_t w_tile_bytes = (uint32_t)p.bn * 128u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages3; ++s) {
      bar_init(smem_addr(&s_full[s]), 1);
      bar_init(smem_addr(&s_ready[s]), kXformWarps);
      bar_init(smem_addr(&s_empty[s]), 1);
    }
    bar_init(smem_addr(&s_tmem_full), 1);
    s_fail = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_addr(&s_tmem_base)),
                 "r"(kTmemCols3)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_acc = s_tmem_base;

  // stage layout: X tile | X_lo (computed) | W tile | W_lo tile
  if (warp == 0 && lane == 0) {
    for (int it = 0; it < total_iters; ++it) {
      const int s = it % kStages3;
      const uint32_t ph = (it / kStages3) & 1;
      if (!bar_wait(smem_addr(&s_empty[s]), ph ^ 1)) {
        s_fail = 1;
        break;
      }
      const uint32_t full = smem_addr(&s_full[s]);
      bar_expect_tx(full, kTileBytes + 2 * w_tile_bytes);
      const int kb = it / p.n_taps, tap = it - kb * p.n_taps;
      const int k0 = kb * kBK;
      const int shift = p.tap_shift[tap];
      const uint32_t x_tile = tiles + s * kStageBytes3, w_tile = x_tile + 2 * kTileBytes,
                     wlo_tile = x_tile + 3 * kTileBytes;
      tma_load_3d(x_tile, &tmX, k0, (int)q0 + shift, 0, full);
      tma_load_3d(w_tile, &tmW, k0, co0, tap, full);
      tma_load_3d(wlo_tile, &tmWlo, k0, co0, tap, full);
    }
  } else if (warp == 1 && lane == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.bn >> 3) << 17) |
                           ((uint32_t)(kBM >> 4) << 24);      // A and B both K-major
    bool failed = false;
    for (int it = 0; it < total_iters; ++it) {
      const int s = it % kStages3;
      const uint32_t ph = (it / kStages3) & 1;
      if (!bar_wait(smem_addr(&s_ready[s]), ph)) {
        s_fail = 1;
        failed = true;
        break;
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t x_tile = tiles + s * kStageBytes3, xlo_tile = x_tile + kTileBytes,
                     w_tile = x_tile + 2 * kTileBytes, wlo_tile = x_tile + 3 * kTileBytes;
#pragma unroll
      for (int j = 0; j < kBK / kUmmaK; ++j) {
        const uint64_t da = desc_k_major(x_tile, j), dal = desc_k_major(xlo_tile, j);
        const uint64_t db = desc_k_major(w_tile, j), dbl = desc_k_major(wlo_tile, j);
        const uint32_t acc_lo_first = (it > 0 || j > 0) ? 1u : 0u;
        const uint32_t acc_hi_first = (it > (p.hi_alt ? 1 : 0) || j > 0) ? 1u : 0u;
        const uint32_t acc_hi = tmem_acc + (uint32_t)((p.hi_alt ? (it & 1) : 0) * kBN);
        const uint32_t acc_lo = tmem_acc + 2u * kBN;
        asm volatile(
            "{\n\t.reg .pred p, q, t;\n\t"
            "setp.ne.b32 p, %7, 0;\n\t"
            "setp.ne.b32 q, %8, 0;\n\t"
            "setp.eq.b32 t, %6, %6;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%1], %3, %4, %6, p;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%1], %2, %5, %6, t;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %2, %4, %6, q;\n\t}"
            ::"r"(acc_hi), "r"(acc_lo), "l"(da), "l"(dal), "l"(db), "l"(dbl), "r"(idesc), "r"(acc_lo_first),
              "r"(acc_hi_first)
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
  } else if (warp >= 4) {
    const int t = threadIdx.x - 128;
    for (int it = 0; it < total_iters; ++it) {
      const int s = it % kStages3;
      const uint32_t ph = (it / kStages3) & 1;
      if (!bar_wait(smem_addr(&s_full[s]), ph)) {
        s_fail = 1;
        break;
      }
      const float4* src = reinterpret_cast<const float4*>(tiles_ptr + s * kStageBytes3);
      float4* dst = reinterpret_cast<float4*>(tiles_ptr + s * kStageBytes3 + kTileBytes);
#pragma unroll
      for (int j = 0; j < kTileBytes / 16 / kXformThreads; ++j) {
        const float4 v = src[t + j * kXformThreads];
        dst[t + j * kXformThreads] = make_float4(tf32_residual(v.x), tf32_residual(v.y), tf32_residual(v.z),
                                       tf32_residual(v.w));
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&s_ready[s])) : "memory");
    }
  }
  __syncwarp();

  // ===== epilogue: TMEM lane = padded pixel q, columns = output channels =====
  bool ok = true;
  if (warp < 4) {
    ok = bar_wait(smem_addr(&s_tmem_full), 0);
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
        const int ho = hp - p.origin, wo = wp - p.origin;
        valid = ho >= 0 && ho < p.H && wo >= 0 && wo < p.W;
        out_base = (((long long)img * p.c_out) * p.H + ho) * p.W + wo;
      }
      const long long ch_stride = (long long)p.H * p.W;
      float rlo = INFINITY, rhi = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < p.bn / 32; ++c) {
        uint32_t r[32], r1[32], r2[32];
        const uint32_t taddr = tmem_acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32);
        tmem_ld32(taddr, r);
        if (total_iters > 1 && p.hi_alt) tmem_ld32(taddr + kBN, r1);
        tmem_ld32(taddr + 2u * kBN, r2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (valid) {
          const int cb = co0 + c * 32;
          float* dst = p.Y + out_base + (long long)cb * ch_stride;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (cb + j < p.c_out) {
              float v = __uint_as_float(r[j]);
              if (total_iters > 1 && p.hi_alt) v += __uint_as_float(r1[j]);
              v += __uint_as_float(r2[j]);
              if (p.bias) v += __ldg(p.bias + cb + j);
              if (p.relu) v = fmaxf(v, 0.f);
              dst[(long long)j * ch_stride] = v;
              if (p.Y2) p.Y2[(dst - p.Y) + (long long)j * ch_stride] = relu_keep_nan(v);
              rlo = fminf(rlo, v);
              rhi = fmaxf(rhi, v);
            }
          }
        }
      }
      warp_range_flush(rlo, rhi, p.bmin, p.bmax, p.rmin, p.rmax);
    }
  }
  if (!ok || s_fail) {
    if ((threadIdx.x & 31) == 0 && p.error_flag) atomicExch(p.error_flag, 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(kTmemCols3)
                 : "memory");
  }
}

// X[img][c][H][W] -> Xp[(plane * n_img + img) * Pp + r][c] (channel-last), 32 x 32 tile transpose.
// Plane (a, b) of a stride-s split holds X(s (hp - origin) + a, s (wp - origin) + b), zero outside.
__global__ void __launch_bounds__(256)
pad_plane_kernel(const float* __restrict__ x, float* __restrict__ xp, int n_img, int C, int H, int W, int stride,
                 int origin, int Hp, int Wp) {
  __shared__ float tile[32][33];
  const int plane = Hp * Wp;
  const int img = blockIdx.z % n_img, ab = blockIdx.z / n_img;
  const int a = ab / stride, b = ab - a * stride;
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int r = r0 + tx;
  int off = -1;
  if (r < plane) {
    const int hp = r / Wp, wp = r - hp * Wp;
    const int h = stride * (hp - origin) + a, w = stride * (wp - origin) + b;
    if (hp >= origin && wp >= origin && h < H && w < W) off = h * W + w;
  }
  const float* src = x + (long long)img * C * H * W;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k;
    tile[ty + 8 * k][tx] = (off >= 0 && c < C) ? src[(long long)c * H * W + off] : 0.f;
  }
  __syncthreads();
  float* dst = xp + ((long long)blockIdx.z * plane + r0) * C + c0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int rr = ty + 8 * k;
    if (r0 + rr < plane && c0 + tx < C) dst[(long long)rr * C + tx] = tile[tx][rr];
  }
}


// ---- pixel-major 1x1 convolution, 3xTF32, activations through TMEM ------------------------
// Y[img][co][px] = bias[co] + sum_ci W[co][ci] * X[img][ci][px], straight from / to NCHW.
// Orientation: the PIXELS sit on the M side (TMEM lanes), the output channels on N:
//     D[px][co] += A[px][ci] * B[co][ci]^T
//  * A = activations is the TMEM operand of tcgen05.mma ("TS" form): four transform warps read the
//    landed X tile [32 ci][128 px] from shared memory (thread <-> pixel, consecutive lanes read
//    consecutive words: conflict free, no swizzle needed because the tensor core never reads this
//    tile), split it into the TF32 pattern and its residual in registers and tcgen05.st both
//    into TMEM. Against the shared-memory form this removes the X_lo write, the three re-reads of
//    the X tiles by the MMAs and the generic->async proxy fence: shared-memory traffic per K block
//    drops from 176 KB to 72 KB (the 128 x 128 x 8 TF32 instruction with both operands in shared
//    memory saturates the 128 B/clk port on its own).
//  * B = weights, K-major, 128-byte swizzle, by TMA together with their host-side residual.
//  * Epilogue lanes are consecutive pixels: every store of an output channel is one coalesced
//    128-byte row segment of the NCHW output (the co-on-lanes orientation wrote 16 bytes to each
//    of 32 rows per instruction).
//  * 128 px x 64 co tiles, 256 TMEM columns (2 x {A_hi, A_lo} x 32 + {acc_hi, acc_lo} x 64) and
//    97 KB of shared memory per CTA: two CTAs per SM, so one tile's epilogue overlaps the other's
//    main loop without a persistent scheduler.
constexpr int kPxBN = 64;
constexpr int kPxStages = 3;
constexpr int kPxXBytes = kBK * kBM * 4;                  // 16 KB, [32 ci][128 px], unswizzled
constexpr int kPxWBytes = kPxBN * kBK * 4;                // 8 KB, [64 co][32 ci], SW128
constexpr int kPxStageBytes = kPxXBytes + 2 * kPxWBytes;  // 32 KB
constexpr int kPxThreads = 256;
constexpr int kPxTmemCols = 256;

struct PxParams {
  int n_img, c_in, c_out, hw;
  float* Y;
  float* Y2;            // optional max(Y, 0)
  const float* bias;
  int* error_flag;
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(kPxThreads, 2)
conv1x1_px_tf32x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                         const __grid_constant__ CUtensorMap tmWlo, const PxParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_full[kPxStages], s_empty[kPxStages], s_aready[2], s_aempty[2], s_acc_full;
  __shared__ uint32_t s_tmem_base;
  __shared__ int s_fail;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tiles = (smem_addr(smem_raw) + 1023u) & ~1023u;
  const uint8_t* tiles_ptr = smem_raw + (tiles - smem_addr(smem_raw));
  const int co0 = blockIdx.x * kPxBN, px0 = blockIdx.y * kBM, img = blockIdx.z;
  const int total_iters = (p.c_in + kBK - 1) / kBK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kPxStages; ++s) {
      bar_init(smem_addr(&s_full[s]), 1);
      bar_init(smem_addr(&s_empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      bar_init(smem_addr(&s_aready[a]), 4);   // one arrival per transform warp
      bar_init(smem_addr(&s_aempty[a]), 1);
    }
    bar_init(smem_addr(&s_acc_full), 1);
    s_fail = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_addr(&s_tmem_base)),
                 "r"(kPxTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem_base;
  const uint32_t acc_hi = tmem + 128u, acc_lo = tmem + 192u;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer: X tile (unswizzled), W and W_lo tiles (K-major, SW128) =====
    for (int it = 0; it < total_iters; ++it) {
      const int s = it % kPxStages;
      const uint32_t ph = (it / kPxStages) & 1;
      if (!bar_wait(smem_addr(&s_empty[s]), ph ^ 1)) {
        s_fail = 1;
        break;
      }
      const uint32_t full = smem_addr(&s_full[s]);
      bar_expect_tx(full, kPxStageBytes);
      const int k0 = it * kBK;
      const uint32_t x_tile = tiles + s * kPxStageBytes, w_tile = x_tile + kPxXBytes, wlo_tile = w_tile + kPxWBytes;
      tma_load_3d(x_tile, &tmX, px0, k0, img, full);
      tma_load_3d(w_tile, &tmW, k0, co0, 0, full);
      tma_load_3d(wlo_tile, &tmWlo, k0, co0, 0, full);
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer: A from TMEM, three MMAs per K step =====
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kPxBN >> 3) << 17) |
                           ((uint32_t)(kBM >> 4) << 24);      // f32 accumulate, tf32 x tf32, both K-major
    bool failed = false;
    for (int it = 0; it < total_iters; ++it) {
      const int s = it % kPxStages, a = it & 1;
      if (!bar_wait(smem_addr(&s_aready[a]), (it >> 1) & 1)) {
        s_fail = 1;
        failed = true;
        break;
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t w_tile = tiles + s * kPxStageBytes + kPxXBytes, wlo_tile = w_tile + kPxWBytes;
      const uint32_t a_hi = tmem + (uint32_t)(a * 64), a_lo = a_hi + 32u;
#pragma unroll
      for (int j = 0; j < kBK / kUmmaK; ++j) {
        const uint64_t db = desc_k_major(w_tile, j), dbl = desc_k_major(wlo_tile, j);
        const uint32_t accumulate = (it > 0 || j > 0) ? 1u : 0u;
        // the two small cross terms are summed apart from the leading term
        mma_tf32_ts(acc_lo, a_lo + (uint32_t)(j * kUmmaK), db, idesc, accumulate);
        mma_tf32_ts(acc_lo, a_hi + (uint32_t)(j * kUmmaK), dbl, idesc, 1u);
        mma_tf32_ts(acc_hi, a_hi + (uint32_t)(j * kUmmaK), db, idesc, accumulate);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_addr(&s_empty[s]))
                   : "memory");
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_addr(&s_aempty[a]))
                   : "memory");
    }
    if (!failed)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_addr(&s_acc_full))
                   : "memory");
  } else if (warp >= 4) {
    // ===== transform warps: X tile -> {TF32 pattern, residual} -> TMEM; then the epilogue =====
    const int quarter = warp - 4;                 // TMEM lane quarter this warp may access (warp % 4)
    const int m = quarter * 32 + lane;            // pixel of this thread within the tile
    const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
    bool ok = true;
    for (int it = 0; it < total_iters && ok; ++it) {
      const int s = it % kPxStages, a = it & 1;
      ok = bar_wait(smem_addr(&s_full[s]), (it / kPxStages) & 1) &&
           bar_wait(smem_addr(&s_aempty[a]), ((it >> 1) & 1) ^ 1);
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) break;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const float* xs = reinterpret_cast<const float*>(tiles_ptr + s * kPxStageBytes) + m;
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const float v = xs[k * kBM];
        hi[k] = __float_as_uint(v);
        lo[k] = __float_as_uint(tf32_residual(v));
      }
      tmem_st32(lane_base + (uint32_t)(a * 64), hi);
      tmem_st32(lane_base + (uint32_t)(a * 64 + 32), lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&s_aready[a])) : "memory");
    }
    if (!ok) s_fail = 1;
    // ----- epilogue: TMEM lane = pixel, columns = output channels -----
    if (ok && total_iters > 0) {
      ok = bar_wait(smem_addr(&s_acc_full), 0);
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) s_fail = 1;
    }
    if (ok) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int px = px0 + m;
      const bool valid = px < p.hw;
      const long long plane = p.hw;
      const long long base = ((long long)img * p.c_out + co0) * plane + px;
#pragma unroll 1
      for (int c = 0; c < kPxBN / 32; ++c) {
        uint32_t r[32], r2[32];
        if (total_iters > 0) {
          tmem_ld32(lane_base + 128u + (uint32_t)(c * 32), r);
          tmem_ld32(lane_base + 192u + (uint32_t)(c * 32), r2);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = r2[j] = 0u;
        }
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int co = co0 + c * 32 + j;
            if (co < p.c_out) {
              float v = __uint_as_float(r[j]) + __uint_as_float(r2[j]);
              if (p.bias) v += __ldg(p.bias + co);
              const long long off = base + (long long)(c * 32 + j) * plane;
              p.Y[off] = v;
              if (p.Y2) p.Y2[off] = relu_keep_nan(v);
            }
          }
        }
      }
    }
  }
  __syncwarp();
  if (s_fail) {
    if ((threadIdx.x & 31) == 0 && p.error_flag) atomicExch(p.error_flag, 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kPxTmemCols) : "memory");
  }
}

// ---- tap-table convolution with the activations through TMEM (experimental, DPL_TAPS_TS=1) -----------
// conv_taps_tf32x3_kernel's problem (shifted windows of the channel-last padded copy, one accumulating GEMM
// step per tap and 32-channel block) on conv1x1_px_tf32x3_kernel's machinery: the X tile [128 q][32 ci] lands
// in shared memory (K-major, 128-byte swizzle — the layout the existing tensor map produces), four transform
// warps read their own row of it (thread <-> padded pixel; the swizzle spreads the 32 rows of a warp over all
// banks: 4 wavefronts per 16-byte access, the minimum), split it into the TF32 pattern and its residual and
// tcgen05.st both into TMEM; the MMAs take A from TMEM and only the weight tiles from shared memory.
// Per 32-channel block and tap, shared-memory traffic drops from 176 KB (both operands in shared memory,
// X_lo written back) to 16 (X landing) + 16 (W, W_lo landing) + 16 (transform read) + 24 (MMA reads of W) KB.
// 128 q x 64 co tiles, 256 TMEM columns, 96 KB of shared memory: two CTAs per SM, one tile's epilogue
// overlaps the other's main loop. One leading accumulator (no hi_alt): NOT YET RUN ON HARDWARE.
__global__ void __launch_bounds__(kPxThreads, 2)
conv_taps_ts_tf32x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                           const __grid_constant__ CUtensorMap tmWlo, const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_full[kPxStages], s_empty[kPxStages], s_aready[2], s_aempty[2], s_acc_full;
  __shared__ uint32_t s_tmem_base;
  __shared__ int s_fail;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tiles = (smem_addr(smem_raw) + 1023u) & ~1023u;
  const uint8_t* tiles_ptr = smem_raw + (tiles - smem_addr(smem_raw));
  const int co0 = blockIdx.x * kPxBN;                      // output-channel groups fastest: X tiles shared in L2
  const long long q0 = (long long)blockIdx.y * kBM;
  const int num_kb = (p.c_in + kBK - 1) / kBK;
  const int total_iters = p.n_taps * num_kb;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kPxStages; ++s) {
      bar_init(smem_addr(&s_full[s]), 1);
      bar_init(smem_addr(&s_empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      bar_init(smem_addr(&s_aready[a]), 4);   // one arrival per transform warp
      bar_init(smem_addr(&s_aempty[a]), 1);
    }
    bar_init(smem_addr(&s_acc_full), 1);
    s_fail = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_addr(&s_tmem_base)),
                 "r"(kPxTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem_base;
  const uint32_t acc_hi = tmem + 128u, acc_lo = tmem + 192u;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer: X tile (shifted rows of the padded copy), W and W_lo tiles of this tap =====
    for (int it = 0; it < total_iters; ++it) {
      const int s = it % kPxStages;
      const uint32_t ph = (it / kPxStages) & 1;
      if (!bar_wait(smem_addr(&s_empty[s]), ph ^ 1)) {
        s_fail = 1;
        break;
      }
      const uint32_t full = smem_addr(&s_full[s]);
      bar_expect_tx(full, kPxStageBytes);
      const int kb = it / p.n_taps, tap = it - kb * p.n_taps;
      const int k0 = kb * kBK;
      const uint32_t x_tile = tiles + s * kPxStageBytes, w_tile = x_tile + kPxXBytes, wlo_tile = w_tile + kPxWBytes;
      tma_load_3d(x_tile, &tmX, k0, (int)q0 + p.tap_shift[tap], 0, full);
      tma_load_3d(w_tile, &tmW, k0, co0, tap, full);
      tma_load_3d(wlo_tile, &tmWlo, k0, co0, tap, full);
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer: A from TMEM, three MMAs per K step =====
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kPxBN >> 3) << 17) |
                           ((uint32_t)(kBM >> 4) << 24);      // f32 accumulate, tf32 x tf32, both K-major
    bool failed = false;
    for (int it = 0; it < total_iters; ++it) {
      const int s = it % kPxStages, a = it & 1;
      if (!bar_wait(smem_addr(&s_aready[a]), (it >> 1) & 1)) {
        s_fail = 1;
        failed = true;
        break;
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t w_tile = tiles + s * kPxStageBytes + kPxXBytes, wlo_tile = w_tile + kPxWBytes;
      const uint32_t a_hi = tmem + (uint32_t)(a * 64), a_lo = a_hi + 32u;
#pragma unroll
      for (int j = 0; j < kBK / kUmmaK; ++j) {
        const uint64_t db = desc_k_major(w_tile, j), dbl = desc_k_major(wlo_tile, j);
        const uint32_t accumulate = (it > 0 || j > 0) ? 1u : 0u;
        mma_tf32_ts(acc_lo, a_lo + (uint32_t)(j * kUmmaK), db, idesc, accumulate);
        mma_tf32_ts(acc_lo, a_hi + (uint32_t)(j * kUmmaK), dbl, idesc, 1u);
        mma_tf32_ts(acc_hi, a_hi + (uint32_t)(j * kUmmaK), db, idesc, accumulate);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_addr(&s_empty[s]))
                   : "memory");
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_addr(&s_aempty[a]))
                   : "memory");
    }
    if (!failed)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_addr(&s_acc_full))
                   : "memory");
  } else if (warp >= 4) {
    // ===== transform warps: own row of the X tile -> {TF32 pattern, residual} -> TMEM; then the epilogue =====
    const int quarter = warp - 4;                 // TMEM lane quarter this warp may access (warp % 4)
    const int m = quarter * 32 + lane;            // padded pixel of this thread within the tile
    const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
    bool ok = true;
    for (int it = 0; it < total_iters && ok; ++it) {
      const int s = it % kPxStages, a = it & 1;
      ok = bar_wait(smem_addr(&s_full[s]), (it / kPxStages) & 1) &&
           bar_wait(smem_addr(&s_aempty[a]), ((it >> 1) & 1) ^ 1);
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) break;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // row m of a 128-byte-swizzled tile: logical 16-byte chunk c sits at chunk c ^ (m & 7)
      const uint8_t* row = tiles_ptr + s * kPxStageBytes + m * 128;
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(row + ((c ^ (m & 7)) << 4));
        hi[4 * c + 0] = __float_as_uint(v.x);
        hi[4 * c + 1] = __float_as_uint(v.y);
        hi[4 * c + 2] = __float_as_uint(v.z);
        hi[4 * c + 3] = __float_as_uint(v.w);
        lo[4 * c + 0] = __float_as_uint(tf32_residual(v.x));
        lo[4 * c + 1] = __float_as_uint(tf32_residual(v.y));
        lo[4 * c + 2] = __float_as_uint(tf32_residual(v.z));
        lo[4 * c + 3] = __float_as_uint(tf32_residual(v.w));
      }
      tmem_st32(lane_base + (uint32_t)(a * 64), hi);
      tmem_st32(lane_base + (uint32_t)(a * 64 + 32), lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&s_aready[a])) : "memory");
    }
    if (!ok) s_fail = 1;
    // ----- epilogue: TMEM lane = padded pixel q, columns = output channels -----
    if (ok) {
      ok = bar_wait(smem_addr(&s_acc_full), 0);
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) s_fail = 1;
    }
    if (ok) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const long long q = q0 + m;
      bool valid = q < p.q_total;
      long long out_base = 0;
      if (valid) {
        const int img = (int)(q / p.plane);
        const int r = (int)(q - (long long)img * p.plane);
        const int hp = r / p.Wp, wp = r - hp * p.Wp;
        const int ho = hp - p.origin, wo = wp - p.origin;
        valid = ho >= 0 && ho < p.H && wo >= 0 && wo < p.W;
        out_base = (((long long)img * p.c_out) * p.H + ho) * p.W + wo;
      }
      const long long ch_stride = (long long)p.H * p.W;
      float rlo = INFINITY, rhi = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < kPxBN / 32; ++c) {
        uint32_t r[32], r2[32];
        tmem_ld32(lane_base + 128u + (uint32_t)(c * 32), r);
        tmem_ld32(lane_base + 192u + (uint32_t)(c * 32), r2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (valid) {
          const int cb = co0 + c * 32;
          float* dst = p.Y + out_base + (long long)cb * ch_stride;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (cb + j < p.c_out) {
              float v = __uint_as_float(r[j]) + __uint_as_float(r2[j]);
              if (p.bias) v += __ldg(p.bias + cb + j);
              if (p.relu) v = fmaxf(v, 0.f);
              dst[(long long)j * ch_stride] = v;
              if (p.Y2) p.Y2[(dst - p.Y) + (long long)j * ch_stride] = relu_keep_nan(v);
              rlo = fminf(rlo, v);
              rhi = fmaxf(rhi, v);
            }
          }
        }
      }
      warp_range_flush(rlo, rhi, p.bmin, p.bmax, p.rmin, p.rmax);
    }
  }
  __syncwarp();
  if (s_fail) {
    if ((threadIdx.x & 31) == 0 && p.error_flag) atomicExch(p.error_flag, 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kPxTmemCols) : "memory");
  }
}

// Unswizzled 3-D fp32 map (px, ci, img) with a [128 x 32 x 1] box for the pixel-major kernel.
int make_map_px(CUtensorMap* map, const float* base, uint64_t hw, uint64_t c_in, uint64_t n_img);

// im2col staging for the few-channel stem convolution (ResNet's 7x7 / stride 2 on 3 channels): the
// tap-table kernel wants >= 16 input channels per tap, so here ALL taps and channels of an output
// pixel are laid side by side,
//     Xp[q][k],  q = (img * Ho + ho) * Wo + wo,  k = (c * kh + a) * kw + b  (zero for k >= C kh kw),
// and the convolution is ONE tap with c_in = k_pad over the weight viewed as [c_out][C kh kw].
// One warp per row: the (c, a, b) decomposition of a lane's columns is hoisted out of the row loop.
constexpr int kIm2colMaxK = 256;

__global__ void __launch_bounds__(256)
im2col_kernel(const float* __restrict__ x, float* __restrict__ xp, int n_img, int C, int H, int W, int kh, int kw,
              int stride, int pad, int Ho, int Wo, int k_pad) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int K = C * kh * kw;
  int off[kIm2colMaxK / 32], da[kIm2colMaxK / 32], db[kIm2colMaxK / 32];
#pragma unroll
  for (int j = 0; j < kIm2colMaxK / 32; ++j) {
    const int k = j * 32 + lane;
    const int c = k / (kh * kw), r = k - c * (kh * kw);
    da[j] = r / kw;
    db[j] = r - da[j] * kw;
    off[j] = k < K ? (c * H + da[j]) * W + db[j] : -1;
  }
  const long long rows = (long long)n_img * Ho * Wo;
  for (long long q = warp; q < rows; q += n_warps) {
    const int wo = (int)(q % Wo);
    const int ho = (int)((q / Wo) % Ho);
    const long long img = q / ((long long)Wo * Ho);
    const int h0 = ho * stride - pad, w0 = wo * stride - pad;
    const float* src = x + img * (long long)C * H * W + (long long)h0 * W + w0;
    float* dst = xp + q * k_pad;
#pragma unroll
    for (int j = 0; j < kIm2colMaxK / 32; ++j) {
      const int k = j * 32 + lane;
      if (k < k_pad) {
        float v = 0.f;
        if (off[j] >= 0) {
          const int h = h0 + da[j], w = w0 + db[j];
          if (h >= 0 && h < H && w >= 0 && w < W) v = __ldg(src + off[j]);
        }
        dst[k] = v;
      }
    }
  }
}


int make_map_px(CUtensorMap* map, const float* base, uint64_t hw, uint64_t c_in, uint64_t n_img) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return DPL_E_UNSUPPORTED;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15u) || (hw & 3u)) {
    set_error("activation not TMA-compatible: base 16-byte aligned and H*W a multiple of 4 required");
    return DPL_E_UNSUPPORTED;
  }
  cuuint64_t dims[3] = {hw, c_in, n_img};
  cuuint64_t strides[2] = {hw * 4, c_in * hw * 4};
  cuuint32_t box[3] = {(cuuint32_t)kBM, (cuuint32_t)kBK, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (pixel-major map) failed with CUresult %d", (int)r);
    return DPL_E_UNSUPPORTED;
  }
  return 0;
}

#include "dpl_x3p.cuh"
#include "dpl_x3ts.cuh"

}  // namespace
}  // namespace dpl

using namespace dpl;

// Experiment switch (DPL_X3_ALT=0: one accumulator for the leading 3xTF32 term).
static int x3_hi_alt() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DPL_X3_ALT");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v;
}

// DPL_GEMM_PERSISTENT=0: keep dpl_gemm_tf32's 1x1-convolution shapes on the one-tile-per-CTA kernel.
static int gemm_persistent() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DPL_GEMM_PERSISTENT");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v;
}

// K up to which the 1x1-convolution GEMM takes the persistent kernel (DPL_X3_PERSISTENT_MAX_K, 0 = never).
static int x3_persistent_max_k() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DPL_X3_PERSISTENT_MAX_K");
    v = e ? atoi(e) : 512;
  }
  return v;
}

// Single-pass TF32 tap-table convolution on the persistent kernel (x3p MODE 1, single) for dpl_tap_conv_tf32
// (dpl_recon_conv.cu): same arguments as its one-tile-per-CTA kernel. Opt-in (DPL_TAPCONV_PERSISTENT=1, returns < 0
// otherwise): measured slower than the one-tile kernel with two CTAs per SM on every ResNet-50 shape (forward
// 0.126 vs 0.112 ms at 64 ch / 56 x 56, 0.083 vs 0.061 at 256 ch / 14 x 14) - a single-pass MMA stream is short
// enough that twice the TMA loads in flight matter more than the overlapped epilogue.
namespace dpl {
int tap_conv_tf32_persistent(const float* d_xp, long long total_rows, const float* d_w_taps, int n_w_taps, float* d_y,
                             int n_img, int ck, int cn, int H, int W, int Hp, int Wp, int origin, int out_stride,
                             int out_a, int out_b, int n_taps, const int* tap_shift, const int* tap_w,
                             const float* d_bias, int* d_error_flag, cudaStream_t stream) {
  static const bool on = [] {
    const char* e = getenv("DPL_TAPCONV_PERSISTENT");
    return e && e[0] == '1';
  }();
  if (!on) return -1;
  const int bn = cn <= 64 ? 64 : 128;
  CUtensorMap tmX, tmW;
  int st = make_map(&tmX, d_xp, (uint64_t)ck, (uint64_t)total_rows, 1, (uint64_t)ck, 0, kBM, false);
  if (!st)
    st = make_map(&tmW, d_w_taps, (uint64_t)ck, (uint64_t)cn, (uint64_t)n_w_taps, (uint64_t)ck, (uint64_t)cn * ck,
                  (uint32_t)bn, false);
  if (st) return st;
  X3PParams xp;
  xp.single = 1;
  xp.chunk_iters = 1 << 20;
  xp.g = GemmParams();
  ConvParams& c = xp.c;
  c = ConvParams();
  c.n_img = n_img;
  c.c_in = ck;
  c.c_out = cn;
  c.H = H;
  c.W = W;
  c.Wp = Wp;
  c.plane = Hp * Wp;
  c.origin = origin;
  c.n_taps = n_taps;
  // the kernel loads the weight slice of iteration tap index t from slice t: gather the slices' indices into
  // the shift table's order by passing tap_w through tap_shift's companion below
  for (int t = 0; t < 9; ++t) c.tap_shift[t] = t < n_taps ? tap_shift[t] : 0;
  for (int t = 0; t < 9; ++t) c.tap_w[t] = t < n_taps ? tap_w[t] : 0;
  c.bn = bn;
  c.hi_alt = 0;
  c.q_total = (long long)n_img * c.plane;
  c.Y = d_y;
  c.bias = d_bias;
  c.relu = 0;
  c.error_flag = d_error_flag;
  c.os = out_stride;
  c.oa = out_a;
  c.ob = out_b;
  const long long total = ((c.q_total + kBM - 1) / kBM) * ((cn + bn - 1) / bn);
  int e = launch_x3p<1>(tmW, tmW, tmX, xp, total, stream);
  if (e) return e;
  DPL_LAUNCH_CHECK("x3p_kernel<1> (single)");
  return 0;
}
}  // namespace dpl

// a_major / b_major: 0 = K-major (element (row, k) at row * ld + k), 1 = MN-major (at k * ld + row).
extern "C" int dpl_gemm_tf32(const float* d_a, int a_major, long long lda, long long a_batch_stride,
                             const float* d_b, int b_major, long long ldb, long long b_batch_stride,
                             float* d_d, long long ldd, long long d_batch_stride, int M, int N, int K,
                             int batch, int fold_batch, int split_k, const float* d_bias, int bias_mode,
                             int relu, int* d_error_flag, void* stream) {
  DPL_REQUIRE(d_a && d_b && d_d, "null pointer");
  DPL_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0, "empty problem");
  DPL_REQUIRE(bias_mode == 0 || d_bias, "bias_mode without bias");
  CUtensorMap tmA, tmB;
  int st;
  const uint64_t a_z = a_batch_stride ? (uint64_t)batch : 1;
  if (a_major == 0)
    st = make_map(&tmA, d_a, K, M, a_z, lda, a_batch_stride, kBM, false);
  else
    st = make_map(&tmA, d_a, M, K, a_z, lda, a_batch_stride, kBK, true);
  if (st) return st;
  if (b_major == 0)
    st = make_map(&tmB, d_b, K, N, batch, ldb, b_batch_stride, kBN, false);
  else
    st = make_map(&tmB, d_b, N, K, batch, ldb, b_batch_stride, kBK, true);
  if (st) return st;

  GemmParams p;
  p.M = M;
  p.N = N;
  p.K = K;
  p.batch = batch;
  p.a_batched = a_batch_stride ? 1 : 0;
  p.fold_batch = fold_batch ? 1 : 0;
  p.D = d_d;
  p.ldd = ldd;
  p.d_batch_stride = d_batch_stride;
  p.bias = d_bias;
  p.bias_mode = bias_mode;
  p.relu = relu;
  p.error_flag = d_error_flag;
  p.D2 = nullptr;
  p.bmin = p.bmax = p.rmin = p.rmax = nullptr;
  unsigned gz;
  if (fold_batch) {
    if (split_k < 1) split_k = 1;
    if (split_k > batch) split_k = batch;
    p.z_per_cta = (batch + split_k - 1) / split_k;
    gz = (unsigned)((batch + p.z_per_cta - 1) / p.z_per_cta);
    p.atomic_out = gz > 1 ? 1 : 0;
  } else {
    p.z_per_cta = 1;
    gz = (unsigned)batch;
    p.atomic_out = 0;
  }
  dim3 grid((M + kBM - 1) / kBM, (N + kBN - 1) / kBN, gz);
  const size_t smem = (size_t)kStages * kStageBytes + 1024;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a_major == 0 && b_major == 1 && !fold_batch && !p.a_batched && bias_mode != 2 && gemm_persistent()) {
    // 1x1 convolution forward / data gradient (NCHW activations on the N side): the persistent kernel of
    // dpl_x3p.cuh in single-pass mode - epilogue overlapped with the next tile, coalesced 128-byte row stores
    // through a shared-memory staging tile (the one-tile kernel below writes 16 bytes to each of 32 rows per
    // store instruction; measured 2.8x slower than cuDNN TF32 on the 64 -> 256 @ 56 x 56 layer)
    X3PParams xp;
    xp.single = 1;
    xp.chunk_iters = 1 << 20;        // one accumulation chain per tile, as cuDNN's TF32 kernels
    xp.g = p;
    xp.g.hi_alt = 0;
    xp.c = ConvParams();
    int e = launch_x3p<0>(tmA, tmA, tmB, xp, (long long)grid.x * grid.y * grid.z, s);
    if (e) return e;
    DPL_LAUNCH_CHECK("x3p_kernel<0> (single)");
    return 0;
  }
#define DPL_GEMM_LAUNCH(AMN, BMN)                                                                      \
  do {                                                                                                 \
    static bool attr_done = false; /* one device per process (one rank per GPU) */                    \
    if (!attr_done) {                                                                                  \
      int e = cuda_status(cudaFuncSetAttribute(gemm_tf32_kernel<AMN, BMN>,                             \
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),\
                          "cudaFuncSetAttribute(gemm_tf32_kernel)");                                   \
      if (e) return e;                                                                                 \
      attr_done = true;                                                                                \
    }                                                                                                  \
    gemm_tf32_kernel<AMN, BMN><<<grid, kGemmThreads, smem, s>>>(tmA, tmB, p);                          \
  } while (0)
  if (a_major == 0 && b_major == 0)
    DPL_GEMM_LAUNCH(false, false);
  else if (a_major == 0 && b_major == 1)
    DPL_GEMM_LAUNCH(false, true);
  else if (a_major == 1 && b_major == 0)
    DPL_GEMM_LAUNCH(true, false);
  else
    DPL_GEMM_LAUNCH(true, true);
#undef DPL_GEMM_LAUNCH
  DPL_LAUNCH_CHECK("gemm_tf32_kernel");
  return 0;
}

// lo = x - trunc_tf32(x): the second operand of the 3xTF32 product, for weights (done once).
extern "C" int dpl_tf32_residual_f32(const float* d_x, float* d_lo, uint64_t n, void* stream) {
  DPL_REQUIRE(d_x && d_lo, "null pointer");
  if (n == 0) return 0;
  uint64_t blocks = (n + 1023) / 1024;
  if (blocks > (uint64_t)sm_count() * 8) blocks = (uint64_t)sm_count() * 8;
  tf32_residual_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_x, d_lo, n);
  DPL_LAUNCH_CHECK("tf32_residual_kernel");
  return 0;
}

// Unbiased operand split for weights (done once): hi = RN_tf32(x), lo = RN_tf32(x - hi). Truncating instead
// (what kind::tf32 does to a raw fp32 pattern) shrinks every product by ~2^-22: 2.8e-7 per layer, compounding.
extern "C" int dpl_tf32_split_f32(const float* d_x, float* d_hi, float* d_lo, uint64_t n, void* stream) {
  DPL_REQUIRE(d_x && d_hi && d_lo, "null pointer");
  if (n == 0) return 0;
  uint64_t blocks = (n + 1023) / 1024;
  if (blocks > (uint64_t)sm_count() * 8) blocks = (uint64_t)sm_count() * 8;
  tf32_split_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_x, d_hi, d_lo, n);
  DPL_LAUNCH_CHECK("tf32_split_kernel");
  return 0;
}

// 3xTF32 variant of dpl_gemm_tf32 (fp32-accurate): same operand description plus d_a_lo, the
// residual of A with A's layout. No batch folding / split-K (forward use only).
extern "C" int dpl_gemm_tf32x3(const float* d_a, const float* d_a_lo, int a_major, long long lda,
                               long long a_batch_stride, const float* d_b, int b_major, long long ldb,
                               long long b_batch_stride, float* d_d, long long ldd, long long d_batch_stride,
                               int M, int N, int K, int batch, const float* d_bias, int bias_mode, int relu,
                               float* d_d_relu, float* d_blob_min, float* d_blob_max, float* d_relu_min,
                               float* d_relu_max, int* d_error_flag, void* stream) {
  DPL_REQUIRE(d_a && d_a_lo && d_b && d_d, "null pointer");
  DPL_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0, "empty problem");
  DPL_REQUIRE(bias_mode == 0 || d_bias, "bias_mode without bias");
  CUtensorMap tmA, tmAlo, tmB;
  int st;
  const uint64_t a_z = a_batch_stride ? (uint64_t)batch : 1;
  if (a_major == 0) {
    st = make_map(&tmA, d_a, K, M, a_z, lda, a_batch_stride, kBM, false);
    if (!st) st = make_map(&tmAlo, d_a_lo, K, M, a_z, lda, a_batch_stride, kBM, false);
  } else {
    st = make_map(&tmA, d_a, M, K, a_z, lda, a_batch_stride, kBK, true);
    if (!st) st = make_map(&tmAlo, d_a_lo, M, K, a_z, lda, a_batch_stride, kBK, true);
  }
  if (st) return st;
  if (b_major == 0)
    st = make_map(&tmB, d_b, K, N, batch, ldb, b_batch_stride, kBN, false);
  else
    st = make_map(&tmB, d_b, N, K, batch, ldb, b_batch_stride, kBK, true);
  if (st) return st;
  GemmParams p;
  p.M = M;
  p.N = N;
  p.K = K;
  p.batch = batch;
  p.a_batched = a_batch_stride ? 1 : 0;
  p.fold_batch = 0;
  p.z_per_cta = 1;
  p.D = d_d;
  p.ldd = ldd;
  p.d_batch_stride = d_batch_stride;
  p.bias = d_bias;
  p.bias_mode = bias_mode;
  p.relu = relu;
  p.atomic_out = 0;
  p.hi_alt = x3_hi_alt();
  p.error_flag = d_error_flag;
  p.D2 = d_d_relu;
  p.bmin = d_blob_min;
  p.bmax = d_blob_max;
  p.rmin = d_relu_min;
  p.rmax = d_relu_max;
  dim3 grid((M + kBM - 1) / kBM, (N + kBN - 1) / kBN, (unsigned)batch);
  const size_t smem = (size_t)kStages3 * kStageBytes3 + 1024;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a_major == 0 && b_major == 1 && !p.a_batched && bias_mode != 2 && x3_chunk_iters() > 0) {
    // 1x1 convolution: persistent kernel with chunked accumulation (dpl_x3p.cuh), any K
    X3PParams xp;
    xp.single = 0;
    xp.chunk_iters = x3_chunk_iters();
    xp.g = p;
    xp.c = ConvParams();
    const long long total = (long long)grid.x * grid.y * grid.z;
    int e = launch_x3p<0>(tmA, tmAlo, tmB, xp, total, s);
    if (e) return e;
    DPL_LAUNCH_CHECK("x3p_kernel<0>");
    return 0;
  }
  if (a_major == 0 && b_major == 1 && !p.a_batched && bias_mode != 2 && K <= x3_persistent_max_k()) {
    // short-K 1x1 convolution: persistent kernel, one CTA per SM
    static bool attr_done_p = false;
    if (!attr_done_p) {
      int e = cuda_status(cudaFuncSetAttribute(gemm_tf32x3_persistent_kernel,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem + kStgBytes)),
                          "cudaFuncSetAttribute(gemm_tf32x3_persistent_kernel)");
      if (e) return e;
      attr_done_p = true;
    }
    const long long total = (long long)grid.x * grid.y * grid.z;
    const unsigned ctas = (unsigned)(total < sm_count() ? total : sm_count());
    gemm_tf32x3_persistent_kernel<<<ctas, kGemmPThreads, smem + kStgBytes, s>>>(tmA, tmAlo, tmB, p);
    DPL_LAUNCH_CHECK("gemm_tf32x3_persistent_kernel");
    return 0;
  }
#define DPL_GEMM3_LAUNCH(AMN, BMN)                                                                     \
  do {                                                                                                 \
    static bool attr_done = false;                                                                     \
    if (!attr_done) {                                                                                  \
      int e = cuda_status(cudaFuncSetAttribute(gemm_tf32x3_kernel<AMN, BMN>,                           \
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),\
                          "cudaFuncSetAttribute(gemm_tf32x3_kernel)");                                 \
      if (e) return e;                                                                                 \
      attr_done = true;                                                                                \
    }                                                                                                  \
    gemm_tf32x3_kernel<AMN, BMN><<<grid, kGemm3Threads, smem, s>>>(tmA, tmAlo, tmB, p);                \
  } while (0)
  if (a_major == 0 && b_major == 0)
    DPL_GEMM3_LAUNCH(false, false);
  else if (a_major == 0 && b_major == 1)
    DPL_GEMM3_LAUNCH(false, true);
  else if (a_major == 1 && b_major == 0)
    DPL_GEMM3_LAUNCH(true, false);
  else
    DPL_GEMM3_LAUNCH(true, true);
#undef DPL_GEMM3_LAUNCH
  DPL_LAUNCH_CHECK("gemm_tf32x3_kernel");
  return 0;
}

// Channel-last staging copy for dpl_conv_taps_tf32x3:
//   d_xp[(plane * n_img + img) * Hp * Wp + hp * Wp + wp][c] = X[img][c][stride (hp - origin) + a][stride (wp - origin) + b]
// (zero where that is outside the image or hp / wp < origin), plane = a * stride + b < n_planes.
extern "C" int dpl_pad_plane_f32(const float* d_x, float* d_xp, int n_img, int channels, int H, int W, int stride,
                                 int origin, int Hp, int Wp, int n_planes, void* stream) {
  DPL_REQUIRE(d_x && d_xp, "null pointer");
  DPL_REQUIRE(n_img > 0 && channels > 0 && H > 0 && W > 0 && Hp > 0 && Wp > 0, "empty problem");
  DPL_REQUIRE(stride >= 1 && stride <= 4 && n_planes >= 1 && n_planes <= stride * stride, "stride / n_planes");
  DPL_REQUIRE(origin == 0 || origin == 1, "origin must be 0 or 1");
  DPL_REQUIRE((long long)n_img * n_planes <= 65535 && (channels + 31) / 32 <= 65535, "grid limit");
  const long long plane = (long long)Hp * Wp;
  DPL_REQUIRE(plane < (1ll << 30), "plane too large");
  dim3 grid((unsigned)((plane + 31) / 32), (unsigned)((channels + 31) / 32), (unsigned)(n_img * n_planes));
  pad_plane_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_x, d_xp, n_img, channels, H, W, stride,
                                                                        origin, Hp, Wp);
  DPL_LAUNCH_CHECK("pad_plane_kernel");
  return 0;
}

// Tap-table convolution, fp32-accurate (3xTF32) on the tensor cores:
//   Y[img][co][ho][wo] = bias[co] + sum_tap sum_ci Wt[tap][co][ci] * Xp[q + tap_shift[tap]][ci],
//   q = img * Hp * Wp + (ho + origin) * Wp + (wo + origin)
//   d_xp        staging copy from dpl_pad_plane_f32, [total_rows][c_in]
//   d_w_taps    weights tap-major [n_taps][c_out][c_in]; d_w_taps_lo their dpl_tf32_residual_f32
//   tap_shift   host array of n_taps row offsets (1 <= n_taps <= 9)
extern "C" int dpl_conv_taps_tf32x3(const float* d_xp, long long total_rows, const float* d_w_taps,
                                    const float* d_w_taps_lo, float* d_y, int n_img, int c_in, int c_out, int Ho,
                                    int Wo, int Hp, int Wp, int origin, int n_taps, const int* tap_shift,
                                    const float* d_bias, int relu, float* d_y_relu, float* d_blob_min,
                                    float* d_blob_max, float* d_relu_min, float* d_relu_max, int* d_error_flag,
                                    void* stream) {
  DPL_REQUIRE(d_xp && d_w_taps && d_w_taps_lo && d_y && tap_shift, "null pointer");
  DPL_REQUIRE(n_img > 0 && c_in > 0 && c_out > 0 && Ho > 0 && Wo > 0, "empty problem");
  DPL_REQUIRE(n_taps >= 1 && n_taps <= 9, "1 <= n_taps <= 9");
  DPL_REQUIRE(origin == 0 || origin == 1, "origin must be 0 or 1");
  DPL_REQUIRE(Hp >= Ho + origin && Wp >= Wo + origin, "padded plane smaller than the output");
  const long long plane = (long long)Hp * Wp;
  const long long q_total = (long long)n_img * plane;
  DPL_REQUIRE(total_rows >= q_total && total_rows < (1ll << 31) - 4096, "total_rows out of range");
  if (c_in & 3) {
    set_error("dpl_conv_taps_tf32x3: c_in must be a multiple of 4 (TMA stride alignment)");
    return DPL_E_UNSUPPORTED;
  }
  CUtensorMap tmX, tmW, tmWlo;
  int st = make_map(&tmX, d_xp, (uint64_t)c_in, (uint64_t)total_rows, 1, (uint64_t)c_in, 0, kBM, false);
  if (st) return st;
  // experimental variant with the activations through TMEM (conv_taps_ts_tf32x3_kernel): opt-in
  static const bool taps_ts = [] {
    const char* e = getenv("DPL_TAPS_TS");
    return e && e[0] == '1';
  }();
  const int bn = taps_ts ? kPxBN : (c_out <= 64 ? 64 : 128);
  st = make_map(&tmW, d_w_taps, (uint64_t)c_in, (uint64_t)c_out, (uint64_t)n_taps, (uint64_t)c_in,
                (uint64_t)c_out * c_in, (uint32_t)bn, false);
  if (!st)
    st = make_map(&tmWlo, d_w_taps_lo, (uint64_t)c_in, (uint64_t)c_out, (uint64_t)n_taps, (uint64_t)c_in,
                  (uint64_t)c_out * c_in, (uint32_t)bn, false);
  if (st) return st;
  ConvParams p;
  p.n_img = n_img;
  p.c_in = c_in;
  p.c_out = c_out;
  p.H = Ho;
  p.W = Wo;
  p.Wp = Wp;
  p.plane = (int)plane;
  p.origin = origin;
  p.n_taps = n_taps;
  for (int t = 0; t < 9; ++t) p.tap_shift[t] = t < n_taps ? tap_shift[t] : 0;
  for (int t = 0; t < 9; ++t) p.tap_w[t] = t;
  p.bn = bn;
  p.hi_alt = x3_hi_alt();
  p.q_total = q_total;
  p.Y = d_y;
  p.bias = d_bias;
  p.relu = relu;
  p.error_flag = d_error_flag;
  p.Y2 = d_y_relu;
  p.bmin = d_blob_min;
  p.bmax = d_blob_max;
  p.rmin = d_relu_min;
  p.rmax = d_relu_max;
  p.os = 1;
  p.oa = p.ob = 0;
  if (!taps_ts && x3_chunk_iters() > 0 && x3_ts()) {
    // persistent kernel, chunked accumulation, activations through TMEM (dpl_x3ts.cuh)
    const long long total = ((q_total + kBM - 1) / kBM) * ((c_out + bn - 1) / bn);
    int e = launch_x3ts<0>(tmX, tmW, tmWlo, p, total, static_cast<cudaStream_t>(stream));
    if (e) return e;
    DPL_LAUNCH_CHECK("x3ts_kernel<0>");
    return 0;
  }
  if (!taps_ts && x3_chunk_iters() > 0) {
    // persistent kernel with chunked accumulation (dpl_x3p.cuh)
    X3PParams xp;
    xp.single = 0;
    xp.chunk_iters = x3_chunk_iters();
    xp.g = GemmParams();
    xp.c = p;
    const long long total = ((q_total + kBM - 1) / kBM) * ((c_out + bn - 1) / bn);
    int e = launch_x3p<1>(tmW, tmWlo, tmX, xp, total, static_cast<cudaStream_t>(stream));
    if (e) return e;
    DPL_LAUNCH_CHECK("x3p_kernel<1>");
    return 0;
  }
  if (taps_ts) {
    DPL_REQUIRE((q_total + kBM - 1) / kBM <= 65535, "grid limit (DPL_TAPS_TS)");
    dim3 grid_ts((unsigned)((c_out + kPxBN - 1) / kPxBN), (unsigned)((q_total + kBM - 1) / kBM), 1);
    const size_t smem_ts = (size_t)kPxStages * kPxStageBytes + 1024;
    static bool attr_ts_done = false;
    if (!attr_ts_done) {
      int e = cuda_status(cudaFuncSetAttribute(conv_taps_ts_tf32x3_kernel,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ts),
                          "cudaFuncSetAttribute(conv_taps_ts_tf32x3_kernel)");
      if (e) return e;
      attr_ts_done = true;
    }
    conv_taps_ts_tf32x3_kernel<<<grid_ts, kPxThreads, smem_ts, static_cast<cudaStream_t>(stream)>>>(tmX, tmW, tmWlo,
                                                                                                   p);
    DPL_LAUNCH_CHECK("conv_taps_ts_tf32x3_kernel");
    return 0;
  }
  dim3 grid((unsigned)((q_total + kBM - 1) / kBM), (unsigned)((c_out + bn - 1) / bn), 1);
  const size_t smem = (size_t)kStages3 * kStageBytes3 + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    int e = cuda_status(cudaFuncSetAttribute(conv_taps_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem),
                        "cudaFuncSetAttribute(conv_taps_tf32x3_kernel)");
    if (e) return e;
    attr_done = true;
  }
  conv_taps_tf32x3_kernel<<<grid, kGemm3Threads, smem, static_cast<cudaStream_t>(stream)>>>(tmX, tmW, tmWlo, p);
  DPL_LAUNCH_CHECK("conv_taps_tf32x3_kernel");
  return 0;
}

// Pixel-major 1x1 convolution straight from / to NCHW (see conv1x1_px_tf32x3_kernel).
//   d_x [n_img][c_in][hw], d_w [c_out][c_in] with residual d_w_lo, d_y [n_img][c_out][hw];
//   hw and c_in multiples of 4 (TMA strides), else DPL_E_UNSUPPORTED.
extern "C" int dpl_conv1x1_px_tf32x3(const float* d_x, const float* d_w, const float* d_w_lo, float* d_y, int n_img,
                                     int c_in, int c_out, int hw, const float* d_bias, float* d_y_relu,
                                     float* d_blob_min, float* d_blob_max, float* d_relu_min, float* d_relu_max,
                                     int* d_error_flag, void* stream) {
  DPL_REQUIRE(d_x && d_w && d_w_lo && d_y, "null pointer");
  DPL_REQUIRE(n_img > 0 && c_in > 0 && c_out > 0 && hw > 0, "empty problem");
  DPL_REQUIRE(n_img <= 65535 && (hw + kBM - 1) / kBM <= 65535, "grid limit");
  CUtensorMap tmX, tmW, tmWlo;
  int st = make_map_px(&tmX, d_x, (uint64_t)hw, (uint64_t)c_in, (uint64_t)n_img);
  if (st) return st;
  if (x3_chunk_iters() > 0) {
    // persistent kernel, chunked accumulation (dpl_x3ts.cuh): 128 px x (64 | 128) co tiles
    const int bn = c_out <= 64 ? 64 : 128;
    st = make_map(&tmW, d_w, (uint64_t)c_in, (uint64_t)c_out, 1, (uint64_t)c_in, 0, (uint32_t)bn, false);
    if (!st) st = make_map(&tmWlo, d_w_lo, (uint64_t)c_in, (uint64_t)c_out, 1, (uint64_t)c_in, 0, (uint32_t)bn, false);
    if (st) return st;
    ConvParams c = ConvParams();
    c.n_img = n_img;
    c.c_in = c_in;
    c.c_out = c_out;
    c.H = hw;
    c.W = 1;
    c.Wp = 1;
    c.plane = hw;
    c.origin = 0;
    c.n_taps = 1;
    c.bn = bn;
    c.q_total = (long long)n_img * hw;
    c.Y = d_y;
    c.bias = d_bias;
    c.relu = 0;
    c.error_flag = d_error_flag;
    c.Y2 = d_y_relu;
    c.bmin = d_blob_min;
    c.bmax = d_blob_max;
    c.rmin = d_relu_min;
    c.rmax = d_relu_max;
    c.os = 1;
    const long long total = (long long)((hw + kBM - 1) / kBM) * n_img * ((c_out + bn - 1) / bn);
    int e = launch_x3ts<1>(tmX, tmW, tmWlo, c, total, static_cast<cudaStream_t>(stream));
    if (e) return e;
    DPL_LAUNCH_CHECK("x3ts_kernel<1>");
    return 0;
  }
  DPL_REQUIRE(!d_blob_min && !d_blob_max && !d_relu_min && !d_relu_max,
              "fused range statistics need the chunked kernel (DPL_X3_CHUNK > 0, DPL_X3_TS = 1)");
  st = make_map(&tmW, d_w, (uint64_t)c_in, (uint64_t)c_out, 1, (uint64_t)c_in, 0, (uint32_t)kPxBN, false);
  if (!st) st = make_map(&tmWlo, d_w_lo, (uint64_t)c_in, (uint64_t)c_out, 1, (uint64_t)c_in, 0, (uint32_t)kPxBN, false);
  if (st) return st;
  PxParams p;
  p.n_img = n_img;
  p.c_in = c_in;
  p.c_out = c_out;
  p.hw = hw;
  p.Y = d_y;
  p.Y2 = d_y_relu;
  p.bias = d_bias;
  p.error_flag = d_error_flag;
  // output-channel groups fastest: the CTAs that share an X tile run back to back (L2 reuse)
  dim3 grid((unsigned)((c_out + kPxBN - 1) / kPxBN), (unsigned)((hw + kBM - 1) / kBM), (unsigned)n_img);
  const size_t smem = (size_t)kPxStages * kPxStageBytes + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    int e = cuda_status(cudaFuncSetAttribute(conv1x1_px_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem),
                        "cudaFuncSetAttribute(conv1x1_px_tf32x3_kernel)");
    if (e) return e;
    attr_done = true;
  }
  conv1x1_px_tf32x3_kernel<<<grid, kPxThreads, smem, static_cast<cudaStream_t>(stream)>>>(tmX, tmW, tmWlo, p);
  DPL_LAUNCH_CHECK("conv1x1_px_tf32x3_kernel");
  return 0;
}

// im2col staging copy for the stem convolution (see im2col_kernel): d_xp [n_img * Ho * Wo][k_pad],
// k_pad >= C * kh * kw, a multiple of 4 and <= 256.
extern "C" int dpl_im2col_f32(const float* d_x, float* d_xp, int n_img, int channels, int H, int W, int kh, int kw,
                              int stride, int pad, int Ho, int Wo, int k_pad, void* stream) {
  DPL_REQUIRE(d_x && d_xp, "null pointer");
  DPL_REQUIRE(n_img > 0 && channels > 0 && H > 0 && W > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0 && Ho > 0 &&
                  Wo > 0,
              "bad geometry");
  DPL_REQUIRE(k_pad >= channels * kh * kw && k_pad <= kIm2colMaxK && (k_pad & 3) == 0, "k_pad out of range");
  const long long rows = (long long)n_img * Ho * Wo;
  long long blocks = (rows + 7) / 8;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  im2col_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_x, d_xp, n_img, channels, H, W, kh,
                                                                                 kw, stride, pad, Ho, Wo, k_pad);
  DPL_LAUNCH_CHECK("im2col_kernel");
  return 0;
}
