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

#include "dpl_common.cuh"

namespace dpl {
namespace {

constexpr int kBM = 128, kBN = 128, kBK = 32;       // tile; kBK floats = one 128-byte swizzle row
constexpr int kStages = 3;                           // 96 KB: two CTAs per SM overlap epilogue and main loop
constexpr int kTileBytes = kBM * kBK * 4;           // 16 KB per operand per stage
constexpr int kStageBytes = 2 * kTileBytes;
constexpr int kUmmaK = 8;                           // tf32: 32 bytes of K per instruction
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
  int* error_flag;
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait (~2 s at 2 GHz): returns false on timeout.
__device__ __forceinline__ bool bar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  uint32_t ok = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return true;
    if (clock64() - t0 > 4000000000ll) return false;
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}

// Shared-memory matrix descriptor (sm_100 format, cute::UMMA::SmemDescriptor): start >> 4 in
// [0,14), leading byte offset >> 4 in [16,30), stride byte offset >> 4 in [32,46), version 1 in
// [46,48), layout type in [61,64) (2 = 128-byte swizzle).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
// K-major, 128B swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart; K step = 32 bytes.
__device__ __forceinline__ uint64_t desc_k_major(uint32_t tile, int kstep) {
  return make_smem_desc(tile + kstep * (kUmmaK * 4), 16, 1024, 2 /* SWIZZLE_128B */);
}
// MN-major: 32-bit operands only exist in the "128-byte swizzle, 32-byte atom" layout
// (UMMA LayoutType 1, TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; cutlass sm100_common.inl: "for
// mn-major tf32 operands, SW128_32B is the only available smem layout"). The tile is 4 column
// blocks of [kBK k-rows][32 elements = 128 bytes]: 128-byte chunks along MN are 4096 bytes
// apart (LBO); the swizzle atom spans 4 k-rows, so groups along K are 512 bytes apart (SBO);
// one K = 8 instruction covers two groups = 1024 bytes.
__device__ __forceinline__ uint64_t desc_mn_major(uint32_t tile, int kstep) {
  return make_smem_desc(tile + kstep * 1024, kBK * 128, 512, 1 /* SWIZZLE_128B_BASE32B */);
}

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
constexpr int kGemm3Threads = 256;
constexpr int kTmemCols3 = 512;                     // three 128-column accumulators (power of two)

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
}

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
      bar_init(smem_addr(&s_ready[s]), 4);   // one arrival per transform warp
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
        const uint32_t acc_hi_first = (it > 1 || j > 0) ? 1u : 0u;
        const uint32_t acc_hi = tmem_acc + (uint32_t)((it & 1) * kBN);
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
      for (int j = 0; j < kTileBytes / 16 / 128; ++j) {
        const float4 v = src[t + j * 128];
        dst[t + j * 128] = make_float4(tf32_residual(v.x), tf32_residual(v.y), tf32_residual(v.z),
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
#pragma unroll 1
      for (int c = 0; c < kBN / 32; ++c) {
        uint32_t r[32], r1[32], r2[32];
        const uint32_t taddr = tmem_acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32);
        tmem_ld32(taddr, r);
        tmem_ld32(taddr + 2u * kBN, r2);
        if (total_iters > 1) tmem_ld32(taddr + kBN, r1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (m < p.M) {
          const int nc = n0 + c * 32;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float acc = __uint_as_float(r[j]);
            if (total_iters > 1) acc += __uint_as_float(r1[j]);
            v[j] = (acc + __uint_as_float(r2[j])) + bias_m;
            if (p.bias_mode == 2 && nc + j < p.N) v[j] += p.bias[nc + j];
            if (p.relu) v[j] = fmaxf(v[j], 0.f);
          }
          float* dst = drow + nc;
          if (nc + 32 <= p.N && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nc + j < p.N) dst[j] = v[j];
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(kTmemCols3)
                 : "memory");
  }
}

__global__ void __launch_bounds__(256)
tf32_residual_kernel(const float* __restrict__ x, float* __restrict__ lo, uint64_t n) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    lo[i] = tf32_residual(x[i]);
}

// ---- host: tensor maps ----------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 3-D fp32 tensor (inner, outer, batch) with a [32 x box_outer x 1] box and 128-byte swizzle.
int make_map(CUtensorMap* map, const float* base, uint64_t inner, uint64_t outer, uint64_t batch,
             uint64_t outer_stride_elems, uint64_t batch_stride_elems, uint32_t box_outer, bool mn_major) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return DPL_E_UNSUPPORTED;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15u) || (outer_stride_elems & 3u) ||
      (batch > 1 && (batch_stride_elems & 3u))) {
    set_error("operand not TMA-compatible: base 16-byte aligned and strides multiples of 4 floats required");
    return DPL_E_UNSUPPORTED;
  }
  cuuint64_t dims[3] = {inner, outer, batch ? batch : 1};
  cuuint64_t strides[2] = {outer_stride_elems * 4, (batch_stride_elems ? batch_stride_elems : outer * outer_stride_elems) * 4};
  cuuint32_t box[3] = {32, box_outer, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return DPL_E_UNSUPPORTED;
  }
  return 0;
}

}  // namespace
}  // namespace dpl

using namespace dpl;

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

// 3xTF32 variant of dpl_gemm_tf32 (fp32-accurate): same operand description plus d_a_lo, the
// residual of A with A's layout. No batch folding / split-K (forward use only).
extern "C" int dpl_gemm_tf32x3(const float* d_a, const float* d_a_lo, int a_major, long long lda,
                               long long a_batch_stride, const float* d_b, int b_major, long long ldb,
                               long long b_batch_stride, float* d_d, long long ldd, long long d_batch_stride,
                               int M, int N, int K, int batch, const float* d_bias, int bias_mode, int relu,
                               int* d_error_flag, void* stream) {
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
  p.error_flag = d_error_flag;
  dim3 grid((M + kBM - 1) / kBM, (N + kBN - 1) / kBN, (unsigned)batch);
  const size_t smem = (size_t)kStages3 * kStageBytes3 + 1024;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
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
