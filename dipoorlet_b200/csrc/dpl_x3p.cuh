// Persistent 3xTF32 contraction with CHUNKED accumulation (included by dpl_gemm.cu, inside namespace dpl::{anon}).
//
// Why: tcgen05.mma adds every K step into the fp32 accumulator with TRUNCATION (the bits shifted out when the
// products are aligned to the accumulator are dropped), so a long accumulation chain shrinks the result
// systematically - measured 2e-8 relative per K = 8 step (tools/x3_accuracy.py: -6.7e-6 after the 576 steps of a
// 3x3 / 512-channel convolution on same-sign data). One such layer is harmless, but ResNet-50 stacks 53 of them
// and the shrink compounds: the calibration forward drifted 1e-6 per layer, 5e-5 at the last blob, against the
// reference's fp32 CPU forward (tests/test_gpu_fullsize_parity.py) - beyond the 1e-5 the clip file is compared at.
// Here the tensor core only ever accumulates `chunk_iters` K blocks (default 2 = 8 K steps) in TMEM; four
// dedicated epilogue warps drain every chunk into REGISTER accumulators with round-to-nearest fp32 adds while
// the MMA thread fills the other TMEM buffer. The operand split is made unbiased as well: the weights' leading
// part is rounded to nearest on the host (w_hi = RN_tf32(w), w_lo = RN_tf32(w - w_hi)); the activations keep
// the truncated fp32 pattern as their leading part (that is what kind::tf32 reads) and the transform warps
// round the residual x - trunc(x) to TF32 to nearest, so the tensor core's own truncation of it is exact.
//
// One CTA per SM walks a static tile list (persistent): TMA ring and TMEM buffers run across tile
// boundaries, so a tile's epilogue (global stores) overlaps the next tile's main loop.
//   warps 0-3   control: warp 0 / lane 0 TMA producer, warp 1 / lane 0 MMA issuer, warp 2 TMEM allocation
//   warps 4-7   transform: lo tile of the activation operand, in shared memory (generic -> async proxy fence)
//   warps 8-15  chunk drain + epilogue: TMEM lane quarter = warp % 4, warps 8-11 own the first half of the tile's
//               columns and warps 12-15 the second (64 accumulator registers per thread; with four warps the
//               epilogue of the short-K layers - 64 KB of stores per tile - was the bottleneck, 10 K clk per tile);
//               setmaxnreg moves registers from the other two warp groups to these two.
// MODE 0 (1x1 convolution / GEMM): D[z][m][n] = A[m][k] (weights, K-major, hi + lo by TMA) x B[z][k][n]
//         (activations, MN-major: NCHW pixels contiguous), output channels on the TMEM lanes.
// MODE 1 (tap-table convolution): D[q][co] = sum_tap Xp[q + shift(tap)][ci] (activations, K-major) x
//         Wt[tap][co][ci] (weights, hi + lo by TMA), padded-plane pixels on the TMEM lanes (see ConvParams).

constexpr int kX3PThreads = 512;
constexpr int kX3PStages = 3;
constexpr int kX3PStageBytes = 4 * kTileBytes;                 // 64 KB
constexpr int kX3PStgBytes = 8 * 32 * 32 * 4;                  // MODE 0 store staging, one XOR-swizzled 32 x 32 tile per warp

struct X3PParams {
  int probe_skip_wlo;   // timing probe only (DPL_X3_PROBE_SKIP_WLO=1): do not load the W_lo tiles - WRONG results
  int single;           // 1: single-pass TF32 (leading term only, no residual tiles) - the reconstruction loop's numerics
  int chunk_iters;      // K blocks (of 32) accumulated in TMEM before a drain
  GemmParams g;         // MODE 0
  ConvParams c;         // MODE 1
};

__device__ __forceinline__ float tf32_round_nearest(float x) {
  uint32_t u = __float_as_uint(x);
  u += 0xFFFu + ((u >> 13) & 1u);
  return __uint_as_float(u & 0xFFFFE000u);
}
// residual of the truncated pattern, itself rounded to TF32 (to nearest): the tensor core reads it exactly
__device__ __forceinline__ float tf32_residual_rn(float x) {
  return tf32_round_nearest(x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u));
}

template <int MODE>
__global__ void __launch_bounds__(kX3PThreads, 1)
x3p_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmWlo,
           const __grid_constant__ CUtensorMap tmX, const X3PParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_full[kX3PStages], s_ready[kX3PStages], s_empty[kX3PStages], s_acc_full[2],
      s_acc_empty[2];
  __shared__ uint32_t s_tmem_base;
  __shared__ int s_fail;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tiles = (smem_addr(smem_raw) + 1023u) & ~1023u;
  uint8_t* tiles_ptr = smem_raw + (tiles - smem_addr(smem_raw));

  // ---- work decomposition -------------------------------------------------------------------------------
  int num_kb, iters, total_tiles, m_tiles = 1, n_tiles = 1, bn;
  if (MODE == 0) {
    num_kb = (p.g.K + kBK - 1) / kBK;
    iters = num_kb;
    m_tiles = (p.g.M + kBM - 1) / kBM;
    n_tiles = (p.g.N + kBN - 1) / kBN;
    total_tiles = m_tiles * n_tiles * p.g.batch;
    bn = kBN;
  } else {
    num_kb = (p.c.c_in + kBK - 1) / kBK;
    iters = p.c.n_taps * num_kb;
    bn = p.c.bn;
    m_tiles = (int)((p.c.q_total + kBM - 1) / kBM);
    n_tiles = (p.c.c_out + bn - 1) / bn;
    total_tiles = m_tiles * n_tiles;
  }
  const int chunk_iters = p.chunk_iters;
  int* error_flag = MODE == 0 ? p.g.error_flag : p.c.error_flag;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kX3PStages; ++s) {
      bar_init(smem_addr(&s_full[s]), 1);
      bar_init(smem_addr(&s_ready[s]), 4);     // one arrival per transform warp
      bar_init(smem_addr(&s_empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      bar_init(smem_addr(&s_acc_full[b]), 1);
      bar_init(smem_addr(&s_acc_empty[b]), 8);   // one arrival per drain warp
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

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;" ::: "memory");
    if (warp == 0 && lane == 0) {
      // ===== TMA producer =====
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles && !s_fail; tile += gridDim.x) {
        int m0, n0, z = 0;
        if (MODE == 0) {
          const int mt = tile % m_tiles, nt = (tile / m_tiles) % n_tiles;
          z = tile / (m_tiles * n_tiles);
          m0 = mt * kBM;
          n0 = nt * kBN;
        } else {
          const int nt = tile % n_tiles, mt = tile / n_tiles;      // output-channel groups of one pixel tile back to back
          m0 = mt * kBM;
          n0 = nt * bn;
        }
        for (int i = 0; i < iters; ++i, ++it) {
          const int s = it % kX3PStages;
          const uint32_t ph = (it / kX3PStages) & 1;
          if (!bar_wait(smem_addr(&s_empty[s]), ph ^ 1)) {
            s_fail = 1;
            break;
          }
          const uint32_t full = smem_addr(&s_full[s]);
          const uint32_t t0 = tiles + s * kX3PStageBytes, t1 = t0 + kTileBytes, t2 = t0 + 2 * kTileBytes,
                         t3 = t0 + 3 * kTileBytes;
          if (MODE == 0) {
            bar_expect_tx(full, ((p.single || p.probe_skip_wlo) ? 2 : 3) * kTileBytes);
            const int k0 = i * kBK;
            tma_load_3d(t0, &tmW, k0, m0, 0, full);
            if (!p.single && !p.probe_skip_wlo) tma_load_3d(t1, &tmWlo, k0, m0, 0, full);
#pragma unroll
            for (int j = 0; j < kBN / 32; ++j) tma_load_3d(t2 + j * (kBK * 128), &tmX, n0 + 32 * j, k0, z, full);
          } else {
            bar_expect_tx(full, kTileBytes + ((p.single || p.probe_skip_wlo) ? 1u : 2u) * (uint32_t)bn * 128u);
            const int kb = i / p.c.n_taps, tap = i - kb * p.c.n_taps;
            const int k0 = kb * kBK;
            tma_load_3d(t0, &tmX, k0, m0 + p.c.tap_shift[tap], 0, full);
            tma_load_3d(t2, &tmW, k0, n0, p.c.tap_w[tap], full);
            if (!p.single && !p.probe_skip_wlo)
              tma_load_3d(t2 + (uint32_t)bn * 128u, &tmWlo, k0, n0, p.c.tap_w[tap], full);   // right behind W_hi
          }
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ===== MMA issuer: chunks of chunk_iters K blocks alternate between the two TMEM buffers =====
      // TWO instructions per K step instead of three: the residual tile of the N-side operand sits right behind its
      // leading tile in shared memory (MODE 0: X | X_lo, MODE 1: W_hi | W_lo), so
      //     A_hi x [B_hi | B_lo]   is ONE MMA of N = 2 bn into the accumulator pair [acc_hi | acc_lo], and
      //     A_lo x  B_hi           a second one of N = bn into acc_lo.
      // Same products, same accumulators, but A_hi is read from shared memory once instead of twice (20 KB instead
      // of 24 KB per K step through the port that bounds this kernel) and a third fewer instructions to issue.
      // Descriptors: tiles are 1024-byte aligned and every operand offset is a multiple of 16 bytes below 256 KB,
      // so a descriptor is base + (offset >> 4) on its low word - no per-instruction field packing.
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((MODE == 0 ? 1u : 0u) << 16) |
                             ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
      const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((MODE == 0 ? 1u : 0u) << 16) |
                              ((uint32_t)((2 * bn) >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
      const uint64_t base_k = desc_k_major(tiles, 0), base_mn = desc_mn_major(tiles, 0);
      const uint64_t base_b = MODE == 0 ? base_mn : base_k;
      constexpr uint32_t kStepA = (kUmmaK * 4) >> 4;                       // K-major: 32 bytes per K step
      constexpr uint32_t kStepB = MODE == 0 ? (1024u >> 4) : kStepA;       // MN-major: 1024 bytes per K step
      int it = 0, g = 0;
      bool failed = false;
      for (int tile = blockIdx.x; tile < total_tiles && !failed && !s_fail; tile += gridDim.x) {
        for (int i = 0; i < iters && !failed;) {
          const int buf = g & 1;
          if (!bar_wait(smem_addr(&s_acc_empty[buf]), ((g >> 1) & 1) ^ 1)) {
            failed = true;
            break;
          }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t acc_hi = tmem_acc + (uint32_t)(buf * 256), acc_lo = acc_hi + (uint32_t)bn;
          const int cend = min(iters, i + chunk_iters);
          const int cbeg = i;
          for (; i < cend; ++i, ++it) {
            const int s = it % kX3PStages;
            const uint32_t ph = (it / kX3PStages) & 1;
            if (!bar_wait(smem_addr(p.single ? &s_full[s] : &s_ready[s]), ph)) {   // single: no transform step
              failed = true;
              break;
            }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t soff = (uint32_t)s * (uint32_t)(kX3PStageBytes >> 4);
            const uint64_t d0 = base_k + soff;                                   // tile 0: A leading part
            const uint64_t d1 = base_k + soff + (uint32_t)(kTileBytes >> 4);     // tile 1: A residual
            const uint64_t d2 = base_b + soff + (uint32_t)(2 * kTileBytes >> 4); // tile 2 (+ 3): B leading part | residual
            if (p.single) {
#pragma unroll
              for (int j = 0; j < kBK / kUmmaK; ++j) {
                const uint32_t accumulate = (i > cbeg || j > 0) ? 1u : 0u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "setp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                    ::"r"(acc_hi), "l"(d0 + j * kStepA), "l"(d2 + j * kStepB), "r"(idesc), "r"(accumulate)
                    : "memory");
              }
            } else {
#pragma unroll
              for (int j = 0; j < kBK / kUmmaK; ++j) {
                const uint32_t accumulate = (i > cbeg || j > 0) ? 1u : 0u;
                asm volatile(
                    "{\n\t.reg .pred p, t;\n\t"
                    "setp.ne.b32 p, %7, 0;\n\t"
                    "setp.eq.b32 t, %5, %5;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %2, %4, %6, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%1], %3, %4, %5, t;\n\t}"
                    ::"r"(acc_hi), "r"(acc_lo), "l"(d0 + j * kStepA), "l"(d1 + j * kStepA), "l"(d2 + j * kStepB),
                      "r"(idesc), "r"(idesc2), "r"(accumulate)
                    : "memory");
              }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                             smem_addr(&s_empty[s]))
                         : "memory");
          }
          if (failed) break;
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                           smem_addr(&s_acc_full[buf]))
                       : "memory");
          ++g;
        }
      }
      if (failed) s_fail = 1;
    }
  } else if (warp < 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;" ::: "memory");
    // ===== transform warps: lo tile of the activation operand =====
    const int tt = threadIdx.x - 128;
    const int src_off = MODE == 0 ? 2 * kTileBytes : 0, dst_off = MODE == 0 ? 3 * kTileBytes : kTileBytes;
    int it = 0;
    bool failed = false;
    for (int tile = blockIdx.x; tile < total_tiles && !failed && !s_fail && !p.single; tile += gridDim.x) {
      for (int i = 0; i < iters; ++i, ++it) {
        const int s = it % kX3PStages;
        const uint32_t ph = (it / kX3PStages) & 1;
        if (!bar_wait(smem_addr(&s_full[s]), ph)) {
          failed = true;
          break;
        }
        const float4* src = reinterpret_cast<const float4*>(tiles_ptr + s * kX3PStageBytes + src_off);
        float4* dst = reinterpret_cast<float4*>(tiles_ptr + s * kX3PStageBytes + dst_off);
#pragma unroll
        for (int j = 0; j < kTileBytes / 16 / 128; ++j) {
          const float4 v = src[tt + j * 128];
          dst[tt + j * 128] = make_float4(tf32_residual_rn(v.x), tf32_residual_rn(v.y), tf32_residual_rn(v.z),
                                          tf32_residual_rn(v.w));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0)
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&s_ready[s])) : "memory");
      }
    }
    if (failed) s_fail = 1;
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 168;" ::: "memory");
    // ===== drain + epilogue warps =====
    const int wq = warp & 3, half = (warp - 8) >> 2;
    const uint32_t lane_base = tmem_acc + ((uint32_t)(wq * 32) << 16);
    const int n_chunks = (iters + chunk_iters - 1) / chunk_iters;
    const int g_per = (bn / 32) / 2;          // column groups of 32 per warp: 2 (bn = 128) or 1 (bn = 64)
    const int g_first = half * g_per;
    int g = 0;
    float rlo = INFINITY, rhi = -INFINITY;   // range of everything this thread stores, flushed once
    float* stg = reinterpret_cast<float*>(tiles_ptr + kX3PStages * kX3PStageBytes) + (warp - 8) * (32 * 32);
    bool failed = false;
    for (int tile = blockIdx.x; tile < total_tiles && !failed; tile += gridDim.x) {
      float acc[64];
      for (int c = 0; c < n_chunks; ++c, ++g) {
        const int buf = g & 1;
        bool ok = bar_wait(smem_addr(&s_acc_full[buf]), (g >> 1) & 1);
        ok = __all_sync(0xffffffffu, ok);
        if (!ok) {
          failed = true;
          break;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int gi = 0; gi < 2; ++gi) {
          if (gi < g_per) {
            const int grp = g_first + gi;
            uint32_t r[32], r2[32];
            tmem_ld32(lane_base + (uint32_t)(buf * 256 + grp * 32), r);
            if (!p.single) {
              tmem_ld32(lane_base + (uint32_t)(buf * 256 + bn + grp * 32), r2);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) r2[j] = 0u;
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (c == 0) {
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[gi * 32 + j] = __uint_as_float(r[j]) + __uint_as_float(r2[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                acc[gi * 32 + j] += __uint_as_float(r[j]) + __uint_as_float(r2[j]);
            }
          }
        }
        // all TMEM reads of this buffer are done: hand it back
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0)
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&s_acc_empty[buf])) : "memory");
      }
      if (failed) break;

      if (MODE == 0) {
        const GemmParams& q = p.g;
        const int mt = tile % m_tiles, nt = (tile / m_tiles) % n_tiles, z = tile / (m_tiles * n_tiles);
        const int m0 = mt * kBM, n0 = nt * kBN;
        const int m = m0 + wq * 32 + lane;
        float* drow = q.D + (long long)z * q.d_batch_stride + (long long)m * q.ldd;
        const float bias_m = (q.bias_mode == 1 && m < q.M) ? q.bias[m] : 0.f;
#pragma unroll
        for (int gi = 0; gi < 2; ++gi) {
          const int nc = n0 + (g_first + gi) * 32;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = acc[gi * 32 + j] + bias_m;
            if (q.relu) v[j] = fmaxf(v[j], 0.f);
            if (m < q.M && nc + j < q.N) {
              rlo = fminf(rlo, v[j]);
              rhi = fmaxf(rhi, v[j]);
            }
          }
          // through an XOR-swizzled shared-memory tile: 4 rows x 128 contiguous bytes per store instruction
          float* blk = q.D + (long long)z * q.d_batch_stride + (long long)(m0 + wq * 32) * q.ldd + nc;
          float* blk2 = q.D2 ? q.D2 + (blk - q.D) : nullptr;
          const bool fullblk = (m0 + wq * 32 + 32 <= q.M) && (nc + 32 <= q.N) && ((q.ldd & 3) == 0) &&
                               ((reinterpret_cast<uintptr_t>(blk) & 15u) == 0) &&
                               ((reinterpret_cast<uintptr_t>(blk2) & 15u) == 0);
          if (fullblk) {
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(stg + lane * 32 + (((j >> 2) ^ (lane & 7)) << 2)) =
                  make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            __syncwarp();
            const int rr = lane >> 3, c8 = lane & 7;
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) {
              const int row = 4 * qq + rr;
              const float4 t = *reinterpret_cast<const float4*>(stg + row * 32 + ((c8 ^ (row & 7)) << 2));
              *reinterpret_cast<float4*>(blk + (long long)row * q.ldd + c8 * 4) = t;
              if (blk2)
                *reinterpret_cast<float4*>(blk2 + (long long)row * q.ldd + c8 * 4) =
                    make_float4(relu_keep_nan(t.x), relu_keep_nan(t.y), relu_keep_nan(t.z), relu_keep_nan(t.w));
            }
          } else if (m < q.M) {
            float* dst = drow + nc;
            float* dst2 = q.D2 ? q.D2 + (drow - q.D) + nc : nullptr;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nc + j < q.N) {
                dst[j] = v[j];
                if (dst2) dst2[j] = relu_keep_nan(v[j]);
              }
          }
        }
      } else {
        const ConvParams& q = p.c;
        const int nt = tile % n_tiles, mt = tile / n_tiles;
        const long long pq = (long long)mt * kBM + wq * 32 + lane;
        const int co0 = nt * bn;
        bool valid = pq < q.q_total;
        long long out_base = 0;
        if (valid) {
          const int img = (int)(pq / q.plane);
          const int r = (int)(pq - (long long)img * q.plane);
          const int hp = r / q.Wp, wp = r - hp * q.Wp;
          const int hq = hp - q.origin, wq2 = wp - q.origin;
          const int ho = hq * q.os + q.oa, wo = wq2 * q.os + q.ob;
          valid = hq >= 0 && wq2 >= 0 && ho < q.H && wo < q.W;
          out_base = (((long long)img * q.c_out) * q.H + ho) * q.W + wo;
        }
        const long long ch_stride = (long long)q.H * q.W;
        if (valid) {
#pragma unroll
          for (int gi = 0; gi < 2; ++gi) {
            if (gi < g_per) {
              const int cb = co0 + (g_first + gi) * 32;
              float* dst = q.Y + out_base + (long long)cb * ch_stride;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (cb + j < q.c_out) {
                  float v = acc[gi * 32 + j];
                  if (q.bias) v += __ldg(q.bias + cb + j);
                  if (q.relu) v = fmaxf(v, 0.f);
                  dst[(long long)j * ch_stride] = v;
                  if (q.Y2) q.Y2[(dst - q.Y) + (long long)j * ch_stride] = relu_keep_nan(v);
                  rlo = fminf(rlo, v);
                  rhi = fmaxf(rhi, v);
                }
              }
            }
          }
        }
      }
    }
    if (failed) s_fail = 1;
    if (MODE == 0)
      warp_range_flush(rlo, rhi, p.g.bmin, p.g.bmax, p.g.rmin, p.g.rmax);
    else
      warp_range_flush(rlo, rhi, p.c.bmin, p.c.bmax, p.c.rmin, p.c.rmax);
  }
  __syncwarp();
  if (s_fail) {
    if ((threadIdx.x & 31) == 0 && error_flag) atomicExch(error_flag, 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(512) : "memory");
  }
}

// K blocks per TMEM chunk (DPL_X3_CHUNK, default 2 = 8 K steps of the tensor core; 0 = the one-accumulator kernels).
inline int x3_chunk_iters() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DPL_X3_CHUNK");
    v = e ? atoi(e) : 2;
    if (v < 0) v = 0;
  }
  return v;
}

template <int MODE>
int launch_x3p(const CUtensorMap& tmW, const CUtensorMap& tmWlo, const CUtensorMap& tmX, const X3PParams& p_in,
               long long total_tiles, cudaStream_t s) {
  static const int probe = [] {
    const char* e = getenv("DPL_X3_PROBE_SKIP_WLO");
    return (e && e[0] == '1') ? 1 : 0;
  }();
  X3PParams p = p_in;
  p.probe_skip_wlo = probe;
  const size_t smem = (size_t)kX3PStages * kX3PStageBytes + kX3PStgBytes + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    int e = cuda_status(cudaFuncSetAttribute(x3p_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                        "cudaFuncSetAttribute(x3p_kernel)");
    if (e) return e;
    attr_done = true;
  }
  const unsigned ctas = (unsigned)(total_tiles < sm_count() ? total_tiles : sm_count());
  x3p_kernel<MODE><<<ctas, kX3PThreads, smem, s>>>(tmW, tmWlo, tmX, p);
  return 0;
}
