// Persistent 3xTF32 convolution with the ACTIVATIONS THROUGH TMEM and chunked accumulation (included by
// dpl_gemm.cu inside namespace dpl::{anon}, after dpl_x3p.cuh). Pixels on the TMEM lanes (M), output channels on N:
//     D[px][co] += A[px][ci] (TMEM operand)  x  B[co][ci]^T (weights, K-major, 128-byte swizzle, hi + lo by TMA)
//
// dpl_x3p.cuh made the forward unbiased (chunked accumulation, round-to-nearest split) but both operands still
// come from shared memory: per 32-channel K block the tensor core reads 96 KB (three MMAs x two operands), TMA
// writes 48 KB and the transform warps read + write 32 KB - 176 KB through a 128 B/clk port = 1375 clk against
// 768 clk of tensor work (ncu: tensor pipe 41 - 47 % active, profiles/r2_x3p_ncu.md). Here the transform warps
// read their own row of the landed activation tile once, split it in registers into the truncated TF32 pattern
// and its rounded residual and tcgen05.st both into TMEM; the MMAs take A from TMEM and only the weight tiles
// from shared memory: 48 (TMA) + 16 (transform read) + 48 (MMA reads of W) = 112 KB per K block.
//
// One accumulator per chunk: the chunk's small cross terms (A_lo W_hi + A_hi W_lo) are issued FIRST, into the
// fresh accumulator, then its leading terms - the truncation of every accumulate step scales with the
// accumulator's magnitude, so the cross terms cost nothing while it is still small and the leading terms see
// the same chain as with a separate accumulator. TMEM: 2 x {A_hi, A_lo} x 32 + 2 x bn accumulator columns.
// A chunk is at most two K blocks (= the two A buffers).
//   warps 0-3   control (TMA producer, MMA issuer, TMEM allocation), 4-stage ring of 48 KB
//   warps 4-7   transform: activation row -> {hi, lo} -> TMEM
//   warps 8-11  chunk drain into 128 registers (round-to-nearest adds) + epilogue; consecutive lanes are
//               consecutive pixels, so every store of an output channel is a coalesced NCHW row segment
// PX = 0: tap-table convolution over the channel-last staging copy (ConvParams; X tile [128 q][32 ci], SW128)
// PX = 1: 1x1 / stride 1 convolution straight from NCHW (X tile [32 ci][128 px], unswizzled; q_total, plane =
//         H * W, Wp = W, origin = 0, one tap with shift 0 in ConvParams; pixel tiles do not cross images)

constexpr int kTsStages = 4;
constexpr int kTsStageBytes = 3 * kTileBytes;    // X | W_hi | W_lo
constexpr int kTsThreads = 384;

template <int PX>
__global__ void __launch_bounds__(kTsThreads, 1)
x3ts_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
            const __grid_constant__ CUtensorMap tmWlo, const ConvParams p, const int chunk_iters) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_full[kTsStages], s_empty[kTsStages], s_aready[2], s_aempty[2], s_acc_full[2],
      s_acc_empty[2];
  __shared__ uint32_t s_tmem_base;
  __shared__ int s_fail;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tiles = (smem_addr(smem_raw) + 1023u) & ~1023u;
  const uint8_t* tiles_ptr = smem_raw + (tiles - smem_addr(smem_raw));
  const int bn = p.bn;
  const int num_kb = (p.c_in + kBK - 1) / kBK;
  const int iters = p.n_taps * num_kb;
  const int n_tiles = (p.c_out + bn - 1) / bn;
  // PX = 1: pixel tiles per image (a tile never crosses an image); PX = 0: tiles over the whole padded index
  const int px_tiles = PX ? (p.plane + kBM - 1) / kBM : 0;
  const long long m_tiles = PX ? (long long)px_tiles * p.n_img : (p.q_total + kBM - 1) / kBM;
  const long long total_tiles = m_tiles * n_tiles;
  const uint32_t w_bytes = (uint32_t)bn * 128u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTsStages; ++s) {
      bar_init(smem_addr(&s_full[s]), 1);
      bar_init(smem_addr(&s_empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      bar_init(smem_addr(&s_aready[b]), 4);      // one arrival per transform warp
      bar_init(smem_addr(&s_aempty[b]), 1);
      bar_init(smem_addr(&s_acc_full[b]), 1);
      bar_init(smem_addr(&s_acc_empty[b]), 4);   // one arrival per drain warp
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
  const uint32_t tmem = s_tmem_base;             // columns 0..127: A buffers; 128..: two accumulators of bn columns

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;" ::: "memory");
    if (warp == 0 && lane == 0) {
      // ===== TMA producer =====
      long long it = 0;
      for (long long tile = blockIdx.x; tile < total_tiles && !s_fail; tile += gridDim.x) {
        const int nt = (int)(tile % n_tiles);
        const long long mt = tile / n_tiles;
        const int co0 = nt * bn;
        int img = 0, px0 = 0;
        long long q0 = 0;
        if (PX) {
          img = (int)(mt / px_tiles);
          px0 = (int)(mt - (long long)img * px_tiles) * kBM;
        } else {
          q0 = mt * kBM;
        }
        for (int i = 0; i < iters; ++i, ++it) {
          const int s = (int)(it % kTsStages);
          const uint32_t ph = (uint32_t)((it / kTsStages) & 1);
          if (!bar_wait(smem_addr(&s_empty[s]), ph ^ 1)) {
            s_fail = 1;
            break;
          }
          const uint32_t full = smem_addr(&s_full[s]);
          bar_expect_tx(full, kTileBytes + 2u * w_bytes);
          const int kb = i / p.n_taps, tap = i - kb * p.n_taps;
          const int k0 = kb * kBK;
          const uint32_t x_tile = tiles + s * kTsStageBytes, w_tile = x_tile + kTileBytes, wlo_tile = w_tile + kTileBytes;
          if (PX)
            tma_load_3d(x_tile, &tmX, px0, k0, img, full);
          else
            tma_load_3d(x_tile, &tmX, k0, (int)q0 + p.tap_shift[tap], 0, full);
          tma_load_3d(w_tile, &tmW, k0, co0, tap, full);
          tma_load_3d(wlo_tile, &tmWlo, k0, co0, tap, full);
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ===== MMA issuer: per chunk the cross terms of all its K blocks first, then the leading terms =====
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) |
                             ((uint32_t)(kBM >> 4) << 24);      // f32 accumulate, tf32 x tf32, B K-major
      long long it = 0;
      int g = 0;
      bool failed = false;
      for (long long tile = blockIdx.x; tile < total_tiles && !failed && !s_fail; tile += gridDim.x) {
        for (int i = 0; i < iters && !failed;) {
          const int buf = g & 1;
          if (!bar_wait(smem_addr(&s_acc_empty[buf]), ((g >> 1) & 1) ^ 1)) {
            failed = true;
            break;
          }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t acc = tmem + 128u + (uint32_t)(buf * bn);
          const int n_it = min(chunk_iters, iters - i);
          // phase 1: cross terms
          for (int c = 0; c < n_it && !failed; ++c) {
            const long long t = it + c;
            const int s = (int)(t % kTsStages), a = (int)(t & 1);
            if (!bar_wait(smem_addr(&s_aready[a]), (uint32_t)((t >> 1) & 1))) {
              failed = true;
              break;
            }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t w_tile = tiles + s * kTsStageBytes + kTileBytes, wlo_tile = w_tile + kTileBytes;
            const uint32_t a_hi = tmem + (uint32_t)(a * 64), a_lo = a_hi + 32u;
#pragma unroll
            for (int j = 0; j < kBK / kUmmaK; ++j) {
              const uint64_t db = desc_k_major(w_tile, j), dbl = desc_k_major(wlo_tile, j);
              mma_tf32_ts(acc, a_lo + (uint32_t)(j * kUmmaK), db, idesc, (c > 0 || j > 0) ? 1u : 0u);
              mma_tf32_ts(acc, a_hi + (uint32_t)(j * kUmmaK), dbl, idesc, 1u);
            }
          }
          if (failed) break;
          // phase 2: leading terms; each K block's stage and A buffer are released behind its last MMA
          for (int c = 0; c < n_it; ++c) {
            const long long t = it + c;
            const int s = (int)(t % kTsStages), a = (int)(t & 1);
            const uint32_t w_tile = tiles + s * kTsStageBytes + kTileBytes;
            const uint32_t a_hi = tmem + (uint32_t)(a * 64);
#pragma unroll
            for (int j = 0; j < kBK / kUmmaK; ++j)
              mma_tf32_ts(acc, a_hi + (uint32_t)(j * kUmmaK), desc_k_major(w_tile, j), idesc, 1u);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                             smem_addr(&s_empty[s]))
                         : "memory");
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                             smem_addr(&s_aempty[a]))
                         : "memory");
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                           smem_addr(&s_acc_full[buf]))
                       : "memory");
          i += n_it;
          it += n_it;
          ++g;
        }
      }
      if (failed) s_fail = 1;
    }
  } else if (warp < 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 112;" ::: "memory");
    // ===== transform warps: own row of the X tile -> {TF32 pattern, rounded residual} -> TMEM =====
    const int quarter = warp - 4;                 // TMEM lane quarter (warp % 4)
    const int m = quarter * 32 + lane;            // pixel of this thread within the tile
    const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
    long long it = 0;
    bool ok = true;
    for (long long tile = blockIdx.x; tile < total_tiles && ok && !s_fail; tile += gridDim.x) {
      for (int i = 0; i < iters && ok; ++i, ++it) {
        const int s = (int)(it % kTsStages), a = (int)(it & 1);
        ok = bar_wait(smem_addr(&s_full[s]), (uint32_t)((it / kTsStages) & 1)) &&
             bar_wait(smem_addr(&s_aempty[a]), (uint32_t)(((it >> 1) & 1) ^ 1));
        ok = __all_sync(0xffffffffu, ok);
        if (!ok) break;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t hi[32], lo[32];
        if (PX) {
          // [32 ci][128 px] unswizzled: consecutive lanes read consecutive words
          const float* xs = reinterpret_cast<const float*>(tiles_ptr + s * kTsStageBytes) + m;
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float v = xs[k * kBM];
            hi[k] = __float_as_uint(v);
            lo[k] = __float_as_uint(tf32_residual_rn(v));
          }
        } else {
          // row m of a 128-byte-swizzled tile: logical 16-byte chunk c sits at chunk c ^ (m & 7)
          const uint8_t* row = tiles_ptr + s * kTsStageBytes + m * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(row + ((c ^ (m & 7)) << 4));
            hi[4 * c + 0] = __float_as_uint(v.x);
            hi[4 * c + 1] = __float_as_uint(v.y);
            hi[4 * c + 2] = __float_as_uint(v.z);
            hi[4 * c + 3] = __float_as_uint(v.w);
            lo[4 * c + 0] = __float_as_uint(tf32_residual_rn(v.x));
            lo[4 * c + 1] = __float_as_uint(tf32_residual_rn(v.y));
            lo[4 * c + 2] = __float_as_uint(tf32_residual_rn(v.z));
            lo[4 * c + 3] = __float_as_uint(tf32_residual_rn(v.w));
          }
        }
        tmem_st32(lane_base + (uint32_t)(a * 64), hi);
        tmem_st32(lane_base + (uint32_t)(a * 64 + 32), lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0)
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&s_aready[a])) : "memory");
      }
    }
    if (!ok) s_fail = 1;
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;" ::: "memory");
    // ===== drain + epilogue warps (TMEM lane quarter = warp - 8) =====
    const int wq = warp - 8;
    const uint32_t lane_base = tmem + ((uint32_t)(wq * 32) << 16);
    const int n_chunks = (iters + chunk_iters - 1) / chunk_iters;
    const int n_grp = bn / 32;
    int g = 0;
    float rlo = INFINITY, rhi = -INFINITY;
    bool failed = false;
    for (long long tile = blockIdx.x; tile < total_tiles && !failed; tile += gridDim.x) {
      float acc[128];
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
        for (int grp = 0; grp < 4; ++grp) {
          if (grp < n_grp) {
            uint32_t r[32];
            tmem_ld32(lane_base + 128u + (uint32_t)(buf * bn + grp * 32), r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (c == 0) {
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[grp * 32 + j] = __uint_as_float(r[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[grp * 32 + j] += __uint_as_float(r[j]);
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0)
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&s_acc_empty[buf])) : "memory");
      }
      if (failed) break;

      const int nt = (int)(tile % n_tiles);
      const long long mt = tile / n_tiles;
      const int co0 = nt * bn;
      bool valid;
      long long out_base;
      if (PX) {
        const int img = (int)(mt / px_tiles);
        const int px = (int)(mt - (long long)img * px_tiles) * kBM + wq * 32 + lane;
        valid = px < p.plane;
        out_base = (long long)img * p.c_out * p.plane + px;
      } else {
        const long long q = mt * kBM + wq * 32 + lane;
        valid = q < p.q_total;
        out_base = 0;
        if (valid) {
          const int img = (int)(q / p.plane);
          const int r = (int)(q - (long long)img * p.plane);
          const int hp = r / p.Wp, wp = r - hp * p.Wp;
          const int ho = hp - p.origin, wo = wp - p.origin;
          valid = ho >= 0 && ho < p.H && wo >= 0 && wo < p.W;
          out_base = (((long long)img * p.c_out) * p.H + ho) * p.W + wo;
        }
      }
      const long long ch_stride = (long long)p.H * p.W;
      if (valid) {
#pragma unroll
        for (int grp = 0; grp < 4; ++grp) {
          if (grp < n_grp) {
            const int cb = co0 + grp * 32;
            float* dst = p.Y + out_base + (long long)cb * ch_stride;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (cb + j < p.c_out) {
                float v = acc[grp * 32 + j];
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
      }
    }
    if (failed) s_fail = 1;
    warp_range_flush(rlo, rhi, p.bmin, p.bmax, p.rmin, p.rmax);
  }
  __syncwarp();
  if (s_fail) {
    if ((threadIdx.x & 31) == 0 && p.error_flag) atomicExch(p.error_flag, 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// Activations through TMEM for the tap-table convolutions: DPL_X3_TS=1 (opt-in; measured equal to slightly
// slower than dpl_x3p.cuh on ResNet-50's 3x3 layers: with one transform warp per scheduler the read - split -
// tcgen05.st - handshake chain of a K block is longer than its MMAs). dpl_conv1x1_px_tf32x3 always takes this kernel.
inline int x3_ts() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DPL_X3_TS");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v;
}

template <int PX>
int launch_x3ts(const CUtensorMap& tmX, const CUtensorMap& tmW, const CUtensorMap& tmWlo, const ConvParams& p,
                long long total_tiles, cudaStream_t s) {
  const size_t smem = (size_t)kTsStages * kTsStageBytes + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    int e = cuda_status(cudaFuncSetAttribute(x3ts_kernel<PX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                        "cudaFuncSetAttribute(x3ts_kernel)");
    if (e) return e;
    attr_done = true;
  }
  int chunk = x3_chunk_iters();
  if (chunk > 2) chunk = 2;      // a chunk spans at most the two A buffers
  if (chunk < 1) chunk = 1;
  const unsigned ctas = (unsigned)(total_tiles < sm_count() ? total_tiles : sm_count());
  x3ts_kernel<PX><<<ctas, kTsThreads, smem, s>>>(tmX, tmW, tmWlo, p, chunk);
  return 0;
}
