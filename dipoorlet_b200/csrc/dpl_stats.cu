// Activation-statistics kernels of the calibration hot path (sm_100a):
//   K1 dpl_segstats_f32     per-segment min / max / sum|x| / nnz      (HBM-read bound)
//   K2 dpl_hist_abs_f32     np.histogram(|x|, bins, (0, data_max))     (HBM-read bound)
//   K3 dpl_hist_percentile  percentile clip search over the histograms (latency bound)
// Every kernel walks ALL blobs of a batch in one launch through the dpl_blob table, so
// that a batch of 123 ResNet-50 blobs costs one launch instead of 123.
//
// Reference semantics restated here (and in oracle/stats.py):
//   dipoorlet/forward_net.py:220-235, 265-280; tensor_cali/basic_algorithm.py:37-53.

#include <math.h>

#include "dpl_common.cuh"

namespace dpl {
namespace {

// ===========================================================================
// K1: segment statistics
// ===========================================================================
constexpr int kSegThreads = 256;
constexpr int kSegWarps = kSegThreads / 32;
constexpr int kSegVec = DPL_SEG_TILE / 4 / kSegThreads;  // float4 per thread per tile = 8

struct SegAcc {
  float mn, mx, sum;
  uint32_t nnz;
  __device__ __forceinline__ void init() {
    mn = INFINITY;
    mx = -INFINITY;
    sum = 0.f;
    nnz = 0;
  }
  __device__ __forceinline__ void add(float x) {
    mn = fminf(mn, x);
    mx = fmaxf(mx, x);
    float a = fabsf(x);
    sum += a;
    nnz += (a > 0.f) ? 1u : 0u;
  }
};

// Tiles are assigned to CTAs in contiguous chunks, so the (blob, segment, tile-in-
// segment) cursor only ever moves forward and is tracked incrementally.
__global__ void __launch_bounds__(kSegThreads)
segstats_tiles_kernel(const dpl_blob* __restrict__ blobs, int n_blobs, uint64_t n_tiles,
                      float* __restrict__ tmin, float* __restrict__ tmax,
                      double* __restrict__ tsum, uint32_t* __restrict__ tnnz) {
  __shared__ float s_mn[kSegWarps], s_mx[kSegWarps];
  __shared__ double s_sum[kSegWarps];
  __shared__ uint32_t s_nnz[kSegWarps];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint64_t tile = n_tiles * blockIdx.x / gridDim.x;
  const uint64_t tile_end = n_tiles * (blockIdx.x + 1ull) / gridDim.x;
  if (tile >= tile_end) return;

  int b = find_blob<4>(blobs, n_blobs, tile);
  const float* base = nullptr;
  uint64_t seg_len = 0, tps = 1, seg = 0, t = 0;
  auto load_blob = [&](uint64_t local_tile) {
    base = reinterpret_cast<const float*>(blobs[b].ptr);
    seg_len = blobs[b].seg_len;
    tps = (seg_len + DPL_SEG_TILE - 1) / DPL_SEG_TILE;
    seg = local_tile / tps;
    t = local_tile - seg * tps;
  };
  load_blob(tile - blobs[b].seg_tile_begin);
  uint64_t blob_tile_end = (b + 1 < n_blobs) ? blobs[b + 1].seg_tile_begin : n_tiles;

  for (; tile < tile_end; ++tile) {
    while (tile >= blob_tile_end) {  // advance to the next non-empty blob
      ++b;
      blob_tile_end = (b + 1 < n_blobs) ? blobs[b + 1].seg_tile_begin : n_tiles;
      if (tile < blob_tile_end) load_blob(0);
    }
    const uint64_t off = t * DPL_SEG_TILE;
    const uint32_t len = (uint32_t)min((uint64_t)DPL_SEG_TILE, seg_len - off);
    const float* p = base + seg * seg_len + off;

    SegAcc acc;
    acc.init();
    // peel to 16-byte alignment, vector body, scalar tail
    const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(p) >> 2) & 3u);
    uint32_t head = (4u - mis) & 3u;
    if (head > len) head = len;
    const uint32_t n4 = (len - head) >> 2;
    const uint32_t tail = len - head - (n4 << 2);
    if ((uint32_t)tid < head) acc.add(ldg_stream1(p + tid));
    if ((uint32_t)tid < tail) acc.add(ldg_stream1(p + head + (n4 << 2) + tid));
    const float4* p4 = reinterpret_cast<const float4*>(p + head);
    float4 v[kSegVec];
#pragma unroll
    for (int u = 0; u < kSegVec; ++u) {
      const uint32_t i = u * kSegThreads + tid;
      if (i < n4) v[u] = ldg_stream4(p4 + i);
    }
#pragma unroll
    for (int u = 0; u < kSegVec; ++u) {
      const uint32_t i = u * kSegThreads + tid;
      if (i < n4) {
        acc.add(v[u].x);
        acc.add(v[u].y);
        acc.add(v[u].z);
        acc.add(v[u].w);
      }
    }
    // block reduction
    float mn = warp_min(acc.mn), mx = warp_max(acc.mx);
    double sm = warp_sum((double)acc.sum);
    uint32_t nz = __reduce_add_sync(0xffffffffu, acc.nnz);
    if (lane == 0) {
      s_mn[warp] = mn;
      s_mx[warp] = mx;
      s_sum[warp] = sm;
      s_nnz[warp] = nz;
    }
    __syncthreads();
    if (warp == 0) {
      mn = lane < kSegWarps ? s_mn[lane] : INFINITY;
      mx = lane < kSegWarps ? s_mx[lane] : -INFINITY;
      sm = lane < kSegWarps ? s_sum[lane] : 0.0;
      nz = lane < kSegWarps ? s_nnz[lane] : 0u;
      mn = warp_min(mn);
      mx = warp_max(mx);
      sm = warp_sum(sm);
      nz = __reduce_add_sync(0xffffffffu, nz);
      if (lane == 0) {
        tmin[tile] = mn;
        tmax[tile] = mx;
        tsum[tile] = sm;
        tnnz[tile] = nz;
      }
    }
    __syncthreads();
    if (++t == tps) {
      t = 0;
      ++seg;
    }
  }
}

// One warp per segment folds that segment's tile partials in a fixed order
// (deterministic sum), writes the per-segment outputs and folds the segment into the
// running per-blob extrema.
__global__ void __launch_bounds__(256)
segstats_finalize_kernel(const dpl_blob* __restrict__ blobs, int n_blobs, uint64_t n_segments,
                         const float* __restrict__ tmin, const float* __restrict__ tmax,
                         const double* __restrict__ tsum, const uint32_t* __restrict__ tnnz,
                         float* __restrict__ out_min, float* __restrict__ out_max,
                         double* __restrict__ out_abssum, uint64_t* __restrict__ out_nnz,
                         float* __restrict__ blob_min, float* __restrict__ blob_max) {
  const int lane = threadIdx.x & 31;
  const uint64_t s = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= n_segments) return;
  const int b = find_blob<3>(blobs, n_blobs, s);
  const uint64_t seg = s - blobs[b].seg_out_base;
  const uint64_t seg_len = blobs[b].seg_len;
  const uint64_t tps = (seg_len + DPL_SEG_TILE - 1) / DPL_SEG_TILE;
  const uint64_t t0 = blobs[b].seg_tile_begin + seg * tps;
  float mn = INFINITY, mx = -INFINITY;
  double sm = 0.0;
  uint64_t nz = 0;
  for (uint64_t i = lane; i < tps; i += 32) {
    mn = fminf(mn, tmin[t0 + i]);
    mx = fmaxf(mx, tmax[t0 + i]);
    sm += tsum[t0 + i];
    nz += tnnz[t0 + i];
  }
  mn = warp_min(mn);
  mx = warp_max(mx);
  sm = warp_sum(sm);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nz += __shfl_xor_sync(0xffffffffu, nz, o);
  if (lane == 0) {
    out_min[s] = mn;
    out_max[s] = mx;
    if (out_abssum) out_abssum[s] = sm;
    if (out_nnz) out_nnz[s] = nz;
    if (blob_min && seg_len > 0) atomic_min_f32(blob_min + blobs[b].stat_index, mn);
    if (blob_max && seg_len > 0) atomic_max_f32(blob_max + blobs[b].stat_index, mx);
  }
}

__global__ void absmax_kernel(const float* __restrict__ bmin, const float* __restrict__ bmax,
                              float* __restrict__ dm, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    // Python max(np.max(maxs), -np.min(mins)): keeps the first unless the second is greater
    float a = bmax[i], c = -bmin[i];
    dm[i] = (c > a) ? c : a;
  }
}

// ===========================================================================
// K2: |x| histogram, NumPy-exact binning
// ===========================================================================
// Bin of |x| in np.histogram(|x|, bins, (0, dm)) with float32 edges
//   edge(i) = fl32(fl32(i) * step),  step = fl32(dm / bins),  edge(bins) = dm
// is the unique i with edge(i) <= |x| < edge(i+1) (right edge inclusive in the last
// bin). r = round_to_nearest(|x| * inv_step) is within {i, i+1} because the estimate
// is off by < 1e-3 bin, so a single "decrement if |x| < edge(r)" lands on i — the same
// index NumPy's estimate-and-correct produces (numpy/lib/_histograms_impl.py).
struct BinParams {
  float dm, step, inv_step;
  int last_bin;
  int zero_bin;  // >= 0: data_max == 0, every kept element (|x| == 0) goes here
};

__device__ __forceinline__ BinParams make_bin_params(float dm, int bins) {
  BinParams bp;
  bp.dm = dm;
  bp.last_bin = bins - 1;
  bp.zero_bin = -1;
  if (dm == 0.f) {
    // NumPy widens an empty range to (-0.5, 0.5): edges = i * fl32(1/bins) - 0.5
    const float st = __fdiv_rn(1.0f, (float)bins);
    int idx = (int)(0.5f * (float)bins);
    if (idx == bins) --idx;
    if (0.f < __fadd_rn(__fmul_rn((float)idx, st), -0.5f)) --idx;
    if (idx != bins - 1) {
      const float e1 = (idx + 1 == bins) ? 0.5f : __fadd_rn(__fmul_rn((float)(idx + 1), st), -0.5f);
      if (0.f >= e1) ++idx;
    }
    bp.zero_bin = idx;
    bp.step = 0.f;
    bp.inv_step = 0.f;
  } else {
    bp.step = __fdiv_rn(dm, (float)bins);
    bp.inv_step = __fdiv_rn((float)bins, dm);
  }
  return bp;
}

// Returns the bin, or -1 when the element is outside [0, dm] (NumPy's keep mask; also
// catches NaN).
__device__ __forceinline__ int bin_of(float x, const BinParams& bp) {
  const float a = fabsf(x);
  const float tm = __fmaf_rn(a, bp.inv_step, 8388608.0f);  // 2^23 + rn(|x| * inv_step)
  const float rf = __fadd_rn(tm, -8388608.0f);
  const float e0 = __fmul_rn(rf, bp.step);
  int r = __float_as_int(tm) - 0x4B000000;
  r -= (a < e0) ? 1 : 0;
  r = min(r, bp.last_bin);
  return (a <= bp.dm) ? r : -1;
}

// ---- variant 1/3: lane-column histogram -------------------------------------
// Shared-memory layout: word (bin >> 1) * 32 + lane holds two 16-bit counters (bins
// 2w and 2w+1) private to the warp LANE, so every shared atomic of a warp touches 32
// distinct banks: no bank conflicts, and no same-address serialisation however
// skewed the data is (the post-ReLU zero spike costs the same as Gaussian data).
// Counters are flushed to the 64-bit global histogram before any can reach 2^16.
constexpr int kHistThreads = 1024;
constexpr int kHistWarps = kHistThreads / 32;
constexpr int kHistUnroll = 4;                       // float4 in flight per thread
constexpr int kLaneColMaxBins = 3072;                // 192 KB of shared memory
// worst case every element of a lane column lands in one counter:
// per outer iteration a lane column receives kHistWarps * kHistUnroll * 4 elements
constexpr uint32_t kColPerIter = kHistWarps * kHistUnroll * 4;
constexpr uint32_t kFlushIters = 65535u / kColPerIter;  // 127

__device__ __forceinline__ void lanecol_add(uint32_t* sh_lane, int bin) {
  if (bin >= 0) atomicAdd(sh_lane + ((bin >> 1) << 5), (bin & 1) ? 65536u : 1u);
}

__device__ __forceinline__ void lanecol_flush(uint32_t* sh, int nwords, int bins,
                                              unsigned long long* __restrict__ g) {
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int w = warp; w < nwords; w += kHistWarps) {
    const uint32_t v = sh[(w << 5) + lane];
    sh[(w << 5) + lane] = 0;
    const uint32_t lo = __reduce_add_sync(0xffffffffu, v & 0xffffu);
    const uint32_t hi = __reduce_add_sync(0xffffffffu, v >> 16);
    if (lane == 0 && lo) atomicAdd(g + 2 * w, (unsigned long long)lo);
    if (lane == 1 && hi && 2 * w + 1 < bins) atomicAdd(g + 2 * w + 1, (unsigned long long)hi);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kHistThreads, 1)
hist_lanecol_kernel(const dpl_blob* __restrict__ blobs, int n_blobs, uint64_t n_tiles,
                    const float* __restrict__ data_max, int bins,
                    unsigned long long* __restrict__ counts) {
  extern __shared__ __align__(128) uint32_t sh[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int nwords = (bins + 1) >> 1;
  for (int i = tid; i < nwords * 32; i += kHistThreads) sh[i] = 0;
  __syncthreads();
  uint32_t* sh_lane = sh + lane;

  uint64_t tile = n_tiles * blockIdx.x / gridDim.x;
  const uint64_t tile_end = n_tiles * (blockIdx.x + 1ull) / gridDim.x;
  if (tile >= tile_end) return;
  int b = find_blob<5>(blobs, n_blobs, tile);

  while (tile < tile_end) {
    uint64_t blob_tile_end = (b + 1 < n_blobs) ? blobs[b + 1].flat_tile_begin : n_tiles;
    if (tile >= blob_tile_end) {
      ++b;
      continue;
    }
    const uint64_t te = min(tile_end, blob_tile_end);
    const uint64_t n_elems = blobs[b].n_seg * blobs[b].seg_len;
    const uint64_t e0 = (tile - blobs[b].flat_tile_begin) * DPL_FLAT_TILE;
    const uint64_t e1 = min(n_elems, (te - blobs[b].flat_tile_begin) * DPL_FLAT_TILE);
    const float* p = reinterpret_cast<const float*>(blobs[b].ptr) + e0;
    const uint64_t len = e1 - e0;
    const int stat = (int)blobs[b].stat_index;
    const BinParams bp = make_bin_params(data_max[stat], bins);
    unsigned long long* g = counts + (uint64_t)stat * bins;
    const bool usable = (bp.dm >= 0.f) && (bp.dm < INFINITY);  // false for NaN / inf / negative

    if (usable && bp.zero_bin >= 0) {
      // degenerate range: only exact zeros are inside [0, 0]
      uint32_t z = 0;
      for (uint64_t i = tid; i < len; i += kHistThreads) z += (fabsf(ldg_stream1(p + i)) == 0.f);
      z = __reduce_add_sync(0xffffffffu, z);
      if (lane == 0 && z) atomicAdd(g + bp.zero_bin, (unsigned long long)z);
    } else if (usable) {
      const bool aligned = (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
      const uint64_t n4 = aligned ? (len >> 2) : 0;
      const float4* p4 = reinterpret_cast<const float4*>(p);
      uint32_t iters = 0;
      for (uint64_t i0 = 0; i0 < n4; i0 += (uint64_t)kHistThreads * kHistUnroll) {
        float4 v[kHistUnroll];
        bool ok[kHistUnroll];
#pragma unroll
        for (int u = 0; u < kHistUnroll; ++u) {
          const uint64_t i = i0 + (uint64_t)u * kHistThreads + tid;
          ok[u] = i < n4;
          if (ok[u]) v[u] = ldg_stream4(p4 + i);
        }
#pragma unroll
        for (int u = 0; u < kHistUnroll; ++u) {
          if (ok[u]) {
            lanecol_add(sh_lane, bin_of(v[u].x, bp));
            lanecol_add(sh_lane, bin_of(v[u].y, bp));
            lanecol_add(sh_lane, bin_of(v[u].z, bp));
            lanecol_add(sh_lane, bin_of(v[u].w, bp));
          }
        }
        if (++iters == kFlushIters) {
          lanecol_flush(sh, nwords, bins, g);
          iters = 0;
        }
      }
      // scalar remainder (len % 4, or the whole range when the blob is not 16-byte
      // aligned); uniform trip count so that every thread reaches the flush barriers
      const uint64_t rem0 = n4 << 2;
      const uint64_t trips = (len - rem0 + kHistThreads - 1) / kHistThreads;
      uint32_t sc = 0;
      for (uint64_t j = 0; j < trips; ++j) {
        const uint64_t i = rem0 + j * kHistThreads + tid;
        if (i < len) lanecol_add(sh_lane, bin_of(ldg_stream1(p + i), bp));
        if (++sc == 2047u) {  // 2047 * 32 warps = 65504 per lane column
          lanecol_flush(sh, nwords, bins, g);
          sc = 0;
        }
      }
      lanecol_flush(sh, nwords, bins, g);
    }
    tile = te;
    ++b;
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- variants 4/5/6: lane-column histogram, tuned inner loop -----------------
// Same idea as variant 1 with a cheaper per-element path (variant 1 issues ~23 SASS
// instructions per element and is issue-bound at 68% of HBM peak):
//   * word w holds bins w (low 16 bits) and w + nwords (high 16 bits): the word index is
//     a mask / compare instead of shift + parity logic;
//   * the edge correction is an integer mask from `set.lt` folded into one add;
//   * out-of-range elements (NumPy's keep mask) are steered to a trash row instead of
//     branching around the atomic; the atomic is a `red.shared` on a 32-bit shared address;
//   * full iterations run without per-load predicates; the ragged end is a separate loop.
template <bool POW2>
__device__ __forceinline__ void lc2_add(uint32_t lane_base, float x, const BinParams& bp,
                                        uint32_t nwords, uint32_t trash_addr) {
  const float a = fabsf(x);
  const float tm = __fmaf_rn(a, bp.inv_step, 8388608.0f);
  const float rf = __fadd_rn(tm, -8388608.0f);
  const float e0 = __fmul_rn(rf, bp.step);
  int mask;
  asm("set.lt.s32.f32 %0, %1, %2;" : "=r"(mask) : "f"(a), "f"(e0));  // -1 when |x| < edge(r)
  uint32_t r = (uint32_t)(__float_as_int(tm) + mask - 0x4B000000);
  r = min(r, (uint32_t)bp.last_bin);
  const bool hi = r >= nwords;
  const uint32_t w = POW2 ? (r & (nwords - 1u)) : (hi ? r - nwords : r);
  uint32_t addr = lane_base + w * 128u;
  addr = (a <= bp.dm) ? addr : trash_addr;
  const uint32_t inc = hi ? 65536u : 1u;
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(inc) : "memory");
}

__device__ __forceinline__ void lc2_flush(uint32_t* sh, int nwords, int bins,
                                          unsigned long long* __restrict__ g) {
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int w = warp; w < nwords; w += kHistWarps) {
    const uint32_t v = sh[(w << 5) + lane];
    sh[(w << 5) + lane] = 0;
    const uint32_t lo = __reduce_add_sync(0xffffffffu, v & 0xffffu);
    const uint32_t hi = __reduce_add_sync(0xffffffffu, v >> 16);
    if (lane == 0 && lo) atomicAdd(g + w, (unsigned long long)lo);
    if (lane == 1 && hi && w + nwords < bins) atomicAdd(g + w + nwords, (unsigned long long)hi);
  }
  __syncthreads();
}

template <bool POW2>
__device__ __forceinline__ void lc2_add4(uint32_t lane_base, const float4& v, const BinParams& bp,
                                         uint32_t nwords, uint32_t trash) {
  lc2_add<POW2>(lane_base, v.x, bp, nwords, trash);
  lc2_add<POW2>(lane_base, v.y, bp, nwords, trash);
  lc2_add<POW2>(lane_base, v.z, bp, nwords, trash);
  lc2_add<POW2>(lane_base, v.w, bp, nwords, trash);
}

// PREFETCH: issue the next iteration's loads before binning the current registers.
template <bool POW2, bool PREFETCH>
__global__ void __launch_bounds__(kHistThreads, 1)
hist_lc2_kernel(const dpl_blob* __restrict__ blobs, int n_blobs, uint64_t n_tiles,
                const float* __restrict__ data_max, int bins,
                unsigned long long* __restrict__ counts) {
  extern __shared__ __align__(128) uint32_t sh[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int nwords = (bins + 1) >> 1;
  for (int i = tid; i < (nwords + 1) * 32; i += kHistThreads) sh[i] = 0;  // + trash row
  __syncthreads();
  const uint32_t lane_base = smem_u32(sh) + lane * 4u;
  const uint32_t trash = lane_base + (uint32_t)nwords * 128u;

  uint64_t tile = n_tiles * blockIdx.x / gridDim.x;
  const uint64_t tile_end = n_tiles * (blockIdx.x + 1ull) / gridDim.x;
  if (tile >= tile_end) return;
  int b = find_blob<5>(blobs, n_blobs, tile);

  while (tile < tile_end) {
    uint64_t blob_tile_end = (b + 1 < n_blobs) ? blobs[b + 1].flat_tile_begin : n_tiles;
    if (tile >= blob_tile_end) {
      ++b;
      continue;
    }
    const uint64_t te = min(tile_end, blob_tile_end);
    const uint64_t n_elems = blobs[b].n_seg * blobs[b].seg_len;
    const uint64_t e0 = (tile - blobs[b].flat_tile_begin) * DPL_FLAT_TILE;
    const uint64_t e1 = min(n_elems, (te - blobs[b].flat_tile_begin) * DPL_FLAT_TILE);
    const float* p = reinterpret_cast<const float*>(blobs[b].ptr) + e0;
    const uint64_t len = e1 - e0;
    const int stat = (int)blobs[b].stat_index;
    const BinParams bp = make_bin_params(data_max[stat], bins);
    unsigned long long* g = counts + (uint64_t)stat * bins;
    const bool usable = (bp.dm >= 0.f) && (bp.dm < INFINITY);

    if (usable && bp.zero_bin >= 0) {
      uint32_t z = 0;
      for (uint64_t i = tid; i < len; i += kHistThreads) z += (fabsf(ldg_stream1(p + i)) == 0.f);
      z = __reduce_add_sync(0xffffffffu, z);
      if (lane == 0 && z) atomicAdd(g + bp.zero_bin, (unsigned long long)z);
    } else if (usable) {
      const bool aligned = (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
      const uint64_t n4 = aligned ? (len >> 2) : 0;
      const float4* p4 = reinterpret_cast<const float4*>(p);
      constexpr uint64_t kStep = (uint64_t)kHistThreads * kHistUnroll;
      const uint64_t full = n4 / kStep;  // iterations with every load in range
      uint32_t iters = 0;
      if (PREFETCH) {
        float4 cur[kHistUnroll], nxt[kHistUnroll];
        if (full > 0) {
#pragma unroll
          for (int u = 0; u < kHistUnroll; ++u) cur[u] = ldg_stream4(p4 + (uint64_t)u * kHistThreads + tid);
        }
        for (uint64_t it = 0; it < full; ++it) {
          if (it + 1 < full) {
            const float4* q = p4 + (it + 1) * kStep + tid;
#pragma unroll
            for (int u = 0; u < kHistUnroll; ++u) nxt[u] = ldg_stream4(q + (uint64_t)u * kHistThreads);
          }
#pragma unroll
          for (int u = 0; u < kHistUnroll; ++u) lc2_add4<POW2>(lane_base, cur[u], bp, nwords, trash);
#pragma unroll
          for (int u = 0; u < kHistUnroll; ++u) cur[u] = nxt[u];
          if (++iters == kFlushIters) {
            lc2_flush(sh, nwords, bins, g);
            iters = 0;
          }
        }
      } else {
        for (uint64_t it = 0; it < full; ++it) {
          const float4* q = p4 + it * kStep + tid;
          float4 v[kHistUnroll];
#pragma unroll
          for (int u = 0; u < kHistUnroll; ++u) v[u] = ldg_stream4(q + (uint64_t)u * kHistThreads);
#pragma unroll
          for (int u = 0; u < kHistUnroll; ++u) lc2_add4<POW2>(lane_base, v[u], bp, nwords, trash);
          if (++iters == kFlushIters) {
            lc2_flush(sh, nwords, bins, g);
            iters = 0;
          }
        }
      }
      // ragged end of the vector body: < kStep float4, at most kHistUnroll per thread
      for (uint64_t i = full * kStep + tid; i < n4; i += kHistThreads)
        lc2_add4<POW2>(lane_base, ldg_stream4(p4 + i), bp, nwords, trash);
      // (iters < kFlushIters and this adds <= 16 per thread: still below 2^16 per column
      //  because kFlushIters * kColPerIter + kColPerIter <= 65535 + 512 would overflow, so flush)
      lc2_flush(sh, nwords, bins, g);
      // scalar remainder, uniform trip count
      const uint64_t rem0 = n4 << 2;
      const uint64_t trips = (len - rem0 + kHistThreads - 1) / kHistThreads;
      uint32_t sc = 0;
      for (uint64_t j = 0; j < trips; ++j) {
        const uint64_t i = rem0 + j * kHistThreads + tid;
        if (i < len) lc2_add<POW2>(lane_base, ldg_stream1(p + i), bp, nwords, trash);
        if (++sc == 2047u) {
          lc2_flush(sh, nwords, bins, g);
          sc = 0;
        }
      }
      if (trips) lc2_flush(sh, nwords, bins, g);
    }
    tile = te;
    ++b;
  }
}

// ---- variant 7: half-warp columns, 32-bit counters ------------------------------
// word bin * 16 + (lane & 15): lanes l and l + 16 share a column, so a warp's atomic sees
// at most a 2-way bank conflict, but the counters are full 32-bit (no packing logic, no
// periodic flush) and the per-element path is 3 instructions shorter than variant 4.
__device__ __forceinline__ void lc3_add(uint32_t col_base, float x, const BinParams& bp,
                                        uint32_t trash_addr) {
  const float a = fabsf(x);
  const float tm = __fmaf_rn(a, bp.inv_step, 8388608.0f);
  const float rf = __fadd_rn(tm, -8388608.0f);
  const float e0 = __fmul_rn(rf, bp.step);
  int mask;
  asm("set.lt.s32.f32 %0, %1, %2;" : "=r"(mask) : "f"(a), "f"(e0));
  uint32_t r = (uint32_t)(__float_as_int(tm) + mask - 0x4B000000);
  r = min(r, (uint32_t)bp.last_bin);
  uint32_t addr = col_base + r * 64u;
  addr = (a <= bp.dm) ? addr : trash_addr;
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(1u) : "memory");
}

__device__ __forceinline__ void lc3_flush(uint32_t* sh, int bins, unsigned long long* __restrict__ g) {
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b0 = warp * 2; b0 < bins; b0 += kHistWarps * 2) {
    const int b = b0 + (lane >> 4);
    uint32_t v = 0;
    if (b < bins) {
      v = sh[(b << 4) + (lane & 15)];
      sh[(b << 4) + (lane & 15)] = 0;
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((lane & 15) == 0 && v && b < bins) atomicAdd(g + b, (unsigned long long)v);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kHistThreads, 1)
hist_lc3_kernel(const dpl_blob* __restrict__ blobs, int n_blobs, uint64_t n_tiles,
                const float* __restrict__ data_max, int bins,
                unsigned long long* __restrict__ counts) {
  extern __shared__ __align__(128) uint32_t sh[];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < (bins + 1) * 16; i += kHistThreads) sh[i] = 0;
  __syncthreads();
  const uint32_t col_base = smem_u32(sh) + (lane & 15) * 4u;
  const uint32_t trash = col_base + (uint32_t)bins * 64u;

  uint64_t tile = n_tiles * blockIdx.x / gridDim.x;
  const uint64_t tile_end = n_tiles * (blockIdx.x + 1ull) / gridDim.x;
  if (tile >= tile_end) return;
  int b = find_blob<5>(blobs, n_blobs, tile);

  while (tile < tile_end) {
    uint64_t blob_tile_end = (b + 1 < n_blobs) ? blobs[b + 1].flat_tile_begin : n_tiles;
    if (tile >= blob_tile_end) {
      ++b;
      continue;
    }
    const uint64_t te = min(tile_end, blob_tile_end);
    const uint64_t n_elems = blobs[b].n_seg * blobs[b].seg_len;
    const uint64_t e0 = (tile - blobs[b].flat_tile_begin) * DPL_FLAT_TILE;
    const uint64_t e1 = min(n_elems, (te - blobs[b].flat_tile_begin) * DPL_FLAT_TILE);
    const float* p = reinterpret_cast<const float*>(blobs[b].ptr) + e0;
    const uint64_t len = e1 - e0;
    const int stat = (int)blobs[b].stat_index;
    const BinParams bp = make_bin_params(data_max[stat], bins);
    unsigned long long* g = counts + (uint64_t)stat * bins;
    const bool usable = (bp.dm >= 0.f) && (bp.dm < INFINITY);

    if (usable && bp.zero_bin >= 0) {
      uint32_t z = 0;
      for (uint64_t i = tid; i < len; i += kHistThreads) z += (fabsf(ldg_stream1(p + i)) == 0.f);
      z = __reduce_add_sync(0xffffffffu, z);
      if (lane == 0 && z) atomicAdd(g + bp.zero_bin, (unsigned long long)z);
    } else if (usable) {
      const bool aligned = (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
      const uint64_t n4 = aligned ? (len >> 2) : 0;
      const float4* p4 = reinterpret_cast<const float4*>(p);
      constexpr uint64_t kStep = (uint64_t)kHistThreads * kHistUnroll;
      const uint64_t full = n4 / kStep;
      for (uint64_t it = 0; it < full; ++it) {
        const float4* q = p4 + it * kStep + tid;
        float4 v[kHistUnroll];
#pragma unroll
        for (int u = 0; u < kHistUnroll; ++u) v[u] = ldg_stream4(q + (uint64_t)u * kHistThreads);
#pragma unroll
        for (int u = 0; u < kHistUnroll; ++u) {
          lc3_add(col_base, v[u].x, bp, trash);
          lc3_add(col_base, v[u].y, bp, trash);
          lc3_add(col_base, v[u].z, bp, trash);
          lc3_add(col_base, v[u].w, bp, trash);
        }
      }
      for (uint64_t i = full * kStep + tid; i < n4; i += kHistThreads) {
        const float4 v = ldg_stream4(p4 + i);
        lc3_add(col_base, v.x, bp, trash);
        lc3_add(col_base, v.y, bp, trash);
        lc3_add(col_base, v.z, bp, trash);
        lc3_add(col_base, v.w, bp, trash);
      }
      for (uint64_t i = (n4 << 2) + tid; i < len; i += kHistThreads)
        lc3_add(col_base, ldg_stream1(p + i), bp, trash);
      lc3_flush(sh, bins, g);  // one CTA's share of a blob is far below 2^32 elements
    }
    tile = te;
    ++b;
  }
}

// ---- variant 3: lane-column histogram fed by a TMA-staged tile ring ---------
// Same counters as variant 1; the blob is streamed into a shared-memory ring with 1-D
// bulk async copies (cp.async.bulk, SASS UBLKCP) completing on mbarriers, so no load
// registers are held and the LSU only sees conflict-free LDS.128 + ATOMS.
constexpr int kStageFloats = 8192;  // 32 KB per stage: 2 float4 per thread
constexpr int kStageBytes = kStageFloats * 4;

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

__global__ void __launch_bounds__(kHistThreads, 1)
hist_lanecol_tma_kernel(const dpl_blob* __restrict__ blobs, int n_blobs, uint64_t n_tiles,
                        const float* __restrict__ data_max, int bins, int n_stages,
                        unsigned long long* __restrict__ counts) {
  extern __shared__ __align__(128) uint32_t sh[];
  __shared__ __align__(8) uint64_t s_full[4], s_empty[4];
  const int tid = threadIdx.x, lane = tid & 31;
  const int nwords = (bins + 1) >> 1;
  float* ring = reinterpret_cast<float*>(sh + (size_t)(nwords + 1) * 32);  // after the trash row
  for (int i = tid; i < (nwords + 1) * 32; i += kHistThreads) sh[i] = 0;
  if (tid == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(smem_u32(&s_full[s]), 1);
      mbar_init(smem_u32(&s_empty[s]), kHistWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t lane_base = smem_u32(sh) + lane * 4u;
  const uint32_t trash = lane_base + (uint32_t)nwords * 128u;

  uint64_t tile = n_tiles * blockIdx.x / gridDim.x;
  const uint64_t tile_end = n_tiles * (blockIdx.x + 1ull) / gridDim.x;
  if (tile >= tile_end) return;
  int b = find_blob<5>(blobs, n_blobs, tile);
  uint32_t it = 0;  // chunks consumed so far by this CTA (stage = it % n_stages)

  while (tile < tile_end) {
    uint64_t blob_tile_end = (b + 1 < n_blobs) ? blobs[b + 1].flat_tile_begin : n_tiles;
    if (tile >= blob_tile_end) {
      ++b;
      continue;
    }
    const uint64_t te = min(tile_end, blob_tile_end);
    const uint64_t n_elems = blobs[b].n_seg * blobs[b].seg_len;
    const uint64_t e0 = (tile - blobs[b].flat_tile_begin) * DPL_FLAT_TILE;
    const uint64_t e1 = min(n_elems, (te - blobs[b].flat_tile_begin) * DPL_FLAT_TILE);
    const float* p = reinterpret_cast<const float*>(blobs[b].ptr) + e0;
    const uint64_t len = e1 - e0;
    const int stat = (int)blobs[b].stat_index;
    const BinParams bp = make_bin_params(data_max[stat], bins);
    unsigned long long* g = counts + (uint64_t)stat * bins;
    const bool usable = (bp.dm >= 0.f) && (bp.dm < INFINITY);

    if (usable && bp.zero_bin >= 0) {
      uint32_t z = 0;
      for (uint64_t i = tid; i < len; i += kHistThreads) z += (fabsf(ldg_stream1(p + i)) == 0.f);
      z = __reduce_add_sync(0xffffffffu, z);
      if (lane == 0 && z) atomicAdd(g + bp.zero_bin, (unsigned long long)z);
    } else if (usable) {
      const bool aligned = (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
      const uint64_t n4 = aligned ? (len >> 2) : 0;
      const uint32_t nchunks = (uint32_t)((n4 * 4 + kStageFloats - 1) / kStageFloats);
      auto chunk_bytes = [&](uint32_t c) -> uint32_t {
        const uint64_t left = n4 * 16 - (uint64_t)c * kStageBytes;
        return (uint32_t)min((uint64_t)kStageBytes, left);
      };
      // prologue: fill the ring. Every stage was fully drained by the previous blob
      // range (all chunks issued were consumed), so no empty-wait is needed here.
      if (tid == 0) {
        for (uint32_t c = 0; c < (uint32_t)n_stages && c < nchunks; ++c) {
          const uint32_t s = (it + c) % n_stages;
          const uint32_t bytes = chunk_bytes(c);
          mbar_expect_tx(smem_u32(&s_full[s]), bytes);
          bulk_g2s(smem_u32(ring + (size_t)s * kStageFloats), p + (size_t)c * kStageFloats, bytes,
                   smem_u32(&s_full[s]));
        }
      }
      uint32_t since_flush = 0;
      for (uint32_t c = 0; c < nchunks; ++c, ++it) {
        const uint32_t s = it % n_stages;
        const uint32_t ph = (it / n_stages) & 1u;
        mbar_wait(smem_u32(&s_full[s]), ph);
        const uint32_t c4 = chunk_bytes(c) >> 4;
        const float4* r4 = reinterpret_cast<const float4*>(ring + (size_t)s * kStageFloats);
        float4 v0, v1;
        const bool ok0 = (uint32_t)tid < c4, ok1 = (uint32_t)tid + kHistThreads < c4;
        if (ok0) v0 = r4[tid];
        if (ok1) v1 = r4[tid + kHistThreads];
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s_empty[s]));  // stage data is in registers now
        if (tid == 0 && c + n_stages < nchunks) {
          mbar_wait(smem_u32(&s_empty[s]), ph);  // all 32 warps released this stage
          const uint32_t bytes = chunk_bytes(c + n_stages);
          mbar_expect_tx(smem_u32(&s_full[s]), bytes);
          bulk_g2s(smem_u32(ring + (size_t)s * kStageFloats),
                   p + (size_t)(c + n_stages) * kStageFloats, bytes, smem_u32(&s_full[s]));
        }
        if (ok0) {
          lc2_add<false>(lane_base, v0.x, bp, nwords, trash);
          lc2_add<false>(lane_base, v0.y, bp, nwords, trash);
          lc2_add<false>(lane_base, v0.z, bp, nwords, trash);
          lc2_add<false>(lane_base, v0.w, bp, nwords, trash);
        }
        if (ok1) {
          lc2_add<false>(lane_base, v1.x, bp, nwords, trash);
          lc2_add<false>(lane_base, v1.y, bp, nwords, trash);
          lc2_add<false>(lane_base, v1.z, bp, nwords, trash);
          lc2_add<false>(lane_base, v1.w, bp, nwords, trash);
        }
        if (++since_flush == 255u) {  // 255 chunks * 256 elements per lane column < 2^16
          lc2_flush(sh, nwords, bins, g);
          since_flush = 0;
        }
      }
      // the last stages' empty barriers must complete their phase before they are reused
      // without a wait in the next prologue: drain them here (cheap, once per blob range)
      __syncthreads();
      const uint64_t rem0 = n4 << 2;
      const uint64_t trips = (len - rem0 + kHistThreads - 1) / kHistThreads;
      uint32_t sc = 0;
      for (uint64_t j = 0; j < trips; ++j) {
        const uint64_t i = rem0 + j * kHistThreads + tid;
        if (i < len) lc2_add<false>(lane_base, ldg_stream1(p + i), bp, nwords, trash);
        if (++sc == 2047u) {
          lc2_flush(sh, nwords, bins, g);
          sc = 0;
        }
      }
      lc2_flush(sh, nwords, bins, g);
    }
    tile = te;
    ++b;
  }
}

// ---- variant 2: plain shared histogram (any bins up to 48 K) -----------------
constexpr int kSimpleThreads = 512;
__global__ void __launch_bounds__(kSimpleThreads)
hist_simple_kernel(const dpl_blob* __restrict__ blobs, int n_blobs, uint64_t n_tiles,
                   const float* __restrict__ data_max, int bins,
                   unsigned long long* __restrict__ counts) {
  extern __shared__ __align__(128) uint32_t sh[];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < bins; i += kSimpleThreads) sh[i] = 0;
  __syncthreads();

  uint64_t tile = n_tiles * blockIdx.x / gridDim.x;
  const uint64_t tile_end = n_tiles * (blockIdx.x + 1ull) / gridDim.x;
  if (tile >= tile_end) return;
  int b = find_blob<5>(blobs, n_blobs, tile);

  while (tile < tile_end) {
    uint64_t blob_tile_end = (b + 1 < n_blobs) ? blobs[b + 1].flat_tile_begin : n_tiles;
    if (tile >= blob_tile_end) {
      ++b;
      continue;
    }
    const uint64_t te = min(tile_end, blob_tile_end);
    const uint64_t n_elems = blobs[b].n_seg * blobs[b].seg_len;
    const uint64_t e0 = (tile - blobs[b].flat_tile_begin) * DPL_FLAT_TILE;
    const uint64_t e1 = min(n_elems, (te - blobs[b].flat_tile_begin) * DPL_FLAT_TILE);
    const float* p = reinterpret_cast<const float*>(blobs[b].ptr) + e0;
    const uint64_t len = e1 - e0;  // one CTA's share of one blob: far below 2^32 elements
    const int stat = (int)blobs[b].stat_index;
    const BinParams bp = make_bin_params(data_max[stat], bins);
    unsigned long long* g = counts + (uint64_t)stat * bins;
    const bool usable = (bp.dm >= 0.f) && (bp.dm < INFINITY);
    if (usable) {
      uint32_t zeros = 0;  // |x| == 0 is always bin 0 (or zero_bin): keep it out of the atomics
      const int zbin = bp.zero_bin >= 0 ? bp.zero_bin : 0;
      const bool aligned = (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
      const uint64_t n4 = (aligned && bp.zero_bin < 0) ? (len >> 2) : 0;
      const float4* p4 = reinterpret_cast<const float4*>(p);
      auto add1 = [&](float x) {
        if (x == 0.f) {
          ++zeros;
        } else if (bp.zero_bin < 0) {
          const int r = bin_of(x, bp);
          if (r >= 0) atomicAdd(sh + r, 1u);
        }
      };
      for (uint64_t i0 = 0; i0 < n4; i0 += (uint64_t)kSimpleThreads * kHistUnroll) {
        float4 v[kHistUnroll];
        bool ok[kHistUnroll];
#pragma unroll
        for (int u = 0; u < kHistUnroll; ++u) {
          const uint64_t i = i0 + (uint64_t)u * kSimpleThreads + tid;
          ok[u] = i < n4;
          if (ok[u]) v[u] = ldg_stream4(p4 + i);
        }
#pragma unroll
        for (int u = 0; u < kHistUnroll; ++u) {
          if (ok[u]) {
            add1(v[u].x);
            add1(v[u].y);
            add1(v[u].z);
            add1(v[u].w);
          }
        }
      }
      for (uint64_t i = (n4 << 2) + tid; i < len; i += kSimpleThreads) add1(ldg_stream1(p + i));
      zeros = __reduce_add_sync(0xffffffffu, zeros);
      if (lane == 0 && zeros) atomicAdd(sh + zbin, zeros);
    }
    __syncthreads();
    for (int i = tid; i < bins; i += kSimpleThreads) {
      const uint32_t c = sh[i];
      if (c) {
        atomicAdd(g + i, (unsigned long long)c);
        sh[i] = 0;
      }
    }
    __syncthreads();
    tile = te;
    ++b;
  }
}

// ===========================================================================
// K3: percentile clip search
// ===========================================================================
constexpr int kPctThreads = 256;
constexpr int kPctChunk = 2048;
__global__ void __launch_bounds__(kPctThreads)
hist_percentile_kernel(const unsigned long long* __restrict__ counts, int bins, double threshold,
                       const float* __restrict__ data_max, const float* __restrict__ bmin,
                       const float* __restrict__ bmax, float* __restrict__ clip,
                       int* __restrict__ out_bin) {
  __shared__ double s_p[kPctChunk];
  __shared__ unsigned long long s_red[kPctThreads / 32];
  __shared__ unsigned long long s_total;
  __shared__ int s_found;
  __shared__ double s_accum;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned long long* h = counts + (uint64_t)blockIdx.x * bins;

  unsigned long long part = 0;
  for (int i = tid; i < bins; i += kPctThreads) part += h[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) s_red[warp] = part;
  if (tid == 0) {
    s_found = -1;
    s_accum = 0.0;
  }
  __syncthreads();
  if (tid == 0) {
    unsigned long long tot = 0;
    for (int w = 0; w < kPctThreads / 32; ++w) tot += s_red[w];
    s_total = tot;
  }
  __syncthreads();
  // hist.astype(np.float32) / hist.sum(): float32 array / int64 scalar -> float64 (NumPy 2)
  const double total = (double)(long long)s_total;
  for (int c0 = 0; c0 < bins && s_found < 0; c0 += kPctChunk) {
    const int n = min(kPctChunk, bins - c0);
    for (int i = tid; i < n; i += kPctThreads)
      s_p[i] = (double)__ull2float_rn(h[c0 + i]) / total;  // 0/0 = NaN when the histogram is empty
    __syncthreads();
    if (tid == 0) {
      double acc = s_accum;
      for (int i = 0; i < n; ++i) {
        acc += s_p[i];
        if (acc >= threshold) {
          s_found = c0 + i;
          break;
        }
      }
      s_accum = acc;
    }
    __syncthreads();
  }
  if (tid == 0) {
    const int b = blockIdx.x;
    const float mn = bmin[b], mx = bmax[b];
    const int i = s_found;
    float lo = mn, hi = mx;
    if (i >= 0) {
      // (i + 0.5) * (data_max / bins), all float32 under NumPy 2
      const float cv = __fmul_rn((float)i + 0.5f, __fdiv_rn(data_max[b], (float)bins));
      lo = (mn > -cv) ? mn : -cv;  // max(-cv, min)
      hi = (mx < cv) ? mx : cv;    // min(cv, max)
    }
    clip[2 * b] = lo;
    clip[2 * b + 1] = hi;
    if (out_bin) out_bin[b] = i;
  }
}

}  // namespace
}  // namespace dpl

// ===========================================================================
// C-ABI
// ===========================================================================
using namespace dpl;

extern "C" int dpl_plan_blobs(dpl_blob* hb, int n_blobs, uint64_t* n_segments,
                              uint64_t* n_seg_tiles, uint64_t* n_flat_tiles) {
  DPL_REQUIRE(hb != nullptr || n_blobs == 0, "null blob table");
  DPL_REQUIRE(n_blobs >= 0, "negative blob count");
  uint64_t segs = 0, st = 0, ft = 0;
  for (int i = 0; i < n_blobs; ++i) {
    hb[i].seg_out_base = segs;
    hb[i].seg_tile_begin = st;
    hb[i].flat_tile_begin = ft;
    const uint64_t tps = (hb[i].seg_len + DPL_SEG_TILE - 1) / DPL_SEG_TILE;
    segs += hb[i].n_seg;
    st += hb[i].n_seg * tps;
    ft += (hb[i].n_seg * hb[i].seg_len + DPL_FLAT_TILE - 1) / DPL_FLAT_TILE;
  }
  if (n_segments) *n_segments = segs;
  if (n_seg_tiles) *n_seg_tiles = st;
  if (n_flat_tiles) *n_flat_tiles = ft;
  return 0;
}

extern "C" size_t dpl_segstats_scratch_bytes(uint64_t n_seg_tiles) {
  // tmin f32 | tmax f32 | tnnz u32 | tsum f64, each section 256-byte aligned
  const size_t a = ((size_t)n_seg_tiles * 4 + 255) & ~(size_t)255;
  const size_t d = ((size_t)n_seg_tiles * 8 + 255) & ~(size_t)255;
  return 3 * a + d + 256;
}

extern "C" int dpl_segstats_f32(const dpl_blob* d_blobs, int n_blobs, uint64_t n_segments,
                                uint64_t n_seg_tiles, float* d_min, float* d_max,
                                double* d_abssum, uint64_t* d_nnz, float* d_blob_min,
                                float* d_blob_max, void* d_scratch, size_t scratch_bytes,
                                int ctas_per_sm, void* stream) {
  DPL_REQUIRE(d_blobs && n_blobs > 0, "empty blob table");
  DPL_REQUIRE(d_min && d_max, "null output");
  DPL_REQUIRE(ctas_per_sm >= 0 && ctas_per_sm <= 8, "ctas_per_sm must be 0 (default) .. 8");
  if (n_segments == 0) return 0;
  if (scratch_bytes < dpl_segstats_scratch_bytes(n_seg_tiles) || !d_scratch) {
    set_error("dpl_segstats_f32: scratch too small (%zu < %zu)", scratch_bytes,
              dpl_segstats_scratch_bytes(n_seg_tiles));
    return DPL_E_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uintptr_t base = (reinterpret_cast<uintptr_t>(d_scratch) + 255) & ~(uintptr_t)255;
  const size_t a = ((size_t)n_seg_tiles * 4 + 255) & ~(size_t)255;
  float* tmin = reinterpret_cast<float*>(base);
  float* tmax = reinterpret_cast<float*>(base + a);
  uint32_t* tnnz = reinterpret_cast<uint32_t*>(base + 2 * a);
  double* tsum = reinterpret_cast<double*>(base + 3 * a);
  if (n_seg_tiles > 0) {
    // 8 CTAs of 256 threads per SM; chunked tile ranges keep every CTA on one HBM stream
    // (fewer when the caller overlaps this pass with tensor-core kernels on another stream and
    // wants it to fit beside their CTAs)
    uint64_t grid = (uint64_t)sm_count() * (ctas_per_sm ? ctas_per_sm : 8);
    if (grid > n_seg_tiles) grid = n_seg_tiles;
    segstats_tiles_kernel<<<(unsigned)grid, kSegThreads, 0, st>>>(d_blobs, n_blobs, n_seg_tiles,
                                                                  tmin, tmax, tsum, tnnz);
    DPL_LAUNCH_CHECK("segstats_tiles_kernel");
  }
  const unsigned fgrid = (unsigned)((n_segments + 7) / 8);
  segstats_finalize_kernel<<<fgrid, 256, 0, st>>>(d_blobs, n_blobs, n_segments, tmin, tmax, tsum,
                                                  tnnz, d_min, d_max, d_abssum, d_nnz, d_blob_min,
                                                  d_blob_max);
  DPL_LAUNCH_CHECK("segstats_finalize_kernel");
  return 0;
}

extern "C" int dpl_absmax_f32(const float* d_blob_min, const float* d_blob_max,
                              float* d_data_max, int n_stats, void* stream) {
  DPL_REQUIRE(d_blob_min && d_blob_max && d_data_max, "null pointer");
  if (n_stats <= 0) return 0;
  absmax_kernel<<<(n_stats + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      d_blob_min, d_blob_max, d_data_max, n_stats);
  DPL_LAUNCH_CHECK("absmax_kernel");
  return 0;
}

extern "C" int dpl_hist_abs_f32(const dpl_blob* d_blobs, int n_blobs, uint64_t n_flat_tiles,
                                const float* d_data_max, int bins, unsigned long long* d_counts,
                                int variant, void* stream) {
  DPL_REQUIRE(d_blobs && n_blobs > 0, "empty blob table");
  DPL_REQUIRE(d_data_max && d_counts, "null pointer");
  DPL_REQUIRE(bins >= 1 && bins <= (1 << 22), "bins out of range");
  if (n_flat_tiles == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (variant == 0) variant = (bins <= kLaneColMaxBins) ? 7 : 2;  // 7: 101% of measured HBM peak on B200
  if (variant == 1) {
    DPL_REQUIRE(bins <= kLaneColMaxBins, "lane-column variant supports bins <= 3072");
    const size_t smem = (size_t)((bins + 1) >> 1) * 128;
    {
      int s = cuda_status(cudaFuncSetAttribute(hist_lanecol_kernel,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               kLaneColMaxBins / 2 * 128),
                          "cudaFuncSetAttribute(hist_lanecol_kernel)");
      if (s) return s;
    }
    uint64_t grid = (uint64_t)sm_count();
    if (grid > n_flat_tiles) grid = n_flat_tiles;
    hist_lanecol_kernel<<<(unsigned)grid, kHistThreads, smem, st>>>(d_blobs, n_blobs, n_flat_tiles,
                                                                    d_data_max, bins, d_counts);
    DPL_LAUNCH_CHECK("hist_lanecol_kernel");
    return 0;
  }
  if (variant == 4 || variant == 5) {
    DPL_REQUIRE(bins <= kLaneColMaxBins, "lane-column variant supports bins <= 3072");
    const int nwords = (bins + 1) >> 1;
    const size_t smem = (size_t)(nwords + 1) * 128;
    const bool pow2 = (nwords & (nwords - 1)) == 0 && (bins & 1) == 0;
    auto kern = pow2 ? (variant == 5 ? hist_lc2_kernel<true, true> : hist_lc2_kernel<true, false>)
                     : (variant == 5 ? hist_lc2_kernel<false, true> : hist_lc2_kernel<false, false>);
    int s = cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (kLaneColMaxBins / 2 + 1) * 128),
                        "cudaFuncSetAttribute(hist_lc2_kernel)");
    if (s) return s;
    uint64_t grid = (uint64_t)sm_count();
    if (grid > n_flat_tiles) grid = n_flat_tiles;
    kern<<<(unsigned)grid, kHistThreads, smem, st>>>(d_blobs, n_blobs, n_flat_tiles, d_data_max, bins,
                                                     d_counts);
    DPL_LAUNCH_CHECK("hist_lc2_kernel");
    return 0;
  }
  if (variant == 7) {
    DPL_REQUIRE(bins <= kLaneColMaxBins, "half-warp-column variant supports bins <= 3072");
    const size_t smem = (size_t)(bins + 1) * 64;
    int s = cuda_status(cudaFuncSetAttribute(hist_lc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (kLaneColMaxBins + 1) * 64),
                        "cudaFuncSetAttribute(hist_lc3_kernel)");
    if (s) return s;
    uint64_t grid = (uint64_t)sm_count();
    if (grid > n_flat_tiles) grid = n_flat_tiles;
    hist_lc3_kernel<<<(unsigned)grid, kHistThreads, smem, st>>>(d_blobs, n_blobs, n_flat_tiles,
                                                                d_data_max, bins, d_counts);
    DPL_LAUNCH_CHECK("hist_lc3_kernel");
    return 0;
  }
  if (variant == 3) {
    DPL_REQUIRE(bins <= 2048, "TMA lane-column variant supports bins <= 2048 (ring + histogram share 227 KB)");
    const size_t hist_bytes = (size_t)(((bins + 1) >> 1) + 1) * 128;
    const int n_stages = 3;
    const size_t smem = hist_bytes + (size_t)n_stages * kStageBytes;
    int s = cuda_status(cudaFuncSetAttribute(hist_lanecol_tma_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem),
                        "cudaFuncSetAttribute(hist_lanecol_tma_kernel)");
    if (s) return s;
    uint64_t grid = (uint64_t)sm_count();
    if (grid > n_flat_tiles) grid = n_flat_tiles;
    hist_lanecol_tma_kernel<<<(unsigned)grid, kHistThreads, smem, st>>>(
        d_blobs, n_blobs, n_flat_tiles, d_data_max, bins, n_stages, d_counts);
    DPL_LAUNCH_CHECK("hist_lanecol_tma_kernel");
    return 0;
  }
  if (variant == 2) {
    DPL_REQUIRE(bins <= 49152, "shared-histogram variant supports bins <= 49152");
    const size_t smem = (size_t)bins * 4;
    {
      int s = cuda_status(cudaFuncSetAttribute(hist_simple_kernel,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               49152 * 4),
                          "cudaFuncSetAttribute(hist_simple_kernel)");
      if (s) return s;
    }
    int per_sm = (int)((200 * 1024) / (smem + 1024));
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    uint64_t grid = (uint64_t)sm_count() * per_sm;
    if (grid > n_flat_tiles) grid = n_flat_tiles;
    hist_simple_kernel<<<(unsigned)grid, kSimpleThreads, smem, st>>>(d_blobs, n_blobs, n_flat_tiles,
                                                                     d_data_max, bins, d_counts);
    DPL_LAUNCH_CHECK("hist_simple_kernel");
    return 0;
  }
  set_error("dpl_hist_abs_f32: unknown variant %d", variant);
  return DPL_E_UNSUPPORTED;
}

extern "C" int dpl_hist_percentile(const unsigned long long* d_counts, int n_stats, int bins,
                                   double threshold, const float* d_data_max,
                                   const float* d_blob_min, const float* d_blob_max,
                                   float* d_clip, int* d_bin, void* stream) {
  DPL_REQUIRE(d_counts && d_data_max && d_blob_min && d_blob_max && d_clip, "null pointer");
  DPL_REQUIRE(bins >= 1, "bins out of range");
  if (n_stats <= 0) return 0;
  hist_percentile_kernel<<<n_stats, kPctThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      d_counts, bins, threshold, d_data_max, d_blob_min, d_blob_max, d_clip, d_bin);
  DPL_LAUNCH_CHECK("hist_percentile_kernel");
  return 0;
}
