// Tensor-core plumbing shared by the tcgen05 kernels (dpl_gemm.cu, dpl_recon_conv.cu): mbarrier / TMA /
// UMMA descriptor helpers and the host-side tensor-map encoder. sm_100a only.
#pragma once

#include <cuda.h>
#include <stdint.h>

#include "dpl_common.cuh"

namespace dpl {
namespace {

constexpr int kBM = 128, kBN = 128, kBK = 32;       // tile; kBK floats = one 128-byte swizzle row
constexpr int kTileBytes = kBM * kBK * 4;           // 16 KB per operand per stage
constexpr int kUmmaK = 8;                           // tf32: 32 bytes of K per instruction

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
