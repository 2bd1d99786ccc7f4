// Probe: which inner start coordinates does a swizzled TMA tile load accept? (one load per process)
// usage: tma_probe <swizzle: 0 none,3 128B,4 128B_ATOM_32B> <coord>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap tm, int c0, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  uint32_t dst = ((uint32_t)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(32 * 32 * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(&tm), "r"(c0), "r"(0), "r"(0), "r"(b) : "memory");
    uint32_t ok = 0;
    long long t0 = clock64();
    while (!ok && clock64() - t0 < 2000000000ll)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(b), "r"(0) : "memory");
    out[32 * 32] = ok ? 1.f : 0.f;
  }
  __syncthreads();
  const float* s = reinterpret_cast<const float*>(smem + (dst - (uint32_t)__cvta_generic_to_shared(smem)));
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) out[i] = s[i];
}

int main(int argc, char** argv) {
  int swz = atoi(argv[1]), coord = atoi(argv[2]);
  const int inner = 4096, outer = 32;
  std::vector<float> h(inner * outer);
  for (int k = 0; k < outer; ++k)
    for (int i = 0; i < inner; ++i) h[k * inner + i] = k * 10000 + i;
  float *d, *out;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&out, (32 * 32 + 1) * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap tm;
  cuuint64_t dims[3] = {inner, outer, 1}, strides[2] = {inner * 4, (cuuint64_t)inner * outer * 4};
  cuuint32_t box[3] = {32, 32, 1}, es[3] = {1, 1, 1};
  CUresult r = ((Fn)fnp)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         (CUtensorMapSwizzle)swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r) { printf("swz %d coord %d: encode failed %d\n", swz, coord, (int)r); return 1; }
  probe<<<1, 128, 32 * 32 * 4 + 1024>>>(tm, coord, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e) { printf("swz %d coord %d: %s\n", swz, coord, cudaGetErrorString(e)); return 2; }
  std::vector<float> o(32 * 32 + 1);
  cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
  // k-row 0 and k-row 1 as stored (first 8 floats of each 128-byte row)
  printf("swz %d coord %d: done=%g row0 [%g %g %g %g %g ... %g] row1 [%g %g ...] row5 [%g %g %g %g %g]\n", swz, coord, o[1024],
         o[0], o[1], o[2], o[3], o[4], o[31], o[32], o[33], o[160], o[161], o[162], o[163], o[164]);
  return 0;
}
