// Stand-alone probe of the two tensor-map shapes the staged integrator uses (development tool, not part of the library):
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tools/tma_probe.cu && ./tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void k_probe(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int c3, int c4, uint32_t bytes, float* out, int nout) {
  extern __shared__ __align__(128) uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  uint8_t* smem = raw + ((128u - (smem_u32(raw) & 127u)) & 127u);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    if (RANK == 5)
      asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                   ::"r"(smem_u32(smem)), "l"((uint64_t)&map), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
    if (RANK == 4)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                   ::"r"(smem_u32(smem)), "l"((uint64_t)&map), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(smem_u32(smem)), "l"((uint64_t)&map), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
  }
  uint32_t ok = 0;
  long long t0 = clock64();
  while (!ok && clock64() - t0 < (1ll << 28))
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
  for (int i = threadIdx.x; i < nout; i += blockDim.x) out[i] = ok ? reinterpret_cast<float*>(smem)[i] : -12345.0f;
}

int main(int argc, char** argv) {
  const int argT = argc > 1 ? atoi(argv[1]) : 36, argV = argc > 2 ? atoi(argv[2]) : 0, argX = argc > 3 ? atoi(argv[3]) : 101;
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)f;
  if (!enc) { printf("no encoder\n"); return 1; }
  // (a) 5-D float32: (4, IX, IY, IZ, N), box (4, 7, 9, 9, 4)
  {
    const int IX = 128, IY = 128, IZ = 256, N = 4, BX = 7, BY = 9, BZ = 9;
    std::vector<float> h((size_t)N * IZ * IY * IX * 4);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 1000003);
    float* d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap m;
    const cuuint64_t dims[5] = {4, IX, IY, IZ, N};
    const cuuint64_t strides[4] = {16, (cuuint64_t)IX * 16, (cuuint64_t)IX * IY * 16, (cuuint64_t)IX * IY * IZ * 16};
    const cuuint32_t box[5] = {4, BX, BY, BZ, N}, es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const uint32_t bytes = 16u * BX * BY * BZ * N;
    float* out; cudaMalloc(&out, bytes);
    cudaFuncSetAttribute(k_probe<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k_probe<5><<<1, 128, bytes + 256>>>(m, 0, 3, 5, 7, 0, bytes, out, bytes / 4);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> o(bytes / 4);
    cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost);
    // element (s=1, z=1, y=2, x=3, c=2) of the box
    const size_t bi = ((((size_t)1 * BZ + 1) * BY + 2) * BX + 3) * 4 + 2;
    const size_t gi = ((((size_t)1 * IZ + 8) * IY + 7) * IX + 6) * 4 + 2;
    printf("5d: encode %d, run %s, box[%zu] = %g, expected %g\n", (int)r, cudaGetErrorString(e), bi, o[bi], h[gi]);
    if (e != cudaSuccess) return 2;
  }
  // (b) pair image: (pitch, H+2, N) pixels of 8 bytes, box (T, T, 1)
  {
    const int T = argT, variant = argV;
    const int pitch = 514, H2 = 426, N = 4;
    std::vector<float> h((size_t)N * H2 * pitch * 2);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 1000003);
    float* d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap m;
    // variant 0: float32 image of width 2*pitch (x coordinate and box width doubled); variant 1: 8-byte elements
    const cuuint64_t dims[3] = {variant ? (cuuint64_t)pitch : 2 * (cuuint64_t)pitch, H2, N};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * 8, (cuuint64_t)pitch * H2 * 8};
    const cuuint32_t box[3] = {variant ? (cuuint32_t)T : 2u * T, (cuuint32_t)T, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&m, variant ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const uint32_t bytes = 8u * T * T;
    float* out; cudaMalloc(&out, bytes);
    cudaFuncSetAttribute(k_probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k_probe<3><<<1, 128, bytes + 256>>>(m, variant ? argX : 2 * argX, 57, 2, 0, 0, bytes, out, bytes / 4);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> o(bytes / 4);
    cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost);
    const size_t bi = ((size_t)5 * T + 9) * 2 + 1;
    const size_t gi = (((size_t)2 * H2 + 62) * pitch + argX + 9) * 2 + 1;
    printf("3d %s T=%d: encode %d, run %s, box[%zu] = %g, expected %g\n", variant ? "f64" : "f32x2", T, (int)r, cudaGetErrorString(e), bi, o[bi], h[gi]);

  }
  return 0;
}
