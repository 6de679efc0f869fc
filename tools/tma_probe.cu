// Probe: which FP64 tensor-map encodings does the driver accept, and what lands in shared memory?
//   (a) k-major operand: 4-D map (k, row, inner-batch, outer-batch), box (16,128,1,1), SWIZZLE_128B
//   (b) m-major operand as a 5-D "panel view" (m%8, k, m/8, inner, outer) with NON-monotonic strides
//       (8 B implicit, ld*8, 64, ...), box (8,16,16,1,1), no swizzle -> smem [16 panels][16 k][8]
//   (c) m-major operand as plain 2-D swizzled boxes (16 m, 16 k), SWIZZLE_128B
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe_kernel(const __grid_constant__ CUtensorMap map, int rank, int c0, int c1, int c2, int c3, int c4,
                             double* out, int ndoubles) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bar;
  double* buf = reinterpret_cast<double*>(sm);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(ndoubles * 8));
    if (rank == 4)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3,%4,%5,%6}], [%2];"
                   ::"r"(smem_u32(buf)), "l"(&map), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    else if (rank == 5)
      asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3,%4,%5,%6,%7}], [%2];"
                   ::"r"(smem_u32(buf)), "l"(&map), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3,%4}], [%2];"
                   ::"r"(smem_u32(buf)), "l"(&map), "r"(smem_u32(&bar)), "r"(c0), "r"(c1) : "memory");
  }
  uint32_t done = 0;
  int spins = 0;
  while (!done && spins < (1 << 22)) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(smem_u32(&bar)));
    ++spins;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ndoubles; i += blockDim.x) out[i] = done ? buf[i] : -777.0;
}

int main() {
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q) != cudaSuccess || !encode) {
    printf("no cuTensorMapEncodeTiled entry point\n");
    return 1;
  }
  const int ld = 512, rows = 512, nb_in = 2, nb_out = 2;
  const long long iS = 256LL * (ld + 1), oS = (long long)ld * rows;       // inner / outer batch strides (elements)
  std::vector<double> h((size_t)oS * nb_out);
  for (size_t i = 0; i < h.size(); i++) h[i] = (double)i;                 // value == linear element index
  double *d, *out;
  cudaMalloc(&d, h.size() * 8);
  cudaMalloc(&out, 2048 * 8);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  std::vector<double> ho(2048);
  int bad_total = 0;

  {  // (a) k-major, swizzle 128B: element (r, k) of the tile at byte r*128 + (((k>>1) ^ (r&7))<<4) + (k&1)*8
    CUtensorMap m;
    cuuint64_t dims[4] = {256, 256, (cuuint64_t)nb_in, (cuuint64_t)nb_out};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 8, (cuuint64_t)iS * 8, (cuuint64_t)oS * 8};
    cuuint32_t box[4] = {16, 128, 1, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("(a) encode kmajor 4D swizzle128: %d\n", (int)r);
    if (r == CUDA_SUCCESS) {
      const int k0 = 32, i0 = 128, zi = 1, zo = 1;
      probe_kernel<<<1, 128, 32768>>>(m, 4, k0, i0, zi, zo, 0, out, 2048);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(ho.data(), out, 2048 * 8, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int rr = 0; rr < 128; rr++)
        for (int k = 0; k < 16; k++) {
          const int idx = (rr * 128 + ((((k >> 1) ^ (rr & 7))) << 4) + (k & 1) * 8) / 8;
          const double expect = (double)(zo * oS + zi * iS + (long long)(i0 + rr) * ld + k0 + k);
          if (ho[idx] != expect) { if (bad < 4) printf("  mismatch r=%d k=%d got %.0f want %.0f\n", rr, k, ho[idx], expect); bad++; }
        }
      printf("(a) run: %s, mismatches %d\n", cudaGetErrorString(e), bad);
      bad_total += bad;
    }
  }
  {  // (b) m-major 5-D panel view
    CUtensorMap m;
    cuuint64_t dims[5] = {8, 256, 256 / 8, (cuuint64_t)nb_in, (cuuint64_t)nb_out};
    cuuint64_t strides[4] = {(cuuint64_t)ld * 8, 64, (cuuint64_t)iS * 8, (cuuint64_t)oS * 8};
    cuuint32_t box[5] = {8, 16, 16, 1, 1}, es[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("(b) encode mmajor 5D panel view: %d\n", (int)r);
    if (r == CUDA_SUCCESS) {
      const int k0 = 48, i0 = 128, zi = 1, zo = 1;
      probe_kernel<<<1, 128, 32768>>>(m, 5, 0, k0, i0 / 8, zi, zo, out, 2048);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(ho.data(), out, 2048 * 8, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int mm = 0; mm < 128; mm++)
        for (int k = 0; k < 16; k++) {
          const int idx = (mm / 8) * 16 * 8 + k * 8 + (mm & 7);
          const double expect = (double)(zo * oS + zi * iS + (long long)(k0 + k) * ld + i0 + mm);
          if (ho[idx] != expect) { if (bad < 4) printf("  mismatch m=%d k=%d got %.0f want %.0f\n", mm, k, ho[idx], expect); bad++; }
        }
      printf("(b) run: %s, mismatches %d\n", cudaGetErrorString(e), bad);
      bad_total += bad;
    }
  }
  {  // (c) m-major 2-D swizzled box (16 m x 16 k): element (k, m) at byte k*128 + (((m>>1) ^ (k&7))<<4) + (m&1)*8
    CUtensorMap m;
    cuuint64_t dims[2] = {256, 256};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {16, 16}, es[2] = {1, 1};
    CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("(c) encode mmajor 2D swizzled box: %d\n", (int)r);
    if (r == CUDA_SUCCESS) {
      const int k0 = 16, i0 = 32;
      probe_kernel<<<1, 128, 32768>>>(m, 2, i0, k0, 0, 0, 0, out, 256);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(ho.data(), out, 256 * 8, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int k = 0; k < 16; k++)
        for (int mm = 0; mm < 16; mm++) {
          const int idx = (k * 128 + ((((mm >> 1) ^ (k & 7))) << 4) + (mm & 1) * 8) / 8;
          const double expect = (double)((long long)(k0 + k) * ld + i0 + mm);
          if (ho[idx] != expect) { if (bad < 4) printf("  mismatch m=%d k=%d got %.0f want %.0f\n", mm, k, ho[idx], expect); bad++; }
        }
      printf("(c) run: %s, mismatches %d\n", cudaGetErrorString(e), bad);
      bad_total += bad;
    }
  }
  // host cost of one encode
  {
    CUtensorMap m;
    cuuint64_t dims[4] = {256, 256, 2, 2};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 8, (cuuint64_t)iS * 8, (cuuint64_t)oS * 8};
    cuuint32_t box[4] = {16, 128, 1, 1}, es[4] = {1, 1, 1, 1};
    cudaEvent_t a, b;
    (void)a; (void)b;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 0; i < 10000; i++)
      encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    printf("encode host cost: %.1f ns\n", ((t1.tv_sec - t0.tv_sec) * 1e9 + (t1.tv_nsec - t0.tv_nsec)) / 10000.0);
  }
  printf("TOTAL mismatches %d\n", bad_total);
  return 0;
}
