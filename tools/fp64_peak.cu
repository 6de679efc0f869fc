// Microbenchmark: FP64 pipe peak on B200 through DMMA.8x8x4 vs DFMA, per warps/SM and ILP.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dmma_loop(double* out, int iters, double a0, double b0) {
  double c[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; i++) { c[i][0] = 0; c[i][1] = 0; }
  double a = a0 + threadIdx.x, b = b0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dfma_loop(double* out, int iters, double a0, double b0) {
  double c[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) c[i] = i;
  double a = a0, b = b0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("device %s SMs %d clock %d kHz\n", p.name, sms, p.clockRate);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    int threads = warps * 32;
    {
      float ms = time_ms([&] { dmma_loop<8><<<sms, threads>>>(out, iters, 1.0, 1e-9); });
      double fl = 2.0 * 256 * 8 * (double)iters * warps * sms;
      printf("DMMA ilp8  warps/SM %2d: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
    }
    {
      float ms = time_ms([&] { dmma_loop<16><<<sms, threads>>>(out, iters, 1.0, 1e-9); });
      double fl = 2.0 * 256 * 16 * (double)iters * warps * sms;
      printf("DMMA ilp16 warps/SM %2d: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
    }
    {
      float ms = time_ms([&] { dmma_loop<2><<<sms, threads>>>(out, iters, 1.0, 1e-9); });
      double fl = 2.0 * 256 * 2 * (double)iters * warps * sms;
      printf("DMMA ilp2  warps/SM %2d: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
    }
    {
      float ms = time_ms([&] { dfma_loop<16><<<sms, threads>>>(out, iters, 1.000001, 1e-9); });
      double fl = 2.0 * 32 * 16 * (double)iters * warps * sms;
      printf("DFMA ilp16 warps/SM %2d: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
    }
  }
  // single-warp latency of dependent DMMA chain
  {
    float ms = time_ms([&] { dmma_loop<1><<<1, 32>>>(out, iters, 1.0, 1e-9); });
    printf("DMMA dependent chain: %.1f ns per DMMA\n", ms * 1e6 / iters);
    float ms2 = time_ms([&] { dfma_loop<1><<<1, 32>>>(out, iters, 1.000001, 1e-9); });
    printf("DFMA dependent chain: %.1f ns per DFMA\n", ms2 * 1e6 / iters);
  }
  return 0;
}
