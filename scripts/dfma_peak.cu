// Measured float64 FMA peak of the CUDA cores (the roofline of the exact scan, search_exact.cu):
// every thread runs 16 independent DFMA chains from registers.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/dfma_peak scripts/dfma_peak.cu && scripts/dfma_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double x, double y) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int ctas = sms * 8, iters = 1 << 16;
  double* out;
  cudaMalloc(&out, sizeof(double) * ctas * 256);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    dfma_kernel<<<ctas, 256>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 16 * (double)iters * ctas * 256;
    printf("dfma: %d CTAs x 256 threads, %.3f ms, %.2f TFLOP/s f64 (%.1f DFMA / clk / SM at 1.9 GHz)\n", ctas, ms,
           flops / ms / 1e9, flops / 2 / (ms * 1e-3) / sms / 1.9e9);
  }
  return cudaGetLastError() != cudaSuccess;
}
