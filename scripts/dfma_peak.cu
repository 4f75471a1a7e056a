// Measured float64 peaks (the rooflines of the exact scan, search_exact.cu): DFMA on the CUDA cores, and how it
// depends on the warps resident per SM sub-partition (every thread runs ILP independent chains from registers),
// and DMMA.8x8x4 on the FP64 tensor cores.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/dfma_peak scripts/dfma_peak.cu && scripts/dfma_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double x, double y) {
  double a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FP64 tensor cores: DMMA.8x8x4, CH independent accumulator fragments per warp
template <int CH>
__global__ void dmma_kernel(double* out, int iters, double x, double y) {
  double c[CH][2];
#pragma unroll
  for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(x), "d"(y));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
void run_mma(int sms, int ctas_per_sm, int threads, double* out) {
  const int ctas = sms * ctas_per_sm, iters = (1 << 18) / CH;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    dmma_kernel<CH><<<ctas, threads>>>(out, iters, 1e-3, 1e-3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double flops = 2.0 * 256 * CH * (double)iters * ctas * (threads / 32);   // 8 x 8 x 4 FMAs per DMMA
  printf("dmma: %2d warps / SM sub-partition, %2d fragments / warp: %8.3f ms  %6.2f TFLOP/s f64\n",
         ctas_per_sm * threads / 128, CH, best, flops / best / 1e9);
}

template <int ILP>
void run(int sms, int ctas_per_sm, int threads, double* out) {
  const int ctas = sms * ctas_per_sm, iters = (1 << 20) / ILP;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    dfma_kernel<ILP><<<ctas, threads>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double flops = 2.0 * ILP * (double)iters * ctas * threads;
  printf("dfma: %2d warps / SM sub-partition, %2d chains / thread: %8.3f ms  %6.2f TFLOP/s f64\n",
         ctas_per_sm * threads / 128, ILP, best, flops / best / 1e9);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
  run<16>(sms, 8, 256, out);   // 16 warps per sub-partition: the peak
  run<32>(sms, 1, 128, out);   // 1 warp
  run<32>(sms, 1, 256, out);   // 2 warps (the register-blocked scan: 256 threads, 1 CTA / SM)
  run<32>(sms, 1, 512, out);   // 4 warps
  run<16>(sms, 1, 512, out);
  run<32>(sms, 2, 512, out);   // 8 warps
  run_mma<16>(sms, 1, 256, out);   // 2 warps per sub-partition (the DMMA scan)
  run_mma<16>(sms, 2, 512, out);   // 8 warps
  run_mma<8>(sms, 1, 256, out);
  return cudaGetLastError() != cudaSuccess;
}
