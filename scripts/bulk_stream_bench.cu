// Microbenchmark (not part of libtsim): how fast can persistent CTAs stream a contiguous buffer from
// HBM into shared memory with 1-D bulk async copies (cp.async.bulk, UBLKCP), as a function of the copy
// size, ring depth and CTAs per SM?  Used to size the ring of the pooling kernel K1 (pool_norm.cu).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o bulk_stream_bench bulk_stream_bench.cu
//   ./bulk_stream_bench <total MB> <chunk bytes> <stages> <ctas per SM> [item bytes = 98304]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// items of item_bytes are dealt round-robin to CTAs; each item is copied in chunks
__global__ void __launch_bounds__(64) stream_kernel(const unsigned char* x, int64_t nitems, int item_bytes, int chunk,
                                                    int stages, unsigned long long* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = (uint64_t*)(smem + (size_t)stages * chunk);
  uint64_t* empty = full + stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int st = 0; uint32_t ph = 0;
    for (int64_t it = blockIdx.x; it < nitems; it += gridDim.x)
      for (int off = 0; off < item_bytes; off += chunk) {
        const int n = min(chunk, item_bytes - off);
        mbar_wait(smem_u32(&empty[st]), ph ^ 1);
        uint32_t fb = smem_u32(&full[st]);
        mbar_expect(fb, n);
        bulk(smem_u32(smem + (size_t)st * chunk), x + it * item_bytes + off, n, fb);
        if (++st == stages) { st = 0; ph ^= 1; }
      }
  } else if (threadIdx.x == 32) {
    int st = 0; uint32_t ph = 0;
    unsigned long long acc = 0;
    for (int64_t it = blockIdx.x; it < nitems; it += gridDim.x)
      for (int off = 0; off < item_bytes; off += chunk) {
        mbar_wait(smem_u32(&full[st]), ph);
        acc += *(volatile unsigned int*)(smem + (size_t)st * chunk);
        mbar_arrive(smem_u32(&empty[st]));
        if (++st == stages) { st = 0; ph ^= 1; }
      }
    if (acc == 0x12345678ull) *sink = acc;
  }
}

int main(int argc, char** argv) {
  double mb = argc > 1 ? atof(argv[1]) : 400;
  int chunk = argc > 2 ? atoi(argv[2]) : 15360;
  int stages = argc > 3 ? atoi(argv[3]) : 5;
  int per_sm = argc > 4 ? atoi(argv[4]) : 2;
  int item_bytes = argc > 5 ? atoi(argv[5]) : 98304;
  int64_t nitems = (int64_t)(mb * 1e6 / item_bytes);
  size_t total = (size_t)nitems * item_bytes;
  unsigned char* x; CK(cudaMalloc(&x, total)); CK(cudaMemset(x, 0, total));
  unsigned char* flush; CK(cudaMalloc(&flush, 512u << 20));
  unsigned long long* sink; CK(cudaMalloc(&sink, 8));
  size_t smem = (size_t)stages * chunk + stages * 16 + 64;
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int it = 0; it < 6; ++it) {
    CK(cudaMemsetAsync(flush, it, 512u << 20));     // evict L2
    CK(cudaEventRecord(e0));
    stream_kernel<<<sms * per_sm, 64, smem>>>(x, nitems, item_bytes, chunk, stages, sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (it > 0 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  printf("total=%.0f MB chunk=%d stages=%d ctas/SM=%d (%.0f KB in flight/SM) item=%d: %.1f us  %.1f GB/s\n", total / 1e6, chunk,
         stages, per_sm, (double)stages * chunk * per_sm / 1024.0, item_bytes, best * 1e3, total / 1e9 / (best * 1e-3));
  return 0;
}
