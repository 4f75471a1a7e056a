// Microbenchmark (not part of libtsim): how fast can one persistent CTA per SM stream a row-major
// [N, D] bf16 matrix from HBM through TMA into shared memory, as a function of the box shape and
// pipeline depth?  Used to choose the small-query (HBM-bound) kernel's tile shape.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_stream_bench tma_stream_bench.cu
//   ./tma_stream_bench <rows> <D> <box_rows> <stages> [kb_per_row_tile_first=0]
#include <cuda.h>
#include <cuda_bf16.h>
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
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

// mode 0: for tile { for kb { load [box_rows x 64] } }   (what search_tc does)
// mode 1: same order but the consumer is trivial either way; kept for future patterns
__global__ void __launch_bounds__(128, 1) stream_kernel(const __grid_constant__ CUtensorMap map, int64_t rows, int kblocks,
                                                        int box_rows, int stages, int stage_bytes, unsigned long long* sink) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = (uint64_t*)(smem + (size_t)stages * stage_bytes);
  uint64_t* empty = full + stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int64_t ntiles = (rows + box_rows - 1) / box_rows;
  if (threadIdx.x == 0) {
    int st = 0; uint32_t ph = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x)
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(smem_u32(&empty[st]), ph ^ 1);
        uint32_t fb = smem_u32(&full[st]);
        mbar_expect(fb, stage_bytes);
        tma2d(smem_u32(smem + (size_t)st * stage_bytes), &map, fb, kb * 64, (int)(t * box_rows));
        if (++st == stages) { st = 0; ph ^= 1; }
      }
  } else if (threadIdx.x == 32) {
    int st = 0; uint32_t ph = 0;
    unsigned long long acc = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x)
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(smem_u32(&full[st]), ph);
        acc += *(volatile unsigned int*)(smem + (size_t)st * stage_bytes);
        mbar_arrive(smem_u32(&empty[st]));
        if (++st == stages) { st = 0; ph ^= 1; }
      }
    if (acc == 0x12345678ull) *sink = acc;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  int64_t rows = argc > 1 ? atoll(argv[1]) : 4000000;
  int D = argc > 2 ? atoi(argv[2]) : 768;
  int box_rows = argc > 3 ? atoi(argv[3]) : 256;
  int stages = argc > 4 ? atoi(argv[4]) : 4;
  int l2promo = argc > 5 ? atoi(argv[5]) : 2;
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fnp;
  __nv_bfloat16* x; CK(cudaMalloc(&x, (size_t)rows * D * 2)); CK(cudaMemset(x, 0, (size_t)rows * D * 2));
  unsigned long long* sink; CK(cudaMalloc(&sink, 8));
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows}; cuuint64_t str[1] = {(cuuint64_t)D * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
  CUtensorMapL2promotion promo = l2promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : l2promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  int stage_bytes = box_rows * 128;
  size_t smem = 1024 + (size_t)stages * stage_bytes + stages * 16 + 64;
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  int kblocks = (D + 63) / 64;
  float best = 1e30f;
  for (int it = 0; it < 6; ++it) {
    CK(cudaEventRecord(e0));
    stream_kernel<<<sms, 128, smem>>>(map, rows, kblocks, box_rows, stages, stage_bytes, sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (it > 0 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  double gb = (double)rows * D * 2 / 1e9;
  printf("rows=%lld D=%d box_rows=%d stages=%d (%.0f KB in flight/SM) l2promo=%d: %.3f ms  %.1f GB/s\n", (long long)rows, D, box_rows,
         stages, stages * stage_bytes / 1024.0, l2promo, best, gb / (best * 1e-3));
  return 0;
}
