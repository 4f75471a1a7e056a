// K2 (exact scan): float64 brute-force cosine + per-slice top-k on the CUDA cores.
//
// This is the path for fp32 inputs (tolerance 1e-5 rules out TF32 tensor cores; the reference's
// own CPU-runnable case, BASELINE config 1, is 10k x 384 fp32), for shapes the TMA path cannot
// take, and the fallback for the few queries whose tensor-core candidates could not be proven
// complete (select_merge.cu).  It restates, in float64, exactly what the reference computes per
// query: F.cosine_similarity(q.expand_as(C), C, -1) (search_pipeline.py:76-77) then the k
// largest (:78), with the north_star tie rule (lower index first).
//
// Layout: grid = (S corpus slices, query groups); one warp per query of a group of 8, one lane
// per corpus row of a 32-row tile staged through shared memory as float64 (coalesced global
// loads, conflict-free column reads).  Each warp keeps a sorted top-k list in shared memory and
// inserts cooperatively.  Lists go to the workspace as [slot][slice][k]; merge_exact_lists
// finishes.  Bytes: N*D*e per group of 8 queries -- this path is not the roofline path.
#include "tsim_common.cuh"

namespace tsim {
namespace {

constexpr int kExThreads = 256;  // 8 warps = 8 queries per group
constexpr int kDC = 128;         // D chunk staged per step
constexpr int kRows = 32;        // corpus rows per tile (one per lane)

struct ExArgs {
  const void* q; int q_dt; int64_t q_stride;
  const void* corpus; int c_dt; int64_t c_stride;
  int64_t Q, N, D; int k; int self_on; int64_t self_off;  // skip row == self_off + query
  int S; int64_t slice_rows;
  const int32_t* flag_cnt; const int32_t* flag_list;
  double* ex_score; uint32_t* ex_idx;
};

// insert (s, r) into a descending list ls/li of length *cnt (capacity k); whole warp calls
__device__ __forceinline__ void warp_list_insert(double* ls, uint32_t* li, int& cnt, int k, double s,
                                                 uint32_t r) {
  const int lane = threadIdx.x & 31;
  if (cnt == k && !(s > ls[k - 1])) return;  // equal score, later row: loses the tie
  int pos = 0;
  for (int base = 0; base < cnt; base += 32) {
    int j = base + lane;
    pos += __popc(__ballot_sync(0xffffffffu, j < cnt && ls[j] >= s));
  }
  const int newcnt = min(cnt + 1, k);
  // shift [pos, newcnt-1) down by one, highest chunk first
  for (int top = newcnt - 1; top > pos; top -= 32) {
    int j = top - lane;
    double vs = 0.0; uint32_t vi = 0;
    bool act = j > pos;
    if (act) { vs = ls[j - 1]; vi = li[j - 1]; }
    __syncwarp();
    if (act) { ls[j] = vs; li[j] = vi; }
    __syncwarp();
  }
  if (lane == 0) { ls[pos] = s; li[pos] = r; }
  cnt = newcnt;
  __syncwarp();
}

__global__ void __launch_bounds__(kExThreads) search_exact_kernel(ExArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* qs = (double*)smem_raw;                     // [8][kDC]
  double* tile = qs + 8 * kDC;                        // [kRows][kDC + 1]
  double* norm2 = tile + kRows * (kDC + 1);           // [kRows]
  double* ls_all = norm2 + kRows;                     // [8][k]
  uint32_t* li_all = (uint32_t*)(ls_all + 8 * a.k);   // [8][k]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t nq = a.flag_cnt ? (int64_t)*a.flag_cnt : a.Q;
  const int64_t ngroups = (nq + 7) / 8;
  const int slice = blockIdx.x;
  const int64_t row_begin = (int64_t)slice * a.slice_rows;
  const int64_t row_end = min(a.N, row_begin + a.slice_rows);
  const int qsz = dtype_size(a.q_dt), csz = dtype_size(a.c_dt);
  double* ls = ls_all + warp * a.k;
  uint32_t* li = li_all + warp * a.k;

  for (int64_t g = blockIdx.y; g < ngroups; g += gridDim.y) {
    const int64_t slot = g * 8 + warp;
    const bool qvalid = slot < nq;
    const int64_t qid = qvalid ? (a.flag_list ? (int64_t)a.flag_list[slot] : slot) : 0;
    const char* qrow = (const char*)a.q + (size_t)qid * a.q_stride * qsz;
    // ||q||^2 in float64 (lane-strided, fixed butterfly)
    double qq = 0.0;
    for (int64_t d = lane; d < a.D; d += 32) {
      double v = (double)load_elem(qrow, a.q_dt, d);
      qq = fma(v, v, qq);
    }
    qq = warp_sum_f64(qq);
    const double qn = fmax(sqrt(qq), kCosEps);
    int cnt = 0;

    for (int64_t r0 = row_begin; r0 < row_end; r0 += kRows) {
      double acc = 0.0;
      if (tid < kRows) norm2[tid] = 0.0;
      for (int64_t d0 = 0; d0 < a.D; d0 += kDC) {
        const int dc = (int)min((int64_t)kDC, a.D - d0);
        __syncthreads();
        // stage the query chunk: warp w stages its own query
        for (int d = lane; d < dc; d += 32) qs[warp * kDC + d] = qvalid ? (double)load_elem(qrow, a.q_dt, d0 + d) : 0.0;
        // stage the row tile: 8 rows per pass, 32 lanes across the chunk
        for (int rr = warp; rr < kRows; rr += 8) {
          const int64_t row = r0 + rr;
          double sq = 0.0;
          if (row < row_end) {
            const char* crow = (const char*)a.corpus + (size_t)row * a.c_stride * csz;
            for (int d = lane; d < dc; d += 32) {
              double v = (double)load_elem(crow, a.c_dt, d0 + d);
              tile[rr * (kDC + 1) + d] = v;
              sq = fma(v, v, sq);
            }
          } else {
            for (int d = lane; d < dc; d += 32) tile[rr * (kDC + 1) + d] = 0.0;
          }
          sq = warp_sum_f64(sq);
          if (lane == 0) norm2[rr] += sq;
        }
        __syncthreads();
        const double* qv = qs + warp * kDC;
        const double* tv = tile + lane * (kDC + 1);
#pragma unroll 4
        for (int d = 0; d < dc; ++d) acc = fma(qv[d], tv[d], acc);
      }
      __syncthreads();
      // lane <-> row r0 + lane
      const int64_t row = r0 + lane;
      const double cn = fmax(sqrt(norm2[lane]), kCosEps);
      const double score = acc / (qn * cn);
      bool want = qvalid && row < row_end && !(a.self_on && row == a.self_off + qid);
      want = want && !(score != score);  // NaN rows are never returned
      unsigned mask = __ballot_sync(0xffffffffu, want);
      while (mask) {
        int src = __ffs(mask) - 1;
        mask &= mask - 1;
        double s = __shfl_sync(0xffffffffu, score, src);
        warp_list_insert(ls, li, cnt, a.k, s, (uint32_t)(r0 + src));
      }
      __syncthreads();
    }
    // write this (slot, slice) list
    if (qvalid) {
      double* os = a.ex_score + ((size_t)slot * a.S + slice) * a.k;
      uint32_t* oi = a.ex_idx + ((size_t)slot * a.S + slice) * a.k;
      for (int j = lane; j < a.k; j += 32) {
        os[j] = j < cnt ? ls[j] : 0.0;
        oi[j] = j < cnt ? li[j] : 0xffffffffu;
      }
    }
    __syncthreads();
  }
}

}  // namespace

int launch_search_exact(const void* q, int q_dt, int64_t q_stride, const void* corpus, int c_dt,
                        int64_t c_stride, int64_t Q, int64_t N, int64_t D, int k,
                        int self_on, int64_t self_off, const SearchPlan& p, const int32_t* flag_cnt,
                        const int32_t* flag_list, double* ex_score, uint32_t* ex_idx,
                        cudaStream_t st) {
  ExArgs a;
  a.q = q; a.q_dt = q_dt; a.q_stride = q_stride;
  a.corpus = corpus; a.c_dt = c_dt; a.c_stride = c_stride;
  a.Q = Q; a.N = N; a.D = D; a.k = k; a.self_on = self_on; a.self_off = self_off;
  a.S = p.S; a.slice_rows = p.slice_rows;
  a.flag_cnt = flag_cnt; a.flag_list = flag_list; a.ex_score = ex_score; a.ex_idx = ex_idx;
  size_t smem = sizeof(double) * (8 * kDC + kRows * (kDC + 1) + kRows + 8 * (size_t)k) + sizeof(uint32_t) * 8 * (size_t)k;
  if (smem > 48 * 1024)
    TSIM_CUDA(cudaFuncSetAttribute(search_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t groups = (Q + 7) / 8;
  // fallback launches do not know the flagged count on the host: a bounded number of group
  // CTAs stride over however many groups there turn out to be (usually none -> they exit)
  int gy = (int)(flag_cnt ? (groups < 8 ? groups : 8) : (groups < 4096 ? groups : 4096));
  dim3 grid((unsigned)p.S, (unsigned)gy);
  search_exact_kernel<<<grid, kExThreads, smem, st>>>(a);
  TSIM_CUDA(cudaGetLastError());
  count_launch();
  return TSIM_OK;
}

}  // namespace tsim
