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


// ---- register-blocked variant: whole-call scans of >= 33 queries ------------------------------
// The kernel above issues two shared-memory loads per DFMA (3 wavefronts per warp-DFMA: the shared
// memory port caps it at 1/6 of the DFMA rate).  Here a thread owns an 8-query x 4-row block of
// accumulators: per d it loads 8 query values (4 broadcast LDS.128) and 4 row values for 32 DFMAs,
// 12 wavefronts per 16 issue cycles -- DFMA-bound.  A CTA takes 64 queries (warp w: queries 8w..8w+7
// of the group) against 128-row tiles (lane l: rows l, l+32, l+64, l+96), so the corpus slice is
// also streamed once per 64 queries instead of once per 8.  The next (tile, D-chunk) is fetched into
// registers while the current one is multiplied.  Same per-warp lists, same insertion order (rows
// ascending), same output layout as the kernel above; merge_exact_lists re-scores canonically.
constexpr int kBQ = 8;           // queries per warp
constexpr int kBR = 4;           // rows per lane
constexpr int kBGroup = 8 * kBQ; // queries per CTA
constexpr int kBRows = 32 * kBR; // rows per tile
constexpr int kBDC = 64;         // D chunk
constexpr int kBOwn = kBRows / 8;  // tile rows staged by each warp

// rows warp + 8 i of the tile at r0, elements d0 + lane and d0 + lane + 32, as floats (exact for every dtype)
template <int DT>
__device__ __forceinline__ void fetch_rows(float (&pre)[kBOwn][2], const ExArgs& a, int64_t r0, int64_t row_end,
                                           int64_t d0, int warp, int lane) {
#pragma unroll
  for (int i = 0; i < kBOwn; ++i) {
    const int64_t row = r0 + warp + 8 * i;
    const char* crow = (const char*)a.corpus + (size_t)row * a.c_stride * dtype_size(DT);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t d = d0 + lane + 32 * h;
      pre[i][h] = (row < row_end && d < a.D) ? Elem<DT>::ld(crow, d) : 0.f;
    }
  }
}

// one copy of the insertion code for the 32 (query, row group) call sites of the blocked kernel
__device__ __noinline__ void warp_list_insert_call(double* ls, uint32_t* li, int* cnt, int k, double s, uint32_t r) {
  int c = *cnt;
  warp_list_insert(ls, li, c, k, s, r);
  *cnt = c;
}

__global__ void __launch_bounds__(kExThreads) search_exact_blocked_kernel(ExArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* qs = (double*)smem_raw;                       // [kBDC][kBGroup]
  double* tile = qs + kBDC * kBGroup;                   // [kBRows][kBDC + 1]
  double* norm2 = tile + kBRows * (kBDC + 1);           // [kBRows]
  double* qn_s = norm2 + kBRows;                        // [kBGroup] 1 / query norm
  double* ls_all = qn_s + kBGroup;                      // [kBGroup][k]
  uint32_t* li_all = (uint32_t*)(ls_all + (size_t)kBGroup * a.k);  // [kBGroup][k]
  int* cnt_s = (int*)(li_all + (size_t)kBGroup * a.k);  // [kBGroup] list lengths (warp-private)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ngroups = (a.Q + kBGroup - 1) / kBGroup;
  const int slice = blockIdx.x;
  const int64_t row_begin = (int64_t)slice * a.slice_rows;
  const int64_t row_end = min(a.N, row_begin + a.slice_rows);
  const int qsz = dtype_size(a.q_dt), csz = dtype_size(a.c_dt);
  const int nchunks = (int)((a.D + kBDC - 1) / kBDC);
  const int64_t ntiles = row_end > row_begin ? (row_end - row_begin + kBRows - 1) / kBRows : 0;
  const int64_t total = ntiles * nchunks;

  for (int64_t g = blockIdx.y; g < ngroups; g += gridDim.y) {
    const int64_t slot0 = g * kBGroup + warp * kBQ;     // this warp's first query
#pragma unroll 1
    for (int j = 0; j < kBQ; ++j) {
      const int64_t qid = min(slot0 + j, a.Q - 1);
      const char* qrow = (const char*)a.q + (size_t)qid * a.q_stride * qsz;
      double qq = 0.0;
      for (int64_t d = lane; d < a.D; d += 32) {
        double v = (double)load_elem(qrow, a.q_dt, d);
        qq = fma(v, v, qq);
      }
      qq = warp_sum_f64(qq);
      if (lane == 0) {
        qn_s[warp * kBQ + j] = 1.0 / fmax(sqrt(qq), kCosEps);
        cnt_s[warp * kBQ + j] = 0;
      }
    }
    __syncwarp();

    float pre[kBOwn][2];  // the next (tile, chunk): rows warp + 8 i, elements lane and lane + 32
    auto fetch = [&](int64_t it) {
      const int64_t r0 = row_begin + (it / nchunks) * kBRows;
      const int64_t d0 = (int64_t)(it % nchunks) * kBDC;
      switch (a.c_dt) {
        case TSIM_F32: fetch_rows<TSIM_F32>(pre, a, r0, row_end, d0, warp, lane); break;
        case TSIM_F16: fetch_rows<TSIM_F16>(pre, a, r0, row_end, d0, warp, lane); break;
        case TSIM_BF16: fetch_rows<TSIM_BF16>(pre, a, r0, row_end, d0, warp, lane); break;
        default: fetch_rows<TSIM_E4M3>(pre, a, r0, row_end, d0, warp, lane); break;
      }
    };
    if (total) fetch(0);

    double acc[kBQ][kBR];
    for (int64_t it = 0; it < total; ++it) {
      const int c = (int)(it % nchunks);
      const int64_t r0 = row_begin + (it / nchunks) * kBRows;
      const int64_t d0 = (int64_t)c * kBDC;
      const int dc = (int)min((int64_t)kBDC, a.D - d0);
      __syncthreads();  // everyone is done reading the previous chunk
      if (c == 0) {
#pragma unroll
        for (int j = 0; j < kBQ; ++j)
#pragma unroll
          for (int r = 0; r < kBR; ++r) acc[j][r] = 0.0;
      }
#pragma unroll
      for (int i = 0; i < kBOwn; ++i) {
        const int rr = warp + 8 * i;
        const double v0 = (double)pre[i][0], v1 = (double)pre[i][1];
        tile[rr * (kBDC + 1) + lane] = v0;
        tile[rr * (kBDC + 1) + lane + 32] = v1;
        const double sq = warp_sum_f64(fma(v0, v0, v1 * v1));
        if (lane == 0) norm2[rr] = (c == 0 ? 0.0 : norm2[rr]) + sq;   // row rr belongs to this warp alone
      }
#pragma unroll 1
      for (int j = 0; j < kBQ; ++j) {
        const int64_t qid = min(slot0 + j, a.Q - 1);
        const char* qrow = (const char*)a.q + (size_t)qid * a.q_stride * qsz;
        for (int d = lane; d < dc; d += 32) qs[d * kBGroup + warp * kBQ + j] = (double)load_elem(qrow, a.q_dt, d0 + d);
      }
      __syncthreads();
      if (it + 1 < total) fetch(it + 1);   // in flight while this chunk is multiplied

      const double* qv = qs + warp * kBQ;
      const double* tv = tile + lane * (kBDC + 1);
#pragma unroll 2
      for (int d = 0; d < dc; ++d) {
        double x[kBQ], y[kBR];
#pragma unroll
        for (int j = 0; j < kBQ; j += 2) {
          const double2 t = *reinterpret_cast<const double2*>(qv + d * kBGroup + j);
          x[j] = t.x; x[j + 1] = t.y;
        }
#pragma unroll
        for (int r = 0; r < kBR; ++r) y[r] = tv[r * 32 * (kBDC + 1) + d];
#pragma unroll
        for (int j = 0; j < kBQ; ++j)
#pragma unroll
          for (int r = 0; r < kBR; ++r) acc[j][r] = fma(x[j], y[r], acc[j][r]);
      }

      if (c == nchunks - 1) {
        // this tile is complete: lane <-> rows r0 + lane + 32 r; insert in ascending row order
        double cn[kBR];   // 1 / row norm
#pragma unroll
        for (int r = 0; r < kBR; ++r) cn[r] = 1.0 / fmax(sqrt(norm2[lane + 32 * r]), kCosEps);
#pragma unroll
        for (int j = 0; j < kBQ; ++j) {
          const int64_t qid = slot0 + j;
          if (qid >= a.Q) break;    // warp-uniform
          double* ls = ls_all + (size_t)(warp * kBQ + j) * a.k;
          uint32_t* li = li_all + (size_t)(warp * kBQ + j) * a.k;
          int* cnt = cnt_s + warp * kBQ + j;
          const double qn = qn_s[warp * kBQ + j];   // 1 / query norm
#pragma unroll
          for (int r = 0; r < kBR; ++r) {
            const int64_t row = r0 + lane + 32 * r;
            // nomination only (merge_exact_lists re-scores canonically): reciprocals instead of 32 divisions
            const double score = acc[j][r] * qn * cn[r];
            bool want = row < row_end && !(a.self_on && row == a.self_off + qid);
            want = want && !(score != score);  // NaN rows are never returned
            want = want && (*cnt < a.k || score > ls[a.k - 1]);   // rows of one ballot are re-checked on insertion
            unsigned mask = __ballot_sync(0xffffffffu, want);
            while (mask) {
              const int src = __ffs(mask) - 1;
              mask &= mask - 1;
              const double s = __shfl_sync(0xffffffffu, score, src);
              warp_list_insert_call(ls, li, cnt, a.k, s, (uint32_t)(r0 + src + 32 * r));
            }
          }
        }
      }
    }
    // write this warp's (slot, slice) lists
#pragma unroll 1
    for (int j = 0; j < kBQ; ++j) {
      const int64_t slot = slot0 + j;
      if (slot >= a.Q) break;
      const double* ls = ls_all + (size_t)(warp * kBQ + j) * a.k;
      const uint32_t* li = li_all + (size_t)(warp * kBQ + j) * a.k;
      const int cnt = cnt_s[warp * kBQ + j];
      double* os = a.ex_score + ((size_t)slot * a.S + slice) * a.k;
      uint32_t* oi = a.ex_idx + ((size_t)slot * a.S + slice) * a.k;
      for (int i = lane; i < a.k; i += 32) {
        os[i] = i < cnt ? ls[i] : 0.0;
        oi[i] = i < cnt ? li[i] : 0xffffffffu;
      }
    }
    __syncwarp();
  }
}

size_t blocked_smem_bytes(int k) {
  return sizeof(double) * ((size_t)kBDC * kBGroup + (size_t)kBRows * (kBDC + 1) + kBRows + kBGroup + (size_t)kBGroup * k) +
         sizeof(uint32_t) * (size_t)kBGroup * k + sizeof(int) * kBGroup;
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
  // whole-call scans of more than a warp's worth of queries: register-blocked kernel (64 queries per CTA)
  const char* noblk = getenv("TSIM_NO_BLOCKED_SCAN");   // experiment knob
  const int64_t bgroups = (Q + kBGroup - 1) / kBGroup;
  if (!flag_cnt && Q > 32 && (int64_t)p.S * bgroups >= 64 && blocked_smem_bytes(k) <= 227 * 1024 &&
      !(noblk && noblk[0] == '1')) {
    const size_t bsmem = blocked_smem_bytes(k);
    TSIM_CUDA(cudaFuncSetAttribute(search_exact_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem));
    dim3 bgrid((unsigned)p.S, (unsigned)(bgroups < 4096 ? bgroups : 4096));
    search_exact_blocked_kernel<<<bgrid, kExThreads, bsmem, st>>>(a);
    TSIM_CUDA(cudaGetLastError());
    count_launch();
    return TSIM_OK;
  }
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
