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
// accumulators: per d it loads 8 query values (4 broadcast LDS.128) and 4 row values for 32 DFMAs.
// A CTA takes 64 queries (warp w: queries 8w..8w+7 of the group) against 128-row tiles (lane l:
// rows l, l+32, l+64, l+96), so the corpus slice is also streamed once per 64 queries instead of
// once per 8.  D is walked in chunks of 32 (lane <-> element while staging); the next chunk's rows
// AND query values are fetched into registers while the current one is multiplied.  Staging stores
// are bank-conflict free: rows with an odd stride (33 doubles), queries as 16-byte pairs with a
// stride of 66 doubles (the first version wrote them 32-way conflicted: 36 % of all shared-memory
// wavefronts, ncu).  Same per-warp lists, same insertion order (rows ascending), same output
// layout as the kernel above; merge_exact_lists re-scores canonically.
constexpr int kBQ = 8;           // queries per warp
constexpr int kBR = 4;           // rows per lane
constexpr int kBGroup = 8 * kBQ; // queries per CTA
constexpr int kBRows = 32 * kBR; // rows per tile
constexpr int kBDC = 32;         // D chunk: one element per lane while staging
constexpr int kBOwn = kBRows / 8;  // tile rows staged by each warp
constexpr int kBQStride = kBGroup + 2;  // doubles per d in the query stage (even: LDS.128 / STS.128 stay aligned)
constexpr int kBTStride = kBDC + 1;     // doubles per row in the tile

// rows warp + 8 i of the tile at r0, element d0 + lane, as floats (exact for every dtype)
template <int DT>
__device__ __forceinline__ void fetch_rows(float (&pre)[kBOwn], const ExArgs& a, int64_t r0, int64_t row_end,
                                           int64_t d, int warp) {
#pragma unroll
  for (int i = 0; i < kBOwn; ++i) {
    const int64_t row = r0 + warp + 8 * i;
    const char* crow = (const char*)a.corpus + (size_t)row * a.c_stride * dtype_size(DT);
    pre[i] = (row < row_end && d < a.D) ? Elem<DT>::ld(crow, d) : 0.f;
  }
}
// element d of the warp's 8 queries
template <int DT>
__device__ __forceinline__ void fetch_queries(float (&qpre)[kBQ], const ExArgs& a, int64_t slot0, int64_t d) {
#pragma unroll
  for (int j = 0; j < kBQ; ++j) {
    const int64_t qid = min(slot0 + j, a.Q - 1);
    const char* qrow = (const char*)a.q + (size_t)qid * a.q_stride * dtype_size(DT);
    qpre[j] = d < a.D ? Elem<DT>::ld(qrow, d) : 0.f;
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
  double* qs = (double*)smem_raw;                       // [kBDC][kBQStride]
  double* tile = qs + kBDC * kBQStride;                 // [kBRows][kBTStride]
  double* norm2 = tile + kBRows * kBTStride;            // [kBRows]
  double* qn_s = norm2 + kBRows;                        // [kBGroup] 1 / query norm
  double* ls_all = qn_s + kBGroup;                      // [kBGroup][k]
  uint32_t* li_all = (uint32_t*)(ls_all + (size_t)kBGroup * a.k);  // [kBGroup][k]
  int* cnt_s = (int*)(li_all + (size_t)kBGroup * a.k);  // [kBGroup] list lengths (warp-private)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ngroups = (a.Q + kBGroup - 1) / kBGroup;
  const int slice = blockIdx.x;
  const int64_t row_begin = (int64_t)slice * a.slice_rows;
  const int64_t row_end = min(a.N, row_begin + a.slice_rows);
  const int qsz = dtype_size(a.q_dt);
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

    float pre[kBOwn];   // the next (tile, chunk): rows warp + 8 i, element lane
    float qpre[kBQ];    // ... and element lane of this warp's queries
    int fc = 0;                   // (tile, chunk) the next fetch reads -- counters, no 64-bit divisions per chunk
    int64_t fr0 = row_begin;
    auto fetch = [&]() {
      const int64_t d = (int64_t)fc * kBDC + lane;
      switch (a.c_dt) {
        case TSIM_F32: fetch_rows<TSIM_F32>(pre, a, fr0, row_end, d, warp); break;
        case TSIM_F16: fetch_rows<TSIM_F16>(pre, a, fr0, row_end, d, warp); break;
        case TSIM_BF16: fetch_rows<TSIM_BF16>(pre, a, fr0, row_end, d, warp); break;
        default: fetch_rows<TSIM_E4M3>(pre, a, fr0, row_end, d, warp); break;
      }
      switch (a.q_dt) {
        case TSIM_F32: fetch_queries<TSIM_F32>(qpre, a, slot0, d); break;
        case TSIM_F16: fetch_queries<TSIM_F16>(qpre, a, slot0, d); break;
        case TSIM_BF16: fetch_queries<TSIM_BF16>(qpre, a, slot0, d); break;
        default: fetch_queries<TSIM_E4M3>(qpre, a, slot0, d); break;
      }
      if (++fc == nchunks) { fc = 0; fr0 += kBRows; }
    };
    if (total) fetch();

    double acc[kBQ][kBR];
    int c = 0;                    // (tile, chunk) being multiplied
    int64_t r0 = row_begin;
    for (int64_t it = 0; it < total; ++it) {
      const int dc = (int)min((int64_t)kBDC, a.D - (int64_t)c * kBDC);
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
        const double v = (double)pre[i];
        tile[rr * kBTStride + lane] = v;
        const double sq = warp_sum_f64(v * v);
        if (lane == 0) norm2[rr] = (c == 0 ? 0.0 : norm2[rr]) + sq;   // row rr belongs to this warp alone
      }
#pragma unroll
      for (int j = 0; j < kBQ; j += 2)   // lane <-> d: 16-byte stores 528 bytes apart, conflict-free
        *reinterpret_cast<double2*>(qs + lane * kBQStride + warp * kBQ + j) = make_double2((double)qpre[j], (double)qpre[j + 1]);
      __syncthreads();
      if (it + 1 < total) fetch();   // in flight while this chunk is multiplied

      const double* qv = qs + warp * kBQ;
      const double* tv = tile + lane * kBTStride;
#pragma unroll 2
      for (int d = 0; d < dc; ++d) {
        double x[kBQ], y[kBR];
#pragma unroll
        for (int j = 0; j < kBQ; j += 2) {
          const double2 t = *reinterpret_cast<const double2*>(qv + d * kBQStride + j);
          x[j] = t.x; x[j + 1] = t.y;
        }
#pragma unroll
        for (int r = 0; r < kBR; ++r) y[r] = tv[r * 32 * kBTStride + d];
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
      if (++c == nchunks) { c = 0; r0 += kBRows; }
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

// ---- FP64 tensor-core variant (DMMA.8x8x4) ----------------------------------------------------
// The register-blocked kernel above is bound by shared-memory wavefronts (16 per 32 DFMAs, ncu: LSU data
// pipe and FP64 pipe co-limited at 33 %).  mma.sync.m8n8k4.f64 multiplies an 8-query x 4-d fragment by a
// 4-d x 8-row fragment from ONE 8-byte load per lane each, so a warp tile of (8 MF) queries x 64 rows
// costs MF + 8 loads per 8 MF DMMAs (64 MF DFMA-equivalents): 0.16-0.28 wavefronts per DFMA instead
// of 0.5.  B200 keeps full-rate FP64 tensor cores (B300 does not), so this is where the exact float64
// path belongs.  CTA = 8 warps x (8 MF) queries against 64-row tiles; queries and rows are staged as
// [row][32 d + 4] doubles (fragment loads and staging stores both conflict-free); D in chunks of 32
// with register prefetch as above.  C fragments: lane holds query (lane >> 2), rows 2 (lane & 3) + {0, 1}
// of each 8 x 8 block; insertion walks them in ascending row order per query.
constexpr int kMRows = 64;             // corpus rows per tile
constexpr int kMNF = kMRows / 8;       // row fragments per tile
constexpr int kMDC = 32;               // D chunk
constexpr int kMStride = kMDC + 4;     // doubles per staged row (4 mod 16: the 16 lanes of a half-warp hit 16 bank pairs)
constexpr int kMOwn = kMRows / 8;      // tile rows staged by each warp

__device__ __forceinline__ void dmma_8x8x4(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int DT, int NR>
__device__ __forceinline__ void fetch_col(float (&pre)[NR], const void* base, int64_t stride, int64_t first, int step,
                                          int64_t limit, bool clamp, int64_t d, int64_t D) {
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    int64_t row = first + (int64_t)step * i;
    const bool ok = clamp || row < limit;
    if (clamp) row = min(row, limit - 1);
    const char* p = (const char*)base + (size_t)row * stride * dtype_size(DT);
    pre[i] = (ok && d < D) ? Elem<DT>::ld(p, d) : 0.f;
  }
}
template <int NR>
__device__ __forceinline__ void fetch_col_dt(float (&pre)[NR], int dt, const void* base, int64_t stride, int64_t first,
                                             int step, int64_t limit, bool clamp, int64_t d, int64_t D) {
  switch (dt) {
    case TSIM_F32: fetch_col<TSIM_F32, NR>(pre, base, stride, first, step, limit, clamp, d, D); break;
    case TSIM_F16: fetch_col<TSIM_F16, NR>(pre, base, stride, first, step, limit, clamp, d, D); break;
    case TSIM_BF16: fetch_col<TSIM_BF16, NR>(pre, base, stride, first, step, limit, clamp, d, D); break;
    default: fetch_col<TSIM_E4M3, NR>(pre, base, stride, first, step, limit, clamp, d, D); break;
  }
}

template <int MF, int MINB>
__global__ void __launch_bounds__(kExThreads, MINB) search_exact_mma_kernel(ExArgs a) {
  constexpr int QW = 8 * MF;        // queries per warp
  constexpr int QG = 8 * QW;        // queries per CTA
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* qs = (double*)smem_raw;                       // [QG][kMStride]
  double* tile = qs + QG * kMStride;                    // [kMRows][kMStride]
  double* norm2 = tile + kMRows * kMStride;             // [kMRows] running squared norms
  double* rinv = norm2 + kMRows;                        // [kMRows] 1 / row norm of the finished tile
  double* qn_s = rinv + kMRows;                         // [QG] 1 / query norm
  double* ls_all = qn_s + QG;                           // [QG][k]
  uint32_t* li_all = (uint32_t*)(ls_all + (size_t)QG * a.k);  // [QG][k]
  int* cnt_s = (int*)(li_all + (size_t)QG * a.k);       // [QG] list lengths (warp-private)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int fr = lane >> 2, fk = lane & 3;              // fragment row / k index of this lane
  const int64_t ngroups = (a.Q + QG - 1) / QG;
  const int slice = blockIdx.x;
  const int64_t row_begin = (int64_t)slice * a.slice_rows;
  const int64_t row_end = min(a.N, row_begin + a.slice_rows);
  const int qsz = dtype_size(a.q_dt);
  const int nchunks = (int)((a.D + kMDC - 1) / kMDC);
  const int64_t ntiles = row_end > row_begin ? (row_end - row_begin + kMRows - 1) / kMRows : 0;
  const int64_t total = ntiles * nchunks;

  for (int64_t g = blockIdx.y; g < ngroups; g += gridDim.y) {
    const int64_t slot0 = g * QG + warp * QW;           // this warp's first query
#pragma unroll 1
    for (int j = 0; j < QW; ++j) {
      const int64_t qid = min(slot0 + j, a.Q - 1);
      const char* qrow = (const char*)a.q + (size_t)qid * a.q_stride * qsz;
      double qq = 0.0;
      for (int64_t d = lane; d < a.D; d += 32) {
        double v = (double)load_elem(qrow, a.q_dt, d);
        qq = fma(v, v, qq);
      }
      qq = warp_sum_f64(qq);
      if (lane == 0) {
        qn_s[warp * QW + j] = 1.0 / fmax(sqrt(qq), kCosEps);
        cnt_s[warp * QW + j] = 0;
      }
    }
    __syncwarp();

    float pre[kMOwn];   // the next (tile, chunk): rows warp + 8 i, element lane
    float qpre[QW];     // ... and element lane of this warp's queries
    int fc = 0;
    int64_t fr0 = row_begin;
    auto fetch = [&]() {
      const int64_t d = (int64_t)fc * kMDC + lane;
      fetch_col_dt<kMOwn>(pre, a.c_dt, a.corpus, a.c_stride, fr0 + warp, 8, row_end, false, d, a.D);
      fetch_col_dt<QW>(qpre, a.q_dt, a.q, a.q_stride, slot0, 1, a.Q, true, d, a.D);
      if (++fc == nchunks) { fc = 0; fr0 += kMRows; }
    };
    if (total) fetch();

    double acc[MF][kMNF][2];
    int c = 0;
    int64_t r0 = row_begin;
    for (int64_t it = 0; it < total; ++it) {
      __syncthreads();  // everyone is done reading the previous chunk
      if (c == 0) {
#pragma unroll
        for (int mi = 0; mi < MF; ++mi)
#pragma unroll
          for (int ni = 0; ni < kMNF; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
      }
#pragma unroll
      for (int i = 0; i < kMOwn; ++i) {
        const int rr = warp + 8 * i;
        const double v = (double)pre[i];
        tile[rr * kMStride + lane] = v;
        const double sq = warp_sum_f64(v * v);
        if (lane == 0) norm2[rr] = (c == 0 ? 0.0 : norm2[rr]) + sq;   // row rr belongs to this warp alone
      }
#pragma unroll
      for (int j = 0; j < QW; ++j) qs[(warp * QW + j) * kMStride + lane] = (double)qpre[j];
      if (c == nchunks - 1) {
        __syncwarp();
        if (lane < kMOwn) rinv[warp + 8 * lane] = 1.0 / fmax(sqrt(norm2[warp + 8 * lane]), kCosEps);
      }
      __syncthreads();
      if (it + 1 < total) fetch();   // in flight while this chunk is multiplied

      const double* qf = qs + (warp * QW + fr) * kMStride + fk;
      const double* tf = tile + fr * kMStride + fk;
#pragma unroll 2
      for (int ks = 0; ks < kMDC / 4; ++ks) {
        double af[MF], bf[kMNF];
#pragma unroll
        for (int mi = 0; mi < MF; ++mi) af[mi] = qf[mi * 8 * kMStride + ks * 4];
#pragma unroll
        for (int ni = 0; ni < kMNF; ++ni) bf[ni] = tf[ni * 8 * kMStride + ks * 4];
#pragma unroll
        for (int mi = 0; mi < MF; ++mi)
#pragma unroll
          for (int ni = 0; ni < kMNF; ++ni) dmma_8x8x4(acc[mi][ni], af[mi], bf[ni]);
      }

      if (c == nchunks - 1) {
        // this tile is complete.  Lane holds query (mi, fr), rows r0 + 8 ni + 2 fk + {0, 1}; the two ballots of
        // a block are walked by ascending (lane, element), i.e. ascending row for every query.
#pragma unroll
        for (int mi = 0; mi < MF; ++mi) {
          const int ql = warp * QW + mi * 8 + fr;          // this lane's query within the CTA
          const int64_t qid = g * QG + ql;
          const double qn = qn_s[ql];
#pragma unroll
          for (int ni = 0; ni < kMNF; ++ni) {
            const int cnt = cnt_s[ql];
            const double thr = cnt < a.k ? -INFINITY : ls_all[(size_t)ql * a.k + a.k - 1];
            const int col = ni * 8 + 2 * fk;
            const double2 ri = *reinterpret_cast<const double2*>(rinv + col);
            const double s0 = acc[mi][ni][0] * qn * ri.x, s1 = acc[mi][ni][1] * qn * ri.y;
            const int64_t row = r0 + col;
            const bool live = qid < a.Q;
            // NaN scores fail `>`: never returned.  Rows of one block are re-checked on insertion.
            const bool w0 = live && row < row_end && !(a.self_on && row == a.self_off + qid) && s0 > thr;
            const bool w1 = live && row + 1 < row_end && !(a.self_on && row + 1 == a.self_off + qid) && s1 > thr;
            unsigned m0 = __ballot_sync(0xffffffffu, w0), m1 = __ballot_sync(0xffffffffu, w1);
            while (m0 | m1) {
              const int src = __ffs(m0 | m1) - 1;
              const bool first = (m0 >> src) & 1u;
              if (first) m0 &= ~(1u << src); else m1 &= ~(1u << src);
              const double s = __shfl_sync(0xffffffffu, first ? s0 : s1, src);
              const int qsrc = warp * QW + mi * 8 + (src >> 2);
              warp_list_insert_call(ls_all + (size_t)qsrc * a.k, li_all + (size_t)qsrc * a.k, cnt_s + qsrc, a.k, s,
                                    (uint32_t)(r0 + ni * 8 + 2 * (src & 3) + (first ? 0 : 1)));
            }
          }
        }
      }
      if (++c == nchunks) { c = 0; r0 += kMRows; }
    }
    // write this warp's (slot, slice) lists
#pragma unroll 1
    for (int j = 0; j < QW; ++j) {
      const int64_t slot = slot0 + j;
      if (slot >= a.Q) break;
      const double* ls = ls_all + (size_t)(warp * QW + j) * a.k;
      const uint32_t* li = li_all + (size_t)(warp * QW + j) * a.k;
      const int cnt = cnt_s[warp * QW + j];
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

size_t mma_smem_bytes(int k, int MF) {
  const size_t QG = 64 * (size_t)MF;
  return sizeof(double) * ((QG + kMRows) * kMStride + 2 * kMRows + QG + QG * k) + sizeof(uint32_t) * QG * k + sizeof(int) * QG;
}

size_t blocked_smem_bytes(int k) {
  return sizeof(double) * ((size_t)kBDC * kBQStride + (size_t)kBRows * kBTStride + kBRows + kBGroup + (size_t)kBGroup * k) +
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
  // whole-call scans of more than a warp's worth of queries: FP64 tensor cores.  Default: 64 queries per CTA,
  // two CTAs per SM (one stages while the other multiplies).  Measured on 1M x 768 fp32 (scripts/ab_exact.py),
  // Q = 1024: k = 10 92.9 ms / k = 100 103.4 ms, against 90.7 / 111.3 ms with 128 queries per CTA and one CTA per
  // SM, 116.4 / 128.5 ms with 64 queries and one CTA; Q = 64: 6.4 ms against 7.9-12.4 ms.
  const char* nomma = getenv("TSIM_NO_MMA_SCAN");       // experiment knob
  if (!flag_cnt && Q > 32 && !(nomma && nomma[0] == '1')) {
    int MF = 1, minb = 2;
    if (const char* v = getenv("TSIM_MMA_VARIANT")) {   // experiment knob: "<8-query fragments per warp>x<CTAs per SM>"
      if (v[0] == '1' && v[1] && v[2] == '1') minb = 1;
      if (v[0] == '2' && mma_smem_bytes(k, 2) <= 227 * 1024) { MF = 2; minb = 1; }
    }
    const size_t msmem = mma_smem_bytes(k, MF);
    const int64_t mgroups = (Q + 64 * MF - 1) / (64 * MF);
    if (msmem <= 227 * 1024 && (int64_t)p.S * mgroups >= 64) {
      dim3 mgrid((unsigned)p.S, (unsigned)(mgroups < 4096 ? mgroups : 4096));
      auto kern = MF == 2 ? search_exact_mma_kernel<2, 1> : minb == 2 ? search_exact_mma_kernel<1, 2> : search_exact_mma_kernel<1, 1>;
      TSIM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
      kern<<<mgrid, kExThreads, msmem, st>>>(a);
      TSIM_CUDA(cudaGetLastError());
      count_launch();
      return TSIM_OK;
    }
  }
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
