// K2x (exact scan): float64 brute-force cosine + per-slice top-k.
//
// This is the path for calls the tcgen05 kernel cannot take (k > 100, k > 24 on fp32 / fp16 rows,
// odd widths, misaligned or mixed-dtype inputs; fp32 tolerance 1e-5 rules out TF32) and the
// fallback for the few queries whose tensor-core candidates could not be proven complete
// (select_merge.cu).  It restates, in float64, exactly what the reference computes per query:
// F.cosine_similarity(q.expand_as(C), C, -1) (search_pipeline.py:76-77) then the k largest (:78),
// with the north_star tie rule (lower index first).
//
// Two kernels share the layout grid = (S corpus slices, query groups), per-warp sorted top-k
// lists in shared memory with cooperative insertion, and the output [slot][slice][k] that
// merge_exact_lists (select_merge.cu) re-scores canonically and merges:
//  * search_exact_mma_kernel -- FP64 tensor cores (DMMA.8x8x4), 64 queries per CTA; more than 32
//    queries, k <= 252, >= 64 slices.  See the comment above it.
//  * search_exact_kernel -- one warp per query of a group of 8, one lane per corpus row of a
//    32-row tile staged through shared memory as float64 (DFMA); everything else.
// Bytes: N*D*e per query group -- compute-bound on the FP64 pipes, not the HBM roofline path.
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
  const double* rinv;   // [N] 1 / max(||row||, eps), then [N] max(||row||, eps) itself; tensor-core scan only
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
  pdl_trigger();
  pdl_wait();
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


// one copy of the insertion code for the many call sites of the tensor-core kernel
__device__ __noinline__ void warp_list_insert_call(double* ls, uint32_t* li, int* cnt, int k, double s, uint32_t r) {
  int c = *cnt;
  warp_list_insert(ls, li, c, k, s, r);
  *cnt = c;
}

// ---- FP64 tensor-core variant (DMMA.8x8x4): whole-call scans of more than 32 queries -----------
// The kernel above issues two shared-memory loads per DFMA: the shared-memory port caps it at ~3 TFLOP/s.
// A register-blocked DFMA version (8 queries x 4 rows per thread, removed) reached 11.3 TFLOP/s, still
// bound by shared-memory wavefronts (16 per 32 DFMAs; ncu: LSU data pipe and FP64 pipe co-limited at 33 %).
// mma.sync.m8n8k4.f64 multiplies an 8-query x 4-d fragment by a 4-d x 8-row fragment from ONE 8-byte
// load per lane each, so a warp tile of (8 MF) queries x 64 rows costs MF + 8 loads per 8 MF DMMAs
// (64 MF DFMA-equivalents): 0.16-0.28 wavefronts per DFMA instead of 0.5.  B200 keeps full-rate FP64
// tensor cores (measured 37 TFLOP/s, scripts/dfma_peak.cu; B300 does not), so this is where the exact
// float64 path belongs.  CTA = 8 warps x (8 MF) queries against 64-row tiles; queries and rows are staged
// as [row][32 d + 4] doubles (fragment loads and staging stores both conflict-free); D in chunks of 32,
// the next chunk's rows and query values fetched into registers while the current one is multiplied.
// Inverse row norms come from one pre-pass per call (row_rinv_f64_kernel), not once per query group.
// C fragments: lane holds query (lane >> 2), rows 2 (lane & 3) + {0, 1} of each 8 x 8 block; insertion
// walks them in ascending row order per query.  Same per-warp lists and output layout as the kernel
// above; merge_exact_lists re-scores canonically, so results are bit-identical.
constexpr int kMRows = 64;             // corpus rows per tile
constexpr int kMNF = kMRows / 8;       // row fragments per tile
constexpr int kMDC = 32;               // D chunk
constexpr int kMStride = kMDC + 4;     // doubles per staged row (4 mod 16: the 16 lanes of a half-warp hit 16 bank pairs)
constexpr int kMOwn = kMRows / 8;      // tile rows staged by each warp

__device__ __forceinline__ void dmma_8x8x4(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// element `d` (already folded into p0) of NR rows `step_bytes` apart; rows i >= nvalid read as 0
template <int DT, int NR>
__device__ __forceinline__ void fetch_col(float (&pre)[NR], const char* p0, int64_t step_bytes, int nvalid) {
#pragma unroll
  for (int i = 0; i < NR; ++i) pre[i] = i < nvalid ? Elem<DT>::ld(p0 + i * step_bytes, 0) : 0.f;
}
template <int NR>
__device__ __forceinline__ void fetch_col_dt(float (&pre)[NR], int dt, const void* base, int64_t stride, int64_t first,
                                             int step, int64_t limit, int64_t d, int64_t D) {
  const int esz = dtype_size(dt);
  const char* p0 = (const char*)base + ((size_t)first * stride + d) * esz;
  const int64_t left = d < D ? (limit - first + step - 1) / step : 0;   // rows first + step i < limit
  const int nvalid = (int)max((int64_t)0, min((int64_t)NR, left));
  const int64_t sb = (int64_t)step * stride * esz;
  switch (dt) {
    case TSIM_F32: fetch_col<TSIM_F32, NR>(pre, p0, sb, nvalid); break;
    case TSIM_F16: fetch_col<TSIM_F16, NR>(pre, p0, sb, nvalid); break;
    case TSIM_BF16: fetch_col<TSIM_BF16, NR>(pre, p0, sb, nvalid); break;
    default: fetch_col<TSIM_E4M3, NR>(pre, p0, sb, nvalid); break;
  }
}

// element `d` of NR query rows given by row numbers (shared memory; a flagged-query launch scans scattered rows)
template <int DT, int NR>
__device__ __forceinline__ void fetch_rows_at(float (&pre)[NR], const char* base, int64_t row_bytes, const int32_t* rows,
                                              int nvalid, int64_t d) {
#pragma unroll
  for (int i = 0; i < NR; ++i) pre[i] = i < nvalid ? Elem<DT>::ld(base + rows[i] * row_bytes, d) : 0.f;
}
template <int NR>
__device__ __forceinline__ void fetch_rows_at_dt(float (&pre)[NR], int dt, const void* base, int64_t stride, const int32_t* rows,
                                                 int nvalid, int64_t d, int64_t D) {
  if (d >= D) nvalid = 0;
  const int64_t rb = stride * dtype_size(dt);
  switch (dt) {
    case TSIM_F32: fetch_rows_at<TSIM_F32, NR>(pre, (const char*)base, rb, rows, nvalid, d); break;
    case TSIM_F16: fetch_rows_at<TSIM_F16, NR>(pre, (const char*)base, rb, rows, nvalid, d); break;
    case TSIM_BF16: fetch_rows_at<TSIM_BF16, NR>(pre, (const char*)base, rb, rows, nvalid, d); break;
    default: fetch_rows_at<TSIM_E4M3, NR>(pre, (const char*)base, rb, rows, nvalid, d); break;
  }
}

// 1 / max(||row||, eps) in float64, one warp per row (lane-strided sums, fixed butterfly); a fallback launch
// (flag_cnt given) leaves at once when no query is flagged
__global__ void __launch_bounds__(256) row_rinv_f64_kernel(const void* corpus, int c_dt, int64_t c_stride, int64_t N,
                                                           int64_t D, const int32_t* flag_cnt, double* rinv) {
  pdl_trigger();
  pdl_wait();
  if (flag_cnt && *flag_cnt == 0) return;
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < N; row += warps) {
    const char* crow = (const char*)corpus + (size_t)row * c_stride * dtype_size(c_dt);
    double sq = 0.0;
    for (int64_t d = lane; d < D; d += 32) {
      const double v = (double)load_elem(crow, c_dt, d);
      sq = fma(v, v, sq);
    }
    sq = warp_sum_f64(sq);
    if (lane == 0) { const double cn = fmax(sqrt(sq), kCosEps); rinv[row] = 1.0 / cn; rinv[N + row] = cn; }
  }
}

// The tensor-core scans scale a tile's dot products by RECIPROCAL norms (acc * (1 / ||q||) * (1 / ||c||): two
// multiplies per score), the canonical routine (tsim_common.cuh) divides (dot / (||q|| ||c||)): one ulp apart now and
// then.  For rows whose sums are exact in any order -- small integers, one-hot / count vectors: the rows where two
// DISTINCT rows can have the same cosine, 3 / sqrt(18) = 1 / sqrt(2) -- that ulp was the only difference between a
// scan's score and the canonical one, and it could drop the lower-index row of a tie at the k-th place before the
// canonical re-score saw it (scripts/fuzz_parity.py, `ints`).  So the reciprocal form only FILTERS, with a margin far
// above its rounding: what passes is inserted with the divided form, whose bits equal the canonical score whenever
// the sums are exact (and differ from it only by summation-order noise on rows that cannot tie unless identical).
constexpr double kScanBand = 1e-9;

template <int MF, int MINB>
__global__ void __launch_bounds__(kExThreads, MINB) search_exact_mma_kernel(ExArgs a) {
  constexpr int QW = 8 * MF;        // queries per warp
  constexpr int QG = 8 * QW;        // queries per CTA
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* qs = (double*)smem_raw;                       // [QG][kMStride]
  double* tile = qs + QG * kMStride;                    // [kMRows][kMStride]
  double* rinv = tile + kMRows * kMStride;              // [kMRows] 1 / row norm of the tile being finished
  double* qn_s = rinv + kMRows;                         // [QG] 1 / query norm
  double* qnc_s = qn_s + QG;                            // [QG] the query norm itself (scores that enter a list are divided)
  double* ls_all = qnc_s + QG;                          // [QG][k]
  uint32_t* li_all = (uint32_t*)(ls_all + (size_t)QG * a.k);  // [QG][k]
  int* cnt_s = (int*)(li_all + (size_t)QG * a.k);       // [QG] list lengths (warp-private)
  int32_t* qid_s = cnt_s + QG;                          // [QG] query number of the slot (its row, self exclusion)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int fr = lane >> 2, fk = lane & 3;              // fragment row / k index of this lane
  pdl_trigger();
  pdl_wait();
  // whole-call scan: slot = query.  Fallback launch: slot b = the b-th flagged query, count read on the device
  const int64_t nq = a.flag_cnt ? (int64_t)*a.flag_cnt : a.Q;
  const int64_t ngroups = (nq + QG - 1) / QG;
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
      const int64_t slot = min(slot0 + j, nq - 1);
      const int64_t qid = a.flag_list ? (int64_t)a.flag_list[slot] : slot;
      const char* qrow = (const char*)a.q + (size_t)qid * a.q_stride * qsz;
      double qq = 0.0;
      for (int64_t d = lane; d < a.D; d += 32) {
        double v = (double)load_elem(qrow, a.q_dt, d);
        qq = fma(v, v, qq);
      }
      qq = warp_sum_f64(qq);
      if (lane == 0) {
        qn_s[warp * QW + j] = 1.0 / fmax(sqrt(qq), kCosEps);
        qnc_s[warp * QW + j] = fmax(sqrt(qq), kCosEps);
        cnt_s[warp * QW + j] = 0;
        qid_s[warp * QW + j] = (int32_t)qid;
      }
    }
    __syncwarp();
    const int qlive = (int)max((int64_t)0, min((int64_t)QW, nq - slot0));   // live queries of this warp

    float pre[kMOwn];   // the next (tile, chunk): rows warp + 8 i, element lane
    float qpre[QW];     // ... and element lane of this warp's queries
    int fc = 0;
    int64_t fr0 = row_begin;
    auto fetch = [&]() {
      const int64_t d = (int64_t)fc * kMDC + lane;
      fetch_col_dt<kMOwn>(pre, a.c_dt, a.corpus, a.c_stride, fr0 + warp, 8, row_end, d, a.D);
      fetch_rows_at_dt<QW>(qpre, a.q_dt, a.q, a.q_stride, qid_s + warp * QW, qlive, d, a.D);
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
      for (int i = 0; i < kMOwn; ++i) tile[(warp + 8 * i) * kMStride + lane] = (double)pre[i];
#pragma unroll
      for (int j = 0; j < QW; ++j) qs[(warp * QW + j) * kMStride + lane] = (double)qpre[j];
      if (c == nchunks - 1 && tid < kMRows) rinv[tid] = r0 + tid < row_end ? a.rinv[r0 + tid] : 0.0;
      __syncthreads();
      if (it + 1 < total) fetch();   // in flight while this chunk is multiplied

      const double* qf = qs + (warp * QW + fr) * kMStride + fk;
      const double* tf = tile + fr * kMStride + fk;
#pragma unroll 2   // deeper unrolling measured no faster (scripts/ab_exact.py)
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
          const int ql = warp * QW + mi * 8 + fr;          // this lane's query slot within the CTA
          const int64_t qid = qid_s[ql];
          const bool live = g * QG + ql < nq;
          const double qn = qn_s[ql];
#pragma unroll
          for (int ni = 0; ni < kMNF; ++ni) {
            const int cnt = cnt_s[ql];
            const double thr = (cnt < a.k ? -INFINITY : ls_all[(size_t)ql * a.k + a.k - 1]) - kScanBand;   // a filter (kScanBand)
            const int col = ni * 8 + 2 * fk;
            const double2 ri = *reinterpret_cast<const double2*>(rinv + col);
            const double s0 = acc[mi][ni][0] * qn * ri.x, s1 = acc[mi][ni][1] * qn * ri.y;
            const int64_t row = r0 + col;
            // NaN scores fail `>`: never returned.  Rows of one block are re-checked on insertion.
            const bool w0 = live && row < row_end && !(a.self_on && row == a.self_off + qid) && s0 > thr;
            const bool w1 = live && row + 1 < row_end && !(a.self_on && row + 1 == a.self_off + qid) && s1 > thr;
            unsigned m0 = __ballot_sync(0xffffffffu, w0), m1 = __ballot_sync(0xffffffffu, w1);
            while (m0 | m1) {
              const int src = __ffs(m0 | m1) - 1;
              const bool first = (m0 >> src) & 1u;
              if (first) m0 &= ~(1u << src); else m1 &= ~(1u << src);
              const double dsrc = __shfl_sync(0xffffffffu, first ? acc[mi][ni][0] : acc[mi][ni][1], src);
              const int qsrc = warp * QW + mi * 8 + (src >> 2);
              const int64_t rsrc = r0 + ni * 8 + 2 * (src & 3) + (first ? 0 : 1);
              const double s = dsrc / (qnc_s[qsrc] * a.rinv[a.N + rsrc]);      // the canonical routine's last step
              warp_list_insert_call(ls_all + (size_t)qsrc * a.k, li_all + (size_t)qsrc * a.k, cnt_s + qsrc, a.k, s, (uint32_t)rsrc);
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
      if (slot >= nq) break;
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

// ---- the same scan with an asynchronous raw-row pipeline (round 2) ---------------------------------------------------
// search_exact_mma_kernel above fetches the next (tile, chunk) into registers, widens it to float64 and stores it to
// shared memory between TWO block barriers per 32-wide chunk: ncu showed the DMMA sub-pipe 46 % busy, 583 staging
// instructions per 64 DMMAs, and the kernel stopped at 68 % of the measured DMMA peak.  Here the rows travel RAW:
// every thread issues 16-byte cp.async copies (LDGSTS) of the chunk after next into a ring of kAStages stages -- no
// registers, no conversion, no store instructions -- there is ONE barrier per chunk, and a fragment element is
// widened when it is loaded (one 2/4-byte shared-memory load + one conversion instead of one 8-byte load: half the
// shared-memory wavefronts).  Row stride in a stage = 32 elements + 16 bytes: 144 / 80 / 48 bytes for 4 / 2 / 1-byte
// elements puts the 8 fragment rows x 4 k of a load on 32 distinct banks (2- and 1-byte elements share words within
// a row: broadcast).  Needs q_dt == c_dt, 16-byte aligned bases and row pitches, D * element size % 16 == 0; anything
// else takes the kernel above.  Same lists, same output, bit-identical results.
template <int DT> struct AsyncCfg {
  static constexpr int ESZ = DT == TSIM_F32 ? 4 : DT == TSIM_E4M3 ? 1 : 2;
  static constexpr int ROWB = kMDC * ESZ + 16;            // bytes per staged row
  static constexpr int PPR = kMDC * ESZ / 16;             // 16-byte pieces per row
};

__device__ __forceinline__ void cp_async_16_zfill(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}

// MF = query fragments per warp (8 MF queries per warp, 64 MF per CTA): with two, a k-step is 2 + 8 fragment loads
// and conversions for 16 DMMAs instead of 1 + 8 for 8.
template <int DT, int kAStages, int MF>
__global__ void __launch_bounds__(kExThreads, 2) search_exact_mma_async_kernel(ExArgs a) {
  using C = AsyncCfg<DT>;
  constexpr int QW = 8 * MF, QG = 64 * MF;
  constexpr int SROWS = kMRows + QG;                       // rows per stage: 64 corpus rows + the CTA's queries
  constexpr int STAGE = SROWS * C::ROWB;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* ring = smem_raw;                                        // [kAStages][128 rows][ROWB]
  double* rinv = (double*)(ring + (size_t)kAStages * STAGE);          // [2][kMRows] 1 / row norm, by tile parity
  double* qn_s = rinv + 2 * kMRows;                                      // [QG] 1 / query norm
  double* qnc_s = qn_s + QG;                                             // [QG] the query norm itself
  double* ls_all = qnc_s + QG;                                           // [QG][k]
  uint32_t* li_all = (uint32_t*)(ls_all + (size_t)QG * a.k);             // [QG][k]
  int* cnt_s = (int*)(li_all + (size_t)QG * a.k);                        // [QG] list lengths (warp-private)
  int32_t* qid_s = cnt_s + QG;                                           // [QG] query number of the slot

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int fr = lane >> 2, fk = lane & 3;
  pdl_trigger();
  pdl_wait();
  const int64_t nq = a.flag_cnt ? (int64_t)*a.flag_cnt : a.Q;
  const int64_t ngroups = (nq + QG - 1) / QG;
  const int slice = blockIdx.x;
  const int64_t row_begin = (int64_t)slice * a.slice_rows;
  const int64_t row_end = min(a.N, row_begin + a.slice_rows);
  const int nchunks = (int)((a.D + kMDC - 1) / kMDC);
  const int64_t ntiles = row_end > row_begin ? (row_end - row_begin + kMRows - 1) / kMRows : 0;
  const int64_t total = ntiles * nchunks;
  const int64_t row_bytes = a.D * C::ESZ;
  const int64_t c_pitch = a.c_stride * C::ESZ, q_pitch = a.q_stride * C::ESZ;

  for (int64_t g = blockIdx.y; g < ngroups; g += gridDim.y) {
    const int64_t slot0 = g * QG + warp * QW;
#pragma unroll 1
    for (int j = 0; j < QW; ++j) {
      const int64_t slot = min(slot0 + j, nq - 1);
      const int64_t qid = a.flag_list ? (int64_t)a.flag_list[slot] : slot;
      const char* qrow = (const char*)a.q + (size_t)qid * q_pitch;
      double qq = 0.0;
      for (int64_t d = lane; d < a.D; d += 32) {
        double v = (double)Elem<DT>::ld(qrow, d);
        qq = fma(v, v, qq);
      }
      qq = warp_sum_f64(qq);
      if (lane == 0) {
        qn_s[warp * QW + j] = 1.0 / fmax(sqrt(qq), kCosEps);
        qnc_s[warp * QW + j] = fmax(sqrt(qq), kCosEps);
        cnt_s[warp * QW + j] = 0;
        qid_s[warp * QW + j] = (int32_t)qid;
      }
    }
    __syncthreads();      // qid_s is read by every thread's copies

    // issue the copies of iteration `it` (tile it / nchunks, chunk it % nchunks) into its stage
    auto issue = [&](int64_t it) {
      if (it < total) {
        const int64_t tile = it / nchunks;
        const int c = (int)(it - tile * nchunks);
        const int64_t r0 = row_begin + tile * kMRows;
        unsigned char* st = ring + (size_t)(it % kAStages) * STAGE;
#pragma unroll
        for (int p = tid; p < SROWS * C::PPR; p += kExThreads) {
          const int row = p / C::PPR, piece = p - row * C::PPR;
          const int64_t boff = (int64_t)c * (kMDC * C::ESZ) + piece * 16;      // byte offset within the source row
          const char* src;
          bool ok = boff < row_bytes;
          if (row < kMRows) {
            ok = ok && r0 + row < row_end;
            src = (const char*)a.corpus + (size_t)(ok ? r0 + row : row_begin) * c_pitch + (ok ? boff : 0);
          } else {
            ok = ok && g * QG + (row - kMRows) < nq;
            src = (const char*)a.q + (size_t)qid_s[row - kMRows] * q_pitch + (ok ? boff : 0);
          }
          cp_async_16_zfill(st + (size_t)row * C::ROWB + piece * 16, src, ok);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int i = 0; i < kAStages - 1; ++i) issue(i);

    double acc[MF][kMNF][2];
    int c = 0;
    int64_t r0 = row_begin, tile_no = 0;
    for (int64_t it = 0; it < total; ++it) {
      asm volatile("cp.async.wait_group %0;" ::"n"(kAStages - 2) : "memory");
      if (c == 0 && tid < kMRows) rinv[(tile_no & 1) * kMRows + tid] = r0 + tid < row_end ? a.rinv[r0 + tid] : 0.0;
      __syncthreads();          // this chunk has landed for everybody; everybody is done with the stage refilled next
      issue(it + kAStages - 1);
      if (c == 0) {
#pragma unroll
        for (int mi = 0; mi < MF; ++mi)
#pragma unroll
          for (int ni = 0; ni < kMNF; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
      }
      const unsigned char* st = ring + (size_t)(it % kAStages) * STAGE;
      const unsigned char* tf = st + (size_t)fr * C::ROWB;                          // corpus rows fr + 8 ni
      const unsigned char* qf = st + (size_t)(kMRows + warp * QW + fr) * C::ROWB;   // this warp's queries fr + 8 mi
#pragma unroll
      for (int ks = 0; ks < kMDC / 4; ++ks) {
        double af[MF], bf[kMNF];
#pragma unroll
        for (int mi = 0; mi < MF; ++mi) af[mi] = (double)Elem<DT>::ld(qf + (size_t)mi * 8 * C::ROWB, ks * 4 + fk);
#pragma unroll
        for (int ni = 0; ni < kMNF; ++ni) bf[ni] = (double)Elem<DT>::ld(tf + (size_t)ni * 8 * C::ROWB, ks * 4 + fk);
#pragma unroll
        for (int mi = 0; mi < MF; ++mi)
#pragma unroll
          for (int ni = 0; ni < kMNF; ++ni) dmma_8x8x4(acc[mi][ni], af[mi], bf[ni]);
      }

      if (c == nchunks - 1) {
        const double* ri_t = rinv + (tile_no & 1) * kMRows;
#pragma unroll
        for (int mi = 0; mi < MF; ++mi) {
        const int ql = warp * QW + mi * 8 + fr;
        const int64_t qid = qid_s[ql];
        const bool live = g * QG + ql < nq;
        const double qn = qn_s[ql];
#pragma unroll
        for (int ni = 0; ni < kMNF; ++ni) {
          const int cnt = cnt_s[ql];
          const double thr = (cnt < a.k ? -INFINITY : ls_all[(size_t)ql * a.k + a.k - 1]) - kScanBand;   // a filter (kScanBand)
          const int col = ni * 8 + 2 * fk;
          const double2 ri = *reinterpret_cast<const double2*>(ri_t + col);
          const double s0 = acc[mi][ni][0] * qn * ri.x, s1 = acc[mi][ni][1] * qn * ri.y;
          const int64_t row = r0 + col;
          const bool w0 = live && row < row_end && !(a.self_on && row == a.self_off + qid) && s0 > thr;
          const bool w1 = live && row + 1 < row_end && !(a.self_on && row + 1 == a.self_off + qid) && s1 > thr;
          unsigned m0 = __ballot_sync(0xffffffffu, w0), m1 = __ballot_sync(0xffffffffu, w1);
          while (m0 | m1) {
            const int src = __ffs(m0 | m1) - 1;
            const bool first = (m0 >> src) & 1u;
            if (first) m0 &= ~(1u << src); else m1 &= ~(1u << src);
            const double dsrc = __shfl_sync(0xffffffffu, first ? acc[mi][ni][0] : acc[mi][ni][1], src);
            const int qsrc = warp * QW + mi * 8 + (src >> 2);
            const int64_t rsrc = r0 + ni * 8 + 2 * (src & 3) + (first ? 0 : 1);
            const double s = dsrc / (qnc_s[qsrc] * a.rinv[a.N + rsrc]);        // the canonical routine's last step
            warp_list_insert_call(ls_all + (size_t)qsrc * a.k, li_all + (size_t)qsrc * a.k, cnt_s + qsrc, a.k, s, (uint32_t)rsrc);
          }
        }
        }
      }
      if (++c == nchunks) { c = 0; r0 += kMRows; ++tile_no; }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // write this warp's (slot, slice) lists
#pragma unroll 1
    for (int j = 0; j < QW; ++j) {
      const int64_t slot = slot0 + j;
      if (slot >= nq) break;
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
    __syncthreads();      // the next group rewrites qid_s / the ring
  }
}

size_t mma_async_smem_bytes(int k, int dt, int stages, int mf) {
  const int esz = dtype_size(dt);
  const size_t qg = 64 * (size_t)mf;
  const size_t stage = (kMRows + qg) * (kMDC * esz + 16);
  return stages * stage + sizeof(double) * (2 * kMRows + 2 * qg + qg * k) + sizeof(uint32_t) * qg * k + 2 * sizeof(int) * qg;
}

size_t mma_smem_bytes(int k, int MF) {
  const size_t QG = 64 * (size_t)MF;
  return sizeof(double) * ((QG + kMRows) * kMStride + kMRows + 2 * QG + QG * k) + sizeof(uint32_t) * QG * k + 2 * sizeof(int) * QG;
}

}  // namespace

int launch_search_exact(const void* q, int q_dt, int64_t q_stride, const void* corpus, int c_dt,
                        int64_t c_stride, int64_t Q, int64_t N, int64_t D, int k,
                        int self_on, int64_t self_off, const SearchPlan& p, const int32_t* flag_cnt,
                        const int32_t* flag_list, double* ex_score, uint32_t* ex_idx,
                        double* ex_rinv, cudaStream_t st) {
  ExArgs a;
  a.q = q; a.q_dt = q_dt; a.q_stride = q_stride;
  a.corpus = corpus; a.c_dt = c_dt; a.c_stride = c_stride;
  a.Q = Q; a.N = N; a.D = D; a.k = k; a.self_on = self_on; a.self_off = self_off;
  a.S = p.S; a.slice_rows = p.slice_rows;
  a.flag_cnt = flag_cnt; a.flag_list = flag_list; a.ex_score = ex_score; a.ex_idx = ex_idx; a.rinv = ex_rinv;
  // whole-call scans of more than a warp's worth of queries: FP64 tensor cores.  Default: 64 queries per CTA,
  // two CTAs per SM (one stages while the other multiplies).  Measured on 1M x 768 fp32 (scripts/ab_exact.py),
  // Q = 1024: k = 10 92.9 ms / k = 100 103.4 ms, against 90.7 / 111.3 ms with 128 queries per CTA and one CTA per
  // SM, 116.4 / 128.5 ms with 64 queries and one CTA; Q = 64: 6.4 ms against 7.9-12.4 ms.
  // (a fallback launch does not know how many queries are flagged: up to 8 group CTAs per slice stride over them)
  if (ex_rinv && Q > 32 && N > 0 && !knob_on("TSIM_NO_MMA_SCAN")) {   // experiment knob
    int MF = 1, minb = 2;
    // experiment knob: 11 = one 8-query fragment per warp x one CTA per SM, 21 = two fragments x one CTA
    const int variant = knob_int("TSIM_MMA_VARIANT", 12);
    if (variant == 11) minb = 1;
    if (variant == 21 && mma_smem_bytes(k, 2) <= 227 * 1024) { MF = 2; minb = 1; }
    const size_t msmem = mma_smem_bytes(k, MF);
    int64_t mgroups = (Q + 64 * MF - 1) / (64 * MF);
    if (flag_cnt && mgroups > 8) mgroups = 8;
    if (msmem <= 227 * 1024 && (int64_t)p.S * mgroups >= 64) {
      const int64_t nb = (N + 7) / 8;
      TSIM_CUDA(launch_pdl(row_rinv_f64_kernel, dim3((unsigned)(nb < 8 * 148 ? nb : 8 * 148)), dim3(256), 0, st, corpus, c_dt,
                           c_stride, N, D, flag_cnt, ex_rinv));
      count_launch();
      dim3 mgrid((unsigned)p.S, (unsigned)(mgroups < 4096 ? mgroups : 4096));
      // asynchronous raw-row pipeline when the operands allow 16-byte cp.async copies
      const int esz = dtype_size(c_dt);
      // (4 stages, or 3 when that is what lets two CTAs share an SM -- one multiplies while the other waits at its
      // barrier; with one CTA per SM -- long lists, k > ~75 on 4-byte rows -- the kernel above measured faster)
      const size_t two_per_sm = (227 * 1024) / 2 - 1024;
      // two query fragments per warp (128 queries per CTA) when there are enough queries and the lists still leave
      // room for two CTAs per SM; experiment knob TSIM_ASYNC_MF=1 keeps one
      int amf = (Q >= 1024 && !flag_cnt && mma_async_smem_bytes(k, c_dt, 3, 2) <= two_per_sm) ? 2 : 1;
      if (knob_int("TSIM_ASYNC_MF", 0) == 1) amf = 1;
      const int nst = mma_async_smem_bytes(k, c_dt, 4, amf) <= two_per_sm ? 4 : 3;
      const size_t asmem = mma_async_smem_bytes(k, c_dt, nst, amf);
      if (amf == 2) {
        const int64_t g2 = (Q + 127) / 128;
        mgrid.y = (unsigned)(g2 < 4096 ? g2 : 4096);
      }
      const bool async_ok = q_dt == c_dt && MF == 1 && variant == 12 && (((uintptr_t)q | (uintptr_t)corpus) & 15) == 0 &&
                            (q_stride * esz) % 16 == 0 && (c_stride * esz) % 16 == 0 && (D * esz) % 16 == 0 &&
                            asmem <= two_per_sm && !knob_on("TSIM_NO_ASYNC_SCAN");
      if (async_ok) {
#define TSIM_ASYNC_KERN(NST, MFA) (c_dt == TSIM_F32 ? search_exact_mma_async_kernel<TSIM_F32, NST, MFA>        \
                                   : c_dt == TSIM_F16 ? search_exact_mma_async_kernel<TSIM_F16, NST, MFA>      \
                                   : c_dt == TSIM_BF16 ? search_exact_mma_async_kernel<TSIM_BF16, NST, MFA>    \
                                                       : search_exact_mma_async_kernel<TSIM_E4M3, NST, MFA>)
        auto akern = amf == 2 ? (nst == 4 ? TSIM_ASYNC_KERN(4, 2) : TSIM_ASYNC_KERN(3, 2))
                              : (nst == 4 ? TSIM_ASYNC_KERN(4, 1) : TSIM_ASYNC_KERN(3, 1));
#undef TSIM_ASYNC_KERN
        TSIM_CUDA(cudaFuncSetAttribute(akern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)asmem));
        TSIM_CUDA(launch_pdl(akern, mgrid, dim3(kExThreads), asmem, st, a));
        count_launch();
        return TSIM_OK;
      }
      auto kern = MF == 2 ? search_exact_mma_kernel<2, 1> : minb == 2 ? search_exact_mma_kernel<1, 2> : search_exact_mma_kernel<1, 1>;
      TSIM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
      TSIM_CUDA(launch_pdl(kern, mgrid, dim3(kExThreads), msmem, st, a));
      count_launch();
      return TSIM_OK;
    }
  }
  size_t smem = sizeof(double) * (8 * kDC + kRows * (kDC + 1) + kRows + 8 * (size_t)k) + sizeof(uint32_t) * 8 * (size_t)k;
  if (smem > 48 * 1024)
    TSIM_CUDA(cudaFuncSetAttribute(search_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t groups = (Q + 7) / 8;
  // fallback launches do not know the flagged count on the host: a bounded number of group
  // CTAs stride over however many groups there turn out to be (usually none -> they exit)
  int gy = (int)(flag_cnt ? (groups < 8 ? groups : 8) : (groups < 4096 ? groups : 4096));
  dim3 grid((unsigned)p.S, (unsigned)gy);
  TSIM_CUDA(launch_pdl(search_exact_kernel, grid, dim3(kExThreads), smem, st, a));
  count_launch();
  return TSIM_OK;
}

}  // namespace tsim
