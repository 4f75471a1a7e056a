// K3: second-pass selection over candidate lists.
//
//  * select_rescore_kernel  -- per query: gather the packed (approx score, row) keys that the
//    tensor-core pass left in its per-unit lists, keep the best KP, PROVE that no row outside
//    them can be in the exact top-k (else flag the query for the float64 scan), recompute the
//    survivors' cosine in float64 with the canonical warp routine and emit the top-k ranked by
//    (score desc, row asc).
//  * merge_exact_lists_kernel -- same ending for the lists written by the exact scan.
//  * merge_topk_kernel -- tsim_merge_topk: merge of per-shard / per-chunk result lists (the
//    merge the reference lacks: search_pipeline.py:83,88 overwrites).
//
// HBM-bound, tiny: algorithmic bytes = Q * n_lists * k * 12 in + Q * k * 12 out.
#include "tsim_common.cuh"

namespace tsim {
namespace {

constexpr int kSelThreads = 256;
constexpr int kKeyCap = 2048;   // packed keys held in smem by select_rescore
constexpr int kPairCap = 2048;  // (double, int64) pairs held in smem by the pair merges

__device__ __forceinline__ int next_pow2(int n) {
  int p = 1;
  while (p < n) p <<= 1;
  return p;
}

// Small inputs (n <= blockDim.x) are sorted by RANK: every thread counts the elements that precede its own (n reads
// of shared memory, broadcast) and drops it at that position -- two barriers instead of the log^2(n) barrier
// stages of the bitonic network (28 for n = 128; the two sorts were ~20 % of select_rescore's samples at k = 100).
// block-wide bitonic sort, DESCENDING, n a power of two, keys in shared memory
__device__ void sort_keys_desc(uint64_t* a, int n) {
  if (n <= (int)blockDim.x) {
    const int t = threadIdx.x;
    uint64_t mine = 0;
    int rank = 0;
    if (t < n) {
      mine = a[t];
      for (int j = 0; j < n; ++j) {
        const uint64_t o = a[j];
        rank += (o > mine || (o == mine && j < t)) ? 1 : 0;    // equal keys (empty slots) keep their order
      }
    }
    __syncthreads();
    if (t < n) a[rank] = mine;
    __syncthreads();
    return;
  }
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          uint64_t x = a[i], y = a[ixj];
          bool desc = ((i & k) == 0);
          if (desc ? (x < y) : (x > y)) { a[i] = y; a[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }
}

// Keep only the keys whose approximate SCORE (upper 32 bits) is at least the `keep`-th best score:
// MSB-first radix select over the 32 score bits (4 passes of 8 bits, shared-memory histogram, one
// warp picks the bucket), then a compaction through `out` (kSelOut entries).  Returns the new count
// m (>= keep, ties at the cut included; keys[0..m) in arbitrary order), or n unchanged when the
// survivors would not fit `out` (massive ties: the caller's full sort handles that).  Replaces a
// 2048-element bitonic sort (66 barrier stages) whenever only the best few dozen of ~2000 keys matter.
constexpr int kSelOut = 512;
__device__ int keep_top_scores(uint64_t* keys, int n, int keep, uint64_t* out, uint32_t* hist, int* sh) {
  if (n <= keep) return n;
  const int tid = threadIdx.x;
  // The scores of one list share their upper bits (cosines of a narrow band), and a first pass that drops every key into
  // one bucket serialises on that bucket's shared-memory atomic.  Select over (score - worst) instead, starting at the top
  // bit of the span: the first pass then spreads over 128+ buckets and passes whose bits are all equal are never run.
  if (tid == 0) { hist[0] = 0u; hist[1] = 0xffffffffu; }
  __syncthreads();
  {
    uint32_t mx = 0u, mn = 0xffffffffu;
    for (int i = tid; i < n; i += blockDim.x) {
      const uint32_t sc = (uint32_t)(keys[i] >> 32);
      mx = max(mx, sc); mn = min(mn, sc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if ((tid & 31) == 0) { atomicMax(&hist[0], mx); atomicMin(&hist[1], mn); }
  }
  __syncthreads();
  const uint32_t worst = hist[1];
  int hi = 32 - __clz(hist[0] - worst);   // bits >= hi of (score - worst) are zero in every key
  __syncthreads();
  uint32_t prefix = 0;   // bits of (k-th best - worst) decided so far (upper bits)
  int need = keep;
  while (hi > 0) {
    const int shift = hi > 8 ? hi - 8 : 0;
    const uint32_t bmask = (1u << (hi - shift)) - 1u;
    const uint32_t pre_hi = hi < 32 ? prefix >> hi : 0u;
    for (int i = tid; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
      const uint32_t rel = (uint32_t)(keys[i] >> 32) - worst;
      if ((hi < 32 ? rel >> hi : 0u) == pre_hi) atomicAdd(&hist[(rel >> shift) & bmask], 1u);
    }
    __syncthreads();
    if (tid < 32) {
      // lane l owns buckets 255 - 8l .. 248 - 8l (descending); find where the running count reaches `need`
      uint32_t c[8], tot = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { c[j] = hist[255 - 8 * tid - j]; tot += c[j]; }
      uint32_t incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (tid >= o) incl += v;
      }
      const uint32_t before = incl - tot;
      if (before < (uint32_t)need && incl >= (uint32_t)need) {
        uint32_t run = before;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (run < (uint32_t)need && run + c[j] >= (uint32_t)need) { sh[0] = 255 - 8 * tid - j; sh[1] = need - (int)run; }
          run += c[j];
        }
      }
    }
    __syncthreads();
    prefix |= (uint32_t)sh[0] << shift;
    need = sh[1];
    hi = shift;
    __syncthreads();
  }
  prefix += worst;
  // prefix = the keep-th best score; everything at or above it survives
  if (tid == 0) sh[2] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) {
    const uint64_t key = keys[i];
    if ((uint32_t)(key >> 32) >= prefix) {
      const int p = atomicAdd(&sh[2], 1);
      if (p < kSelOut) out[p] = key;
    }
  }
  __syncthreads();
  const int m = sh[2];
  if (m > kSelOut) return n;
  for (int i = tid; i < m; i += blockDim.x) keys[i] = out[i];
  __syncthreads();
  return m;
}

// (score desc, index asc) order on pairs; idx < 0 (padding) ranks last
__device__ __forceinline__ bool pair_before(double s1, int64_t i1, double s2, int64_t i2) {
  bool p1 = i1 < 0, p2 = i2 < 0;
  if (p1 != p2) return p2;
  if (p1) return false;
  if (s1 != s2) return s1 > s2;
  return i1 < i2;
}

__device__ void sort_pairs(double* s, int64_t* ix, int n, int rank_max = 64) {
  // (a pair comparison is ~4x the instructions of a key comparison: the rank sort's n^2 compares only beat the
  // bitonic network's barriers for short inputs)
  if (n <= rank_max && n <= (int)blockDim.x) {       // rank sort, see sort_keys_desc
    const int t = threadIdx.x;
    double ms = 0.0;
    int64_t mi = -1;
    int rank = 0;
    if (t < n) {
      ms = s[t]; mi = ix[t];
      for (int j = 0; j < n; ++j) {
        const double os = s[j];
        const int64_t oi = ix[j];
        const bool same = (oi < 0 && mi < 0) || (oi == mi && os == ms);   // padding, or the same pair twice
        rank += (same ? j < t : pair_before(os, oi, ms, mi)) ? 1 : 0;
      }
    }
    __syncthreads();
    if (t < n) { s[rank] = ms; ix[rank] = mi; }
    __syncthreads();
    return;
  }
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          double s1 = s[i], s2 = s[ixj];
          int64_t i1 = ix[i], i2 = ix[ixj];
          bool first_block = ((i & k) == 0);
          bool swap = first_block ? pair_before(s2, i2, s1, i1) : pair_before(s1, i1, s2, i2);
          if (swap) { s[i] = s2; s[ixj] = s1; ix[i] = i2; ix[ixj] = i1; }
        }
      }
      __syncthreads();
    }
  }
}

// Streaming top-`keep` over `total` source pairs using a kPairCap smem buffer.
// load(i, &s, &ix) fetches source pair i.  On return s/ix[0..n) is sorted, n <= kPairCap.
template <class Load>
__device__ int stream_top_pairs(double* s, int64_t* ix, int* n_sh, int64_t total, int keep, Load load) {
  if (threadIdx.x == 0) *n_sh = 0;
  __syncthreads();
  int64_t cursor = 0;
  int n = 0;
  while (cursor < total) {
    int64_t take = min((int64_t)(kPairCap - n), total - cursor);
    for (int64_t i = threadIdx.x; i < take; i += blockDim.x) {
      double v; int64_t id;
      load(cursor + i, &v, &id);
      if (id >= 0) {
        int p = atomicAdd(n_sh, 1);
        s[p] = v; ix[p] = id;
      }
    }
    cursor += take;
    __syncthreads();
    n = *n_sh;
    if (cursor < total && n > kPairCap / 2) {
      int np = next_pow2(n);
      for (int i = n + threadIdx.x; i < np; i += blockDim.x) { s[i] = 0.0; ix[i] = -1; }
      __syncthreads();
      sort_pairs(s, ix, np);
      n = min(n, keep);
      if (threadIdx.x == 0) *n_sh = n;
      __syncthreads();
    }
  }
  int np = next_pow2(max(n, 1));
  for (int i = n + threadIdx.x; i < np; i += blockDim.x) { s[i] = 0.0; ix[i] = -1; }
  __syncthreads();
  sort_pairs(s, ix, np);
  return n;
}

// Re-evaluate rows ix[0..m) against query row q with the canonical float64 routine, sort by
// (score desc, row asc) and write the best k.  ix holds LOCAL corpus rows.
// `stage` (null: read the rows in place): kRescoreDepth row slots of `stage_row_bytes` per warp, in shared memory.
// The candidate rows are scattered over the shard (DRAM, not L2), and a warp that reads one row at a time spends
// a DRAM round trip per row (ncu: 55 % of select_rescore's samples sat on the first use of the loaded row).  With
// slots, every warp keeps kRescoreDepth rows in flight as cp.async copies and evaluates from shared memory: the
// gather runs at DRAM bandwidth instead of DRAM latency.  Same routine, same bits either way.
constexpr int kRescoreDepth = 3;
__device__ void rescore_and_emit(double* s, int64_t* ix, int m, const void* qrow, int q_dt,
                                 const void* corpus, int c_dt, int64_t c_stride, int64_t D, int k,
                                 int64_t idx_base, float* out_score, double* out_score64,
                                 int64_t* out_idx, unsigned char* stage = nullptr, int stage_row_bytes = 0,
                                 double* qd = nullptr, int pair_rank_max = 64) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int csz = dtype_size(c_dt);
  const int64_t rowb = D * csz;
  // qd (shared memory, D doubles, or null): float64 copy of the query row, widened once for all candidate rows
  if (qd) {
    for (int64_t e = threadIdx.x; e < qd_len(D); e += blockDim.x) qd[qd_slot(e)] = e < D ? query_elem_f64(qrow, q_dt, e) : 0.0;
    __syncthreads();
  }
  const double qn = warp_query_norm(qrow, q_dt, D);
  if (stage) {
    unsigned char* mine = stage + (size_t)warp * kRescoreDepth * stage_row_bytes;
    const int rounds = m > warp ? (m - warp + nw - 1) / nw : 0;
    auto issue = [&](int i) {     // row of round i -> slot i % depth (an empty group past the end keeps the counts aligned)
      if (i < rounds) {
        const char* r = (const char*)corpus + (size_t)ix[warp + i * nw] * c_stride * csz;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(mine + (size_t)(i % kRescoreDepth) * stage_row_bytes);
        for (int o = lane * 16; o < rowb; o += 32 * 16)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + o), "l"(r + o) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int i = 0; i < kRescoreDepth; ++i) issue(i);
    for (int i = 0; i < rounds; ++i) {
      asm volatile("cp.async.wait_group %0;" ::"n"(kRescoreDepth - 1) : "memory");
      __syncwarp();
      const double v = warp_exact_cosine(qrow, q_dt, qd, qn, mine + (size_t)(i % kRescoreDepth) * stage_row_bytes, c_dt, D);
      if (lane == 0) s[warp + i * nw] = v == v ? v : -INFINITY;   // (NaN would break the sort's total order)
      __syncwarp();           // every lane is done with the slot before it is refilled
      issue(i + kRescoreDepth);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else {
    for (int j = warp; j < m; j += nw) {
      const char* crow = (const char*)corpus + (size_t)ix[j] * c_stride * csz;
      double v = warp_exact_cosine(qrow, q_dt, qd, qn, crow, c_dt, D);
      if (lane == 0) s[j] = v == v ? v : -INFINITY;
    }
  }
  int np = next_pow2(max(m, 1));
  __syncthreads();
  for (int i = m + threadIdx.x; i < np; i += blockDim.x) { s[i] = 0.0; ix[i] = -1; }
  __syncthreads();
  sort_pairs(s, ix, np, pair_rank_max);
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    bool ok = j < m;
    out_score[j] = ok ? (float)s[j] : -INFINITY;
    if (out_score64) out_score64[j] = ok ? s[j] : -INFINITY;
    out_idx[j] = ok ? idx_base + ix[j] : -1;
  }
}

struct SelArgs {
  const void* q; int q_dt; int64_t q_stride;
  const void* corpus; int c_dt; int64_t c_stride;
  int64_t Q, N, D; int k; int64_t idx_base;
  int KP; int64_t NC; float eps;
  const uint64_t* cand; const uint32_t* thr;
  int32_t* flag_cnt; int32_t* flag_list;
  float* out_score; double* out_score64; int64_t* out_idx; int32_t* out_flags;
  // retry stage (SelRetry, tsim_common.cuh): r_in_list != null -> this launch IS the retry pass
  const int32_t* r_in_cnt; const int32_t* r_in_list; int r_cap; int r_skip; int r_forward;
  void* r_q; int64_t r_q_stride;
  // append mode: the main pass's survivors, app_keys[q][0 .. min(app_cnt[q], app_cap)) (null: lists only)
  const uint64_t* app_keys; const uint32_t* app_cnt; int app_cap;
  int stage_row_bytes;   // > 0: re-score through shared-memory row slots of this many bytes (rescore_and_emit)
  int qd_offset;         // ... and keep a float64 copy of the query row at this byte offset of the dynamic window
  int pair_rank_max;     // final (score, row) sort: rank sort up to this many pairs, bitonic network beyond
};

__global__ void __launch_bounds__(kSelThreads, 4) select_rescore_kernel(SelArgs a) {
  // dynamic shared memory: [keys kKeyCap | sel_out kSelOut] while the candidates are selected, then (the keys are
  // dead by then) the row slots of the re-score, a.stage_row_bytes each (0: rows are read in place)
  extern __shared__ __align__(16) unsigned char sel_dyn[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(sel_dyn);
  uint64_t* sel_out = keys + kKeyCap;
  __shared__ uint32_t sel_hist[256];
  __shared__ int sel_sh[4];
  __shared__ double es[128];
  __shared__ int64_t ei[128];
  __shared__ int n_sh;
  __shared__ double red[kSelThreads / 32];
  const int tid = threadIdx.x;
  pdl_trigger();
  pdl_wait();
  const int64_t slot = blockIdx.x;   // position of this query's lists and threshold in the workspace
  int64_t q = slot;                  // query of the call
  if (a.r_in_list) {
    // retry pass: compact slot b <-> query in_list[b]; the overflow goes straight to the float64 scan
    const int n_in = *a.r_in_cnt;
    if (slot == 0 && a.r_forward)
      for (int i = a.r_cap + tid; i < n_in; i += blockDim.x) a.flag_list[atomicAdd(a.flag_cnt, 1)] = a.r_in_list[i];
    if (a.r_skip + slot >= n_in || a.r_skip + slot >= a.r_cap) return;
    q = a.r_in_list[a.r_skip + slot];
  }
  const uint32_t thr_ord = a.thr[slot];  // 0: no unit list ever filled -> nothing was dropped
  if (tid == 0) n_sh = 0;
  __syncthreads();

  // ---- gather surviving keys; sort; keep the best KP --------------------------------------
  // two sources: the per-unit lists of the candidate passes, and (append mode) the main pass's global list
  uint32_t app_n = 0;
  bool app_overflow = false;
  if (a.app_keys) {
    app_n = a.app_cnt[slot];
    app_overflow = app_n > (uint32_t)a.app_cap;      // rows were dropped unseen: the proof cannot hold
    if (app_overflow) app_n = (uint32_t)a.app_cap;
  }
  int n = 0;
  bool truncated = false;   // (block-uniform) this kernel itself dropped keys: more than KP candidates reached it
  for (int part = 0; part < 2; ++part) {
    const uint64_t* src = part == 0 ? a.cand + (size_t)slot * a.NC * a.KP : a.app_keys + (size_t)slot * a.app_cap;
    const int64_t total = part == 0 ? a.NC * a.KP : (int64_t)app_n;
    int64_t cursor = 0;
    while (cursor < total) {
      int64_t take = min((int64_t)(kKeyCap - n), total - cursor);
      for (int64_t i0 = tid; i0 < take; i0 += (int64_t)blockDim.x * 8) {   // eight independent loads in flight
        uint64_t key[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int64_t i = i0 + (int64_t)u * blockDim.x;
          key[u] = i < take ? __ldcg(src + cursor + i) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (key[u] != 0 && (uint32_t)(key[u] >> 32) >= thr_ord) keys[atomicAdd(&n_sh, 1)] = key[u];
      }
      cursor += take;
      __syncthreads();
      n = n_sh;
      if ((cursor < total || part == 0) && n > kKeyCap / 2) {
        truncated = true;             // (kKeyCap / 2 > KP: something goes)
        const int kept = keep_top_scores(keys, n, a.KP, sel_out, sel_hist, sel_sh);
        if (kept == n) {              // massive ties: sort and truncate
          int np = next_pow2(n);
          for (int i = n + tid; i < np; i += blockDim.x) keys[i] = 0;
          __syncthreads();
          sort_keys_desc(keys, np);
          n = min(n, a.KP);
        } else {
          n = kept;
        }
        if (tid == 0) n_sh = n;
        __syncthreads();
      }
    }
  }
  {
    // only the best KP matter: radix-select them (ties at the cut included), then sort those few -- a bitonic
    // sort of every survivor was 3/4 of this kernel's instructions at Q = 1 (148 lists, 69 us; ncu)
    truncated = truncated || n > a.KP;
    n = keep_top_scores(keys, n, a.KP, sel_out, sel_hist, sel_sh);
    int np = next_pow2(max(n, 1));
    for (int i = n + tid; i < np; i += blockDim.x) keys[i] = 0;
    __syncthreads();
    sort_keys_desc(keys, np);
  }
  const int m = min(n, a.KP);

  // ---- ||q|| (float64) for the safety margin -----------------------------------------------
  const char* qrow = (const char*)a.q + (size_t)q * a.q_stride * dtype_size(a.q_dt);
  double qq = 0.0;
  for (int64_t d = tid; d < a.D; d += blockDim.x) {
    double v = (double)load_elem(qrow, a.q_dt, d);
    qq = fma(v, v, qq);
  }
  qq = warp_sum_f64(qq);
  if ((tid & 31) == 0) red[tid >> 5] = qq;
  __syncthreads();
  qq = 0.0;
  for (int i = 0; i < kSelThreads / 32; ++i) qq += red[i];
  const double qn = fmax(sqrt(qq), kCosEps);

  // ---- completeness proof -------------------------------------------------------------------
  // Every row that is NOT among the candidates has approx <= a_KP (the KP-th best candidate):
  // it was either evicted from a full unit list or rejected by a threshold that was the minimum
  // of a full list.  With |approx - exact * ||q||| <= eps * ||q||, a gap a_k - a_KP > 2 eps ||q||
  // proves such a row (and any candidate ranked beyond KP) is below the exact k-th best.
  // The proof is owed whenever ANY row was left out on its approximate score: by a threshold of the candidate passes
  // (thr_ord != 0) or by the truncation to KP right here (`truncated`: lists that never filled -- a shard of a few
  // tiles per worker, KP = 112 -- hand over every row, and near-duplicate clusters put hundreds of them within
  // 2 eps of the k-th best; scripts/fuzz_parity.py found those answers wrong and unflagged).
  bool flagged = false;
  if (thr_ord != 0 || truncated) {
    int kk = min(a.k, m);
    float a_k = key_score(keys[kk - 1]);
    float a_kp = key_score(keys[m - 1]);
    flagged = !((double)a_k - (double)a_kp > 2.0 * (double)a.eps * qn) || (m < a.KP);
  }
  flagged = flagged || app_overflow;

  for (int j = tid; j < m; j += blockDim.x) { ei[j] = (int64_t)key_idx(keys[j]); es[j] = 0.0; }
  __syncthreads();          // keys / sel_out are dead from here on: their memory becomes the row slots
  rescore_and_emit(es, ei, m, qrow, a.q_dt, a.corpus, a.c_dt, a.c_stride, a.D, a.k, a.idx_base,
                   a.out_score + q * a.k, a.out_score64 ? a.out_score64 + q * a.k : nullptr,
                   a.out_idx + q * a.k, a.stage_row_bytes ? sel_dyn : nullptr, a.stage_row_bytes,
                   a.stage_row_bytes ? reinterpret_cast<double*>(sel_dyn + a.qd_offset) : nullptr, a.pair_rank_max);
  // out_flags: 0 answered by the first tensor pass, 2 by the wide retry pass, 1 by the float64 scan
  // (a flagged query keeps / gets 1 here; whichever later stage answers it overwrites that)
  if (tid == 0) {
    if (a.out_flags) a.out_flags[q] = flagged ? 1 : (a.r_in_list ? 2 : 0);
    if (flagged) n_sh = atomicAdd(a.flag_cnt, 1);
  }
  if (!flagged) return;               // block-uniform
  __syncthreads();
  const int pos = n_sh;
  if (tid == 0) a.flag_list[pos] = (int32_t)q;
  if (a.r_q && !a.r_in_list && pos < a.r_cap) {
    // first pass with a retry stage behind it: leave the query row where the wide pass reads it
    const int esz = dtype_size(a.q_dt);
    const uint4* s4 = (const uint4*)qrow;   // tensor-path rows: 16-byte aligned, D * esz % 16 == 0
    uint4* d4 = (uint4*)((char*)a.r_q + (size_t)pos * a.r_q_stride * esz);
    for (int64_t i = tid; i < a.D * esz / 16; i += blockDim.x) d4[i] = s4[i];
  }
}

// Between the bootstrap launch and the main launch: the KP-th best key of the union of a query's
// bootstrap lists (slots 0..Gq-1) is a lower bound of its KP-th best over the whole corpus.
__global__ void __launch_bounds__(kSelThreads) tighten_kernel(const uint64_t* cand, uint32_t* thr, int64_t NC, int KP,
                                                              int nslots, uint32_t* ladder) {
  __shared__ uint64_t keys[kKeyCap];
  __shared__ uint64_t sel_out[kSelOut];
  __shared__ uint32_t sel_hist[256];
  __shared__ int sel_sh[4];
  __shared__ int n_sh;
  pdl_trigger();
  pdl_wait();
  const int64_t q = blockIdx.x;
  const int tid = threadIdx.x;
  const uint64_t* src = cand + (size_t)q * NC * KP;
  const int64_t total = (int64_t)nslots * KP;
  const uint32_t thr_ord = thr[q];
  // A full list's minimum is a lower bound of the union's KP-th best: only keys at or above the best
  // such bound can matter (at least KP of them exist), which leaves a few dozen keys to sort instead of
  // nslots * KP.  Thread per key, eight independent loads in flight; per-list minimum and fill count
  // through shared-memory atomics.
  constexpr int kMaxSlots = 1024;
  __shared__ unsigned long long slot_min[kMaxSlots];
  __shared__ int slot_cnt[kMaxSlots];
  __shared__ unsigned long long bound_sh;
  const bool use_bound = nslots <= kMaxSlots;
  if (tid == 0) { n_sh = 0; bound_sh = 0ull; }
  if (use_bound) {
    for (int sl = tid; sl < nslots; sl += blockDim.x) { slot_min[sl] = ~0ull; slot_cnt[sl] = 0; }
    __syncthreads();
    for (int64_t i0 = tid; i0 < total; i0 += (int64_t)blockDim.x * 8) {
      uint64_t key[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t i = i0 + (int64_t)u * blockDim.x;
        key[u] = i < total ? __ldcg(src + i) : 0ull;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (key[u] != 0) {
          const int sl = (int)((i0 + (int64_t)u * blockDim.x) / KP);
          atomicMin(&slot_min[sl], (unsigned long long)key[u]);
          atomicAdd(&slot_cnt[sl], 1);
        }
      }
    }
    __syncthreads();
    for (int sl = tid; sl < nslots; sl += blockDim.x)
      if (slot_cnt[sl] == KP) atomicMax(&bound_sh, slot_min[sl]);
  }
  __syncthreads();
  const uint64_t bound = bound_sh;
  int64_t cursor = 0;
  int n = 0;
  while (cursor < total) {
    int64_t take = min((int64_t)(kKeyCap - n), total - cursor);
    for (int64_t i0 = tid; i0 < take; i0 += (int64_t)blockDim.x * 4) {
      uint64_t key[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t i = i0 + (int64_t)u * blockDim.x;
        key[u] = i < take ? __ldcg(src + cursor + i) : 0ull;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (key[u] != 0 && key[u] >= bound && (uint32_t)(key[u] >> 32) >= thr_ord) keys[atomicAdd(&n_sh, 1)] = key[u];
    }
    cursor += take;
    __syncthreads();
    n = n_sh;
    if (cursor < total && n > kKeyCap / 2) {
      const int m = keep_top_scores(keys, n, KP, sel_out, sel_hist, sel_sh);
      if (m == n) {                 // massive ties: sort and truncate
        int np = next_pow2(n);
        for (int i = n + tid; i < np; i += blockDim.x) keys[i] = 0;
        __syncthreads();
        sort_keys_desc(keys, np);
        n = min(n, KP);
      } else {
        n = m;
      }
      if (tid == 0) n_sh = n;
      __syncthreads();
    }
  }
  n = keep_top_scores(keys, n, KP, sel_out, sel_hist, sel_sh);
  int np = next_pow2(max(n, 1));
  for (int i = n + tid; i < np; i += blockDim.x) keys[i] = 0;
  __syncthreads();
  sort_keys_desc(keys, np);
  if (tid == 0 && n >= KP) atomicMax(thr + q, (uint32_t)(keys[KP - 1] >> 32));
  if (!ladder) return;
  // Threshold ladder for the following launches: 16 evenly spaced levels from the sample's KP-th best
  // (level 0 = the threshold just set) through its best (level 8) to as far again above it -- for
  // Gaussian-like tails even steps in score are roughly geometric steps in rank, and the full corpus's
  // KP-th best usually ends up near the sample's best.  ANY ascending levels are valid; they only decide
  // how closely the threshold can follow the data (search_tc.cu, struct Ladder).  Counts start from the
  // sample's own top-KP rows.
  __shared__ uint32_t cnt[kLadder];
  uint32_t* lad = ladder + (size_t)q * 2 * kLadder;
  if (tid < kLadder) cnt[tid] = 0;
  __syncthreads();
  float base = INFINITY, step = 0.f, inv = 0.f;      // n < KP: no threshold was set -> a ladder that never fires
  if (n >= KP) {
    base = key_score(keys[KP - 1]);
    step = (key_score(keys[0]) - base) * (1.f / 8.f);
    if (!(step > 0.f) || !(step < INFINITY)) step = 0.f;
    inv = step > 0.f ? 1.f / step : 0.f;
    for (int i = tid; i < KP; i += blockDim.x) {
      const int j = ladder_level(base, step, inv, key_score(keys[i]));
      if (j >= 0) atomicAdd(&cnt[j], 1u);
    }
  }
  __syncthreads();
  if (tid < kLadder) {
    lad[kLadder + tid] = cnt[tid];
    lad[tid] = tid == 0 ? __float_as_uint(base) : tid == 1 ? __float_as_uint(step) : tid == 2 ? __float_as_uint(inv) : 0u;
  }
}

// The same job on an append plan's global lists, one WARP per query (eight queries per block, no block barrier):
// a warp-private MSB-first radix select (4 passes x 8 bits over the upper 32 key bits, 256-bin histogram in shared
// memory) finds the KP-th best approximate score among the min(app_cnt, cap) keys appended so far -- a lower bound
// of the query's final KP-th best, since every appended row reaches select_rescore -- raises thr[q] to it and lays
// out the threshold ladder exactly as tighten_kernel does (counts = the appended rows at or above each level).
// Q = 4096, ~1000 keys per query: ~10 us against 77 us for the block-per-query kernel above.
__global__ void __launch_bounds__(kSelThreads) tighten_app_kernel(const uint64_t* app_keys, const uint32_t* app_cnt,
                                                                  int cap, int64_t Q, int KP, uint32_t* thr,
                                                                  uint32_t* ladder) {
  __shared__ uint32_t hist_all[kSelThreads / 32][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * (kSelThreads / 32) + warp;
  pdl_trigger();
  pdl_wait();
  if (q >= Q) return;
  const uint32_t n = min(app_cnt[q], (uint32_t)cap);
  warp_tighten<false>(app_keys + (size_t)q * cap, n, KP, thr + q, ladder ? ladder + (size_t)q * 2 * kLadder : nullptr,
                      hist_all[warp]);
}

// First kernel of a search (replaces a memset of the control words, a memset of the query padding and a 2-D copy).
__global__ void __launch_bounds__(256) search_prep_kernel(uint4* zero_base, size_t zero_vecs, const char* q_src,
                                                          size_t q_src_stride_bytes, char* q_dst, size_t row_vecs,
                                                          int64_t Q, int64_t q_rows_padded, int q_span) {
  pdl_trigger();
  pdl_wait();
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (size_t i = tid; i < zero_vecs; i += nthr) zero_base[i] = z;
  if (!q_dst) return;
  const size_t total = (size_t)q_rows_padded * row_vecs;
  for (size_t i = tid; i < total; i += nthr) {
    const size_t r = i / row_vecs, c = i - r * row_vecs;
    const size_t src_row = q_span > 0 ? r % (size_t)q_span : r;      // replicated query block (qrep plans)
    reinterpret_cast<uint4*>(q_dst)[i] = (int64_t)src_row < Q ? reinterpret_cast<const uint4*>(q_src + src_row * q_src_stride_bytes)[c] : z;
  }
}

struct ExMergeArgs {
  const void* q; int q_dt; int64_t q_stride;
  const void* corpus; int c_dt; int64_t c_stride;
  int64_t Q, D; int k; int64_t idx_base; int S;
  const int32_t* flag_cnt; const int32_t* flag_list;  // null: every query, slot == query
  const double* ex_score; const uint32_t* ex_idx;     // [slot][S][k]
  float* out_score; double* out_score64; int64_t* out_idx; int32_t* out_flags;
};

__global__ void __launch_bounds__(kSelThreads, 3) merge_exact_lists_kernel(ExMergeArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s = (double*)smem_raw;
  int64_t* ix = (int64_t*)(smem_raw + sizeof(double) * kPairCap);
  __shared__ int n_sh;
  pdl_trigger();
  pdl_wait();
  const int64_t slot = blockIdx.x;
  int64_t q = slot;
  if (a.flag_cnt) {
    if (slot >= *a.flag_cnt) return;
    q = a.flag_list[slot];
  }
  const double* ss = a.ex_score + (size_t)slot * a.S * a.k;
  const uint32_t* si = a.ex_idx + (size_t)slot * a.S * a.k;
  int n = stream_top_pairs(s, ix, &n_sh, (int64_t)a.S * a.k, a.k,
                           [&](int64_t i, double* v, int64_t* id) {
                             uint32_t r = si[i];
                             *v = ss[i];
                             *id = (r == 0xffffffffu) ? -1 : (int64_t)r;
                           });
  const int m = min(n, a.k);
  __syncthreads();
  const char* qrow = (const char*)a.q + (size_t)q * a.q_stride * dtype_size(a.q_dt);
  rescore_and_emit(s, ix, m, qrow, a.q_dt, a.corpus, a.c_dt, a.c_stride, a.D, a.k, a.idx_base,
                   a.out_score + q * a.k, a.out_score64 ? a.out_score64 + q * a.k : nullptr,
                   a.out_idx + q * a.k);
  if (threadIdx.x == 0 && a.out_flags) a.out_flags[q] = 1;
}

// Lists that are already ordered best-first (what tsim_search_topk emits, padding last) are merged by
// RANK: the final position of an element is the number of elements of every list that precede it, found
// by one binary search per list -- no barriers, ~n_lists * log2(k_in) shared-memory probes per element
// (8 lists of 100: 332 -> ~60 us at Q = 4096 against the 55-stage bitonic sort).  Equal (score, index)
// pairs in two lists (overlapping shards; not expected) are ordered by list number so ranks stay unique.
// Anything else (unsorted input) takes the bitonic sort.
// Input addressing: entry j of list l of query q sits at [l * list_stride + q * query_stride + j] of `sc` / `ix_in`
// (query-major rows: list_stride = k_in, query_stride = n_lists * k_in; the rank-major receive buffer of an
// all-gather is read in place with list_stride = one rank's message).
__global__ void __launch_bounds__(kSelThreads) merge_topk_kernel(const double* sc, const int64_t* ix_in,
                                                                int64_t total, int k_in, int k_out,
                                                                int64_t list_stride, int64_t query_stride,
                                                                float* out_score, double* out_score64,
                                                                int64_t* out_idx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int np = next_pow2((int)max(total, (int64_t)1));
  double* s = (double*)smem_raw;
  int64_t* ix = (int64_t*)(smem_raw + sizeof(double) * np);
  const int64_t q = blockIdx.x;
  for (int i = threadIdx.x; i < np; i += blockDim.x) {
    bool ok = i < total;
    const int64_t at = ok ? (int64_t)(i / k_in) * list_stride + q * query_stride + (i % k_in) : 0;
    s[i] = ok ? sc[at] : 0.0;
    ix[i] = ok ? ix_in[at] : -1;
  }
  __syncthreads();
  int unsorted = 0, valid = 0;
  for (int i = threadIdx.x; i < (int)total; i += blockDim.x) {
    if (ix[i] >= 0) ++valid;
    if ((i + 1) % k_in != 0 && pair_before(s[i + 1], ix[i + 1], s[i], ix[i])) unsorted = 1;
  }
  unsorted = __syncthreads_or(unsorted);
  if (!unsorted) {
    const int n_lists = (int)(total / k_in);
    float* os = out_score + q * k_out;
    double* os64 = out_score64 ? out_score64 + q * k_out : nullptr;
    int64_t* oi = out_idx + q * k_out;
    __shared__ int n_valid_sh;
    if (threadIdx.x == 0) n_valid_sh = 0;
    __syncthreads();
    if (valid) atomicAdd(&n_valid_sh, valid);
    // Any list that holds >= k_out entries bounds the answer from below by its k_out-th score: an element strictly
    // below the best such bound has >= k_out elements ahead of it and needs no rank (8 lists of 100: ~110 of the 800
    // elements are left to rank; 340 -> ~120 us at Q = 4096).
    __shared__ unsigned long long bound_sh;
    if (threadIdx.x == 0) bound_sh = 0ull;
    __syncthreads();
    if (k_in >= k_out)
      for (int l = threadIdx.x; l < n_lists; l += blockDim.x)
        if (ix[l * k_in + k_out - 1] >= 0) atomicMax(&bound_sh, (unsigned long long)f64_to_ord(s[l * k_in + k_out - 1]));
    __syncthreads();
    const uint64_t bound = bound_sh;
    for (int e = threadIdx.x; e < (int)total; e += blockDim.x) {
      const double se = s[e];
      const int64_t ie = ix[e];
      if (ie < 0 || f64_to_ord(se) < bound) continue;
      const int le = e / k_in;
      int rank = 0;
      for (int l = 0; l < n_lists && rank < k_out; ++l) {
        if (l == le) { rank += e - le * k_in; continue; }
        // number of elements of list l that come before e (ties: the lower list number first)
        const double* sl = s + l * k_in;
        const int64_t* il = ix + l * k_in;
        int lo = 0, hi = k_in;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          const bool before = l < le ? !pair_before(se, ie, sl[mid], il[mid]) : pair_before(sl[mid], il[mid], se, ie);
          if (before) lo = mid + 1; else hi = mid;
        }
        rank += lo;
      }
      if (rank < k_out) {
        os[rank] = (float)se;
        if (os64) os64[rank] = se;
        oi[rank] = ie;
      }
    }
    __syncthreads();
    for (int j = n_valid_sh + threadIdx.x; j < k_out; j += blockDim.x) {
      os[j] = -INFINITY;
      if (os64) os64[j] = -INFINITY;
      oi[j] = -1;
    }
    return;
  }
  sort_pairs(s, ix, np);
  for (int j = threadIdx.x; j < k_out; j += blockDim.x) {
    bool ok = j < total && ix[j] >= 0;
    out_score[q * k_out + j] = ok ? (float)s[j] : -INFINITY;
    if (out_score64) out_score64[q * k_out + j] = ok ? s[j] : -INFINITY;
    out_idx[q * k_out + j] = ok ? ix[j] : -1;
  }
}

}  // namespace

int launch_select_rescore(const void* q, int q_dt, int64_t q_stride, const void* corpus, int c_dt,
                          int64_t c_stride, int64_t Q, int64_t N, int64_t D, int k,
                          int64_t idx_base, const SearchPlan& p, const uint64_t* cand,
                          const uint32_t* thr, int32_t* flag_cnt, int32_t* flag_list,
                          float* out_score, double* out_score64, int64_t* out_idx,
                          int32_t* out_flags, cudaStream_t st, const SelRetry* retry,
                          const uint64_t* app_keys, const uint32_t* app_cnt) {
  SelArgs a;
  a.q = q; a.q_dt = q_dt; a.q_stride = q_stride;
  a.corpus = corpus; a.c_dt = c_dt; a.c_stride = c_stride;
  a.Q = Q; a.N = N; a.D = D; a.k = k; a.idx_base = idx_base;
  a.KP = p.KP; a.NC = p.NC; a.eps = p.eps; a.cand = cand; a.thr = thr;
  a.flag_cnt = flag_cnt; a.flag_list = flag_list;
  a.out_score = out_score; a.out_score64 = out_score64; a.out_idx = out_idx; a.out_flags = out_flags;
  a.r_in_cnt = retry ? retry->in_cnt : nullptr; a.r_in_list = retry ? retry->in_list : nullptr;
  a.r_cap = retry ? retry->cap : 0; a.r_q = retry ? retry->r_q : nullptr; a.r_q_stride = retry ? retry->r_q_stride : 0;
  a.r_skip = retry ? retry->skip : 0; a.r_forward = retry ? retry->forward : 0;
  a.app_keys = app_keys; a.app_cnt = app_cnt; a.app_cap = p.app_cap;
  // re-score through shared-memory row slots when a row is at most 2 KB and 16-byte granular (cp.async)
  const int64_t rowb = D * dtype_size(c_dt);
  const bool stage_ok = rowb <= 2048 && rowb % 16 == 0 && ((uintptr_t)corpus & 15) == 0 &&
                        (c_stride * dtype_size(c_dt)) % 16 == 0 && !knob_on("TSIM_NO_RESCORE_STAGE");
  a.stage_row_bytes = stage_ok ? (int)rowb : 0;
  a.pair_rank_max = knob_int("TSIM_PAIR_RANK_MAX", 64);   // experiment knob
  size_t smem = (size_t)(kKeyCap + kSelOut) * sizeof(uint64_t);
  const size_t stage_bytes = (size_t)(kSelThreads / 32) * kRescoreDepth * a.stage_row_bytes;
  a.qd_offset = (int)stage_bytes;
  if (stage_ok && stage_bytes + (size_t)qd_len(D) * sizeof(double) > smem) smem = stage_bytes + (size_t)qd_len(D) * sizeof(double);
  // static + dynamic shared memory beyond 48 KB needs the opt-in (the kernel also has ~4 KB of static arrays)
  if (smem > 40 * 1024)
    TSIM_CUDA(cudaFuncSetAttribute(select_rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // a retry pass launches one block per compact slot: Q = the slot capacity there
  TSIM_CUDA(launch_pdl(select_rescore_kernel, dim3((unsigned)Q), dim3(kSelThreads), smem, st, a));
  count_launch();
  return TSIM_OK;
}

int launch_tighten(int64_t Q, const SearchPlan& p, int nslots, const uint64_t* cand, uint32_t* thr,
                   uint32_t* ladder, cudaStream_t st) {
  TSIM_CUDA(launch_pdl(tighten_kernel, dim3((unsigned)Q), dim3(kSelThreads), 0, st, cand, thr, p.NC, p.KP, nslots, ladder));
  count_launch();
  return TSIM_OK;
}

int launch_search_prep(void* zero_base, size_t zero_bytes, const void* q_src, size_t q_src_stride_bytes, void* q_dst,
                       size_t row_bytes, int64_t Q, int64_t q_rows_padded, int q_span, cudaStream_t st) {
  // zero_base is 256-byte aligned and zero_bytes a multiple of 256 (workspace layout); rows are 16-byte multiples
  const size_t work = zero_bytes / 16 + (q_dst ? (size_t)q_rows_padded * (row_bytes / 16) : 0);
  size_t blocks = (work + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 1184) blocks = 1184;
  TSIM_CUDA(launch_pdl(search_prep_kernel, dim3((unsigned)blocks), dim3(256), 0, st, (uint4*)zero_base, zero_bytes / 16,
                       (const char*)q_src, q_src_stride_bytes, (char*)q_dst, row_bytes / 16, Q, q_rows_padded, q_span));
  count_launch();
  return TSIM_OK;
}

int launch_tighten_app(int64_t Q, const SearchPlan& p, const uint64_t* app_keys, const uint32_t* app_cnt,
                       uint32_t* thr, uint32_t* ladder, cudaStream_t st) {
  const int per = kSelThreads / 32;
  TSIM_CUDA(launch_pdl(tighten_app_kernel, dim3((unsigned)((Q + per - 1) / per)), dim3(kSelThreads), 0, st, app_keys, app_cnt,
                       p.app_cap, Q, p.KP, thr, ladder));
  count_launch();
  return TSIM_OK;
}

int launch_merge_exact_lists(const void* q, int q_dt, int64_t q_stride, const void* corpus,
                             int c_dt, int64_t c_stride, int64_t Q, int64_t D, int k,
                             int64_t idx_base, const SearchPlan& p, const int32_t* flag_cnt,
                             const int32_t* flag_list, const double* ex_score,
                             const uint32_t* ex_idx, float* out_score, double* out_score64,
                             int64_t* out_idx, int32_t* out_flags, cudaStream_t st) {
  ExMergeArgs a;
  a.q = q; a.q_dt = q_dt; a.q_stride = q_stride;
  a.corpus = corpus; a.c_dt = c_dt; a.c_stride = c_stride;
  a.Q = Q; a.D = D; a.k = k; a.idx_base = idx_base; a.S = p.S;
  a.flag_cnt = flag_cnt; a.flag_list = flag_list; a.ex_score = ex_score; a.ex_idx = ex_idx;
  a.out_score = out_score; a.out_score64 = out_score64; a.out_idx = out_idx; a.out_flags = out_flags;
  size_t smem = (size_t)kPairCap * (sizeof(double) + sizeof(int64_t));
  TSIM_CUDA(launch_pdl(merge_exact_lists_kernel, dim3((unsigned)Q), dim3(kSelThreads), smem, st, a));
  count_launch();
  return TSIM_OK;
}

int launch_merge_topk(const double* sc, const int64_t* ix, int64_t Q, int64_t n_lists, int k_in,
                      int k_out, int64_t list_stride, int64_t query_stride, float* out_score, double* out_score64,
                      int64_t* out_idx, cudaStream_t st) {
  int64_t total = n_lists * (int64_t)k_in;
  int np = 1;
  while (np < total) np <<= 1;
  size_t smem = (size_t)np * (sizeof(double) + sizeof(int64_t));
  if (smem > 48 * 1024)
    TSIM_CUDA(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  merge_topk_kernel<<<(unsigned)Q, kSelThreads, smem, st>>>(sc, ix, total, k_in, k_out, list_stride, query_stride,
                                                            out_score, out_score64, out_idx);
  TSIM_CUDA(cudaGetLastError());
  count_launch();
  return TSIM_OK;
}

}  // namespace tsim
