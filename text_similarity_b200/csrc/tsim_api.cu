// C-ABI entry points of libtsim.so (see include/tsim.h), launch planning and error reporting.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "tsim_common.cuh"

namespace tsim {

static thread_local char g_err[512] = "";
static thread_local cudaEvent_t g_ev_start = nullptr, g_ev_stop = nullptr;
static unsigned long long g_launches = 0, g_encodes = 0, g_env_reads = 0, g_plans = 0;
void count_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }
void count_map_encode() { __atomic_add_fetch(&g_encodes, 1ull, __ATOMIC_RELAXED); }

#ifdef TSIM_EXPERIMENT
int knob_int(const char* name, int dflt) {
  __atomic_add_fetch(&g_env_reads, 1ull, __ATOMIC_RELAXED);
  const char* v = getenv(name);
  return (v && v[0]) ? atoi(v) : dflt;
}
#endif

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int device_sm_count() {
  // per-device cache; 148 (B200) when no device is visible so that planning still works on a
  // CPU-only build box
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 148; }
  if (dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = 148;
    }
    cached[dev] = n;
  }
  return cached[dev];
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static bool tensor_shape_ok(int64_t Q, int64_t N, int64_t D, int k, int q_dt, int c_dt) {
  const bool bf16 = q_dt == TSIM_BF16 && c_dt == TSIM_BF16 && D % 8 == 0 && D >= 8;
  const bool e4m3 = q_dt == TSIM_E4M3 && c_dt == TSIM_E4M3 && D % 16 == 0 && D >= 16;
  return (bf16 || e4m3) && k <= 100 && Q > 0 && N > 0 && N < (int64_t)0x7fffff00 && Q < (int64_t)0x7fffff00;
}

int make_search_plan(int64_t Q, int64_t N, int64_t D, int k, int q_dt, int c_dt, int mode,
                     bool need_invnorm, int shadow_kind, SearchPlan* p) {
  const bool shadow = shadow_kind == 1;      // the rounded shadow: wide margin, widest lists, k <= 24
  const bool split = shadow_kind == 2;       // the split shadow: D here is 3 x the segment width, ordinary lists
  memset(p, 0, sizeof(*p));
  p->ksplit = split ? (int)(D / 3 / 64) : 0;
  __atomic_add_fetch(&g_plans, 1ull, __ATOMIC_RELAXED);
  const int sms = device_sm_count();
  // a shadow pass needs the widest lists: the candidates must reach 2 * kShadowEps below the k-th best
  const bool ok = tensor_shape_ok(Q, N, D, k, q_dt, c_dt) && (!shadow || k <= 24);
  p->eps = shadow ? shadow_eps(D) : split ? split_shadow_eps(D / 3) : approx_eps(D, dtype_size(c_dt));
  if (mode == TSIM_MODE_TENSOR && !ok) {
    set_error("search: TSIM_MODE_TENSOR needs bf16 (D %% 8 == 0) or e4m3 (D %% 16 == 0) queries AND corpus, k <= 100 "
              "(got q_dt=%d c_dt=%d D=%lld k=%d)", q_dt, c_dt, (long long)D, k);
    return TSIM_ERR_UNSUPPORTED;
  }
  p->use_tensor = (mode != TSIM_MODE_EXACT) && ok;
  size_t off = 0;
  if (p->use_tensor) {
    p->KP = shadow ? 112 : k <= 10 ? 16 : k <= 24 ? 32 : k <= 52 ? 64 : 112;
    // more than one 128-query block: CTA pairs (cta_group::2) share each corpus tile between two SMs
    p->pair = (Q > 128 && !knob_on("TSIM_NO_PAIR")) ? 1 : 0;
    const int qrows = p->pair ? 256 : 128;
    const int workers = p->pair ? sms / 2 : sms;
    p->QB = (int)((Q + qrows - 1) / qrows);
    const int64_t T = (N + 255) / 256;
    if (p->QB <= workers && (workers % p->QB) * 100 <= 3 * workers) {
      // few query blocks: sticky schedule (see search_tc.cu), one candidate list per CTA
      p->sticky = 1;
      p->Gq = workers / p->QB;
      if (p->Gq > T) p->Gq = (int)T;
      if (p->Gq < 1) p->Gq = 1;
      // Bootstrap: with >= 2 queries the per-CTA lists' warm-up (every early score is a candidate)
      // costs more than two extra launches (measured on the fp8 12.5M x 384 shard: Q = 8 1.23 -> 1.02 ms); scan 1 strided sample tile per worker first and turn
      // their union's KP-th best into every query's starting threshold.
      int64_t minq = knob_int("TSIM_BOOT_MINQ", 2);
      if (minq < 1) minq = 2;
      if (Q >= minq && T >= 32 * (int64_t)p->Gq && !knob_on("TSIM_NO_BOOT")) {
        int64_t tpw = knob_int("TSIM_BOOT_TPW", 1);                     // sample tiles per worker
        if (tpw < 1 || tpw > 16) tpw = 1;
        // at most tpw sample tiles per worker (rounding the stride DOWN gave Gq + 1 tiles: one worker scanned two
        // cold tiles while the other 147 waited for it at the fused pass's grid barrier)
        p->boot_stride = (T + tpw * (int64_t)p->Gq - 1) / (tpw * (int64_t)p->Gq);   // >= 2
        p->boot_tiles = (T + p->boot_stride - 1) / p->boot_stride;      // every multiple of the stride below T
        p->boot_slots = p->Gq;
        p->fused = knob_on("TSIM_NO_FUSED") ? 0 : 1;   // sample, thresholds and main in one cooperative launch
      }
    }
    // Small batches on a large shard (config 4): the swapped-role kernel (search_sw.cu) -- corpus rows on the MMA's M
    // side, the queries resident on the N side -- which under the board's 1 kW power cap is worth 10-20 % of a
    // sustained stream of searches (profiles/README.md, "small batches").  Rows wider than 512 bytes with fewer than 8
    // queries stay on search_tc.cu: its query block is mostly zero rows then (cheap MMAs), and a swapped
    // MMA occupies the tensor pipe for ~146 clocks per 128 rows x 32 bytes whatever its N, which caps that kernel at
    // ~6.6 TB/s on 1536-byte rows where search_tc.cu streams 7.2 TB/s in a burst.
    {
      const int64_t T128 = (N + 127) / 128;
      const int kb = (int)((D * dtype_size(c_dt) + 127) / 128);
      const bool narrow = kb <= 4;
      if (Q <= 32 && (narrow || Q >= 8 || knob_on("TSIM_SWAP_ALL")) && p->KP <= 32 && !shadow && !split && T128 >= 16 * (int64_t)sms &&
          search_sw_stages(kb) > 0 && !knob_on("TSIM_NO_SWAP")) {
        p->swapped = 1;
        p->sticky = 1; p->pair = 0; p->QB = 1; p->Gq = sms; p->fused = 0; p->qrep = 0;
        p->boot_tiles = 0; p->boot_stride = 0; p->boot_slots = 0;
        p->sw_ns = 24;                                   // 24 tiles x 128 rows = 3072 sample rows = 192 lists of 16 per query
        p->sw_stride = (int)(T128 / p->sw_ns);
        p->append = 1;
        p->app_cap = 4096;
      }
    }
    // Query replication (search_tc.cu, TcArgs::qrep) spreads the epilogue of a tile over the four lane quadrants -- what
    // rows of at most 512 bytes need, whose tiles last ~2 us.  On wider rows one warp keeps up, and the replicas cost
    // power: zero rows are cheaper to multiply than copies (10M x 768 bf16 under a sustained stream: Q = 32 2.69 ms
    // with, 2.46 ms without; Q = 8 2.46 vs 2.40 ms).
    if (p->sticky && !p->pair && !p->swapped && !knob_on("TSIM_NO_QREP") &&
        ((D * dtype_size(c_dt) + 127) / 128 <= 4 || knob_on("TSIM_QREP_WIDE")))
      p->qrep = Q <= 32 ? 4 : Q <= 64 ? 2 : 1;
    if (!p->sticky) {
      // Round-robin units (many query blocks).  A unit's list starts empty, so what it filters with is
      // the query's GLOBAL threshold: a strided sample (~1/64 of the tiles) is scanned first and leaves
      // every query a threshold and a threshold ladder (rank counters) that the main launch keeps
      // raising.  With thresholds global, short units (16K rows) win: the units that share a corpus
      // chunk drift less, so the chunk is re-read from DRAM less often.  Small corpora keep one launch
      // and long units.
      const bool boot = T >= 24 * 64 && !knob_on("TSIM_NO_BOOT");
      int64_t nct = (8 * (int64_t)workers + p->QB - 1) / p->QB;  // aim at >= ~8 units per worker
      if (nct < 1) nct = 1;
      int64_t R = (N + nct - 1) / nct;
      R = (R + 255) / 256 * 256;
      if (R < 256) R = 256;
      const int64_t rcap = boot ? 16384 : 65536;
      if (R > rcap) R = rcap;
      if (const int64_t v = knob_int("TSIM_CHUNK_ROWS", 0); v >= 256) R = v / 256 * 256;   // experiment knob: rows per unit
      // bound the candidate buffer (Q * NC * KP * 8 bytes) to ~2 GB
      while ((double)Q * (double)((N + R - 1) / R) * p->KP * 8.0 > 2.0e9 && R < ((int64_t)1 << 30)) R *= 2;
      p->R = R;
      const int64_t tpc = R / 256;
      if (boot && T >= 24 * tpc) {
        // Sample = every boot_stride-th tile, ~1/64 of the corpus.  Its first launch (every
        // mini_mult-th sample tile, about 4 tiles, one-tile units) is the only one that runs with
        // cold lists; the rest of the sample and the main launch start from thresholds + a ladder.
        int64_t div = knob_int("TSIM_BOOT_DIV", 64);                    // sample = 1 / div of the tiles
        if (div < 4 || div > 1024) div = 64;
        int64_t want = T / div;
        if (want < 8) want = 8;
        p->boot_stride = T / want;                                       // >= 2
        p->boot_tiles = (T + p->boot_stride - 1) / p->boot_stride;
        if (!knob_on("TSIM_NO_MINI")) {                                  // experiment knob: one cold sample launch
          p->mini_mult = (p->boot_tiles + 3) / 4;
          if (p->mini_mult < 2) p->mini_mult = 2;
          p->mini_tiles = (p->boot_tiles + p->mini_mult - 1) / p->mini_mult;
          p->mini_slots = p->mini_tiles;                                 // one-tile units
        }
        // the rest of the sample in short units (lists rarely fill, thresholds come from the ladder):
        // the unit length in [8, 48] tiles with the smallest makespan = waves x tiles per unit
        const int64_t rest = p->boot_tiles - p->mini_tiles;
        int64_t best_tpc = 8, best_span = -1;
        for (int64_t t = 8; t <= 48; ++t) {
          const int64_t units = (int64_t)p->QB * ((rest + t - 1) / t);
          const int64_t span = ((units + workers - 1) / workers) * t;
          if (best_span < 0 || span < best_span) { best_span = span; best_tpc = t; }
        }
        p->boot_tpc = best_tpc;
        p->boot_slots = (rest + p->boot_tpc - 1) / p->boot_tpc;
      }
      // Bootstrapped round-robin plans with KP >= 32 keep no candidate lists at all: every pass (mini sample with
      // no threshold, rest of the sample, main) APPENDS the rows that beat the query's threshold to one global list
      // per query (search_tc.cu, AppendList); tighten and select_rescore read that list.
      p->append = (p->boot_tiles && p->KP >= knob_int("TSIM_APPEND_MIN_KP", 32)) ? 1 : 0;
      p->app_cap = Q > 65536 ? 2048 : 4096;
      if (p->append) p->NC = 0;
      else p->NC = p->boot_tiles ? p->mini_slots + p->boot_slots + (T - p->boot_tiles + tpc - 1) / tpc : (T + tpc - 1) / tpc;
    } else if (p->swapped) {
      p->R = 128;
      p->NC = 8 * (int64_t)p->sw_ns;
    } else {
      p->R = 256;
      const int rep = p->qrep > 1 ? p->qrep : 1;          // main-pass lists per worker (sample lists: one)
      p->NC = p->boot_tiles ? p->Gq + (int64_t)rep * p->Gq : (int64_t)rep * p->Gq;
    }
    p->off_cand = off; off = align_up(off + (size_t)Q * p->NC * p->KP * sizeof(uint64_t), 256);
    if (p->append) { p->off_app_keys = off; off = align_up(off + (size_t)Q * p->app_cap * sizeof(uint64_t), 256); }
  }
  // exact scan (whole-call path, or fallback for flagged queries)
  int64_t S = (N + 1023) / 1024;
  if (S < 1) S = 1;
  if (S > 2 * sms) S = 2 * sms;
  while (S > 1 && (double)Q * (double)S * k * 12.0 > 5.0e8) S = (S + 1) / 2;
  if (!p->use_tensor && Q > 32) {
    // whole-call scan of many queries (64 per CTA, search_exact.cu): every slice starts its lists cold, so no more
    // slices than it takes to fill the GPU a few times over (Q = 1024, k = 100: 296 slices cost 133K list
    // insertions per query, 37 slices 24K)
    int64_t cap = (4 * (int64_t)sms + (Q + 63) / 64 - 1) / ((Q + 63) / 64);
    if (const int64_t v = knob_int("TSIM_SCAN_SLICES", 0); v >= 1) cap = v;   // experiment knob
    if (cap < 1) cap = 1;
    if (S > cap) S = cap;
  }
  int64_t sr = (N + S - 1) / S;
  sr = (sr + 31) / 32 * 32;
  if (sr < 32) sr = 32;
  p->slice_rows = sr;
  p->S = (int)((N + sr - 1) / sr);
  if (p->S < 1) p->S = 1;
  p->off_thr = off; off = align_up(off + (size_t)Q * sizeof(uint32_t), 256);
  p->off_flagcnt = off; off += 256;
  // unit-claim areas of the round-robin schedule (search_tc.cu), zeroed by the same memset as thr / flag count
  p->off_sched = off;
  p->sched_area = (p->use_tensor && !p->sticky) ? 256 + (size_t)sms * 32 * sizeof(uint64_t) : 0;
  off += 3 * p->sched_area;
  if (p->append) { p->off_app_cnt = off; off = align_up(off + (size_t)Q * sizeof(uint32_t), 256); }   // zeroed with thr
  if (p->fused) { p->off_gbar = off; off += 256; }                                                      // zeroed with thr
  // retry stage: its thresholds and level-2 flag count sit in the same zeroed span
  p->retry = 0;   // TSIM_NO_RETRY (experiment knob): flagged queries go straight to the float64 scan
  if (p->use_tensor && !shadow && !split && p->KP < kRetryKP && !knob_on("TSIM_NO_RETRY")) {
    const int64_t rounds = (Q + kRetryQ - 1) / kRetryQ;
    p->retry = (int)(rounds < kRetryMaxRounds ? rounds : kRetryMaxRounds);
  }
  if (p->retry) {
    p->off_r_thr = off; off += align_up((size_t)p->retry * kRetryQ * sizeof(uint32_t), 256);
    p->off_r_flagcnt = off; off += 256;
  }
  p->off_flaglist = off; off = align_up(off + (size_t)Q * sizeof(int32_t), 256);
  if (p->use_tensor && (p->boot_tiles || p->swapped)) {
    p->off_ladder = off; off = align_up(off + (size_t)Q * 2 * kLadder * sizeof(uint32_t), 256);
  }
  if (p->use_tensor && (Q % (p->pair ? 256 : 128) != 0 || p->qrep > 1)) {  // zero-padded copy of the queries (TMA OOB fill is slow)
    p->off_qpad = off; off = align_up(off + (size_t)p->QB * (p->pair ? 256 : 128) * D * dtype_size(q_dt), 256);   // (swapped plans use 32 rows of it)
  }
  if (need_invnorm && p->use_tensor) { p->off_invnorm = off; off = align_up(off + (size_t)N * sizeof(float), 256); }
  if (p->retry) {
    const int64_t T = (N + 255) / 256;
    p->r_Gq = (int)(T < sms ? T : sms);
    p->off_r_flaglist = off; off = align_up(off + (size_t)Q * sizeof(int32_t), 256);
    p->off_r_q = off; off = align_up(off + (size_t)p->retry * kRetryQ * D * dtype_size(q_dt), 256);
    p->off_r_cand = off; off = align_up(off + (size_t)kRetryQ * p->r_Gq * kRetryKP * sizeof(uint64_t), 256);
  }
  if (Q > 32 && N > 0) {   // float64 scan on the FP64 tensor cores (whole call, or the flagged-query fallback): row norms once per call
    p->has_ex_rinv = 1;
    p->off_ex_rinv = off; off = align_up(off + 2 * (size_t)N * sizeof(double), 256);   // [N] 1 / norm, then [N] the norms
  }
  p->off_ex_score = off; off = align_up(off + (size_t)Q * p->S * k * sizeof(double), 256);
  p->off_ex_idx = off; off = align_up(off + (size_t)Q * p->S * k * sizeof(uint32_t), 256);
  p->total = off + 256;
  return TSIM_OK;
}

}  // namespace tsim

using namespace tsim;

extern "C" int tsim_version(void) { return TSIM_ABI_VERSION; }
extern "C" const char* tsim_last_error(void) { return g_err; }
extern "C" uint64_t tsim_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
extern "C" void tsim_debug_counters(uint64_t out[4]) {
  out[0] = __atomic_load_n(&g_launches, __ATOMIC_RELAXED);
  out[1] = __atomic_load_n(&g_encodes, __ATOMIC_RELAXED);
  out[2] = __atomic_load_n(&g_env_reads, __ATOMIC_RELAXED);
  out[3] = __atomic_load_n(&g_plans, __ATOMIC_RELAXED);
}
extern "C" int tsim_build_flags(void) {
#ifdef TSIM_EXPERIMENT
  return 1;
#else
  return 0;
#endif
}
extern "C" int tsim_set_timing_events(void* start, void* stop) {
  g_ev_start = (cudaEvent_t)start;
  g_ev_stop = (cudaEvent_t)stop;
  return TSIM_OK;
}

static int check_search_args(int64_t Q, int64_t N, int64_t D, int k, int q_dt, int c_dt, int mode) {
  TSIM_CHECK_ARG(Q >= 0 && N >= 0 && D > 0, "search: bad shape Q=%lld N=%lld D=%lld", (long long)Q, (long long)N, (long long)D);
  TSIM_CHECK_ARG(k >= 1 && k <= 1024, "search: k=%d out of range [1, 1024]", k);
  TSIM_CHECK_ARG(dtype_size(q_dt) && dtype_size(c_dt), "search: bad dtype q=%d c=%d", q_dt, c_dt);
  TSIM_CHECK_ARG(mode >= TSIM_MODE_AUTO && mode <= TSIM_MODE_TENSOR, "search: bad mode %d", mode);
  TSIM_CHECK_ARG(N < (int64_t)0xfffffff0, "search: N=%lld rows per shard exceeds 2^32", (long long)N);
  TSIM_CHECK_ARG(Q < ((int64_t)1 << 31), "search: Q=%lld queries per call exceeds 2^31", (long long)Q);
  return TSIM_OK;
}

extern "C" size_t tsim_search_workspace_bytes(int64_t Q, int64_t N, int64_t D, int k, int q_dt, int c_dt,
                                              int mode) {
  if (check_search_args(Q, N, D, k, q_dt, c_dt, mode) != TSIM_OK) return 0;
  SearchPlan p;
  if (make_search_plan(Q, N, D, k, q_dt, c_dt, mode, /*need_invnorm=*/true, /*shadow=*/0, &p) != TSIM_OK) return 0;
  return p.total;
}

extern "C" size_t tsim_search_shadow_workspace_bytes(int64_t Q, int64_t N, int64_t D, int k, int shadow_dt) {
  if (check_search_args(Q, N, D, k, shadow_dt, shadow_dt, TSIM_MODE_AUTO) != TSIM_OK) return 0;
  SearchPlan p;
  if (make_search_plan(Q, N, D, k, shadow_dt, shadow_dt, TSIM_MODE_AUTO, true, /*shadow=*/1, &p) != TSIM_OK) return 0;
  return p.total;
}

// q / corpus: the rows results are defined on (re-score, exact scan).  tq / tcorpus (dtype t_dt): what the
// tensor pass reads -- the same arrays, or bf16 shadows of fp32 / fp16 rows (`shadow`).
static int search_impl(const void* q, int q_dt, int64_t q_stride, const void* corpus, int c_dt, int64_t c_stride,
                       const void* tq, int64_t tq_stride, const void* tcorpus, int64_t tc_stride, int t_dt, int shadow,
                       const float* corpus_inv_norm, int64_t Q, int64_t N,
                       int64_t D, int k, int64_t idx_base, int64_t exclude_self_base, int mode,
                       float* out_score, double* out_score64, int64_t* out_idx,
                       int32_t* out_flags, void* ws, size_t ws_bytes, void* stream,
                       const SearchPlan* prepared = nullptr, MapCache* maps = nullptr) {
  int rc = prepared ? TSIM_OK : check_search_args(Q, N, D, k, q_dt, c_dt, mode);
  if (rc) return rc;
  if (Q == 0) return TSIM_OK;
  TSIM_CHECK_ARG(q && out_score && out_idx, "search: null pointer");
  TSIM_CHECK_ARG(N == 0 || corpus, "search: null corpus");
  TSIM_CHECK_ARG(q_stride >= D && c_stride >= D, "search: row stride smaller than D");
  cudaStream_t st = (cudaStream_t)stream;
  SearchPlan p;
  // width of what the tensor pass reads: a split shadow's pass is three segments long (the query shadow's width)
  const int64_t Dt = shadow == 2 ? 3 * split_shadow_seg(D) : D;
  if (prepared) p = *prepared;     // a plan handle (tsim_plan_create): no planning, cached TMA descriptors
  else rc = make_search_plan(Q, N, Dt, k, shadow ? t_dt : q_dt, shadow ? t_dt : c_dt, mode, true, shadow, &p);
  if (rc) return rc;
  if (!ws || ws_bytes < p.total) {
    set_error("search: workspace too small (%zu < %zu bytes)", ws_bytes, p.total);
    return TSIM_ERR_WORKSPACE;
  }
  if (p.use_tensor) {
    const int per16 = 16 / dtype_size(t_dt);
    const bool aligned = (((uintptr_t)tq & 15) == 0) && (((uintptr_t)tcorpus & 15) == 0) &&
                         (tq_stride % per16 == 0) && (tc_stride % per16 == 0);
    if (!aligned) {
      if (mode == TSIM_MODE_TENSOR) {
        set_error("search: TMA needs 16-byte aligned bases and row strides");
        return TSIM_ERR_MISALIGNED;
      }
      p.use_tensor = 0;
    }
  }
  char* w = (char*)ws;
  uint32_t* thr = (uint32_t*)(w + p.off_thr);
  int32_t* flag_cnt = (int32_t*)(w + p.off_flagcnt);
  int32_t* flag_list = (int32_t*)(w + p.off_flaglist);
  double* ex_score = (double*)(w + p.off_ex_score);
  uint32_t* ex_idx = (uint32_t*)(w + p.off_ex_idx);
  const int self_on = exclude_self_base >= 0;
  const int64_t self_off = exclude_self_base - idx_base;  // local corpus row of query 0's own row

  if (p.use_tensor) {
    // thr, flag_cnt, the unit-claim areas, the append counters and the retry stage's thr / flag_cnt are adjacent:
    // one prep kernel zeroes them and writes the zero-padded copy of the queries (TMA OOB fill is slow)
    const size_t zero_end = p.retry ? p.off_r_flagcnt + 256
                            : p.fused ? p.off_gbar + 256
                            : p.append ? p.off_app_cnt + align_up((size_t)Q * sizeof(uint32_t), 256)
                                       : p.off_sched + 3 * p.sched_area;
    uint64_t* sched = p.sched_area ? (uint64_t*)(w + p.off_sched) : nullptr;
    const void* qt = tq;
    int64_t qt_stride = tq_stride;
    const int qrows = p.swapped ? 32 : p.pair ? 256 : 128;
    const size_t rowb = (size_t)Dt * dtype_size(t_dt);
    const bool pad = Q % qrows != 0 || p.qrep > 1;
    // (small-batch plans: the append lists, which sit right below thr, are zeroed as well -- search_sw.cu re-makes a
    // crowded query's threshold from its list and must not read a previous call's keys)
    const size_t zero_begin = p.swapped ? p.off_app_keys : p.off_thr;
    rc = launch_search_prep(w + zero_begin, align_up(zero_end - zero_begin, 256), tq, (size_t)tq_stride * dtype_size(t_dt),
                            pad ? w + p.off_qpad : nullptr, rowb, Q, (int64_t)p.QB * qrows, p.qrep > 1 ? 128 / p.qrep : 0, st);
    if (rc) return rc;
    if (pad) { qt = w + p.off_qpad; qt_stride = Dt; }
    const float* c_inv = corpus_inv_norm;
    if (!c_inv) {
      float* tmp = (float*)(w + p.off_invnorm);
      // (a split shadow's rows are [hi | lo | hi]: their own norm is not the row's -- take it from the originals)
      rc = shadow == 2 ? launch_row_inv_norm(corpus, c_dt, N, D, c_stride, tmp, st)
                       : launch_row_inv_norm(tcorpus, t_dt, N, D, tc_stride, tmp, st);
      if (rc) return rc;
      c_inv = tmp;
    }
    const bool timed = g_ev_start && g_ev_stop;
    if (timed) TSIM_CUDA(cudaEventRecord(g_ev_start, st));
    uint64_t* cand = (uint64_t*)(w + p.off_cand);
    uint64_t* app_keys = p.append ? (uint64_t*)(w + p.off_app_keys) : nullptr;
    uint32_t* app_cnt = p.append ? (uint32_t*)(w + p.off_app_cnt) : nullptr;
    uint32_t* ladder = nullptr;
    if (p.swapped) {
      // small batches: sample tiles -> 16-entry lists, thresholds + ladders, then the main pass appends
      uint32_t* lad = (uint32_t*)(w + p.off_ladder);
      rc = launch_search_sw(qt, qt_stride, tcorpus, tc_stride, t_dt, c_inv, Q, N, Dt, self_on, self_off, p, 1, cand, thr, lad,
                            app_keys, app_cnt, st, maps);
      if (rc) return rc;
      rc = launch_sw_tighten(Q, p, cand, thr, lad, app_keys, app_cnt, st);
      if (rc) return rc;
      rc = launch_search_sw(qt, qt_stride, tcorpus, tc_stride, t_dt, c_inv, Q, N, Dt, self_on, self_off, p, 0, cand, thr, lad,
                            app_keys, app_cnt, st, maps);
      if (rc) return rc;
    } else if (p.fused) {
      uint32_t* lad = knob_on("TSIM_NO_LADDER") ? nullptr : (uint32_t*)(w + p.off_ladder);   // experiment knob
      rc = launch_search_tc(qt, qt_stride, tcorpus, tc_stride, t_dt, c_inv, Q, N, Dt, self_on, self_off, p, TC_PASS_FUSED,
                            cand, thr, lad, sched, st, maps, nullptr, nullptr, 0, nullptr, nullptr,
                            (uint32_t*)(w + p.off_gbar));
      if (rc) return rc;
    } else {
    if (p.boot_tiles) {
      uint32_t* lad = knob_on("TSIM_NO_LADDER") ? nullptr : (uint32_t*)(w + p.off_ladder);   // experiment knob
      if (p.mini_mult) {
        rc = launch_search_tc(qt, qt_stride, tcorpus, tc_stride, t_dt, c_inv, Q, N, Dt, self_on, self_off, p,
                              TC_PASS_MINI, cand, thr, nullptr, sched, st, maps, nullptr, nullptr, 0, app_keys, app_cnt);
        if (rc) return rc;
        rc = p.append ? launch_tighten_app(Q, p, app_keys, app_cnt, thr, lad, st)
                      : launch_tighten(Q, p, (int)p.mini_slots, cand, thr, lad, st);
        if (rc) return rc;
        ladder = lad;
      }
      rc = launch_search_tc(qt, qt_stride, tcorpus, tc_stride, t_dt, c_inv, Q, N, Dt, self_on, self_off, p,
                            p.mini_mult ? TC_PASS_SAMPLE_REST : TC_PASS_SAMPLE, cand, thr, ladder, sched, st, maps,
                            nullptr, nullptr, 0, app_keys, app_cnt);
      if (rc) return rc;
      // re-levels the ladder
      rc = p.append ? launch_tighten_app(Q, p, app_keys, app_cnt, thr, lad, st)
                    : launch_tighten(Q, p, (int)(p.mini_slots + p.boot_slots), cand, thr, lad, st);
      if (rc) return rc;
      ladder = lad;
    }
    rc = launch_search_tc(qt, qt_stride, tcorpus, tc_stride, t_dt, c_inv, Q, N, Dt, self_on, self_off, p,
                          p.boot_tiles ? TC_PASS_MAIN : TC_PASS_ALL, cand, thr, ladder, sched, st, maps,
                          nullptr, nullptr, 0, app_keys, app_cnt);
    if (rc) return rc;
    }
    if (timed) TSIM_CUDA(cudaEventRecord(g_ev_stop, st));
    SelRetry first = {nullptr, nullptr, p.retry * kRetryQ, 0, 0, w + p.off_r_q, D};
    SearchPlan psel = p;
    if (p.swapped) psel.NC = 0;        // the sample keys that matter were moved to the append lists (sw_tighten_kernel)
    rc = launch_select_rescore(q, q_dt, q_stride, corpus, c_dt, c_stride, Q, N, D, k, idx_base, psel,
                               (const uint64_t*)(w + p.off_cand), thr, flag_cnt, flag_list,
                               out_score, out_score64, out_idx, out_flags, st, p.retry ? &first : nullptr,
                               app_keys, app_cnt);
    if (rc) return rc;
    if (p.retry) {
      // Queries whose KP candidates could not be proven complete (ties straddling ranks k..KP) are re-run
      // in compact blocks of 128 with KP = 112 lists: per round one sticky single-launch pass over the
      // corpus and a second select_rescore; every kernel of the stage reads the flagged count on the
      // device and leaves at once when its round has nothing to do (the usual case).
      SearchPlan pr;
      memset(&pr, 0, sizeof(pr));
      pr.use_tensor = 1; pr.eps = p.eps; pr.KP = kRetryKP; pr.pair = 0; pr.QB = 1; pr.sticky = 1;
      pr.Gq = p.r_Gq; pr.R = 256; pr.NC = p.r_Gq;
      int32_t* r_flag_cnt = (int32_t*)(w + p.off_r_flagcnt);
      int32_t* r_flag_list = (int32_t*)(w + p.off_r_flaglist);
      uint64_t* r_cand = (uint64_t*)(w + p.off_r_cand);
      const size_t rowb = (size_t)D * dtype_size(t_dt);
      for (int r = 0; r < p.retry; ++r) {
        uint32_t* r_thr = (uint32_t*)(w + p.off_r_thr) + (size_t)r * kRetryQ;
        rc = launch_search_tc(w + p.off_r_q + (size_t)r * kRetryQ * rowb, D, tcorpus, tc_stride, t_dt, c_inv, kRetryQ, N, D,
                              self_on, self_off, pr, TC_PASS_ALL, r_cand, r_thr, nullptr, nullptr, st, maps,
                              flag_cnt, flag_list + (size_t)r * kRetryQ, r * kRetryQ);
        if (rc) return rc;
        SelRetry again = {flag_cnt, flag_list, p.retry * kRetryQ, r * kRetryQ, r == p.retry - 1, nullptr, 0};
        rc = launch_select_rescore(q, q_dt, q_stride, corpus, c_dt, c_stride, kRetryQ, N, D, k, idx_base, pr,
                                   r_cand, r_thr, r_flag_cnt, r_flag_list, out_score, out_score64, out_idx,
                                   out_flags, st, &again);
        if (rc) return rc;
      }
      flag_cnt = r_flag_cnt;      // the float64 scan answers what is left
      flag_list = r_flag_list;
    }
    // queries whose candidate set could not be proven complete: float64 scan (usually none)
    rc = launch_search_exact(q, q_dt, q_stride, corpus, c_dt, c_stride, Q, N, D, k, self_on, self_off, p,
                             flag_cnt, flag_list, ex_score, ex_idx, p.has_ex_rinv ? (double*)(w + p.off_ex_rinv) : nullptr, st);
    if (rc) return rc;
    return launch_merge_exact_lists(q, q_dt, q_stride, corpus, c_dt, c_stride, Q, D, k, idx_base, p,
                                    flag_cnt, flag_list, ex_score, ex_idx, out_score, out_score64,
                                    out_idx, out_flags, st);
  }
  const bool timed = g_ev_start && g_ev_stop;
  if (timed) TSIM_CUDA(cudaEventRecord(g_ev_start, st));
  rc = launch_search_exact(q, q_dt, q_stride, corpus, c_dt, c_stride, Q, N, D, k, self_on, self_off, p,
                           nullptr, nullptr, ex_score, ex_idx, p.has_ex_rinv ? (double*)(w + p.off_ex_rinv) : nullptr, st);
  if (rc) return rc;
  if (timed) TSIM_CUDA(cudaEventRecord(g_ev_stop, st));
  return launch_merge_exact_lists(q, q_dt, q_stride, corpus, c_dt, c_stride, Q, D, k, idx_base, p,
                                  nullptr, nullptr, ex_score, ex_idx, out_score, out_score64, out_idx,
                                  out_flags, st);
}

extern "C" int tsim_search_topk(const void* q, int q_dt, int64_t q_stride, const void* corpus, int c_dt,
                                int64_t c_stride, const float* corpus_inv_norm, int64_t Q, int64_t N,
                                int64_t D, int k, int64_t idx_base, int64_t exclude_self_base, int mode,
                                float* out_score, double* out_score64, int64_t* out_idx,
                                int32_t* out_flags, void* ws, size_t ws_bytes, void* stream) {
  return search_impl(q, q_dt, q_stride, corpus, c_dt, c_stride, q, q_stride, corpus, c_stride, c_dt, 0,
                     corpus_inv_norm, Q, N, D, k, idx_base, exclude_self_base, mode, out_score, out_score64, out_idx,
                     out_flags, ws, ws_bytes, stream);
}

extern "C" int tsim_search_topk_shadow(const void* q, int q_dt, int64_t q_stride, const void* corpus, int c_dt,
                                       int64_t c_stride, const void* q_shadow, int64_t qs_stride,
                                       const void* corpus_shadow, int64_t cs_stride, int shadow_dt,
                                       const float* shadow_inv_norm, int64_t Q, int64_t N, int64_t D, int k,
                                       int64_t idx_base, int64_t exclude_self_base,
                                       float* out_score, double* out_score64, int64_t* out_idx,
                                       int32_t* out_flags, void* ws, size_t ws_bytes, void* stream) {
  TSIM_CHECK_ARG(q_shadow && (N == 0 || corpus_shadow), "search_shadow: null shadow pointer");
  TSIM_CHECK_ARG(shadow_dt == TSIM_BF16, "search_shadow: the shadow must be bf16 (got dtype %d)", shadow_dt);
  TSIM_CHECK_ARG(qs_stride >= D && cs_stride >= D, "search_shadow: row stride smaller than D");
  return search_impl(q, q_dt, q_stride, corpus, c_dt, c_stride, q_shadow, qs_stride, corpus_shadow, cs_stride,
                     shadow_dt, 1, shadow_inv_norm, Q, N, D, k, idx_base, exclude_self_base, TSIM_MODE_AUTO,
                     out_score, out_score64, out_idx, out_flags, ws, ws_bytes, stream);
}

// ---- test hooks ---------------------------------------------------------------------------------
extern "C" float tsim_debug_eps(int64_t D, int dt, int shadow) {
  return shadow == 2 ? split_shadow_eps(D) : shadow ? shadow_eps(D) : approx_eps(D, dtype_size(dt) ? dtype_size(dt) : 2);
}

extern "C" int tsim_debug_tensor_pass(const void* q, int64_t q_stride, const void* corpus, int64_t c_stride, int dt,
                                      const float* corpus_inv_norm, int64_t Q, int64_t N, int64_t D,
                                      uint64_t* out_keys, int64_t* out_lists, uint32_t* thr_scratch, void* stream) {
  TSIM_CHECK_ARG(Q >= 1 && Q <= 128 && N >= 1 && N < (int64_t)0x7fffff00, "debug_tensor_pass: needs 1 <= Q <= 128, N >= 1");
  TSIM_CHECK_ARG(tensor_shape_ok(Q, N, D, 100, dt, dt), "debug_tensor_pass: bf16 (D %% 8 == 0) or e4m3 (D %% 16 == 0) only");
  TSIM_CHECK_ARG(q && corpus && corpus_inv_norm && out_keys && out_lists && thr_scratch, "debug_tensor_pass: null pointer");
  const int per16 = 16 / dtype_size(dt);
  TSIM_CHECK_ARG((((uintptr_t)q | (uintptr_t)corpus) & 15) == 0 && q_stride % per16 == 0 && c_stride % per16 == 0,
                 "debug_tensor_pass: TMA needs 16-byte aligned bases and row strides");
  const int64_t T = (N + 255) / 256;
  const int sms = device_sm_count();
  SearchPlan pr;
  memset(&pr, 0, sizeof(pr));
  pr.use_tensor = 1; pr.eps = approx_eps(D, dtype_size(dt)); pr.KP = kRetryKP; pr.pair = 0; pr.QB = 1; pr.sticky = 1;
  pr.Gq = (int)(T < sms ? T : sms); pr.R = 256; pr.NC = pr.Gq;
  *out_lists = pr.NC;
  cudaStream_t st = (cudaStream_t)stream;
  TSIM_CUDA(cudaMemsetAsync(thr_scratch, 0, (size_t)Q * sizeof(uint32_t), st));
  return launch_search_tc(q, q_stride, corpus, c_stride, dt, corpus_inv_norm, Q, N, D, 0, 0, pr, TC_PASS_ALL, out_keys,
                          thr_scratch, nullptr, nullptr, st);
}

// ---- plan handles ----------------------------------------------------------------------------
struct tsim_plan {
  SearchPlan p;
  int64_t Q, N, D;
  int k, q_dt, c_dt, mode, shadow_dt;   // shadow_dt < 0: no shadow
  int shadow_kind;                      // 0 none, 1 rounded shadow (D wide), 2 split shadow (3 D wide)
  MapCache* maps;
};

static tsim_plan_t* plan_create(int64_t Q, int64_t N, int64_t D, int k, int q_dt, int c_dt, int mode, int shadow_dt,
                                int shadow_kind) {
  const bool shadow = shadow_kind != 0;
  if (shadow && shadow_dt != TSIM_BF16) { set_error("plan: the shadow must be bf16 (got dtype %d)", shadow_dt); return nullptr; }
  if (check_search_args(Q, N, D, k, q_dt, c_dt, shadow ? TSIM_MODE_AUTO : mode) != TSIM_OK) return nullptr;
  tsim_plan* h = new (std::nothrow) tsim_plan();
  if (!h) { set_error("plan: out of host memory"); return nullptr; }
  h->Q = Q; h->N = N; h->D = D; h->k = k; h->q_dt = q_dt; h->c_dt = c_dt;
  h->mode = shadow ? TSIM_MODE_AUTO : mode; h->shadow_dt = shadow ? shadow_dt : -1; h->shadow_kind = shadow_kind;
  if (make_search_plan(Q, N, shadow_kind == 2 ? 3 * split_shadow_seg(D) : D, k, shadow ? shadow_dt : q_dt, shadow ? shadow_dt : c_dt, h->mode, true,
                       shadow_kind, &h->p) != TSIM_OK) {
    delete h;
    return nullptr;
  }
  h->maps = map_cache_create();
  return h;
}

extern "C" tsim_plan_t* tsim_plan_create(int64_t Q, int64_t N, int64_t D, int k, int q_dt, int c_dt, int mode,
                                         int shadow_dt) {
  return plan_create(Q, N, D, k, q_dt, c_dt, mode, shadow_dt, shadow_dt >= 0 ? 1 : 0);
}

extern "C" tsim_plan_t* tsim_plan_create_split_shadow(int64_t Q, int64_t N, int64_t D, int k, int q_dt, int c_dt) {
  return plan_create(Q, N, D, k, q_dt, c_dt, TSIM_MODE_AUTO, TSIM_BF16, 2);
}

extern "C" void tsim_plan_destroy(tsim_plan_t* h) {
  if (!h) return;
  map_cache_destroy(h->maps);
  delete h;
}

extern "C" size_t tsim_plan_workspace_bytes(const tsim_plan_t* h) { return h ? h->p.total : 0; }

extern "C" int tsim_plan_search(tsim_plan_t* h, const void* q, int64_t q_stride, const void* corpus, int64_t c_stride,
                                const float* corpus_inv_norm, const void* q_shadow, int64_t qs_stride,
                                const void* corpus_shadow, int64_t cs_stride, int64_t idx_base,
                                int64_t exclude_self_base, float* out_score, double* out_score64, int64_t* out_idx,
                                int32_t* out_flags, void* ws, size_t ws_bytes, void* stream) {
  TSIM_CHECK_ARG(h, "plan_search: null plan");
  const bool shadow = h->shadow_dt >= 0;
  if (shadow) {
    TSIM_CHECK_ARG(q_shadow && (h->N == 0 || corpus_shadow), "plan_search: this plan needs the bf16 shadows");
    const int64_t seg = h->shadow_kind == 2 ? split_shadow_seg(h->D) : h->D;
    TSIM_CHECK_ARG(qs_stride >= (h->shadow_kind == 2 ? 3 : 1) * seg && cs_stride >= (h->shadow_kind == 2 ? 2 : 1) * seg,
                   "plan_search: shadow row stride smaller than the shadow's width");
    return search_impl(q, h->q_dt, q_stride, corpus, h->c_dt, c_stride, q_shadow, qs_stride, corpus_shadow, cs_stride,
                       h->shadow_dt, h->shadow_kind, corpus_inv_norm, h->Q, h->N, h->D, h->k, idx_base, exclude_self_base,
                       TSIM_MODE_AUTO, out_score, out_score64, out_idx, out_flags, ws, ws_bytes, stream, &h->p, h->maps);
  }
  return search_impl(q, h->q_dt, q_stride, corpus, h->c_dt, c_stride, q, q_stride, corpus, c_stride, h->c_dt, 0,
                     corpus_inv_norm, h->Q, h->N, h->D, h->k, idx_base, exclude_self_base, h->mode, out_score,
                     out_score64, out_idx, out_flags, ws, ws_bytes, stream, &h->p, h->maps);
}

extern "C" int tsim_merge_topk(const double* sc, const int64_t* ix, int64_t Q, int64_t n_lists, int k_in,
                               int k_out, float* out_score, double* out_score64, int64_t* out_idx,
                               void* stream) {
  TSIM_CHECK_ARG(Q >= 0 && n_lists >= 1 && k_in >= 1 && k_out >= 1, "merge_topk: bad shape");
  TSIM_CHECK_ARG(n_lists * (int64_t)k_in <= 4096, "merge_topk: n_lists * k_in = %lld exceeds 4096",
                 (long long)(n_lists * (int64_t)k_in));
  if (Q == 0) return TSIM_OK;
  TSIM_CHECK_ARG(sc && ix && out_score && out_idx, "merge_topk: null pointer");
  return launch_merge_topk(sc, ix, Q, n_lists, k_in, k_out, k_in, n_lists * (int64_t)k_in, out_score, out_score64,
                           out_idx, (cudaStream_t)stream);
}

extern "C" int tsim_merge_topk_strided(const double* sc, const int64_t* ix, int64_t Q, int64_t n_lists, int k_in,
                                       int64_t list_stride, int64_t query_stride, int k_out, float* out_score,
                                       double* out_score64, int64_t* out_idx, void* stream) {
  TSIM_CHECK_ARG(Q >= 0 && n_lists >= 1 && k_in >= 1 && k_out >= 1, "merge_topk_strided: bad shape");
  TSIM_CHECK_ARG(n_lists * (int64_t)k_in <= 4096, "merge_topk_strided: n_lists * k_in = %lld exceeds 4096",
                 (long long)(n_lists * (int64_t)k_in));
  TSIM_CHECK_ARG(list_stride >= 0 && query_stride >= k_in, "merge_topk_strided: bad strides");
  if (Q == 0) return TSIM_OK;
  TSIM_CHECK_ARG(sc && ix && out_score && out_idx, "merge_topk_strided: null pointer");
  return launch_merge_topk(sc, ix, Q, n_lists, k_in, k_out, list_stride, query_stride, out_score, out_score64, out_idx,
                           (cudaStream_t)stream);
}
