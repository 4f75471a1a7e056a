// K2 (tensor-core pass): query x corpus^T on tcgen05 / TMEM, fed by TMA, with the top-k
// candidate selection fused into the epilogue so the [Q, N] score matrix never exists in HBM.
//
// Replaces the per-query loop of SentenceMiningPipeline._search (reference
// src/pipeline/search_pipeline.py:73-79: one F.cosine_similarity pass over the whole corpus
// and one torch.topk PER QUERY) and cos_sim's full score matrix (src/utils/metrics.py:99-101).
//
// Orientation: queries are the MMA M side (128 per CTA = the 128 TMEM lanes), corpus rows the
// N side (256 per tile = 256 fp32 TMEM columns), D the K side in 64-element (128-byte) blocks.
// After a tile's MMAs, TMEM lane i holds query i's 256 dot products; epilogue thread i reads
// them with tcgen05.ld (32x32b: thread <-> lane), scales by the corpus row's inverse norm and
// compares with its query's running threshold -- one FMUL + one compare per score, and a rare
// insertion into that query's private sorted list in shared memory.  Two 256-column
// accumulator stages (all 512 TMEM columns) let the epilogue of tile t overlap the MMAs of t+1.
//
// Work unit = (128-query block, R-row corpus chunk); one persistent CTA per SM claims units from a
// global counter, query block fastest, so the CTAs that share a corpus chunk run side by side and
// the chunk is read from HBM once and from L2 by the rest (get_unit below).
// Each unit leaves its KP best (approx score, row) keys in cand[q][chunk][KP]; full lists raise
// thr[q] (atomicMax) so later units start with a tight threshold.  select_merge.cu finishes.
//
// Roofline: 2*Q*N*D flops on the tensor pipe (Q >~ 240) or N*D*2 bytes from HBM (small Q).
#include <cuda.h>
#include <stdlib.h>

#include "tsim_common.cuh"
#include "tc_ptx.cuh"

namespace tsim {
namespace {

constexpr int BM = 128;        // queries per CTA tile (TMEM lanes)
constexpr int BN = 256;        // corpus rows per tile (TMEM columns per accumulator stage)
constexpr int kCnStride = BN + 16;   // per-warp copy of a tile's inverse norms + (max, min) per 32-column chunk
constexpr int kThreads = 256;  // warps0-3 epilogue, warp4 TMA, warp5 MMA, warp6 TMEM alloc, warp7 idle (the issue
                               // arbiter favours the higher warp id: the two latency-critical single-thread roles win)

constexpr int kEpiThreads = 128;
constexpr int A_BYTES = BM * BK_BYTES;   // 16 KB: 128 query rows x 128 bytes
// B (corpus) bytes per CTA per stage: all 256 rows of the tile for a lone CTA, 128 rows for each CTA
// of a pair (cta_group::2: the MMA reads the other half from the peer's shared memory)
template <bool PAIR> struct Cfg {
  static constexpr int B_ROWS = PAIR ? BN / 2 : BN;
  static constexpr int B_BYTES = B_ROWS * BK_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;       // 48 KB lone, 32 KB per CTA of a pair
  static constexpr int MMA_M = PAIR ? 2 * BM : BM;
  // UMMA instruction descriptor: D = f32 (bit 4), A/B format at bits 7/10 (kind::f16: 1 = bf16;
  // kind::f8f6f4: 0 = e4m3), both K-major, N = 256, M = 128 / 256
  static constexpr uint32_t IDESC_BF16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                         ((uint32_t)(MMA_M >> 4) << 24);
  static constexpr uint32_t IDESC_E4M3 = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(MMA_M >> 4) << 24);
};

struct TcArgs {
  const float* c_inv;   // [N] inverse norms of the stored corpus rows
  int64_t Q, N;
  int kblocks;          // ceil(D * element size / 128)
  int ksplit;           // split shadow: corpus k-block of pass k-block 3 j + r is j (r < 2) or ksplit + j (r == 2); 0: kb
  int QB;               // query blocks (128 queries; 256 when CTA pairs are used)
  int T;                // corpus tiles (of 256 rows) THIS launch scans (see tile_mode); all tile arithmetic is 32-bit
  int tile_mode;        // 0: tiles 0..T-1; 1: the tiles i * tile_stride; 2: every tile that is not a multiple of
                        // tile_stride; 3: the multiples of tile_stride that are not multiples of tile_stride * tile_mult
  int tile_stride;      // distance between sample tiles (modes 1, 2, 3)
  int tile_mult;        // mode 3: every tile_mult-th sample tile belongs to the mini sample
  int slot_base;        // first candidate-list slot this launch writes
  // Fused sticky pass (one launch instead of sample -> tighten -> main): every worker first scans its sample
  // tile(s) (tile_mode / T / slot_base above), the epilogue warps of ALL CTAs meet at a grid barrier, each CTA turns
  // the sample lists of "its" queries into thresholds + ladders (warp_tighten), a second barrier, then the worker's
  // main unit (tile mode 2 over T2 tiles, slots from slot_base2).  TMA and MMA warps never wait: they run ahead
  // into the main unit's tiles while the thresholds are made.
  int fused; int T2; int slot_base2;
  // Query replication (sticky lone-CTA plans with Q <= 64): the padded 128-row query block holds the Q <= 128 / qrep
  // queries qrep times over, so all four TMEM lane quadrants carry the same scores and the four epilogue warps split a
  // tile's 256 columns between them (each thread: query et % (128 / qrep), columns of part et / (128 / qrep)).  With one
  // copy, a batch of <= 32 queries keeps ONE warp busy for all 8 column chunks of every tile while three idle -- and
  // on 384-byte fp8 rows that one warp, not HBM, set the pace.  Sample units are not split (one list per worker).
  int qrep;
  uint32_t* gbar;       // [2] grid-barrier counters (zeroed per call)
  int sticky;           // 1: CTA <-> (query block, tile residue class), one list per CTA lifetime
  int Gq;               // sticky: CTAs per query block
  int tpc;              // round-robin: tiles per corpus chunk
  int64_t NC;           // candidate lists per query (chunks, or Gq when sticky)
  int n_units;          // round-robin: QB * chunks of this launch
  int self_on; int64_t self_off;
  uint64_t* cand;       // [Q][NC][KP] packed keys
  uint32_t* thr;        // [Q] ordered-float global thresholds (0 = none yet)
  uint32_t* ladder;     // [Q][2 * kLadder] threshold ladder (main launch after a bootstrap), or null
  uint64_t* sched;      // round-robin: this launch's claim area (zeroed per call), or null = static dealing:
                        // [0] next-unit counter, then per worker a ring of kSchedRing claim records
  // append mode (main pass of a bootstrapped plan with KP >= 32): candidates go to a per-query global list
  uint64_t* app_keys;   // [Q][app_cap] packed keys
  uint32_t* app_cnt;    // [Q] entries appended so far (may exceed app_cap: the overflow is dropped and the query flagged)
  int app_cap;
  int KP;               // list capacity the ladder counts against (the template's KP is 0 in append mode)
  const int32_t* q_count;  // retry pass: device-side number of live queries (<= Q), or null = Q
  const int32_t* q_map;    // retry pass: compact query -> query of the call (self-exclusion), or null
  int q_skip;              // retry pass: live queries = clamp(*q_count - q_skip, 0, Q)
  int hot_scaled;       // experiment knob (TSIM_HOT_SCALED=1): scale all 32 scores of a chunk before the filter (the old hot path)
  int roles_low;        // experiment knob (TSIM_ROLES_LOW=1): TMA / MMA / alloc on warps 0-2, epilogue on warps 4-7
  int dbg;              // TSIM_DEBUG bits (diagnosis only): 1 skip A loads, 2 skip MMAs, 4 skip epilogue math
};

// Logical tile number of this launch -> tile of the corpus.  A bootstrap launch (mode 1) scans a
// strided sample so every query gets a good threshold before the main launch (mode 2) scans the rest.
__device__ __forceinline__ int actual_tile(const TcArgs& a, int mode, int u) {
  if (mode == 1) return u * a.tile_stride;
  if (mode == 2) return u + u / (a.tile_stride - 1) + 1;
  if (mode == 3) return (u + u / (a.tile_mult - 1) + 1) * a.tile_stride;
  return u;
}

// A unit = one candidate list: a query block and the sequence of corpus tiles scanned into it.
struct Unit { int qb; int slot; int tile0; int tstride; int ntiles; int mode; int split; };   // slot: absolute (first of
                                                                 // qrep when split); mode: tile mode; split: columns dealt over the parts

// Sticky schedule (few query blocks, the HBM-bound regime): CTA c keeps query block c % QB for its
// whole life and takes tiles j, j+Gq, ... (j = c / QB): perfect tile balance, neighbouring CTAs
// stream neighbouring tiles, the QB CTAs that share a tile run side by side (L2), and each query's
// list -- hence its threshold -- persists over everything the CTA sees.
// Round-robin schedule (many query blocks, the tensor-bound regime): unit u = (chunk u / QB,
// query block u % QB); the QB units of one chunk are handed out back to back so that the workers
// scanning a chunk run side by side and the chunk comes from HBM once and from L2 for the rest.
//  * claimed (a.sched): a worker takes the next unit from a global counter when it has issued the
//    loads of its previous one.  The QB units of a chunk then start within QB / workers of a unit's
//    duration of each other, whatever the workers' history: with static dealing every worker's
//    cold-path time random-walks away from its neighbours' over ~130 units, the 16 readers of a chunk
//    end up further apart than L2 holds, and the corpus came from DRAM 3.1 times.
//    One thread per worker claims (the TMA thread; the leader CTA's for a pair) and publishes
//    (iteration, unit) in the worker's ring in global memory; the MMA thread, the epilogue threads and
//    the peer CTA poll that record -- the claimer runs a pipeline depth ahead, so they rarely spin.
//    The ring cannot wrap onto a record still needed: the TMA thread is never more than
//    STAGES k-blocks + 2 accumulator tiles (< 10 one-tile units) ahead of the slowest reader.
//  * static (a.sched == null, TSIM_STATIC_UNITS=1): unit u -> worker u % nw.
// `wid` / `nw` = this worker's index and the number of workers (CTAs, or CTA pairs).
constexpr int kSchedRing = 32;
constexpr uint32_t kNoUnit = 0xffffffffu;

__device__ __forceinline__ bool get_unit(const TcArgs& a, int it, int wid, int nw, bool claimer, Unit& un) {
  if (a.sticky) {
    if (it > (a.fused ? 1 : 0) || wid >= a.Gq * a.QB) return false;
    const int j = wid / a.QB;
    const int T = it ? a.T2 : a.T;           // fused: unit 0 = the worker's sample tiles, unit 1 = its main tiles
    un.mode = it ? 2 : a.tile_mode;
    un.split = (a.qrep > 1 && un.mode != 1) ? 1 : 0;
    un.qb = wid % a.QB; un.slot = (it ? a.slot_base2 : a.slot_base) + (un.split ? j * a.qrep : j); un.tile0 = j; un.tstride = a.Gq;
    un.ntiles = j < T ? (T - j + a.Gq - 1) / a.Gq : 0;
    return true;
  }
  int u;
  if (a.sched) {
    uint64_t* rec = a.sched + 32 + (size_t)wid * kSchedRing + (it & (kSchedRing - 1));
    uint32_t got;
    if (claimer) {
      got = atomicAdd((unsigned int*)a.sched, 1u);
      if (got >= (uint32_t)a.n_units) got = kNoUnit;
      const uint64_t v = ((uint64_t)(uint32_t)(it + 1) << 32) | got;
      asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(rec), "l"(v) : "memory");
    } else {
      uint64_t v;
      do {
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(rec) : "memory");
      } while ((uint32_t)(v >> 32) != (uint32_t)(it + 1));
      got = (uint32_t)v;
    }
    if (got == kNoUnit) return false;
    u = (int)got;
  } else {
    u = wid + it * nw;
    if (u >= a.n_units) return false;
  }
  const int chunk = u / a.QB;
  un.qb = u - chunk * a.QB; un.slot = a.slot_base + chunk; un.tile0 = chunk * a.tpc; un.tstride = 1; un.mode = a.tile_mode;
  un.split = 0;
  un.ntiles = min(a.tpc, a.T - un.tile0);
  return true;
}

// Threshold ladder (layout and level function: tsim_common.cuh; levels and initial counts: tighten_kernel).
// Each epilogue thread bins the rows it inserts into a private histogram (16 x 16-bit counters in
// registers) and, at the end of a tile in which it inserted anything, adds the histogram to its query's
// global counters (fire-and-forget atomics), reads them back once and raises its threshold to the
// highest level that has >= KP rows at or above it.  Why that is a valid bound on the query's final
// KP-th best: every counted row sits in a candidate list, or was evicted from a full list whose KP
// entries are all at least as good -- either way >= KP candidates at or above the level reach
// select_rescore.  (Counts are read slightly stale and rows that are not inserted are never counted:
// both only delay a raise.)  Without this, after the bootstrap a query's threshold rises only when one
// unit's own list fills, i.e. almost never with short units.
struct Ladder {
  float base, step, inv;
  uint32_t hist[kLadder / 2];
  uint32_t* g;
  __device__ __forceinline__ void init(uint32_t* lad) {
    g = lad;
    base = INFINITY; step = 0.f; inv = 0.f;
    if (g) {
      const uint4 h = __ldcg((const uint4*)g);   // (a fused pass reads a ladder another CTA of the same launch wrote)
      base = __uint_as_float(h.x); step = __uint_as_float(h.y); inv = __uint_as_float(h.z);
    }
#pragma unroll
    for (int i = 0; i < kLadder / 2; ++i) hist[i] = 0;
  }
  __device__ __forceinline__ void count(float s) {
    const int j = ladder_level(base, step, inv, s);     // base = +inf without a ladder: j = -1
    if (j < 0) return;
    const uint32_t inc = 1u << ((j & 1) * 16);
    const int w = j >> 1;
#pragma unroll
    for (int i = 0; i < kLadder / 2; ++i) hist[i] += (i == w) ? inc : 0u;
  }
  // at most 256 rows are counted between two flushes (one tile), so the 16-bit counters cannot overflow
  __device__ __forceinline__ float flush(float thr, uint32_t KP) {
    uint32_t any = 0;
#pragma unroll
    for (int i = 0; i < kLadder / 2; ++i) any |= hist[i];
    if (!g || !any) return thr;
    uint32_t bin[kLadder];
#pragma unroll
    for (int i = 0; i < kLadder / 4; ++i) {
      const uint4 b = __ldcg((const uint4*)g + kLadder / 4 + i);
      bin[4 * i] = b.x; bin[4 * i + 1] = b.y; bin[4 * i + 2] = b.z; bin[4 * i + 3] = b.w;
    }
    uint32_t c = 0;
    int best = -1;
#pragma unroll
    for (int i = kLadder - 1; i >= 0; --i) {
      const uint32_t own = (hist[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
      if (own) atomicAdd(g + kLadder + i, own);
      c += bin[i] + own;
      if (best < 0 && c >= KP) best = i;
    }
#pragma unroll
    for (int i = 0; i < kLadder / 2; ++i) hist[i] = 0;
    return best >= 0 ? fmaxf(thr, ladder_value(base, step, best)) : thr;
  }
};

// ---- per-query candidate lists -----------------------------------------------------------------
// Each epilogue thread owns the running top-KP (approx score, row) list of its query, sorted
// descending, empty slots = (-inf, 0xffffffff).  The hot loop never touches the list: it takes the
// max of a 32-column chunk and compares it with `thr`.  Only when that fires does the cold path
// re-read the chunk from TMEM, 8 columns at a time, and insert what qualifies:
//  * KP == 16 (k <= 10, the headline case): the list lives in REGISTERS for the whole unit and an
//    insertion is a branch-free 16-step select network (no memory, no dependent-load chain);
//  * larger KP: the list lives in shared memory as [entry][thread] columns (bank-conflict free)
//    and the cold path is one non-inlined sorted shift-insertion.
// A full list's minimum is a valid lower bound of the query's KP-th best anywhere in the corpus,
// so it is published with atomicMax for every other CTA to filter with.
// Cold-path helpers shared by both list types.  The chunk's 32 scaled scores are in registers (the
// hot path just produced them): the columns that beat the threshold become a bit mask (branch-free)
// and ONE rolled loop walks the set bits, fetching sc[j] through a 5-level select tree (registers
// cannot be indexed dynamically).  The cold path stays a few hundred bytes of code; 32 (or 8) inlined
// insertions overflowed the instruction cache and cost ~4000 cycles per triggered chunk.
__device__ __forceinline__ uint32_t candidate_mask(const float* sc, float thr, int64_t row_base, int64_t self_row,
                                                   int lim) {
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) m |= (sc[j] > thr) ? (1u << j) : 0u;
  if (lim < 32) m &= (1u << lim) - 1u;
  const int64_t ds = self_row - row_base;
  if (ds >= 0 && ds < 32) m &= ~(1u << (int)ds);
  return m;
}
struct RegList16 {
  float a[16]; uint32_t r[16];
  __device__ __forceinline__ RegList16(float*, uint32_t*, int*) {}
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = -INFINITY; r[i] = 0xffffffffu; }
  }
  // sc[0..32) = this lane's scaled scores of the chunk
  __device__ __forceinline__ float slow(const float* sc, float thr, int64_t row_base, int64_t self_row, int lim,
                                        uint32_t* thr_g, Ladder& lad) {
    const float thr_in = thr;
    uint32_t m = candidate_mask(sc, thr, row_base, self_row, lim);
#pragma unroll 1
    while (m) {
      const int j = __ffs(m) - 1;
      m &= m - 1;
      const float s = select32(sc, j);
      if (s > thr) {              // thr may have risen since the mask was built
        const uint32_t row = (uint32_t)(row_base + j);
        lad.count(s);
        // a[] is sorted descending, so (s > a[i]) is monotone in i: entry i becomes the newcomer
        // where the predicate first turns true, the old a[i-1] after that, and stays otherwise.
        // Strict '>' puts the newcomer after equal scores (it has the larger row).
#pragma unroll
        for (int i = 15; i > 0; --i) {
          const bool gt = s > a[i], gtp = s > a[i - 1];
          a[i] = gt ? (gtp ? a[i - 1] : s) : a[i];
          r[i] = gt ? (gtp ? r[i - 1] : row) : r[i];
        }
        if (s > a[0]) { a[0] = s; r[0] = row; }
        thr = fmaxf(thr, a[15]);
      }
    }
    if (thr > thr_in) atomicMax(thr_g, f32_to_ord(thr));
    return thr;
  }
  __device__ __forceinline__ void tile_end() {}
  __device__ __forceinline__ void flush(uint64_t* dst) {
#pragma unroll
    for (int j = 0; j < 16; ++j) dst[j] = a[j] > -INFINITY ? pack_key(a[j], r[j]) : 0ull;
  }
};

// Larger lists (KP = 32 / 64 / 112) live in shared memory UNSORTED: a newcomer overwrites the current
// minimum and the new minimum is found by one pipelined scan of KP independent loads (a sorted
// shift would be a chain of ~KP/2 dependent load/store steps).  *statep packs the fill count (low 16
// bits) and the slot of the minimum (high 16 bits).  Which of several equal minima is evicted does
// not matter: every dropped row still has approx <= the final KP-th best (see select_merge.cu).
// Position and value of the minimum of a full list: one pipelined scan of KP independent loads.
template <int KP>
__device__ __noinline__ uint64_t smem_list_min(const float* ls) {
  float m = ls[0];
  int mp = 0;
#pragma unroll 8
  for (int i = 1; i < KP; ++i) {
    const float v = ls[i * kEpiThreads];
    if (v < m) { m = v; mp = i; }
  }
  return ((uint64_t)(uint32_t)mp << 32) | __float_as_uint(m);
}

template <int KP>
struct SmemList {
  float* ls; uint32_t* li; int* cntp;
  __device__ __forceinline__ SmemList(float* s, uint32_t* i, int* n) : ls(s), li(i), cntp(n) {}
  __device__ __forceinline__ void reset() {
#pragma unroll 4
    for (int j = 0; j < KP; ++j) { ls[j * kEpiThreads] = -INFINITY; li[j * kEpiThreads] = 0xffffffffu; }
    *cntp = 0;
  }
  // sc[0..32) = this lane's scaled scores of the chunk.  One predicated branch per column: a column
  // costs two instructions unless some lane of the warp has a candidate in it, so the cost follows
  // the number of candidates, not the number of chunks that have one.
  __device__ __forceinline__ float slow(const float* sc, float thr, int64_t row_base, int64_t self_row, int lim,
                                        uint32_t* thr_g, Ladder& lad) {
    const float thr_in = thr;
    int cnt = *cntp & 0xffff, minpos = *cntp >> 16;
    const int cnt0 = cnt;
    auto put = [&](float v, int j) {
      const int pos = cnt < KP ? cnt : minpos;
      ls[pos * kEpiThreads] = v;
      li[pos * kEpiThreads] = (uint32_t)(row_base + j);
      if (cnt < KP) ++cnt;
      if (cnt == KP) {                        // full: locate the new minimum
        const uint64_t r = smem_list_min<KP>(ls);
        minpos = (int)(r >> 32);
        thr = fmaxf(thr, __uint_as_float((uint32_t)r));
      }
    };
    uint32_t m = candidate_mask(sc, thr, row_base, self_row, lim);
#pragma unroll 1
    while (m) {
      const int j = __ffs(m) - 1;
      m &= m - 1;
      const float v = select32(sc, j);
      if (v > thr) put(v, j);     // thr may have risen since the mask was built (the list filled up)
    }
    *cntp = cnt | (minpos << 16);
    // ladder: the rows appended while the list was not full sit at [cnt0, cnt) (replacements of a full
    // list's minimum are not counted: that list's own minimum is the threshold then)
#pragma unroll 1
    for (int pos = cnt0; pos < cnt; ++pos) lad.count(ls[pos * kEpiThreads]);
    if (thr > thr_in) atomicMax(thr_g, f32_to_ord(thr));
    return thr;
  }
  __device__ __forceinline__ void tile_end() {}
  __device__ __forceinline__ void flush(uint64_t* dst) {
#pragma unroll 4
    for (int j = 0; j < KP; ++j) {
      const float sc = ls[j * kEpiThreads];
      dst[j] = sc > -INFINITY ? pack_key(sc, li[j * kEpiThreads]) : 0ull;
    }
  }
};

// Append mode (negative KP template argument = staging entries per thread): no list at all.  Once every query has a threshold near its final KP-th
// best (sample passes + ladder), a row that beats it is simply APPENDED to the query's global list: select_rescore
// sorts the few hundred survivors.  A row is dropped only by `score <= thr`, and thr is always a valid lower bound
// of the KP-th best candidate that reaches select_rescore (the sample's KP-th best, or a ladder level with >= KP
// appended rows at or above it), so the completeness proof is unchanged.  Shared memory holds only a staging column
// of a few keys per thread; it is written out with ONE atomicAdd per thread per tile, after the accumulator has
// been handed back (or at once when it fills).  No list in shared memory = all pipeline stages back (6 for a pair).
// Staging entries per thread: 8 for the main pass (~0.2 candidates per query per tile), 40 for the sample passes,
// whose thresholds are still loose (tens of candidates per query per tile: one flush per tile instead of one
// atomic round trip per 8 rows).  Encoded as a NEGATIVE KP template argument of the kernel.
constexpr int kAppStageMain = 8, kAppStageSample = 40;
__device__ __noinline__ void append_write_out(uint64_t* gkeys, uint32_t* gcnt, int cap, const uint64_t* st, int n) {
  const uint32_t pos = atomicAdd(gcnt, (uint32_t)n);
#pragma unroll 1
  for (int i = 0; i < n; ++i)
    if (pos + i < (uint32_t)cap) gkeys[pos + i] = st[i * kEpiThreads];
}
template <int STAGE>
struct AppendList {
  uint64_t* st;         // this thread's staging column: st[i * kEpiThreads]
  int n;
  uint64_t* gkeys; uint32_t* gcnt; int cap;
  __device__ __forceinline__ AppendList(float* s, uint32_t*, int*) : st((uint64_t*)s), n(0), gkeys(nullptr), gcnt(nullptr), cap(0) {}
  __device__ __forceinline__ void bind(uint64_t* keys, uint32_t* cnt, int c) { gkeys = keys; gcnt = cnt; cap = c; }
  __device__ __forceinline__ void reset() { n = 0; }
  __device__ __forceinline__ void write_out() {
    append_write_out(gkeys, gcnt, cap, st, n);
    n = 0;
  }
  __device__ __forceinline__ float slow(const float* sc, float thr, int64_t row_base, int64_t self_row, int lim,
                                        uint32_t*, Ladder& lad) {
    uint32_t m = candidate_mask(sc, thr, row_base, self_row, lim);
#pragma unroll 1
    while (m) {
      const int j = __ffs(m) - 1;
      m &= m - 1;
      const float v = select32(sc, j);
      lad.count(v);
      st[n * kEpiThreads] = pack_key(v, (uint32_t)(row_base + j));
      if (++n == STAGE) write_out();
    }
    return thr;
  }
  __device__ __forceinline__ void tile_end() { if (n) write_out(); }
  __device__ __forceinline__ void flush(uint64_t*) {}
};

template <int KP> struct ListFor { using type = SmemList<KP>; };
template <> struct ListFor<16> { using type = RegList16; };
template <> struct ListFor<-kAppStageMain> { using type = AppendList<kAppStageMain>; };
template <> struct ListFor<-kAppStageSample> { using type = AppendList<kAppStageSample>; };

template <int KP, int STAGES, bool PAIR, bool FP8>
__global__ void __launch_bounds__(kThreads, 1)
search_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c, TcArgs a) {
  extern __shared__ unsigned char smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment: round the dynamic window up (1 KB of slack is
  // requested by the launcher).
  // [STAGES][A 16K | B 32K] | lists | cnorm[2 acc][4 warps][BN + 16] | list fill counts | barriers | tmem ptr
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int B_BYTES = Cfg<PAIR>::B_BYTES;
  constexpr int STAGE_BYTES = Cfg<PAIR>::STAGE_BYTES;
  constexpr int BK = FP8 ? BK_BYTES : BK_BYTES / 2;       // elements per k-block (TMA coordinates are in elements)
  constexpr uint32_t IDESC = FP8 ? Cfg<PAIR>::IDESC_E4M3 : Cfg<PAIR>::IDESC_BF16;
  unsigned char* tiles = smem;
  constexpr int kListWords = KP > 0 ? 2 * KP : -2 * KP;       // 32-bit words per thread: (score, row) lists, or staged keys
  float* list_s = (float*)(smem + (size_t)STAGES * STAGE_BYTES);
  uint32_t* list_i = (uint32_t*)(list_s + (KP > 0 ? KP : 0) * kEpiThreads);
  float* cnorm = list_s + kListWords * kEpiThreads;
  int* list_n = (int*)(cnorm + 2 * 4 * kCnStride);
  uint64_t* bars = (uint64_t*)(list_n + kEpiThreads);
  uint64_t* full_bar = bars;                 // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]       MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;      // [2]       epilogue -> MMA
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);

  // retry pass: the number of live queries is only known on the device; usually none -> leave at once
  int64_t q_live = a.Q;
  pdl_trigger();
  if (a.q_count) {
    pdl_wait();          // the flagged count is the previous kernel's output
    const int64_t c = (int64_t)*a.q_count - a.q_skip;
    q_live = c < a.Q ? c : a.Q;
    if (q_live <= 0) return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int roles_low = TSIM_KNOB_DEV(a.roles_low), dbg = TSIM_KNOB_DEV(a.dbg);   // constants 0 in the release build
  const int kWarpTma = roles_low ? 0 : 4, kWarpMma = roles_low ? 1 : 5, kWarpAlloc = roles_low ? 2 : 6;
  const bool is_epi = roles_low ? warp >= 4 : warp < 4;
  // a pair = cluster of 2 CTAs: rank 0 (leader) issues the MMAs for both, each CTA loads its own 128
  // queries and its own half of the corpus tile, and scans its own 128 TMEM lanes
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int wid = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int nw = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == kWarpTma && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_c);
  }
  if (warp == kWarpMma && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&tfull_bar[s]), 1); mbar_init(smem_u32(&tempty_bar[s]), PAIR ? 8 : 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == kWarpAlloc) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();   // barrier inits visible to the peer before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail; from here on
  // its outputs (thresholds, ladder, padded queries, claim areas) are read
  pdl_wait();

  if (warp == kWarpTma) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      Unit un;
      for (int it = 0; get_unit(a, it, wid, nw, rank == 0, un); ++it) {
        for (int t = 0; t < un.ntiles; ++t) {
          const int row0 = actual_tile(a, un.mode, un.tile0 + t * un.tstride) * BN;
          int cj = 0, cr = 0;
          for (int kb = 0; kb < a.kblocks; ++kb) {
            // corpus k-block: a split shadow's rows are [hi | lo] and the pass walks (hi_j, hi_j, lo_j) against the
            // query shadow's (hi_j, lo_j, hi_j) -- the second read of hi_j comes from L2
            const int ckb = a.ksplit ? (cr < 2 ? cj : a.ksplit + cj) : kb;
            if (++cr == 3) { cr = 0; ++cj; }
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
            const uint32_t sa = smem_u32(tiles + (size_t)stage * STAGE_BYTES);
            if (PAIR) {
              // both CTAs' bytes land on the leader's barrier; only the leader arms it (for both)
              const uint32_t fb = map_to_cta(smem_u32(&full_bar[stage]), 0);
              if (rank == 0) mbar_arrive_expect_tx(smem_u32(&full_bar[stage]), 2 * ((dbg & 1) ? B_BYTES : STAGE_BYTES));
              if (!(dbg & 1)) tma_load_2d_pair(sa, &tmap_q, fb, kb * BK, un.qb * (2 * BM) + (int)rank * BM);
              tma_load_2d_pair(sa + A_BYTES, &tmap_c, fb, ckb * BK, row0 + (int)rank * (BN / 2));
            } else {
              const uint32_t fb = smem_u32(&full_bar[stage]);
              mbar_arrive_expect_tx(fb, (dbg & 1) ? B_BYTES : STAGE_BYTES);
              if (!(dbg & 1)) tma_load_2d(sa, &tmap_q, fb, kb * BK, un.qb * BM);
              tma_load_2d(sa + A_BYTES, &tmap_c, fb, ckb * BK, row0);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == kWarpMma) {
    // ===================== MMA issuer (the pair's leader issues for both CTAs) =====================
    if (lane == 0 && rank == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t aphase = 0;
      Unit un;
      for (int it = 0; get_unit(a, it, wid, nw, false, un); ++it) {
        for (int t = 0; t < un.ntiles; ++t) {
          mbar_wait(smem_u32(&tempty_bar[acc]), aphase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)acc * BN;
          for (int kb = 0; kb < a.kblocks; ++kb) {
            mbar_wait(smem_u32(&full_bar[stage]), phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(tiles + (size_t)stage * STAGE_BYTES);
            const uint64_t adesc = make_umma_desc(sa);
            const uint64_t bdesc = make_umma_desc(sa + A_BYTES);
#pragma unroll
            for (int k = 0; k < BK_BYTES / UMMA_K_BYTES; ++k) {
              if (dbg & 2) break;
              // advance 32 bytes (16 bf16 / 32 e4m3) inside the 128-byte swizzle row: +2 in >>4 units
              if (PAIR) tc_mma_pair<FP8>(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), IDESC, (kb | k) ? 1u : 0u);
              else tc_mma<FP8>(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), IDESC, (kb | k) ? 1u : 0u);
            }
            // frees the smem stage (in both CTAs of a pair) when the MMAs retire
            if (PAIR) tc_commit_pair(smem_u32(&empty_bar[stage])); else tc_commit(smem_u32(&empty_bar[stage]));
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          // accumulator stage complete: wake the epilogue warps (of both CTAs)
          if (PAIR) tc_commit_pair(smem_u32(&tfull_bar[acc])); else tc_commit(smem_u32(&tfull_bar[acc]));
          if (++acc == 2) { acc = 0; aphase ^= 1; }
        }
      }
    }
  } else if (is_epi) {
    // ===================== epilogue: threshold filter + per-query lists =====================
    const int et = threadIdx.x & 127;            // 0..127 = TMEM lane = query within the block
    const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32)) << 16;
    // (append mode: the staging column holds 8-byte keys, [entry][thread])
    typename ListFor<KP>::type list(KP > 0 ? list_s + et : (float*)((uint64_t*)list_s + et), list_i + et, list_n + et);
    int acc = 0; uint32_t aphase = 0;
    Unit un;
    for (int it = 0; get_unit(a, it, wid, nw, false, un); ++it) {
      const int64_t qbase = PAIR ? (int64_t)un.qb * (2 * BM) + rank * BM : (int64_t)un.qb * BM;
      const int qspan = a.qrep > 1 ? BM / a.qrep : BM;      // distinct queries in this CTA's block (replication, TcArgs)
      const int part = et / qspan;                          // which share of a tile's columns this thread examines
      const int64_t qg = qbase + (et & (qspan - 1));
      const bool qvalid = qg < q_live;
      const bool warp_live = __any_sync(0xffffffffu, qvalid);
      if (a.fused && it == 1) {
        // ---- between the sample unit and the main unit of a fused sticky pass ----
        // (1) every CTA's sample lists are in global memory
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
          atomicAdd(a.gbar, 1u);
          uint32_t seen;
          do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(a.gbar) : "memory"); } while (seen < gridDim.x);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // (2) thresholds + ladders: the 128 queries of this CTA's block half are dealt over the Gq workers that share
        // the block (query ql belongs to worker ql % Gq), a worker's queries over its four epilogue warps
        {
          constexpr int KPv = KP > 0 ? KP : 1;
          uint32_t* hist = reinterpret_cast<uint32_t*>(cnorm);   // cnorm is idle between units
          const int j = wid / a.QB;
          const uint32_t nkeys = (uint32_t)(a.Gq * KPv);
          if (nkeys <= 128u * kEpiKeys) {
            // all 128 epilogue threads per query, keys in registers (block-uniform loop: named barriers inside)
            for (int ql = j; ql < BM; ql += a.Gq) {
              const int64_t q2 = qbase + ql;
              if (q2 < q_live)
                epi_tighten(a.cand + ((size_t)q2 * a.NC + a.slot_base) * KPv, nkeys, KPv, a.thr + q2,
                            a.ladder ? a.ladder + (size_t)q2 * (2 * kLadder) : nullptr, hist, et);
            }
          } else {
            for (int ql = j + (warp & 3) * a.Gq; ql < BM; ql += 4 * a.Gq) {
              const int64_t q2 = qbase + ql;
              if (q2 < q_live)
                warp_tighten<true>(a.cand + ((size_t)q2 * a.NC + a.slot_base) * KPv, nkeys, KPv, a.thr + q2,
                                   a.ladder ? a.ladder + (size_t)q2 * (2 * kLadder) : nullptr, hist + (warp & 3) * 256);
            }
          }
        }
        // (3) ... are visible to every CTA before anybody filters with them
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
          atomicAdd(a.gbar + 1, 1u);
          uint32_t seen;
          do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(a.gbar + 1) : "memory"); } while (seen < gridDim.x);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      const int64_t self_row = a.self_on ? a.self_off + ((a.q_map && qvalid) ? (int64_t)a.q_map[qg] : qg) : -1;
      uint32_t* thr_g = a.thr + (qvalid ? qg : 0);
      Ladder lad;
      lad.init((a.ladder && qvalid && !(a.fused && it == 0)) ? a.ladder + (size_t)qg * (2 * kLadder) : nullptr);
      list.reset();
      if constexpr (KP < 0) list.bind(a.app_keys + (size_t)(qvalid ? qg : 0) * a.app_cap, a.app_cnt + (qvalid ? qg : 0), a.app_cap);
      float thr = qvalid ? -INFINITY : INFINITY;   // padded query lanes never insert
      // Tile metadata (the 256 inverse norms, 8 per lane, and the query's shared threshold) is
      // fetched one tile ahead so its global-load latency hides behind the previous tile's work.
      float nreg[8];
      uint32_t gthr = 0;
      auto fetch_meta = [&](int t) {
        const int64_t r0 = (int64_t)actual_tile(a, un.mode, un.tile0 + t * un.tstride) * BN;
        const int nc = (int)min((int64_t)BN, a.N - r0);
#pragma unroll
        for (int i = 0; i < 8; ++i) nreg[i] = (lane + 32 * i < nc) ? __ldg(a.c_inv + r0 + lane + 32 * i) : 0.f;
        gthr = qvalid ? __ldcg(thr_g) : 0u;
      };
      if (un.ntiles > 0) fetch_meta(0);
      for (int t = 0; t < un.ntiles; ++t) {
        const int64_t trow0 = (int64_t)actual_tile(a, un.mode, un.tile0 + t * un.tstride) * BN;
        const int ncols = (int)min((int64_t)BN, a.N - trow0);
        // each epilogue warp keeps a private copy of the tile's inverse norms: no cross-warp barrier
        // ... followed by, per 32-column chunk, the largest and the smallest of its inverse norms (chunk i = nreg[i]
        // across the lanes; inverse norms are >= 0, so their bit patterns order like the values; a NaN / Inf norm
        // reads as "not finite" below and sends its chunk down the exact path)
        float* cn = cnorm + (acc * 4 + (warp & 3)) * kCnStride;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          cn[lane + 32 * i] = nreg[i];
          const uint32_t hi = __reduce_max_sync(0xffffffffu, __float_as_uint(nreg[i]));
          const uint32_t lo = __reduce_min_sync(0xffffffffu, __float_as_uint(nreg[i]));
          if (lane == 0) { cn[BN + 2 * i] = __uint_as_float(hi); cn[BN + 2 * i + 1] = __uint_as_float(lo); }
        }
        if (gthr) thr = fmaxf(thr, ord_to_f32(gthr));
        __syncwarp();
        if (t + 1 < un.ntiles) fetch_meta(t + 1);
        mbar_wait(smem_u32(&tfull_bar[acc]), aphase);
        tc_fence_after();
        const uint32_t tbase = tmem_base + lane_addr + (uint32_t)acc * BN;
        // (a warp whose 32 queries are all padding -- three of the four at Q <= 32 -- has nothing to examine)
        int nchunks = ((dbg & 4) || !warp_live) ? 0 : (ncols + 31) / 32;
        int c0 = 0;                                          // this thread's chunks: [c0, nchunks)
        if (a.qrep > 1) {
          const int per = (BN / 32) / a.qrep;
          c0 = un.split ? part * per : 0;
          nchunks = min(nchunks, un.split ? c0 + per : (part == 0 ? BN / 32 : 0));
        }
        // The TMEM loads are software-pipelined: chunk c + 1 is requested before chunk c is examined, so the
        // tcgen05.ld round trip overlaps the max tree instead of preceding it.  With 384-byte fp8 rows a tile is
        // only 98 KB -- 2.1 us of HBM time -- and the serial ld -> wait -> examine loop took longer than that:
        // the epilogue, not HBM, set the pace of config 4 (skipping it: 0.88 -> 0.71 ms per search).
        auto examine = [&](const uint32_t (&v)[32], int c) {
          // Hot path: FMNMX3 tree over the 32 RAW dot products, one FMUL, one compare.  A scaled score is
          // s_j * inv_j with inv_lo <= inv_j <= inv_hi, so none can exceed  bound = mx * (mx >= 0 ? inv_hi : inv_lo)
          // (float rounding is monotone): bound <= thr proves the chunk holds no candidate without scaling a
          // single score or reading a norm.  With unit-norm rows inv_hi / inv_lo - 1 ~ 4e-3, i.e. ~8 % more chunks
          // reach the exact path than with the exact maximum; the 32 FMULs + 8 LDS.128 per chunk move there.
          float gm[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t* x = v + g * 8;
            gm[g] = fmaxf(fmaxf(fmaxf(__uint_as_float(x[0]), __uint_as_float(x[1])), fmaxf(__uint_as_float(x[2]), __uint_as_float(x[3]))),
                          fmaxf(fmaxf(__uint_as_float(x[4]), __uint_as_float(x[5])), fmaxf(__uint_as_float(x[6]), __uint_as_float(x[7]))));
          }
          const float mxr = fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3]));
          const float2 mm = *reinterpret_cast<const float2*>(cn + BN + 2 * c);
          const bool fire = mxr * (mxr >= 0.f ? mm.x : mm.y) > thr || !(mm.x < INFINITY) || TSIM_KNOB_DEV(a.hot_scaled);
          // Exact path, entered by the whole warp when any lane may have a candidate; rare once the lists are warm.
          if (__any_sync(0xffffffffu, fire)) {
            const float4* cn4 = (const float4*)(cn + c * 32);
            float sc[32];
            float mx = -INFINITY;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float4 w0 = cn4[2 * g], w1 = cn4[2 * g + 1];
              float* x = sc + g * 8;
              x[0] = __uint_as_float(v[g * 8 + 0]) * w0.x; x[1] = __uint_as_float(v[g * 8 + 1]) * w0.y;
              x[2] = __uint_as_float(v[g * 8 + 2]) * w0.z; x[3] = __uint_as_float(v[g * 8 + 3]) * w0.w;
              x[4] = __uint_as_float(v[g * 8 + 4]) * w1.x; x[5] = __uint_as_float(v[g * 8 + 5]) * w1.y;
              x[6] = __uint_as_float(v[g * 8 + 6]) * w1.z; x[7] = __uint_as_float(v[g * 8 + 7]) * w1.w;
              mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])), fmaxf(fmaxf(x[4], x[5]), fmaxf(x[6], x[7]))));
            }
            if (__any_sync(0xffffffffu, mx > thr))
              thr = list.slow(sc, thr, trow0 + c * 32, self_row, ncols - c * 32, thr_g, lad);
          }
        };
        uint32_t va[32], vb[32];
        if (c0 < nchunks) tc_ld32(tbase + c0 * 32, va);
#pragma unroll 1
        for (int c = c0; c < nchunks; c += 2) {
          tc_ld_wait_on(va);
          if (c + 1 < nchunks) tc_ld32(tbase + (c + 1) * 32, vb);
          examine(va, c);
          if (c + 1 < nchunks) {
            tc_ld_wait_on(vb);
            if (c + 2 < nchunks) tc_ld32(tbase + (c + 2) * 32, va);
            examine(vb, c + 1);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_cluster(map_to_cta(smem_u32(&tempty_bar[acc]), 0));   // the leader's MMA thread waits
          else mbar_arrive(smem_u32(&tempty_bar[acc]));
        }
        if (++acc == 2) { acc = 0; aphase ^= 1; }
        // ladder: publish this tile's insertions, pick up everybody else's (after the accumulator
        // stage has been handed back: the MMA of the tile after next does not wait for this)
        list.tile_end();     // append mode: this tile's staged keys -> the query's global list
        {
          const float t0 = thr;
          thr = lad.flush(thr, (uint32_t)(KP > 0 ? KP : a.KP));
          if (thr > t0) atomicMax(thr_g, f32_to_ord(thr));
        }
      }
      // flush this unit's list
      if (qvalid && KP > 0 && (un.split || part == 0))
        list.flush(a.cand + ((size_t)qg * a.NC + un.slot + (un.split ? part : 0)) * KP);
    }
  }

  tc_fence_before();
  // a pair stays alive together: the leader's MMAs read the peer's shared memory until the last commit
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == kWarpAlloc) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D row-major [rows, D] tensor map (bf16 or 1-byte e4m3) with a [box_rows, 128 bytes] box and 128-byte swizzle
int make_map(CUtensorMap* m, const void* base, int64_t rows, int64_t D, int64_t stride_elems, int box_rows, int esz) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return TSIM_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)stride_elems * esz};
  cuuint32_t box[2] = {(cuuint32_t)(BK_BYTES / esz), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return TSIM_ERR_CUDA; }
  count_map_encode();
  return TSIM_OK;
}

template <int KP, int STAGES, bool PAIR, bool FP8>
int launch_cfg(const CUtensorMap& mq, const CUtensorMap& mc, const TcArgs& a, cudaStream_t st) {
  size_t smem = 1024 + (size_t)STAGES * Cfg<PAIR>::STAGE_BYTES + (size_t)(KP > 0 ? KP : -KP) * kEpiThreads * 8 + 2 * 4 * kCnStride * 4 +
                kEpiThreads * 4 + (2 * STAGES + 4) * 8 + 16;
  auto kern = search_tc_kernel<KP, STAGES, PAIR, FP8>;
  TSIM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int sms = device_sm_count();
  const int workers = PAIR ? sms / 2 : sms;
  int64_t nwork = a.sticky ? (int64_t)a.Gq * a.QB : (a.n_units < workers ? a.n_units : workers);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(PAIR ? 2 * nwork : nwork));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[3];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // see pdl_wait() in the kernel
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = knob_on("TSIM_NO_PDL") ? 1 : 2;
  if (a.fused) {
    // the grid barriers of a fused pass need every CTA resident at once: a cooperative launch is gang-scheduled
    // (two such searches on two streams cannot interleave their CTAs and dead-lock each other)
    attr[cfg.numAttrs].id = cudaLaunchAttributeCooperative;
    attr[cfg.numAttrs].val.cooperative = 1;
    ++cfg.numAttrs;
  }
  TSIM_CUDA(cudaLaunchKernelEx(&cfg, kern, mq, mc, a));
  count_launch();
  return TSIM_OK;
}

}  // namespace

// Descriptor cache of a plan handle: a handful of (array, box) combinations per plan (queries or their padded
// copy, the corpus with the lone-CTA and the pair box, the retry stage's compact query block), least recently
// used entry replaced.
struct MapCache {
  struct Entry { const void* base; int64_t rows, D, stride; int box_rows, esz; uint64_t tick; CUtensorMap map; };
  static constexpr int kCap = 16;
  Entry e[kCap];
  int n = 0;
  uint64_t tick = 0;
};
MapCache* map_cache_create() { return new MapCache(); }
void map_cache_destroy(MapCache* c) { delete c; }

int get_tensor_map(MapCache* c, CUtensorMap_st* m, const void* base, int64_t rows, int64_t D, int64_t stride, int box_rows,
                   int esz) {
  if (!c) return make_map(m, base, rows, D, stride, box_rows, esz);
  ++c->tick;
  for (int i = 0; i < c->n; ++i) {
    MapCache::Entry& en = c->e[i];
    if (en.base == base && en.rows == rows && en.D == D && en.stride == stride && en.box_rows == box_rows && en.esz == esz) {
      en.tick = c->tick;
      *m = en.map;
      return TSIM_OK;
    }
  }
  const int rc = make_map(m, base, rows, D, stride, box_rows, esz);
  if (rc) return rc;
  int slot = c->n;
  if (c->n < MapCache::kCap) ++c->n;
  else {
    slot = 0;
    for (int i = 1; i < MapCache::kCap; ++i) if (c->e[i].tick < c->e[slot].tick) slot = i;
  }
  c->e[slot] = {base, rows, D, stride, box_rows, esz, c->tick, *m};
  return TSIM_OK;
}

int launch_search_tc(const void* q, int64_t q_stride, const void* corpus, int64_t c_stride, int dt,
                     const float* c_inv, int64_t Q, int64_t N, int64_t D, int self_on, int64_t self_off,
                     const SearchPlan& p, int pass, uint64_t* cand, uint32_t* thr, uint32_t* ladder, uint64_t* sched,
                     cudaStream_t st, MapCache* maps, const int32_t* q_count, const int32_t* q_map, int q_skip,
                     uint64_t* app_keys, uint32_t* app_cnt, uint32_t* gbar) {
  CUtensorMap mq, mc;
  const int qrows = p.pair ? 2 * BM : BM;
  // q holds QB * qrows rows (the API pads the last query block with zero rows)
  const int esz = dt == TSIM_E4M3 ? 1 : 2;
  int rc = get_tensor_map(maps, &mq, q, (int64_t)p.QB * qrows, D, q_stride, BM, esz);
  if (rc) return rc;
  rc = get_tensor_map(maps, &mc, corpus, N, p.ksplit ? 2 * (int64_t)p.ksplit * 64 : D, c_stride, p.pair ? BN / 2 : BN, esz);
  if (rc) return rc;
  TcArgs a;
  a.c_inv = c_inv; a.Q = Q; a.N = N;
  a.kblocks = (int)((D * esz + BK_BYTES - 1) / BK_BYTES);
  a.ksplit = p.ksplit;
  const int64_t T = (N + BN - 1) / BN;
  a.QB = p.QB; a.sticky = p.sticky; a.Gq = p.Gq; a.tpc = (int)(p.R / BN);
  a.NC = p.NC;
  a.tile_mode = 0; a.tile_stride = 1; a.tile_mult = 2; a.slot_base = 0; a.T = (int)T;
  switch (pass) {
    case TC_PASS_SAMPLE:        // the whole sample, cold
      a.tile_mode = 1; a.tile_stride = (int)p.boot_stride; a.T = (int)p.boot_tiles;
      if (!p.sticky) a.tpc = (int)p.boot_tpc;
      break;
    case TC_PASS_MINI:          // every mini_mult-th sample tile, one-tile units, cold
      a.tile_mode = 1; a.tile_stride = (int)(p.boot_stride * p.mini_mult); a.T = (int)p.mini_tiles; a.tpc = 1;
      break;
    case TC_PASS_SAMPLE_REST:   // the other sample tiles
      a.tile_mode = 3; a.tile_stride = (int)p.boot_stride; a.tile_mult = (int)p.mini_mult; a.T = (int)(p.boot_tiles - p.mini_tiles);
      a.tpc = (int)p.boot_tpc; a.slot_base = (int)p.mini_slots;
      break;
    case TC_PASS_FUSED:         // sticky: the worker's sample tiles, in-kernel thresholds, then its main tiles
      a.tile_mode = 1; a.tile_stride = (int)p.boot_stride; a.T = (int)p.boot_tiles;
      break;
    case TC_PASS_MAIN:          // everything that is not a sample tile
      a.tile_mode = 2; a.tile_stride = (int)p.boot_stride; a.T = (int)(T - p.boot_tiles);
      a.slot_base = (int)(p.mini_slots + p.boot_slots);
      break;
    default: break;
  }
  a.qrep = p.qrep > 1 ? p.qrep : 1;
  a.fused = pass == TC_PASS_FUSED ? 1 : 0;
  a.T2 = a.fused ? (int)(T - p.boot_tiles) : 0;
  a.slot_base2 = a.fused ? (int)(p.mini_slots + p.boot_slots) : 0;
  a.gbar = gbar;
  a.n_units = p.QB * ((a.T + a.tpc - 1) / (a.tpc > 0 ? a.tpc : 1));
  a.self_on = self_on; a.self_off = self_off;
  a.cand = cand; a.thr = thr; a.ladder = ladder;
  a.q_count = q_count; a.q_map = q_map; a.q_skip = q_skip;
  const bool append = p.append != 0;   // every pass of an append plan (mini sample, sample, main) appends
  a.app_keys = append ? app_keys : nullptr; a.app_cnt = append ? app_cnt : nullptr; a.app_cap = p.app_cap; a.KP = p.KP;
  // every tcgen05 launch of a call claims from its own zeroed area (mini sample | sample | main)
  const int area = pass == TC_PASS_MINI ? 0 : (pass == TC_PASS_SAMPLE || pass == TC_PASS_SAMPLE_REST) ? 1 : 2;
  // experiment knobs (compiled out of the release library): static round-robin dealing, diagnosis bits, warp roles
  a.sched = (sched && !p.sticky && !knob_on("TSIM_STATIC_UNITS")) ? sched + (size_t)area * (p.sched_area / 8) : nullptr;
  a.dbg = knob_int("TSIM_DEBUG", 0);
  a.roles_low = knob_on("TSIM_ROLES_LOW") ? 1 : 0;
  a.hot_scaled = knob_on("TSIM_HOT_SCALED") ? 1 : 0;
#define TSIM_DISPATCH(FP8)                                                       \
  if (append) {                                                                  \
    if (pass != TC_PASS_MAIN) {   /* sample passes: loose thresholds, deep staging */  \
      if (p.pair) return launch_cfg<-kAppStageSample, 4, true, FP8>(mq, mc, a, st);      \
      return launch_cfg<-kAppStageSample, 3, false, FP8>(mq, mc, a, st);                 \
    }                                                                            \
    if (p.pair) return launch_cfg<-kAppStageMain, 6, true, FP8>(mq, mc, a, st);          \
    return launch_cfg<-kAppStageMain, 4, false, FP8>(mq, mc, a, st);                     \
  }                                                                              \
  if (p.pair) {                                                                  \
    switch (p.KP) {                                                              \
      case 16: return launch_cfg<16, 6, true, FP8>(mq, mc, a, st);               \
      case 32: return launch_cfg<32, 5, true, FP8>(mq, mc, a, st);               \
      case 64: return launch_cfg<64, 4, true, FP8>(mq, mc, a, st);               \
      case 112: return launch_cfg<112, 3, true, FP8>(mq, mc, a, st);             \
    }                                                                            \
  } else {                                                                       \
    switch (p.KP) {                                                              \
      case 16: return launch_cfg<16, 4, false, FP8>(mq, mc, a, st);              \
      case 32: return launch_cfg<32, 3, false, FP8>(mq, mc, a, st);              \
      case 64: return launch_cfg<64, 3, false, FP8>(mq, mc, a, st);              \
      case 112: return launch_cfg<112, 2, false, FP8>(mq, mc, a, st);            \
    }                                                                            \
  }
  if (dt == TSIM_E4M3) { TSIM_DISPATCH(true) } else { TSIM_DISPATCH(false) }
#undef TSIM_DISPATCH
  set_error("search_tc: bad KP %d", p.KP);
  return TSIM_ERR_INVALID_ARG;
}

}  // namespace tsim
