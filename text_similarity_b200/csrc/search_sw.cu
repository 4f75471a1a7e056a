// K2s (small-batch tensor-core pass): the candidate pass for calls of at most 32 queries -- BASELINE config 4's
// regime (100M x 384 e4m3 over 8 GPUs, 1 ... 32 queries per call), where a search streams the shard from HBM once.
//
// Same job as search_tc.cu (replaces the per-query F.cosine_similarity + torch.topk loop of
// SentenceMiningPipeline._search, reference src/pipeline/search_pipeline.py:73-79), with the MMA roles SWAPPED:
// corpus rows are the M side (128 per tile = the 128 TMEM lanes), the <= 32 queries the N side (32 fp32 TMEM columns
// per accumulator stage), so one tile costs 128 x 32 x D multiply-adds instead of the 128 x 256 x D that
// search_tc.cu spends when 96-127 of its 128 query rows are padding.
//
// Why it exists: power.  Measured on B200 (scripts/ridge_probe.py, profiles/r02b_power_cap_probe.txt): with the
// queries on the M side an e4m3 12.5M x 384 shard search at Q = 1 holds the board at its 1 kW software power cap --
// the padded MMAs alone draw ~450 W -- the SM clock sinks to 0.85-0.98 GHz and the search takes 0.870 ms; with the
// MMAs switched off (diagnosis build) the same TMA stream runs at 1.55 GHz and 0.706 ms = 6.87 TB/s.  An HBM-bound
// kernel has no business spending half a kilowatt on multiplying zeros.
//
// Structure (one persistent CTA per SM, 256 threads: warps 0-3 epilogue, warp 4 TMA, warp 5 MMA, warp 6 TMEM):
//  * the (zero-padded) 32-query block is loaded ONCE into shared memory, all k-blocks of it ([kblocks][32 rows x 128 B],
//    128-byte swizzle): it is the B operand of every MMA of the CTA's life;
//  * everything else of shared memory is a ring of corpus k-blocks ([128 rows x 128 B] = 16 KB per stage, 10-13
//    stages = 160-208 KB in flight per SM), the A operand;
//  * eight accumulator stages (all 512 TMEM columns): the TMA and MMA warps start streaming the corpus the moment the
//    CTA is resident and run up to eight tiles ahead, while the epilogue warps still wait (griddepcontrol.wait) for
//    the thresholds the previous kernel is making.  A stage is TWO partial accumulators of 32 columns, for the even
//    and the odd 32-byte K sub-steps: an N = 32 MMA is 16 clocks of tensor work but its result is ~180 clocks away,
//    and the 4 x kblocks MMAs of a tile on ONE accumulator are one dependent chain -- 48 x 180 clocks = 4.6 us per
//    tile of 768-wide bf16 rows, more than the tile's 4.1 us of HBM time (measured: 6.6 TB/s).  Two interleaved
//    chains halve that; the epilogue adds the two partial sums;
//  * epilogue thread i owns TMEM lane i = corpus row i of the tile: one tcgen05.ld.32x32b.x32 brings its row's 32
//    scores, one FMUL each by the row's inverse norm, one compare each with the queries' thresholds (a warp-private
//    shared-memory copy, refreshed per tile); the accumulator stage is handed back BEFORE the (rare) survivors are
//    appended to their query's global list -- the append lists, ladder counters and thresholds are the ones of
//    search_tc.cu's append mode, read by tighten_kernel / select_rescore (select_merge.cu).
//  * Sample pass (SAMPLE = true, <= 24 CTAs): every row of the strided sample tiles becomes a candidate, written
//    as 16-entry lists (cand[q][8 lists per worker][16]); tighten_kernel turns them into each query's starting
//    threshold + ladder; the main pass scans every other tile.
// Completeness: a row is dropped only by score <= thr[q], and thr[q] is always the KP-th best of rows that reach
// select_rescore or a ladder level with >= KP appended rows at or above it -- the argument of DESIGN.md section 2.
//
// Roofline: N * D * e bytes from HBM (the corpus, once); the MMAs are 1/8 of search_tc.cu's.
#include <cuda.h>
#include <stdlib.h>

#include "tsim_common.cuh"
#include "tc_ptx.cuh"

namespace tsim {
namespace {

constexpr int SW_ROWS = 128;       // corpus rows per tile (TMEM lanes, MMA M)
constexpr int SW_NQ = 32;          // query columns per accumulator stage (MMA N)
constexpr int SW_KSPLIT = 2;       // partial accumulators per stage (see below)
constexpr int SW_ACC_COLS = SW_KSPLIT * SW_NQ;       // TMEM columns per accumulator stage
constexpr int SW_ACC = 512 / SW_ACC_COLS;            // accumulator stages (tiles in flight between the MMA thread and the epilogue)
constexpr int SW_THREADS = 256;
constexpr int SW_MAX_STAGES = 16;
constexpr int SW_A_BYTES = SW_ROWS * BK_BYTES;   // 16 KB corpus k-block
constexpr int SW_Q_BYTES = SW_NQ * BK_BYTES;     // 4 KB query k-block
constexpr int SW_PEND = 64;        // survivors a warp parks in shared memory between two flushes to the global lists
constexpr int SW_RETIGHTEN = 1024; // a query's append list reaching a multiple of this gets its threshold re-made from its
constexpr int SW_RT_KEYS = 512;    // ... most recent SW_RT_KEYS entries

struct SwArgs {
  const float* c_inv;     // [N] inverse norms of the stored corpus rows
  int64_t N; int Q;
  int kblocks;            // ceil(D * element size / 128)
  int stages;             // corpus ring depth
  int T;                  // corpus tiles of 128 rows
  int ns, stride;         // sample tiles: i * stride, i < ns
  int self_on; int64_t self_off;
  uint64_t* cand;         // sample pass: [Q][NC][16] packed keys
  int64_t NC;
  uint32_t* thr;          // [Q] ordered-float thresholds
  uint32_t* ladder;       // [Q][2 * kLadder]
  uint64_t* app_keys; uint32_t* app_cnt; int app_cap;   // per-query append lists
  int KP;
  int dbg;                // TSIM_DEBUG bits (experiment build): 2 skip MMAs, 4 skip the epilogue's compares
};

// A query's threshold follows its ladder (tsim_common.cuh): raise thr[q] to the highest level that has >= KP appended
// rows at or above it -- valid, because every one of them reaches select_rescore.  One lane per query; counts read a
// little stale only delay a raise.
__device__ __forceinline__ void sw_ladder_raise(const SwArgs& a, int q, float base, float step) {
  const uint32_t* cnt = a.ladder + (size_t)q * (2 * kLadder) + kLadder;
  uint32_t c = 0;
  int best = -1;
#pragma unroll
  for (int i = kLadder / 4 - 1; i >= 0; --i) {
    const uint4 b = __ldcg(reinterpret_cast<const uint4*>(cnt) + i);
    const uint32_t w[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int e = 3; e >= 0; --e) {
      c += w[e];
      if (best < 0 && c >= (uint32_t)a.KP) best = 4 * i + e;
    }
  }
  if (best >= 0) atomicMax(a.thr + q, f32_to_ord(ladder_value(base, step, best)));
}

// The KP-th best ordered score among n <= 512 keys (0 = empty slot), by ONE warp: every lane pulls its 16 keys into
// registers with independent loads (one L2 round trip), then an MSB-first radix select of 4 passes x 8 bits runs on the
// registers (hist: 256 words of shared memory private to the warp).  0: fewer than KP keys.
__device__ __noinline__ uint32_t sw_kth_best(const uint64_t* src, uint32_t n, int KP, uint32_t* hist) {
  const int lane = threadIdx.x & 31;
  uint32_t sc[SW_RT_KEYS / 32];
  uint32_t live = 0;
#pragma unroll
  for (int i = 0; i < SW_RT_KEYS / 32; ++i) {
    const uint32_t idx = (uint32_t)lane + 32u * i;
    sc[i] = idx < n ? (uint32_t)(__ldcg(src + idx) >> 32) : 0u;
  }
#pragma unroll
  for (int i = 0; i < SW_RT_KEYS / 32; ++i) live += sc[i] != 0u;
  live = __reduce_add_sync(0xffffffffu, live);
  if (live < (uint32_t)KP) return 0u;
  uint32_t prefix = 0, need = (uint32_t)KP;
  for (int shift = 24; shift >= 0; shift -= 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) hist[lane + 32 * i] = 0u;
    __syncwarp();
    const uint32_t hi_mask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
#pragma unroll
    for (int i = 0; i < SW_RT_KEYS / 32; ++i)
      if (sc[i] != 0u && (sc[i] & hi_mask) == prefix) atomicAdd(&hist[(sc[i] >> shift) & 255u], 1u);
    __syncwarp();
    // lane l owns buckets 255 - 8l .. 248 - 8l (descending): where does the running count reach `need`?
    uint32_t c[8], tot = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j] = hist[255 - 8 * lane - j]; tot += c[j]; }
    uint32_t incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const uint32_t before = incl - tot;
    const bool mine = before < need && incl >= need;
    uint32_t bucket = 0, rest = 0;
    if (mine) {
      uint32_t run = before;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (run < need && run + c[j] >= need) { bucket = 255u - 8u * lane - j; rest = need - run; }
        run += c[j];
      }
    }
    const int owner = __ffs(__ballot_sync(0xffffffffu, mine)) - 1;     // exactly one lane (live >= need)
    bucket = __shfl_sync(0xffffffffu, bucket, owner);
    need = __shfl_sync(0xffffffffu, rest, owner);
    prefix |= bucket << shift;
    __syncwarp();
  }
  return prefix;
}

// the u-th tile of worker `w` (of `nw`): sample pass -> sample tile w; main pass -> w, w + nw, ... skipping sample tiles
__device__ __forceinline__ bool sw_is_sample(const SwArgs& a, int tile) {
  return a.ns > 0 && tile % a.stride == 0 && tile / a.stride < a.ns;
}

template <bool FP8, bool SAMPLE>
__global__ void __launch_bounds__(SW_THREADS, 1)
search_sw_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c, SwArgs a) {
  extern __shared__ unsigned char smem_raw[];
  // [kblocks] query k-blocks 4K | [stages] corpus k-blocks 16K | thresholds [4 warps][32] | ladders [32] | parked
  // survivors [4 warps][64] | barriers | tmem ptr
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int BK = FP8 ? BK_BYTES : BK_BYTES / 2;     // elements per k-block
  // UMMA instruction descriptor: D = f32, A / B = bf16 (kind::f16) or e4m3 (kind::f8f6f4), K-major, N = 32, M = 128
  constexpr uint32_t IDESC = (1u << 4) | (FP8 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(SW_NQ >> 3) << 17) |
                             ((uint32_t)(SW_ROWS >> 4) << 24);
  unsigned char* qtiles = smem;
  unsigned char* ring = smem + (((size_t)a.kblocks * SW_Q_BYTES + 1023) & ~(size_t)1023);
  float* thr_s = reinterpret_cast<float*>(ring + (size_t)a.stages * SW_A_BYTES);   // [4][32]
  float4* lad_s = reinterpret_cast<float4*>(thr_s + 4 * SW_NQ);                    // [32] ladder (base, step, 1 / step) per query
  uint64_t* pend_k = reinterpret_cast<uint64_t*>(lad_s + SW_NQ);                   // [4 warps][SW_PEND] parked survivors: key
  int* pend_q = reinterpret_cast<int*>(pend_k + 4 * SW_PEND);                      //                                  ... query
  uint32_t* rt_hist = reinterpret_cast<uint32_t*>(pend_q + 4 * SW_PEND);           // [4 warps][256] radix histogram (re-tighten)
  uint64_t* bars = reinterpret_cast<uint64_t*>(rt_hist + 4 * 256);
  uint64_t* full_bar = bars;                       // [stages]  TMA -> MMA
  uint64_t* empty_bar = bars + SW_MAX_STAGES;      // [stages]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * SW_MAX_STAGES;  // [SW_ACC]  MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + SW_ACC;       // [SW_ACC]  epilogue -> MMA
  uint64_t* q_bar = tempty_bar + SW_ACC;           // [1]       the query block has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dbg = TSIM_KNOB_DEV(a.dbg);
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_c);
  }
  if (warp == 5 && lane == 0) {
    for (int s = 0; s < a.stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int s = 0; s < SW_ACC; ++s) { mbar_init(smem_u32(&tfull_bar[s]), 1); mbar_init(smem_u32(&tempty_bar[s]), 4); }
    mbar_init(smem_u32(q_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 6) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(SW_ACC * SW_ACC_COLS)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch.  The sample pass reads what search_prep wrote (the padded query block) and lets its
  // successors start only after it has waited for that kernel; the main pass can therefore load the query block and
  // stream the corpus at once (its grid cannot start before the threshold kernel's, which cannot start before every
  // sample CTA is past its wait) -- only its epilogue warps wait, for the thresholds, ladders and append counters.
  if (SAMPLE) { pdl_wait(); pdl_trigger(); }
  else { pdl_trigger(); if (warp < 4) pdl_wait(); }

  const int w = (int)blockIdx.x, nw = (int)gridDim.x;
  // this worker's tiles, the same sequence in every role
  auto first_tile = [&]() { return SAMPLE ? w * a.stride : w; };
  auto next_tile = [&](int t) { return SAMPLE ? a.T : t + nw; };     // (a sample worker scans one tile)
  auto skip = [&](int t) { return !SAMPLE && sw_is_sample(a, t); };

  // The two single-thread roles run WARP-UNIFORM: all 32 lanes walk the loop and wait on the barriers, one elected lane
  // issues the TMA / MMA / commit.  Written as `if (lane == 0) { whole loop }` every operand of the tcgen05 / TMA
  // instructions (uniform registers) went through ELECT + R2UR.BROADCAST sequences inside divergent code, and the MMA
  // thread's loop body -- ~700 clocks per k-block, with no wait in it -- paced the kernel (ncu: 6.0 TB/s, tensor pipe
  // 9 % busy, the epilogue waiting for accumulators).
  if (warp == 4) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(smem_u32(q_bar), (uint32_t)a.kblocks * SW_Q_BYTES);
      for (int kb = 0; kb < a.kblocks; ++kb) tma_load_2d(smem_u32(qtiles + (size_t)kb * SW_Q_BYTES), &tmap_q, smem_u32(q_bar), kb * BK, 0);
    }
    __syncwarp();
    const uint32_t ring0 = smem_u32(ring), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
    int stage = 0; uint32_t phase = 0;
    for (int t = first_tile(); t < a.T; t = next_tile(t)) {
      if (skip(t)) continue;
      for (int kb = 0; kb < a.kblocks; ++kb) {
        mbar_wait(empty0 + 8u * stage, phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(full0 + 8u * stage, SW_A_BYTES);
          tma_load_2d(ring0 + (uint32_t)stage * SW_A_BYTES, &tmap_c, full0 + 8u * stage, kb * BK, t * SW_ROWS);
        }
        __syncwarp();
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    mbar_wait(smem_u32(q_bar), 0);
    tc_fence_after();
    // descriptors advance linearly with the shared-memory address (>> 4): 16 KB per corpus stage, 4 KB per query k-block
    const uint64_t adesc0 = make_umma_desc(smem_u32(ring)), bdesc0 = make_umma_desc(smem_u32(qtiles));
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), tfull0 = smem_u32(tfull_bar), tempty0 = smem_u32(tempty_bar);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t aphase = 0;
    for (int t = first_tile(); t < a.T; t = next_tile(t)) {
      if (skip(t)) continue;
      mbar_wait(tempty0 + 8u * acc, aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * SW_ACC_COLS);
      for (int kb = 0; kb < a.kblocks; ++kb) {
        mbar_wait(full0 + 8u * stage, phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = adesc0 + (uint64_t)(stage * (SW_A_BYTES >> 4));     // corpus rows: M side
          const uint64_t bdesc = bdesc0 + (uint64_t)(kb * (SW_Q_BYTES >> 4));        // queries: N side
#pragma unroll
          for (int k = 0; k < BK_BYTES / UMMA_K_BYTES; ++k) {
            if (dbg & 2) break;
            tc_mma<FP8>(d_tmem + (uint32_t)((k % SW_KSPLIT) * SW_NQ), adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), IDESC,
                        (kb | (k / SW_KSPLIT)) ? 1u : 0u);
          }
          tc_commit(empty0 + 8u * stage);
          if (kb == a.kblocks - 1) tc_commit(tfull0 + 8u * acc);
        }
        __syncwarp();
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
      if (++acc == SW_ACC) { acc = 0; aphase ^= 1; }
    }
  } else if (warp < 4) {
    // ===================== epilogue: thread <-> corpus row =====================
    const int et = threadIdx.x;                                   // 0..127 = TMEM lane = row within the tile
    const uint32_t lane_addr = ((uint32_t)(warp * 32)) << 16;
    float* tw = thr_s + warp * SW_NQ;                             // this warp's copy of the 32 thresholds
    // the queries' ladders (laid out by the threshold kernel; only their counters change during this pass)
    float lbase = INFINITY, lstep = 0.f;
    if (!SAMPLE) {
      uint4 h = make_uint4(__float_as_uint(INFINITY), 0u, 0u, 0u);
      if (lane < a.Q) h = __ldcg(reinterpret_cast<const uint4*>(a.ladder + (size_t)lane * (2 * kLadder)));
      lbase = __uint_as_float(h.x); lstep = __uint_as_float(h.y);
      if (warp == 0) lad_s[lane] = make_float4(lbase, lstep, __uint_as_float(h.z), 0.f);
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    int acc = 0; uint32_t aphase = 0;
    // Tile metadata -- the row's inverse norm and (lane j) query j's threshold -- is fetched kPf tiles ahead: a tile of
    // 384-byte rows lasts ~1 us, about one L2 round trip under load (one tile ahead, the epilogue waited for these
    // loads every tile and paced the whole kernel: 1.2 us per tile)
    constexpr int kPf = 4;
    float inv_q[kPf]; int tile_q[kPf];
    uint32_t thr_q[2] = {0u, 0u};       // thresholds travel two tiles ahead only: a raise should bite soon
    if (!SAMPLE && lane < a.Q) thr_q[0] = thr_q[1] = __ldcg(a.thr + lane);
    auto advance = [&](int t) { t = next_tile(t); while (t < a.T && skip(t)) t = next_tile(t); return t; };
    int tnext = first_tile();
    while (tnext < a.T && skip(tnext)) tnext = next_tile(tnext);
#pragma unroll
    for (int i = 0; i < kPf; ++i) {
      tile_q[i] = tnext; inv_q[i] = 0.f;
      if (tnext < a.T) {
        const int64_t r = (int64_t)tnext * SW_ROWS + et;
        inv_q[i] = r < a.N ? __ldg(a.c_inv + r) : 0.f;
        tnext = advance(tnext);
      }
    }
    // Survivors are parked in shared memory and flushed to their queries' global lists 24+ at a time (or every 8th
    // tile): a flush costs the same two round trips (list position, ladder counters) whether it carries one row or
    // thirty-two, and done per tile those round trips -- ~2 us each time a warp had a survivor -- made a Q = 32 search
    // 45 us slower than a Q = 1 search.
    uint64_t* pk = pend_k + warp * SW_PEND;
    int* pq = pend_q + warp * SW_PEND;
    int npend = 0, tiles_done = 0;
    const int flush_mask = a.kblocks <= 4 ? 7 : 0;
    auto flush = [&]() {
      uint32_t touched = 0, crowded = 0;
      for (int base = 0; base < npend; base += 32) {
        const int i = base + lane;
        if (i < npend) {
          const uint64_t key = pk[i];
          const int j = pq[i];
          const uint32_t pos = atomicAdd(a.app_cnt + j, 1u);
          if (pos < (uint32_t)a.app_cap) a.app_keys[(size_t)j * a.app_cap + pos] = key;
          if ((pos + 1) % SW_RETIGHTEN == 0 && pos < (uint32_t)a.app_cap) crowded |= 1u << j;
          const float4 ld = lad_s[j];
          const int lvl = ladder_level(ld.x, ld.y, ld.z, key_score(key));
          if (lvl >= 0) atomicAdd(a.ladder + (size_t)j * (2 * kLadder) + kLadder + lvl, 1u);
          touched |= 1u << j;
        }
      }
      touched = __reduce_or_sync(0xffffffffu, touched);
      crowded = __reduce_or_sync(0xffffffffu, crowded);
      __syncwarp();
      if ((touched >> lane) & 1u) sw_ladder_raise(a, lane, lbase, lstep);
      // A list that keeps filling although its threshold sits on the ladder's top level -- a dense cluster of rows
      // around the query that the sample could not foresee -- would overflow (4096 entries: the query is flagged and the
      // retry pass answers it, a second scan of the shard).  The warp whose row made the list reach a multiple of 1024
      // re-makes the threshold from the list itself: the KP-th best of its most recent 512 rows -- KP appended rows at or
      // above it exist, so it is valid (slots other CTAs have reserved but not written yet read as 0 = empty -- the
      // prep kernel zeroes the lists of these plans -- which only lowers the result).  At most three times per query per
      // search, a few microseconds of one warp each.
#pragma unroll 1
      while (crowded) {
        const int j = __ffs(crowded) - 1;
        crowded &= crowded - 1;
        const uint32_t cnt = min(__ldcg(a.app_cnt + j), (uint32_t)a.app_cap);
        const uint32_t first = cnt > SW_RT_KEYS ? cnt - SW_RT_KEYS : 0u;
        const uint32_t t = sw_kth_best(a.app_keys + (size_t)j * a.app_cap + first, cnt - first, a.KP, rt_hist + warp * 256);
        if (t && lane == 0) atomicMax(a.thr + j, t);
      }
      npend = 0;
      __syncwarp();
    };
    while (tile_q[0] < a.T) {
      const int t = tile_q[0];
      const int64_t row = (int64_t)t * SW_ROWS + et;
      const float inv = inv_q[0];
      if (!SAMPLE) {
        // a query that has no threshold yet (0) accepts everything; padding columns accept nothing
        tw[lane] = lane < a.Q ? (thr_q[0] ? ord_to_f32(thr_q[0]) : -INFINITY) : INFINITY;
        __syncwarp();
      }
#pragma unroll
      for (int i = 0; i + 1 < kPf; ++i) { tile_q[i] = tile_q[i + 1]; inv_q[i] = inv_q[i + 1]; }
      tile_q[kPf - 1] = tnext; inv_q[kPf - 1] = 0.f;
      thr_q[0] = thr_q[1];
      if (!SAMPLE && lane < a.Q) thr_q[1] = __ldcg(a.thr + lane);
      if (tnext < a.T) {
        const int64_t r = (int64_t)tnext * SW_ROWS + et;
        inv_q[kPf - 1] = r < a.N ? __ldg(a.c_inv + r) : 0.f;
        tnext = advance(tnext);
      }
      mbar_wait(smem_u32(&tfull_bar[acc]), aphase);
      tc_fence_after();
      uint32_t v[32], u[32];
      tc_ld32(tmem_base + lane_addr + (uint32_t)(acc * SW_ACC_COLS), v);
      tc_ld32(tmem_base + lane_addr + (uint32_t)(acc * SW_ACC_COLS + SW_NQ), u);
      tc_ld_wait_on(v);
      tc_ld_wait_on(u);
      // the accumulator stage goes back at once: the scores are in registers
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
      if (++acc == SW_ACC) { acc = 0; aphase ^= 1; }
      float sc[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) sc[j] = (__uint_as_float(v[j]) + __uint_as_float(u[j])) * inv;
      const bool live = row < a.N && !(dbg & 4);
      // self exclusion: query j's own row is self_off + j
      const int64_t jself = a.self_on ? row - a.self_off : -1;
      if (SAMPLE) {
        // every row of a sample tile is a candidate: lane l of warp w fills entry l % 16 of list 8 * worker + 2 * w + l / 16
        const int64_t list = (int64_t)w * 8 + warp * 2 + (lane >> 4);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (j < a.Q) {
            const bool ok = live && jself != j && sc[j] == sc[j] && sc[j] > -INFINITY;   // NaN rows are never returned
            a.cand[((size_t)j * a.NC + list) * 16 + (lane & 15)] = ok ? pack_key(sc[j], (uint32_t)row) : 0ull;
          }
        }
      } else {
        uint32_t m = 0;
        const float4* t4 = reinterpret_cast<const float4*>(tw);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 th = t4[g];
          m |= (sc[4 * g] > th.x ? 1u : 0u) << (4 * g);
          m |= (sc[4 * g + 1] > th.y ? 1u : 0u) << (4 * g + 1);
          m |= (sc[4 * g + 2] > th.z ? 1u : 0u) << (4 * g + 2);
          m |= (sc[4 * g + 3] > th.w ? 1u : 0u) << (4 * g + 3);
        }
        if (!live) m = 0;
        if (jself >= 0 && jself < 32) m &= ~(1u << (int)jself);
        // Cold path (a few hundred rows per query per search): park the survivors (one per lane per round)
#pragma unroll 1
        while (__any_sync(0xffffffffu, m != 0)) {
          const bool has = m != 0;
          const int j = has ? __ffs(m) - 1 : 0;
          m &= m - 1;                                     // (0 stays 0)
          const float s = select32(sc, j);
          const bool ok = has && s < INFINITY;
          const uint32_t bal = __ballot_sync(0xffffffffu, ok);
          if (ok) {
            const int at = npend + __popc(bal & ((1u << lane) - 1u));
            pk[at] = pack_key(s, (uint32_t)row);
            pq[at] = j;
          }
          npend += __popc(bal);
          __syncwarp();
          if (npend > SW_PEND - 32) flush();
        }
        ++tiles_done;
        // (a tile of 1536-byte rows lasts four times as long as one of 384-byte rows: flush sooner there -- measured
        // on 10M x 768 bf16, Q = 32: every tile 2.43 ms, every 8th 2.49 ms; on 12.5M x 384 e4m3: 0.808 vs 0.794 ms)
        if (npend >= 24 || (npend > 0 && ((tiles_done & flush_mask) == 0 || (dbg & 8)))) flush();   // (dbg 8: every tile)
        __syncwarp();     // everybody has read tw before the next tile overwrites it
      }
    }
    if (!SAMPLE && npend > 0) flush();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 6) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(SW_ACC * SW_ACC_COLS)) : "memory");
  }
}

// Between the sample pass and the main pass, one block of 512 threads per query (six sample keys a thread): the KP-th best of the query's
// 192 x 16 sample keys becomes its threshold; a ladder of 16 levels is laid out above it with a step taken from the
// sample's tail slope (epi_tighten, tail_step): the main pass appends rows instead of keeping lists, so the ladder is
// the ONLY thing that raises a threshold and it has to reach the corpus's own KP-th best -- with search_tc.cu's steps of
// an eighth of (sample best - threshold) one query in ten ran out of ladder and overflowed its append list.  The
// sample keys at or above the threshold move to the append list, so that select_rescore reads one short list per query.
constexpr int SW_TT = 512, SW_TK = 6;   // sw_tighten: threads per query, sample keys per thread
__global__ void __launch_bounds__(SW_TT) sw_tighten_kernel(const uint64_t* cand, int64_t NC, int KP, uint32_t* thr,
                                                         uint32_t* ladder, uint64_t* app_keys, uint32_t* app_cnt, int app_cap) {
  __shared__ uint32_t hist[544];
  pdl_trigger();
  pdl_wait();
  const int64_t q = blockIdx.x;
  epi_tighten<SW_TT, SW_TK>(cand + (size_t)q * NC * 16, (uint32_t)(NC * 16), KP, thr + q, ladder + (size_t)q * (2 * kLadder), hist,
              (int)threadIdx.x, 0.25f, app_keys + (size_t)q * app_cap, app_cnt + q, app_cap, /*tail_step=*/true);
}

size_t sw_smem_bytes(int kblocks, int stages) {
  return 1024 + (((size_t)kblocks * SW_Q_BYTES + 1023) & ~(size_t)1023) + (size_t)stages * SW_A_BYTES + 4 * SW_NQ * 4 +
         SW_NQ * 16 + 4 * SW_PEND * 12 + 4 * 256 * 4 + (2 * SW_MAX_STAGES + 2 * SW_ACC + 1) * 8 + 16;
}

template <bool FP8, bool SAMPLE>
int sw_launch(const CUtensorMap& mq, const CUtensorMap& mc, const SwArgs& a, int grid, cudaStream_t st) {
  const size_t smem = sw_smem_bytes(a.kblocks, a.stages);
  auto kern = search_sw_kernel<FP8, SAMPLE>;
  TSIM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(SW_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = knob_on("TSIM_NO_PDL") ? 0 : 1;
  TSIM_CUDA(cudaLaunchKernelEx(&cfg, kern, mq, mc, a));
  count_launch();
  return TSIM_OK;
}

}  // namespace

// Corpus ring depth for a query block of `kblocks` k-blocks (0: the block does not fit beside a useful ring)
int search_sw_stages(int kblocks) {
  int s = SW_MAX_STAGES;
  while (s >= 6 && sw_smem_bytes(kblocks, s) > 232448) --s;
  return s >= 6 ? s : 0;
}

int launch_sw_tighten(int64_t Q, const SearchPlan& p, const uint64_t* cand, uint32_t* thr, uint32_t* ladder,
                      uint64_t* app_keys, uint32_t* app_cnt, cudaStream_t st) {
  static_assert(8 * 24 * 16 <= SW_TT * SW_TK, "sample keys per query exceed what epi_tighten holds in registers");
  TSIM_CUDA(launch_pdl(sw_tighten_kernel, dim3((unsigned)Q), dim3(SW_TT), 0, st, cand, p.NC, p.KP, thr, ladder, app_keys,
                       app_cnt, p.app_cap));
  count_launch();
  return TSIM_OK;
}

// sample == 1: the sample pass (p.sw_ns CTAs, lists into cand); 0: the main pass (appends)
int launch_search_sw(const void* q, int64_t q_stride, const void* corpus, int64_t c_stride, int dt, const float* c_inv,
                     int64_t Q, int64_t N, int64_t D, int self_on, int64_t self_off, const SearchPlan& p, int sample,
                     uint64_t* cand, uint32_t* thr, uint32_t* ladder, uint64_t* app_keys, uint32_t* app_cnt,
                     cudaStream_t st, MapCache* maps) {
  const int esz = dt == TSIM_E4M3 ? 1 : 2;
  CUtensorMap mq, mc;
  int rc = get_tensor_map(maps, &mq, q, SW_NQ, D, q_stride, SW_NQ, esz);     // the padded block: 32 rows
  if (rc) return rc;
  rc = get_tensor_map(maps, &mc, corpus, N, D, c_stride, SW_ROWS, esz);
  if (rc) return rc;
  SwArgs a;
  a.c_inv = c_inv; a.N = N; a.Q = (int)Q;
  a.kblocks = (int)((D * esz + BK_BYTES - 1) / BK_BYTES);
  a.stages = search_sw_stages(a.kblocks);
  a.T = (int)((N + SW_ROWS - 1) / SW_ROWS);
  a.ns = p.sw_ns; a.stride = p.sw_stride;
  a.self_on = self_on; a.self_off = self_off;
  a.cand = cand; a.NC = p.NC; a.thr = thr; a.ladder = ladder;
  a.app_keys = app_keys; a.app_cnt = app_cnt; a.app_cap = p.app_cap; a.KP = p.KP;
  a.dbg = knob_int("TSIM_DEBUG", 0);
  if (a.stages == 0) { set_error("search_sw: query block of %d k-blocks does not fit", a.kblocks); return TSIM_ERR_UNSUPPORTED; }
  const int sms = device_sm_count();
  if (sample) {
    if (dt == TSIM_E4M3) return sw_launch<true, true>(mq, mc, a, p.sw_ns, st);
    return sw_launch<false, true>(mq, mc, a, p.sw_ns, st);
  }
  const int grid = a.T < sms ? a.T : sms;
  if (dt == TSIM_E4M3) return sw_launch<true, false>(mq, mc, a, grid, st);
  return sw_launch<false, false>(mq, mc, a, grid, st);
}

}  // namespace tsim
