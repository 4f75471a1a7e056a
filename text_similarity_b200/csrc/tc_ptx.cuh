// PTX wrappers shared by the tcgen05 kernels (search_tc.cu, search_sw.cu): mbarriers, TMA tensor loads, tcgen05.mma /
// commit / ld, cluster helpers and the UMMA shared-memory descriptor.  Internal to libtsim; every function is
// __device__ __forceinline__.
#pragma once
#include <cuda.h>
#include <stdint.h>

#include "tsim_common.cuh"

namespace tsim {
namespace {

constexpr int BK_BYTES = 128;  // bytes per row per k-block = one 128-byte swizzle atom row (64 bf16 / 128 e4m3)
constexpr int UMMA_K_BYTES = 32;  // one MMA consumes 32 bytes of K per row (K = 16 bf16 / 32 e4m3)

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// one lane of the (converged) warp, the same one every time
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// 2-CTA variants: the peer's TMA signals the LEADER's barrier; the leader's commit reaches both CTAs
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  // default semantics (release at CTA scope), as CUTLASS's ClusterBarrier::arrive(cta_id): a
  // cluster-scope release would cost a full memory barrier per tile; the TMEM hand-off is ordered
  // by tcgen05.fence::before_thread_sync / after_thread_sync around the barrier
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
template <bool FP8>
__device__ __forceinline__ void tc_mma_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  if constexpr (FP8)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <bool FP8>
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  if constexpr (FP8)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// The same wait, naming the 32 registers of an earlier tc_ld32 as in/out operands: the load is asynchronous, and
// nothing else tells the compiler that uses of v must stay BELOW the wait when another load is issued in between.
__device__ __forceinline__ void tc_ld_wait_on(uint32_t* v) {
  asm volatile(
      "tcgen05.wait::ld.sync.aligned;"
      : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
        "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
        "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
        "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
      :
      : "memory");
}

// UMMA shared-memory descriptor, K-major operand, SWIZZLE_128B, 128-byte rows:
//   start address >> 4 | LBO (unused for swizzled K-major) = 1 | SBO = 1024 B (8 rows) >> 4 |
//   version 1 (sm_100) | layout type 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_umma_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3ffff) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// element j (dynamic) of 32 values held in registers: a 5-level select tree (registers cannot be indexed dynamically)
__device__ __forceinline__ float select32(const float* sc, int j) {
  float t16[16], t8[8], t4[4];
#pragma unroll
  for (int i = 0; i < 16; ++i) t16[i] = (j & 16) ? sc[16 + i] : sc[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) t8[i] = (j & 8) ? t16[8 + i] : t16[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) t4[i] = (j & 4) ? t8[4 + i] : t8[i];
  const float u0 = (j & 2) ? t4[2] : t4[0], u1 = (j & 2) ? t4[3] : t4[1];
  return (j & 1) ? u1 : u0;
}

// Thresholds inside a fused sticky pass, by the 128 epilogue threads of a CTA (named barrier 1) for ONE query: the
// same radix select and ladder as warp_tighten (tsim_common.cuh), but every thread first pulls its share of the
// keys into registers with independent loads -- one L2 round trip -- and the four passes then run on registers.
// (A single warp walking 74 keys per lane with a load -> shared-atomic dependency per key took ~45 us, during
// which every CTA of the launch sat at the grid barrier with HBM idle.)  n <= 128 * kEpiKeys keys; hist: 288 words (544 with tail_step).
// NT threads x NK keys each: 128 x 24 inside the fused pass; the stand-alone kernel of search_sw.cu runs 512 x 6 (the
// same 3072 keys: at 24 keys a thread the ~3100 instructions per warp of this latency chain took 17 us, one warp per
// scheduler).
constexpr int kEpiKeys = 24;
template <int NT>
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }
// ladder_frac: ladder step as a fraction of (sample best - KP-th best): 1/8 puts the sample's best at level 8.
// app_keys / app_cnt (optional): also copy every sample key at or above the new threshold to the query's append list
// (the sample lists then need not be read again by select_rescore).
template <int NT = 128, int NK = kEpiKeys>
__device__ __noinline__ void epi_tighten(const uint64_t* src, uint32_t n, int KP, uint32_t* thr_q, uint32_t* lad,
                                         uint32_t* hist, int et, float ladder_frac = 0.125f,
                                         uint64_t* app_keys = nullptr, uint32_t* app_cnt = nullptr, int app_cap = 0,
                                         bool tail_step = false) {
  uint32_t sc[NK];
#pragma unroll
  for (int i = 0; i < NK; ++i) {
    const uint32_t idx = (uint32_t)et + (uint32_t)NT * i;
    sc[i] = idx < n ? (uint32_t)(__ldcg(src + idx) >> 32) : 0u;
  }
  uint32_t* ctl = hist + 256;      // [0] live keys, [1] best score, [2] bucket, [3] need, [4] [5] same for want2, [6] worst
  if (et < 32) ctl[et] = et == 6 ? 0xffffffffu : 0u;
  epi_bar<NT>();
  uint32_t live = 0, best = 0, worst = 0xffffffffu;
#pragma unroll
  for (int i = 0; i < NK; ++i) {
    live += sc[i] != 0u; best = max(best, sc[i]);
    if (sc[i] != 0u) worst = min(worst, sc[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    live += __shfl_xor_sync(0xffffffffu, live, o);
    best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    worst = min(worst, __shfl_xor_sync(0xffffffffu, worst, o));
  }
  if ((et & 31) == 0) { atomicAdd(&ctl[0], live); atomicMax(&ctl[1], best); atomicMin(&ctl[6], worst); }
  epi_bar<NT>();
  live = ctl[0]; best = ctl[1]; worst = ctl[6];
  if (live < (uint32_t)KP) {       // not enough rows for a threshold: a ladder that never fires
    if (lad && et < kLadder) { lad[kLadder + et] = 0u; lad[et] = et == 0 ? __float_as_uint(INFINITY) : 0u; }
    if (app_keys) {                // no threshold: every live sample key stays a candidate
#pragma unroll
      for (int i = 0; i < NK; ++i)
        if (sc[i] != 0u) {
          const uint32_t pos = atomicAdd(app_cnt, 1u);
          if (pos < (uint32_t)app_cap) app_keys[pos] = __ldcg(src + (uint32_t)et + (uint32_t)NT * i);
        }
    }
    epi_bar<NT>();
    return;
  }
  // MSB-first radix select over the registers: the `want`-th best ordered score -- and, in the same passes (second
  // histogram at hist + 288, scanned by the second warp), the `want2`-th best when want2 != 0.  The select runs over
  // (score - worst) from the top bit of the span down: a sample's scores share their upper bits, and a pass that drops
  // every key into one bucket serialises on that bucket's shared-memory atomic.
  uint32_t prefix = 0, lower = 0;
  {
    const uint32_t want2 = (tail_step && live >= 4u * (uint32_t)KP) ? 4u * (uint32_t)KP : 0u;
    uint32_t* hist2 = hist + 288;
    uint32_t need = (uint32_t)KP, need2 = want2;
    int hi = 32 - __clz(best - worst);   // bits >= hi of (score - worst) are zero in every live key
    while (hi > 0) {
      const int shift = hi > 8 ? hi - 8 : 0;
      const uint32_t bmask = (1u << (hi - shift)) - 1u;
      const uint32_t pre_hi = hi < 32 ? prefix >> hi : 0u, low_hi = hi < 32 ? lower >> hi : 0u;
      for (int i = et; i < 256; i += NT) { hist[i] = 0u; if (want2) hist2[i] = 0u; }
      epi_bar<NT>();
#pragma unroll
      for (int i = 0; i < NK; ++i)
        if (sc[i] != 0u) {
          const uint32_t rel = sc[i] - worst, top = hi < 32 ? rel >> hi : 0u;
          if (top == pre_hi) atomicAdd(&hist[(rel >> shift) & bmask], 1u);
          if (want2 && top == low_hi) atomicAdd(&hist2[(rel >> shift) & bmask], 1u);
        }
      epi_bar<NT>();
      if (et < 32 || (want2 && et < 64)) {
        // lane l owns buckets 255 - 8l .. 248 - 8l (descending): where does the running count reach `need`?
        const uint32_t* h = et < 32 ? hist : hist2;
        const uint32_t nd = et < 32 ? need : need2;
        const int l = et & 31;
        uint32_t c[8], tot = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { c[j] = h[255 - 8 * l - j]; tot += c[j]; }
        uint32_t incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
          if (l >= o) incl += v;
        }
        const uint32_t before = incl - tot;
        if (before < nd && incl >= nd) {
          uint32_t run = before;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (run < nd && run + c[j] >= nd) { ctl[et < 32 ? 2 : 4] = 255u - 8u * l - j; ctl[et < 32 ? 3 : 5] = nd - run; }
            run += c[j];
          }
        }
      }
      epi_bar<NT>();
      prefix |= ctl[2] << shift;
      need = ctl[3];
      if (want2) { lower |= ctl[4] << shift; need2 = ctl[5]; }
      hi = shift;
      epi_bar<NT>();
    }
    prefix += worst;
    lower += worst;
    if (!want2) lower = 0;
  }
  // tail_step (append plans of search_sw.cu): the ladder step comes from the sample's own tail slope -- the gap between
  // its KP-th and 4 KP-th best, i.e. ln 4 in rank -- not from its best score: a query whose twin sits in the sample
  // (score 1.0, common: queries are often corpus rows) would stretch an eighth-of-(best - base) ladder so far that no
  // level beyond the first ever collects KP rows.  0.43 = 0.6 / ln 4: 15 steps span ~9 e-foldings of rank.
  const float robust_step = lower ? (ord_to_f32(prefix) - ord_to_f32(lower)) * 0.43f : 0.f;
  if (et == 0) atomicMax(thr_q, prefix);
  if (lad) {
    const float base = ord_to_f32(prefix);
    float step = robust_step > 0.f ? robust_step : (ord_to_f32(best) - base) * ladder_frac;
    if (!(step > 0.f) || !(step < INFINITY)) step = 0.f;
    const float inv = step > 0.f ? 1.f / step : 0.f;
    if (et < kLadder) hist[et] = 0u;
    epi_bar<NT>();
#pragma unroll
    for (int i = 0; i < NK; ++i)
      if (sc[i] >= prefix) {             // (prefix > 0: empty slots never pass)
        const int j = ladder_level(base, step, inv, ord_to_f32(sc[i]));
        if (j >= 0) atomicAdd(&hist[j], 1u);
      }
    epi_bar<NT>();
    if (et < kLadder) {
      lad[kLadder + et] = hist[et];
      lad[et] = et == 0 ? __float_as_uint(base) : et == 1 ? __float_as_uint(step) : et == 2 ? __float_as_uint(inv) : 0u;
    }
  }
  if (app_keys) {
#pragma unroll
    for (int i = 0; i < NK; ++i)
      if (sc[i] >= prefix) {
        const uint32_t pos = atomicAdd(app_cnt, 1u);
        if (pos < (uint32_t)app_cap) app_keys[pos] = __ldcg(src + (uint32_t)et + (uint32_t)NT * i);
      }
  }
  epi_bar<NT>();
}

}  // namespace
}  // namespace tsim
