// Shared device/host helpers for libtsim (B200 / sm_100a).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tsim.h"

namespace tsim {

constexpr double kCosEps = 1e-8;   // F.cosine_similarity eps, reference search_pipeline.py:77
constexpr float kPoolEps = 1e-9f;  // clamp in AvgPoolingStrategy, reference modules.py:168

// Error bound of a tensor-core (bf16 x bf16 or e4m3 x e4m3 -> fp32) cosine, in units of ||q|| * ||c||:
// |approx - exact| <= approx_eps(D, element size).  Candidates are proven complete when the approximate k-th
// and KP-th best differ by more than 2 * eps * ||q|| (select_merge.cu).
// Model: the products are exact in fp32; tcgen05 adds the K = 16 (bf16) / 32 (e4m3) products of one MMA step to
// the fp32 accumulator with truncation, i.e. at most one ulp (2^-23 relative) of the running sum per step, and
// every running sum is bounded by sum |q_i c_i| <= ||q|| ||c||.  So the error is at most steps * 2^-23 (one-sided
// truncation does not average out: near-duplicate rows, whose products are all positive, do approach half of it).
// approx_eps = 4 x that bound, floored at 5e-5 (which also covers the fp32 inverse-norm scaling, ~2^-22).
// MEASURED (tests/test_gpu_eps.py, B200): max |approx - exact| / (||q|| ||c||) over unit-norm, 10^3-dynamic-range,
// cancellation-heavy (+x, -x) and near-duplicate rows stays below 1/4 of approx_eps for D = 64 ... 16384, both dtypes.
constexpr float kApproxEpsFloor = 5e-5f;
__host__ __device__ inline float approx_eps(int64_t D, int esz) {
  const int64_t per_step = esz == 1 ? 32 : 16;
  const float steps = (float)((D + per_step - 1) / per_step);
  const float e = steps * 4.76837158e-7f;   // steps * 2^-21 = 4 * steps * 2^-23
  return e > kApproxEpsFloor ? e : kApproxEpsFloor;
}
// The same bound when the tensor pass runs on a bf16 SHADOW of fp32 / fp16 rows (both operands rounded
// to 8 mantissa bits): |cos_shadow - cos| <= 2 * 2^-9 * 1.002 = 3.92e-3 (each unit vector moves by at most
// its relative rounding error), the shadow query's norm is off by <= 2^-9 (1.95e-3 of a score <= 1), plus
// the tensor-core term (5e-5 up to D = 1664; shadow_eps adds what approx_eps exceeds that by): 5.93e-3, rounded up.
constexpr float kShadowEps = 6e-3f;
__host__ __device__ inline float shadow_eps(int64_t D) { return kShadowEps + (approx_eps(D, 2) - kApproxEpsFloor); }

// Threshold ladder (search_tc.cu / select_merge.cu): per query, kLadder ascending score levels
// level(i) = base + i * step and the number of candidate rows seen so far in [level(i), level(i+1));
// once >= KP rows sit at or above a level, that level is a valid global threshold.  Lets every CTA
// filter with (nearly) the threshold a single sequential top-KP scan of everything seen so far would
// have.  Per-query record: 2 * kLadder words = { base, step, 1/step, pad..., counts[kLadder] }.
constexpr int kLadder = 16;

void set_error(const char* fmt, ...);

// ---- experiment knobs ---------------------------------------------------------------------
// The release library reads NO environment variable: every knob below is its default, folded at compile
// time, and the diagnosis switches inside the kernels (skip the MMAs, skip the epilogue ...) do not exist.
// `python -m text_similarity_b200.build --experiment` builds libtsim_exp.so with -DTSIM_EXPERIMENT, in which
// the knobs are environment variables read per call (scripts/ab_*.py interleave settings in one process).
#ifdef TSIM_EXPERIMENT
int knob_int(const char* name, int dflt);     // atoi(getenv(name)) or dflt; counted (tsim_debug_counters)
#define TSIM_KNOB_DEV(x) (x)
#else
inline int knob_int(const char*, int dflt) { return dflt; }
#define TSIM_KNOB_DEV(x) 0
#endif
inline bool knob_on(const char* name) { return knob_int(name, 0) == 1; }

#define TSIM_CHECK_ARG(cond, ...)          \
  do {                                     \
    if (!(cond)) {                         \
      ::tsim::set_error(__VA_ARGS__);      \
      return TSIM_ERR_INVALID_ARG;         \
    }                                      \
  } while (0)

#define TSIM_CUDA(call)                                                            \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) {                                                      \
      ::tsim::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),   \
                        __FILE__, __LINE__);                                       \
      return TSIM_ERR_CUDA;                                                        \
    }                                                                              \
  } while (0)

__host__ __device__ inline int dtype_size(int dt) {
  switch (dt) {
    case TSIM_F32: return 4;
    case TSIM_F16: return 2;
    case TSIM_BF16: return 2;
    case TSIM_E4M3: return 1;
    default: return 0;
  }
}

// ---- element loads -------------------------------------------------------------------
template <int DT> struct Elem;
template <> struct Elem<TSIM_F32> {
  using T = float;
  static __device__ __forceinline__ float ld(const void* p, int64_t i) { return ((const float*)p)[i]; }
};
template <> struct Elem<TSIM_F16> {
  using T = __half;
  static __device__ __forceinline__ float ld(const void* p, int64_t i) { return __half2float(((const __half*)p)[i]); }
};
template <> struct Elem<TSIM_BF16> {
  using T = __nv_bfloat16;
  static __device__ __forceinline__ float ld(const void* p, int64_t i) { return __bfloat162float(((const __nv_bfloat16*)p)[i]); }
};
template <> struct Elem<TSIM_E4M3> {
  using T = __nv_fp8_e4m3;
  static __device__ __forceinline__ float ld(const void* p, int64_t i) {
    return float(((const __nv_fp8_e4m3*)p)[i]);
  }
};

__device__ __forceinline__ float load_elem(const void* p, int dt, int64_t i) {
  switch (dt) {
    case TSIM_F32: return Elem<TSIM_F32>::ld(p, i);
    case TSIM_F16: return Elem<TSIM_F16>::ld(p, i);
    case TSIM_BF16: return Elem<TSIM_BF16>::ld(p, i);
    default: return Elem<TSIM_E4M3>::ld(p, i);
  }
}

// ---- order-preserving float <-> uint maps (for atomicMax and packed sort keys) ----------
__host__ __device__ __forceinline__ uint32_t f32_to_ord(float f) {
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(f);
#else
  uint32_t b; memcpy(&b, &f, 4);
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord_to_f32(uint32_t o) {
  uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(b);
}
__device__ __forceinline__ uint64_t f64_to_ord(double d) {
  uint64_t b = (uint64_t)__double_as_longlong(d);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

// Packed candidate key: sorting DESCENDING by the 64-bit value ranks by
// (approximate score descending, row index ascending).  0 is the empty slot.
__device__ __forceinline__ uint64_t pack_key(float s, uint32_t idx) {
  return ((uint64_t)f32_to_ord(s) << 32) | (uint64_t)(0xffffffffu - idx);
}
__device__ __forceinline__ float key_score(uint64_t k) { return ord_to_f32((uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t key_idx(uint64_t k) { return 0xffffffffu - (uint32_t)k; }

// ---- threshold ladder ---------------------------------------------------------------------
__device__ __forceinline__ float ladder_value(float base, float step, int i) { return fmaf((float)i, step, base); }
// Highest level whose VALUE is <= s (-1: s is below the ladder).  The value test makes the answer
// safe against rounding in the division-free index estimate: a row is never credited above its score.
__device__ __forceinline__ int ladder_level(float base, float step, float inv, float s) {
  const float t = (s - base) * inv;
  int j = t >= (float)(kLadder - 1) ? kLadder - 1 : (t > 0.f ? (int)t : 0);
  if (ladder_value(base, step, j) > s) --j;
  return j;
}

// ---- warp helpers ------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum_f32(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// The CANONICAL exact cosine of two stored rows, evaluated by one full warp.
// Every score the library returns is produced by this routine, whichever path nominated the
// row, so equal stored rows always yield bit-identical scores (ties -> lower index).
// Products of bf16/fp16/fp8/fp32 values are exact in float64; lane l sums elements
// l, l+32, ... sequentially, then a fixed xor-butterfly adds the 32 partials.
template <int QDT, int CDT>
__device__ __forceinline__ void exact_cosine_sums(const void* q, const void* c, int64_t D, double& dot, double& qq,
                                                  double& cc) {
  int64_t d = threadIdx.x & 31;
  // eight elements per lane per step, loads first (16 in flight), FMAs in the canonical order
  for (; d + 32 * 7 < D; d += 32 * 8) {
    float a[8], b[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { a[u] = Elem<QDT>::ld(q, d + 32 * u); b[u] = Elem<CDT>::ld(c, d + 32 * u); }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const double x = (double)a[u], y = (double)b[u];
      dot = fma(x, y, dot);
      qq = fma(x, x, qq);
      cc = fma(y, y, cc);
    }
  }
  for (; d < D; d += 32) {
    const double x = (double)Elem<QDT>::ld(q, d), y = (double)Elem<CDT>::ld(c, d);
    dot = fma(x, y, dot);
    qq = fma(x, x, qq);
    cc = fma(y, y, cc);
  }
}

__device__ __forceinline__ double warp_exact_cosine(const void* q, int q_dt, const void* c, int c_dt,
                                                    int64_t D) {
  const int lane = threadIdx.x & 31;
  double dot = 0.0, qq = 0.0, cc = 0.0;
  if (q_dt == TSIM_BF16 && c_dt == TSIM_BF16) exact_cosine_sums<TSIM_BF16, TSIM_BF16>(q, c, D, dot, qq, cc);
  else if (q_dt == TSIM_E4M3 && c_dt == TSIM_E4M3) exact_cosine_sums<TSIM_E4M3, TSIM_E4M3>(q, c, D, dot, qq, cc);
  else if (q_dt == TSIM_F32 && c_dt == TSIM_F32) exact_cosine_sums<TSIM_F32, TSIM_F32>(q, c, D, dot, qq, cc);
  else {
    for (int64_t d = lane; d < D; d += 32) {
      double a = (double)load_elem(q, q_dt, d);
      double b = (double)load_elem(c, c_dt, d);
      dot = fma(a, b, dot);
      qq = fma(a, a, qq);
      cc = fma(b, b, cc);
    }
  }
  dot = warp_sum_f64(dot);
  qq = warp_sum_f64(qq);
  cc = warp_sum_f64(cc);
  double qn = fmax(sqrt(qq), kCosEps);
  double cn = fmax(sqrt(cc), kCosEps);
  return dot / (qn * cn);
}

// ---- launch plans shared by the API and the kernels ---------------------------------------
struct SearchPlan {
  // tensor path
  int use_tensor;      // 1: tcgen05 candidate pass + select/rescore
  float eps;           // bound of |approx - exact| in units of ||q|| used by the completeness proof
  int KP;              // per-unit candidate list capacity (16/32/64/128)
  int pair;            // 1: CTA pairs (cta_group::2), query blocks of 256
  int QB;              // query blocks of 128 (256 when pair)
  int sticky;          // 1: each CTA keeps one query block and strides over tiles (few query blocks)
  int Gq;              // sticky: CTAs per query block
  int64_t boot_tiles;  // > 0: a bootstrap launch scans this many strided sample tiles first
  int64_t boot_stride; //      distance between sample tiles
  int64_t boot_slots;  //      candidate-list slots written by the bootstrap launch (the main launch follows)
  int64_t boot_tpc;    //      round-robin: tiles per unit of the bootstrap launch
  // round-robin only, > 0: the sample itself is scanned in two launches -- first every mini_mult-th
  // sample tile (mini_tiles of them, one-tile units, cold lists), then the rest of the sample with
  // thresholds and a ladder from the first.  Cold lists are expensive (every early row is inserted),
  // so only a handful of tiles ever see them.
  int64_t mini_mult, mini_tiles, mini_slots;
  // Append mode (round-robin plans with a bootstrap sample and KP >= 32): the MAIN pass keeps no lists; rows that
  // beat the query's threshold are appended to app_keys[q][0 .. app_cap) (count in app_cnt[q], zeroed with thr).
  // NC then counts the sample passes' list slots only.
  int append;
  int app_cap;
  size_t off_app_keys, off_app_cnt;
  int64_t R;           // round-robin: corpus rows per unit (multiple of 256)
  int64_t NC;          // candidate lists per query: chunks ceil(N / R), or Gq when sticky
  // exact path
  int S;               // corpus slices (CTAs along the corpus) of the exact scan
  int64_t slice_rows;  // rows per slice
  // workspace offsets (bytes)
  size_t off_cand, off_thr, off_flagcnt, off_flaglist, off_invnorm, off_qpad, off_ex_score, off_ex_idx;
  int has_ex_rinv;     // 1: the workspace holds [N] float64 inverse row norms for the tensor-core exact scan
  size_t off_ex_rinv;
  size_t off_ladder;   // [Q][2 * kLadder] u32: per-query threshold ladder (levels | counts), bootstrap plans only
  size_t off_sched;    // round-robin: 3 unit-claim areas (one per tcgen05 launch of a call), zeroed with thr
  size_t sched_area;   // bytes per claim area: 256 (counter) + workers * 32 records * 8
  // Wide second tensor pass for queries whose first-pass candidate set could not be proven complete
  // (ties straddling ranks k..KP, e.g. a sentence duplicated more than KP - k times): up to kRetryQ of
  // them are re-run with KP = kRetryKP lists before anything falls back to the float64 scan.
  int retry;           // > 0: rounds of the retry stage (tensor path, KP < kRetryKP, no shadow): round r re-runs
                       // flagged queries [r * kRetryQ, (r + 1) * kRetryQ); 1 round per 128 queries of the call, <= 4
  int r_Gq;            // retry: workers (= candidate lists per retried query)
  size_t off_r_thr, off_r_flagcnt;   // [rounds][kRetryQ] thresholds, level-2 flag count (zeroed with thr)
  size_t off_r_flaglist;             // [Q] level-2 flag list (queries the float64 scan answers)
  size_t off_r_q;                    // [rounds * kRetryQ][D] compact copy of the flagged queries
  size_t off_r_cand;                 // [kRetryQ][r_Gq][kRetryKP] keys (reused by every round)
  size_t total;
};
constexpr int kRetryQ = 128;
constexpr int kRetryKP = 112;
constexpr int kRetryMaxRounds = 4;

// q_dt / c_dt: the dtypes the tensor pass reads (the shadow's when `shadow`)
int make_search_plan(int64_t Q, int64_t N, int64_t D, int k, int q_dt, int c_dt, int mode,
                     bool need_invnorm, bool shadow, SearchPlan* plan);

// How select_rescore takes part in the retry stage (null: plain call, flagged queries go to flag_list).
//  * first pass (in_list == null): a flagged query also copies its row into r_q[position in flag_list]
//    (positions < cap) so that the wide pass can read the flagged queries as one compact block;
//  * retry pass (in_list != null): block b handles query in_list[skip + b] (skip + b < min(*in_cnt, cap))
//    with the compact slot b; queries it still cannot prove -- and, in the last round (`forward`), the
//    overflow in_list[cap..*in_cnt) -- go to flag_list (the level-2 list the float64 scan reads).
struct SelRetry {
  const int32_t* in_cnt; const int32_t* in_list; int cap; int skip; int forward;
  void* r_q; int64_t r_q_stride;   // first pass only: compact query buffer (elements of q_dt)
};

// kernels' host launchers (defined in the .cu files)
// pass: 0 = the whole corpus in one launch; 1 = bootstrap sample (all of it); 2 = main (everything
// that is not a sample tile); 3 = mini sample (first of two sample launches); 4 = rest of the sample
enum { TC_PASS_ALL = 0, TC_PASS_SAMPLE = 1, TC_PASS_MAIN = 2, TC_PASS_MINI = 3, TC_PASS_SAMPLE_REST = 4 };
// TMA descriptors kept by a plan handle (tsim_plan_create), keyed by (base, rows, D, stride, box, element size):
// a repeated search of the same arrays encodes nothing.  Null: encode per launch.
struct MapCache;
MapCache* map_cache_create();
void map_cache_destroy(MapCache* c);
int launch_search_tc(const void* q, int64_t q_stride, const void* corpus, int64_t c_stride, int dt,
                     const float* c_inv, int64_t Q, int64_t N, int64_t D,
                     int self_on, int64_t self_off, const SearchPlan& p, int pass, uint64_t* cand,
                     uint32_t* thr, uint32_t* ladder, uint64_t* sched, cudaStream_t st, MapCache* maps = nullptr,
                     const int32_t* q_count = nullptr, const int32_t* q_map = nullptr, int q_skip = 0,
                     uint64_t* app_keys = nullptr, uint32_t* app_cnt = nullptr);
int launch_tighten(int64_t Q, const SearchPlan& p, int nslots, const uint64_t* cand, uint32_t* thr,
                   uint32_t* ladder, cudaStream_t st);
int launch_select_rescore(const void* q, int q_dt, int64_t q_stride, const void* corpus, int c_dt,
                          int64_t c_stride, int64_t Q, int64_t N, int64_t D, int k,
                          int64_t idx_base, const SearchPlan& p, const uint64_t* cand,
                          const uint32_t* thr, int32_t* flag_cnt, int32_t* flag_list,
                          float* out_score, double* out_score64, int64_t* out_idx,
                          int32_t* out_flags, cudaStream_t st, const SelRetry* retry = nullptr,
                          const uint64_t* app_keys = nullptr, const uint32_t* app_cnt = nullptr);
int launch_search_exact(const void* q, int q_dt, int64_t q_stride, const void* corpus, int c_dt,
                        int64_t c_stride, int64_t Q, int64_t N, int64_t D, int k,
                        int self_on, int64_t self_off, const SearchPlan& p, const int32_t* flag_cnt,
                        const int32_t* flag_list, double* ex_score, uint32_t* ex_idx,
                        double* ex_rinv, cudaStream_t st);
int launch_merge_exact_lists(const void* q, int q_dt, int64_t q_stride, const void* corpus,
                             int c_dt, int64_t c_stride, int64_t Q, int64_t D, int k,
                             int64_t idx_base, const SearchPlan& p, const int32_t* flag_cnt,
                             const int32_t* flag_list, const double* ex_score,
                             const uint32_t* ex_idx, float* out_score, double* out_score64,
                             int64_t* out_idx, int32_t* out_flags, cudaStream_t st);
int launch_merge_topk(const double* sc, const int64_t* ix, int64_t Q, int64_t n_lists, int k_in,
                      int k_out, int64_t list_stride, int64_t query_stride, float* out_score, double* out_score64,
                      int64_t* out_idx, cudaStream_t st);
int launch_row_inv_norm(const void* x, int dt, int64_t N, int64_t D, int64_t stride, float* out,
                        cudaStream_t st);

int device_sm_count();
void count_launch();       // bumps the counter behind tsim_launch_count()
void count_map_encode();   // ... and the cuTensorMapEncodeTiled count of tsim_debug_counters()

}  // namespace tsim
