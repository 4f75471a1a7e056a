// Shared device/host helpers for libtsim (B200 / sm_100a).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tsim.h"

struct CUtensorMap_st;   // <cuda.h>: CUtensorMap

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
// SPLIT shadow of fp32 / fp16 rows (k up to 100): x = hi + lo + r with hi = bf16(x), lo = bf16(x - hi), |r| <= 2^-18 |x|;
// the tensor pass multiplies [qh | qh | ql] by [ch | cl | ch] (3 D wide), i.e. qh.ch + qh.cl + ql.ch.  What is
// missing from q.c is ql.cl + (qh + ql).rc + rq.c: each at most 2^-18 ||q|| ||c|| (3.8e-6); together 1.15e-5,
// rounded up to 1.5e-5, plus the tensor-core term of a 3 D-wide bf16 dot product.
__host__ __device__ inline float split_shadow_eps(int64_t D) { return approx_eps(3 * D, 2) + 1.5e-5f; }
// segment width of a split shadow: D rounded up to whole 64-element (128-byte) k-blocks, zero padded
__host__ __device__ inline int64_t split_shadow_seg(int64_t D) { return (D + 63) / 64 * 64; }

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

// ---- programmatic dependent launch --------------------------------------------------------
// A search is a chain of ~10 dependent kernels on one stream, several of them tiny (tighten, select, the retry /
// fallback kernels that usually find nothing to do).  Every kernel of the chain is launched with the
// programmatic-stream-serialization attribute, calls pdl_trigger() at its top (the next kernel may be scheduled as
// soon as every CTA of this one has started) and pdl_wait() before it first touches anything its predecessors
// wrote (blocks until they have completed and flushed): launch latency and prologues (barrier init, TMEM
// allocation, descriptor prefetch) overlap the predecessor's tail instead of adding up.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = knob_on("TSIM_NO_PDL") ? 0 : 1;   // experiment knob
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
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
// Products of bf16/fp16/fp8/fp32 values are exact in float64.  Fixed summation order: lane l owns the
// 8-element groups g with g % 32 == l (elements 8g .. 8g + 7), adds them in ascending element order with
// FMAs, then a fixed xor-butterfly adds the 32 partials.  ||q||^2 is summed once per query in the same order
// (warp_query_norm); dot and ||c||^2 per row.
// A group is kept as RAW bits (one 16-byte load for bf16 / fp16 rows, 8 bytes for e4m3, 32 for fp32 when the row
// is aligned; element loads packed the same way otherwise).  The query row is widened ONCE per query into a
// float64 copy in shared memory where the caller has room, so a corpus element costs one conversion and two FMAs.
// Widening to float64 goes through float (a shift for bf16) and ONE F2F.F64.F32 per element (2 issue cycles per
// warp on the XU pipe; doing it with integer bit manipulation costs 6-8 ALU cycles and was measured slower).
// Raw bits of one 8-element group of a row of element type DT.
template <int DT> struct Group8;
template <> struct Group8<TSIM_BF16> {
  uint32_t w[4];
  __device__ __forceinline__ void load_vec(const char* p) { const uint4 v = *reinterpret_cast<const uint4*>(p); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
  __device__ __forceinline__ void load_elems(const char* p, int n) {
    const unsigned short* h = reinterpret_cast<const unsigned short*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = (2 * i < n ? (uint32_t)h[2 * i] : 0u) | ((2 * i + 1 < n ? (uint32_t)h[2 * i + 1] : 0u) << 16);
  }
  __device__ __forceinline__ double get(int i) const {
    return (double)__uint_as_float((i & 1) ? (w[i >> 1] & 0xffff0000u) : (w[i >> 1] << 16));
  }
  static constexpr int kVecAlign = 16, kBytes = 2;
};
template <> struct Group8<TSIM_F16> {
  uint32_t w[4];
  __device__ __forceinline__ void load_vec(const char* p) { const uint4 v = *reinterpret_cast<const uint4*>(p); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
  __device__ __forceinline__ void load_elems(const char* p, int n) {
    const unsigned short* h = reinterpret_cast<const unsigned short*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = (2 * i < n ? (uint32_t)h[2 * i] : 0u) | ((2 * i + 1 < n ? (uint32_t)h[2 * i + 1] : 0u) << 16);
  }
  __device__ __forceinline__ double get(int i) const {
    return (double)__half2float(__ushort_as_half((unsigned short)((w[i >> 1] >> ((i & 1) * 16)) & 0xffffu)));
  }
  static constexpr int kVecAlign = 16, kBytes = 2;
};
template <> struct Group8<TSIM_E4M3> {
  uint32_t w[2];
  __device__ __forceinline__ void load_vec(const char* p) { const uint2 v = *reinterpret_cast<const uint2*>(p); w[0] = v.x; w[1] = v.y; }
  __device__ __forceinline__ void load_elems(const char* p, int n) {
    const unsigned char* h = reinterpret_cast<const unsigned char*>(p);
    w[0] = w[1] = 0u;
#pragma unroll
    for (int i = 0; i < 8; ++i) if (i < n) w[i >> 2] |= (uint32_t)h[i] << ((i & 3) * 8);
  }
  __device__ __forceinline__ double get(int i) const {
    __nv_fp8_e4m3 v;
    v.__x = (__nv_fp8_storage_t)((w[i >> 2] >> ((i & 3) * 8)) & 0xffu);
    return (double)float(v);
  }
  static constexpr int kVecAlign = 8, kBytes = 1;
};
template <> struct Group8<TSIM_F32> {
  uint32_t w[8];
  __device__ __forceinline__ void load_vec(const char* p) {
    const uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + 16);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
  }
  __device__ __forceinline__ void load_elems(const char* p, int n) {
    const uint32_t* h = reinterpret_cast<const uint32_t*>(p);
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = i < n ? h[i] : 0u;
  }
  __device__ __forceinline__ double get(int i) const { return (double)__uint_as_float(w[i]); }
  static constexpr int kVecAlign = 16, kBytes = 4;
};
// group starting at element e0 of `row` (all-zero bits past D: +0.0, which leaves every FMA chain unchanged)
template <int DT>
__device__ __forceinline__ void load_group8(Group8<DT>& g, const void* row, int64_t e0, int64_t D, bool vec) {
  const char* p = (const char*)row + e0 * Group8<DT>::kBytes;
  if (e0 + 8 <= D && vec) g.load_vec(p);
  else g.load_elems(p, e0 < D ? (int)min((int64_t)8, D - e0) : 0);
}

constexpr int kCosGroups = 3;   // 8-element groups per lane in flight (D = 768: the whole row in one round trip)

// sum of squares of one row in the canonical order (all lanes return the warp total)
template <int DT>
__device__ __forceinline__ double canonical_sumsq(const void* r, int64_t D) {
  const bool vec = ((uintptr_t)r & (Group8<DT>::kVecAlign - 1)) == 0;
  double acc = 0.0;
  for (int64_t e0 = (int64_t)(threadIdx.x & 31) * 8; e0 < D; e0 += 256 * kCosGroups) {
    Group8<DT> g[kCosGroups];
#pragma unroll
    for (int u = 0; u < kCosGroups; ++u) load_group8<DT>(g[u], r, e0 + 256 * u, D, vec);
#pragma unroll
    for (int u = 0; u < kCosGroups; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { const double x = g[u].get(i); acc = fma(x, x, acc); }
    }
  }
  return warp_sum_f64(acc);
}
// max(||q||, eps) of a query row: the same for every corpus row, so computed once per query
__device__ __forceinline__ double warp_query_norm(const void* q, int q_dt, int64_t D) {
  double qq;
  switch (q_dt) {
    case TSIM_F32: qq = canonical_sumsq<TSIM_F32>(q, D); break;
    case TSIM_F16: qq = canonical_sumsq<TSIM_F16>(q, D); break;
    case TSIM_BF16: qq = canonical_sumsq<TSIM_BF16>(q, D); break;
    default: qq = canonical_sumsq<TSIM_E4M3>(q, D); break;
  }
  return fmax(sqrt(qq), kCosEps);
}

// element e of a query row as float64 (for filling a shared-memory float64 copy of the query)
__device__ __forceinline__ double query_elem_f64(const void* q, int q_dt, int64_t e) {
  return (double)load_elem(q, q_dt, e);
}

// Position (in doubles) of element e in the lane-interleaved float64 copy of a query row: the pair (2p, 2p + 1) of
// lane l's group U * 32 + l sits at double2 slot (U * 4 + p) * 32 + l, so a warp's 16-byte reads are conflict free
// (a plain [D] layout puts the lanes 64 bytes apart: 16-way bank conflicts, measured +25 % on select_rescore).
// The copy holds qd_len(D) doubles, zero past D.
__host__ __device__ inline int64_t qd_len(int64_t D) { return (D + 255) / 256 * 256; }
__device__ __forceinline__ int64_t qd_slot(int64_t e) {
  const int64_t g = e >> 3, i = e & 7;
  return ((((g >> 5) * 4 + (i >> 1)) * 32 + (g & 31)) << 1) + (i & 1);
}
// dot(q, c) and ||c||^2 in the canonical order.  qd: lane-interleaved float64 copy of the query row (shared
// memory), or null -> the query's elements are widened from `q` on the fly.
template <int QDT, int CDT>
__device__ __forceinline__ void exact_cosine_sums(const void* q, const double* qd, const void* c, int64_t D,
                                                  double& dot, double& cc) {
  const bool qvec = ((uintptr_t)q & (Group8<QDT>::kVecAlign - 1)) == 0;
  const bool cvec = ((uintptr_t)c & (Group8<CDT>::kVecAlign - 1)) == 0;
  for (int64_t e0 = (int64_t)(threadIdx.x & 31) * 8; e0 < D; e0 += 256 * kCosGroups) {
    Group8<CDT> b[kCosGroups];
#pragma unroll
    for (int u = 0; u < kCosGroups; ++u) load_group8<CDT>(b[u], c, e0 + 256 * u, D, cvec);
#pragma unroll
    for (int u = 0; u < kCosGroups; ++u) {
      const int64_t eu = e0 + 256 * u;
      if (eu >= D) continue;     // an all-zero group past the end of the row: every FMA would leave its sum unchanged
      if (qd) {
        // lane-interleaved float64 copy (qd_slot below): consecutive lanes read consecutive 16-byte words
        const double2* q2 = reinterpret_cast<const double2*>(qd) + ((eu >> 8) * 4) * 32 + (threadIdx.x & 31);
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          const double2 x = q2[(i >> 1) * 32];
          const double y0 = b[u].get(i), y1 = b[u].get(i + 1);
          dot = fma(x.x, y0, dot); cc = fma(y0, y0, cc);
          dot = fma(x.y, y1, dot); cc = fma(y1, y1, cc);
        }
      } else {
        Group8<QDT> a;
        load_group8<QDT>(a, q, eu, D, qvec);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const double x = a.get(i), y = b[u].get(i);
          dot = fma(x, y, dot); cc = fma(y, y, cc);
        }
      }
    }
  }
}

template <int QDT>
__device__ __forceinline__ void exact_cosine_sums_c(const void* q, const double* qd, const void* c, int c_dt, int64_t D,
                                                    double& dot, double& cc) {
  switch (c_dt) {
    case TSIM_F32: exact_cosine_sums<QDT, TSIM_F32>(q, qd, c, D, dot, cc); break;
    case TSIM_F16: exact_cosine_sums<QDT, TSIM_F16>(q, qd, c, D, dot, cc); break;
    case TSIM_BF16: exact_cosine_sums<QDT, TSIM_BF16>(q, qd, c, D, dot, cc); break;
    default: exact_cosine_sums<QDT, TSIM_E4M3>(q, qd, c, D, dot, cc); break;
  }
}

// qn = warp_query_norm(q, q_dt, D).  All lanes return the score.
__device__ __forceinline__ double warp_exact_cosine(const void* q, int q_dt, const double* qd, double qn,
                                                    const void* c, int c_dt, int64_t D) {
  double dot = 0.0, cc = 0.0;
  if (q_dt == TSIM_BF16 && c_dt == TSIM_BF16) exact_cosine_sums<TSIM_BF16, TSIM_BF16>(q, qd, c, D, dot, cc);
  else if (q_dt == TSIM_E4M3 && c_dt == TSIM_E4M3) exact_cosine_sums<TSIM_E4M3, TSIM_E4M3>(q, qd, c, D, dot, cc);
  else if (q_dt == TSIM_F32 && c_dt == TSIM_F32) exact_cosine_sums<TSIM_F32, TSIM_F32>(q, qd, c, D, dot, cc);
  else {
    switch (q_dt) {      // mixed dtypes: same order, same bits, one more switch
      case TSIM_F32: exact_cosine_sums_c<TSIM_F32>(q, qd, c, c_dt, D, dot, cc); break;
      case TSIM_F16: exact_cosine_sums_c<TSIM_F16>(q, qd, c, c_dt, D, dot, cc); break;
      case TSIM_BF16: exact_cosine_sums_c<TSIM_BF16>(q, qd, c, c_dt, D, dot, cc); break;
      default: exact_cosine_sums_c<TSIM_E4M3>(q, qd, c, c_dt, D, dot, cc); break;
    }
  }
  dot = warp_sum_f64(dot);
  cc = warp_sum_f64(cc);
  const double cn = fmax(sqrt(cc), kCosEps);
  return dot / (qn * cn);
}

// ---- threshold from a set of candidate keys, by ONE warp ----------------------------------------
// MSB-first radix select (4 passes x 8 bits over the upper 32 key bits; `hist`: 256 words of shared memory private
// to the warp) of the KP-th best approximate score among src[0 .. n) (0 = empty slot, skipped): a valid lower bound
// of the query's final KP-th best whenever every key stands for a row that reaches select_rescore.  Raises *thr_q
// to it and, if `lad` is given, lays out the query's threshold ladder: 16 evenly spaced levels from that score
// (level 0) through the best score seen (level 8) to as far again above it, with the number of keys at or above
// the cut in each level.  Fewer than KP keys: no threshold, a ladder that never fires.
// L2ONLY: the keys were written by other CTAs of the SAME kernel (fused sticky pass): read them through L2.
template <bool L2ONLY>
__device__ __forceinline__ void warp_tighten(const uint64_t* src, uint32_t n, int KP, uint32_t* thr_q, uint32_t* lad,
                                             uint32_t* hist) {
  const int lane = threadIdx.x & 31;
  auto key_at = [&](uint32_t i) -> uint32_t { return (uint32_t)((L2ONLY ? __ldcg(src + i) : __ldg(src + i)) >> 32); };
  uint32_t prefix = 0, need = (uint32_t)KP, best = 0, live = 0;
  bool enough = true;
  for (int shift = 24; shift >= 0 && enough; shift -= 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) hist[lane + 32 * i] = 0;
    __syncwarp();
    const uint32_t hi_mask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
    for (uint32_t i = lane; i < n; i += 32) {
      const uint32_t sc = key_at(i);
      if (sc == 0u) continue;            // empty slot (a real key's ordered score is never 0)
      if (shift == 24) { best = max(best, sc); ++live; }
      if ((sc & hi_mask) == prefix) atomicAdd(&hist[(sc >> shift) & 255u], 1u);
    }
    __syncwarp();
    if (shift == 24) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) live += __shfl_xor_sync(0xffffffffu, live, o);
      enough = live >= (uint32_t)KP;
      if (!enough) break;
    }
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
    const int owner = __ffs(__ballot_sync(0xffffffffu, mine)) - 1;    // exactly one lane (live >= need)
    bucket = __shfl_sync(0xffffffffu, bucket, owner);
    need = __shfl_sync(0xffffffffu, rest, owner);
    prefix |= bucket << shift;
    __syncwarp();
  }
  if (!enough) {
    if (lad && lane < kLadder) {
      lad[kLadder + lane] = 0u;
      lad[lane] = lane == 0 ? __float_as_uint(INFINITY) : 0u;
    }
    return;
  }
  // prefix = the KP-th best approximate score (ordered-float bits)
  if (lane == 0) atomicMax(thr_q, prefix);
  if (!lad) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
  const float base = ord_to_f32(prefix);
  float step = (ord_to_f32(best) - base) * (1.f / 8.f);
  if (!(step > 0.f) || !(step < INFINITY)) step = 0.f;
  const float inv = step > 0.f ? 1.f / step : 0.f;
  if (lane < kLadder) hist[lane] = 0;
  __syncwarp();
  for (uint32_t i = lane; i < n; i += 32) {
    const uint32_t sc = key_at(i);
    if (sc >= prefix) {                  // (prefix > 0: empty slots never pass)
      const int j = ladder_level(base, step, inv, ord_to_f32(sc));
      if (j >= 0) atomicAdd(&hist[j], 1u);
    }
  }
  __syncwarp();
  if (lane < kLadder) {
    lad[kLadder + lane] = hist[lane];
    lad[lane] = lane == 0 ? __float_as_uint(base) : lane == 1 ? __float_as_uint(step) : lane == 2 ? __float_as_uint(inv) : 0u;
  }
  __syncwarp();
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
  // Append mode (round-robin plans with a bootstrap sample and KP >= 32): no candidate lists; in every pass the rows
  // that beat the query's threshold are appended to app_keys[q][0 .. app_cap) (count in app_cnt[q], zeroed with thr).
  // NC = 0 then.
  int qrep;            // sticky lone-CTA plans with Q <= 64: the query block holds the queries qrep (2 / 4) times over and
                       // the epilogue warps split every tile's columns (search_tc.cu, TcArgs::qrep); lists per worker x qrep
  // Small-batch plans (search_sw.cu: Q <= 32 on a large shard): corpus rows on the MMA's M side, the queries resident
  // on the N side.  Tiles of 128 rows; sample tiles i * sw_stride (i < sw_ns) are scanned first into 16-entry lists
  // (cand[q][8 * sw_ns][16]), tighten_kernel makes thresholds + ladders, the main pass appends (app_keys / app_cnt).
  int swapped, sw_ns, sw_stride;
  // split shadow (make_search_plan shadow_kind 2): 64-element k-blocks per segment of the corpus shadow's [hi | lo] rows;
  // k-block 3 j + r of the pass reads corpus block j (r < 2: hi) or ksplit + j (r == 2: lo).  0: plain rows
  int ksplit;
  int fused;           // sticky + bootstrap: one cooperative launch does sample, thresholds and main (TC_PASS_FUSED)
  size_t off_gbar;     // its two grid-barrier counters (256 bytes, zeroed with thr)
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
// shadow: 0 none; 1 rounded bf16 shadow (D wide, k <= 24, 112 candidates); 2 split shadow (pass D = 3 x the rows' width)
int make_search_plan(int64_t Q, int64_t N, int64_t D, int k, int q_dt, int c_dt, int mode,
                     bool need_invnorm, int shadow, SearchPlan* plan);

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
// 5 = sticky plans: sample + in-kernel thresholds + main in ONE cooperative launch (search_tc.cu, TcArgs::fused)
enum { TC_PASS_ALL = 0, TC_PASS_SAMPLE = 1, TC_PASS_MAIN = 2, TC_PASS_MINI = 3, TC_PASS_SAMPLE_REST = 4, TC_PASS_FUSED = 5 };
// TMA descriptors kept by a plan handle (tsim_plan_create), keyed by (base, rows, D, stride, box, element size):
// a repeated search of the same arrays encodes nothing.  Null: encode per launch.
struct MapCache;
MapCache* map_cache_create();
void map_cache_destroy(MapCache* c);
// 2-D row-major [rows, D] tensor map (bf16, or 1-byte e4m3) with a [box_rows, 128 bytes] box and 128-byte swizzle, from
// the cache `c` when it holds one (null: encode)
int get_tensor_map(MapCache* c, CUtensorMap_st* m, const void* base, int64_t rows, int64_t D, int64_t stride, int box_rows,
                   int esz);
int launch_sw_tighten(int64_t Q, const SearchPlan& p, const uint64_t* cand, uint32_t* thr, uint32_t* ladder,
                      uint64_t* app_keys, uint32_t* app_cnt, cudaStream_t st);
int search_sw_stages(int kblocks);   // corpus ring depth of the small-batch kernel for this many query k-blocks (0: no fit)
int launch_search_sw(const void* q, int64_t q_stride, const void* corpus, int64_t c_stride, int dt, const float* c_inv,
                     int64_t Q, int64_t N, int64_t D, int self_on, int64_t self_off, const SearchPlan& p, int sample,
                     uint64_t* cand, uint32_t* thr, uint32_t* ladder, uint64_t* app_keys, uint32_t* app_cnt,
                     cudaStream_t st, MapCache* maps);
int launch_search_tc(const void* q, int64_t q_stride, const void* corpus, int64_t c_stride, int dt,
                     const float* c_inv, int64_t Q, int64_t N, int64_t D,
                     int self_on, int64_t self_off, const SearchPlan& p, int pass, uint64_t* cand,
                     uint32_t* thr, uint32_t* ladder, uint64_t* sched, cudaStream_t st, MapCache* maps = nullptr,
                     const int32_t* q_count = nullptr, const int32_t* q_map = nullptr, int q_skip = 0,
                     uint64_t* app_keys = nullptr, uint32_t* app_cnt = nullptr, uint32_t* gbar = nullptr);
int launch_tighten(int64_t Q, const SearchPlan& p, int nslots, const uint64_t* cand, uint32_t* thr,
                   uint32_t* ladder, cudaStream_t st);
int launch_tighten_app(int64_t Q, const SearchPlan& p, const uint64_t* app_keys, const uint32_t* app_cnt,
                       uint32_t* thr, uint32_t* ladder, cudaStream_t st);
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
// first kernel of a search: zero the control words [zero_base, zero_base + zero_bytes) and, when q_dst is given,
// copy the Q query rows into the zero-padded block the TMA reads (q_rows_padded rows of row_bytes)
// (q_span > 0: padded row r holds query r % q_span -- the replicated block of a qrep plan)
int launch_search_prep(void* zero_base, size_t zero_bytes, const void* q_src, size_t q_src_stride_bytes, void* q_dst,
                       size_t row_bytes, int64_t Q, int64_t q_rows_padded, int q_span, cudaStream_t st);

int device_sm_count();
void count_launch();       // bumps the counter behind tsim_launch_count()
void count_map_encode();   // ... and the cuTensorMapEncodeTiled count of tsim_debug_counters()

}  // namespace tsim
