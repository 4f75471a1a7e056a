// K1: fused masked mean-pool + L2-normalise + cast (fp32 / bf16 / e4m3), and the row
// inverse-norm pass.  Replaces AvgPoolingStrategy.forward (reference
// src/modules/modules.py:158-171; same arithmetic at src/models/sentence_encoder.py:35-38),
// which issues six ATen kernels and materialises a [B,L,D] fp32 mask and product.
//
// HBM-bound: algorithmic bytes per call = B*L*D*e_in + B*L*mask_bytes + B*D*e_out + B*4.
// One pass over the token tensor with 16-byte coalesced loads; trailing padding is not read at
// all and masked tokens never enter the sum.  Grid = B * S CTAs (S = splits of the token axis) sized to >= 2 CTAs per SM;
// split partial sums go through a small fp32 workspace and the last CTA to finish a row
// (ticket counter) adds them in a fixed order, so results are deterministic.
#include <stdlib.h>

#include "tsim_common.cuh"

namespace tsim {

namespace {

__device__ __forceinline__ float mask_value(const void* mask, int mask_dt, int64_t i) {
  switch (mask_dt) {
    case TSIM_I64: return (float)((const int64_t*)mask)[i];
    case TSIM_I32: return (float)((const int32_t*)mask)[i];
    case TSIM_U8: return (float)((const uint8_t*)mask)[i];
    default: return ((const float*)mask)[i];
  }
}

template <int DT, int VEC> struct VecLoad;
template <> struct VecLoad<TSIM_F32, 4> {
  static __device__ __forceinline__ void ld(const void* p, int64_t i, float* o) {
    float4 v = __ldg((const float4*)((const float*)p + i));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
};
template <> struct VecLoad<TSIM_F16, 8> {
  static __device__ __forceinline__ void ld(const void* p, int64_t i, float* o) {
    uint4 v = __ldg((const uint4*)((const __half*)p + i));
    const __half2* h = (const __half2*)&v;
#pragma unroll
    for (int j = 0; j < 4; ++j) { float2 f = __half22float2(h[j]); o[2 * j] = f.x; o[2 * j + 1] = f.y; }
  }
};
template <> struct VecLoad<TSIM_BF16, 8> {
  static __device__ __forceinline__ void ld(const void* p, int64_t i, float* o) {
    uint4 v = __ldg((const uint4*)((const __nv_bfloat16*)p + i));
    const __nv_bfloat162* h = (const __nv_bfloat162*)&v;
#pragma unroll
    for (int j = 0; j < 4; ++j) { float2 f = __bfloat1622float2(h[j]); o[2 * j] = f.x; o[2 * j + 1] = f.y; }
  }
};
template <int DT> struct VecLoad<DT, 1> {
  static __device__ __forceinline__ void ld(const void* p, int64_t i, float* o) { o[0] = Elem<DT>::ld(p, i); }
};

// deterministic block-wide sum (all threads get the result); red = 32 floats of smem
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum_f32(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}

__device__ __forceinline__ float round_store(void* out, int out_dt, int64_t i, float v) {
  switch (out_dt) {
    case TSIM_F32: ((float*)out)[i] = v; return v;
    case TSIM_BF16: {
      __nv_bfloat16 h = __float2bfloat16_rn(v);
      ((__nv_bfloat16*)out)[i] = h;
      return __bfloat162float(h);
    }
    default: {
      __nv_fp8_e4m3 h(v);
      ((__nv_fp8_e4m3*)out)[i] = h;
      return float(h);
    }
  }
}

struct PoolArgs {
  const void* tok; const void* mask; int mask_dt;
  int64_t B, L, D, sb, sl, msb;
  int S, TL;              // token splits, tokens per split
  float* partial;         // [B, S, D]
  int* ticket;            // [B]
  void* out; int out_dt; int64_t out_stride; const int64_t* out_rows;
  float* out_inv; int normalize;
  int dbg;                // TSIM_POOL_DEBUG bits (diagnosis only): 1 = skip the accumulation
};

template <int DT, int VEC>
__global__ void __launch_bounds__(256) pool_norm_kernel(PoolArgs a) {
  extern __shared__ float sm[];           // [rpi][D] partial rows (then reused as pooled[D]) | [TL] token weights
  __shared__ float red[32];
  __shared__ int s_last;
  const int b = blockIdx.x / a.S, sp = blockIdx.x % a.S;
  const int nvec = (int)(a.D / VEC);
  const int rpi = max(1, (int)blockDim.x / nvec);
  const int tid = threadIdx.x;
  const int r = tid / nvec;
  const int l0 = sp * a.TL, l1 = min((int)a.L, l0 + a.TL);

  // ---- phase 1: this CTA's token range, reduced over tokens --------------------------
  // The range's token weights are staged in shared memory and its last non-zero weight located, so
  // the inner loop is branch-free: unconditional 16-byte loads (8 in flight per thread) up to the
  // last real token; trailing padding (the common mask shape 1..1 0..0) is never read.
  float* wts = sm + (size_t)rpi * a.D;    // [TL]
  __shared__ int s_lend;
  if (tid == 0) s_lend = l0;
  __syncthreads();
  for (int l = l0 + tid; l < l1; l += blockDim.x) {
    const float w = mask_value(a.mask, a.mask_dt, (int64_t)b * a.msb + l);
    wts[l - l0] = w;
    if (w != 0.f) atomicMax(&s_lend, l + 1);
  }
  __syncthreads();
  const int lend = s_lend;
  // thread (r, v): tokens l0 + r, l0 + r + rpi, ...; columns v*VEC .. v*VEC+VEC-1 (and, when
  // there are more vector columns than threads, further columns nvec-strided: v += blockDim)
  constexpr int UN = 8;
  for (int v0 = 0; v0 < nvec; v0 += blockDim.x) {
    const int v = v0 + (rpi > 1 ? tid % nvec : tid);
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
    if (v < nvec && r < rpi) {
      const int64_t base = (int64_t)b * a.sb + (int64_t)v * VEC;
      int l = l0 + r;
      for (; l + (UN - 1) * rpi < lend; l += UN * rpi) {
        float x[UN][VEC];
#pragma unroll
        for (int u = 0; u < UN; ++u) VecLoad<DT, VEC>::ld(a.tok, base + (int64_t)(l + u * rpi) * a.sl, x[u]);
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const float w = wts[l + u * rpi - l0];
          if (w != 0.f) {      // a masked token's values must not reach the sum (they may be Inf/NaN)
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] = fmaf(w, x[u][j], acc[j]);
          }
        }
      }
      for (; l < lend; l += rpi) {
        const float w = wts[l - l0];
        if (w != 0.f) {
          float x[VEC];
          VecLoad<DT, VEC>::ld(a.tok, base + (int64_t)l * a.sl, x);
#pragma unroll
          for (int j = 0; j < VEC; ++j) acc[j] = fmaf(w, x[j], acc[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < VEC; ++j) sm[(int64_t)r * a.D + (int64_t)v * VEC + j] = acc[j];
    }
  }
  __syncthreads();
  // fold the rpi sub-rows in a fixed order; result lands in sm[0..D)
  for (int d = tid; d < a.D; d += blockDim.x) {
    float t = sm[d];
    for (int rr = 1; rr < rpi; ++rr) t += sm[(int64_t)rr * a.D + d];
    if (a.S > 1) a.partial[((int64_t)b * a.S + sp) * a.D + d] = t;
    else sm[d] = t;
  }

  // ---- phase 2: the last CTA of row b finalises ---------------------------------------
  if (a.S > 1) {
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&a.ticket[b], 1) == a.S - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int d = tid; d < a.D; d += blockDim.x) {
      float t = 0.f;
      for (int s = 0; s < a.S; ++s) t += __ldcg(&a.partial[((int64_t)b * a.S + s) * a.D + d]);
      sm[d] = t;
    }
  }
  __syncthreads();

  float cnt = 0.f;
  for (int l = tid; l < a.L; l += blockDim.x) cnt += mask_value(a.mask, a.mask_dt, (int64_t)b * a.msb + l);
  cnt = block_sum(cnt, red);
  const float denom = fmaxf(cnt, kPoolEps);  // modules.py:168

  float ss = 0.f, amax = 0.f;
  for (int d = tid; d < a.D; d += blockDim.x) {
    float m = sm[d] / denom;  // modules.py:170
    sm[d] = m;
    ss = fmaf(m, m, ss);
    amax = fmaxf(amax, fabsf(m));
  }
  ss = block_sum(ss, red);
  float scale = 1.f;
  if (a.normalize) scale = 1.f / fmaxf(sqrtf(ss), (float)kCosEps);
  if (a.out_dt == TSIM_E4M3) {
    // per-row power-of-two scale: largest |element| lands in [64, 128) (e4m3 max is 448)
    float m = amax;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = m;
    __syncthreads();
    float bm = 0.f;
    for (int i = 0; i < (int)((blockDim.x + 31) >> 5); ++i) bm = fmaxf(bm, red[i]);
    bm *= scale;
    if (bm > 0.f) {
      int e;
      frexpf(bm, &e);  // bm = f * 2^e, f in [0.5, 1)
      scale *= exp2f((float)(7 - e));
    }
  }

  const int64_t orow = a.out_rows ? a.out_rows[b] : (int64_t)b;
  float ss2 = 0.f;
  for (int d = tid; d < a.D; d += blockDim.x) {
    float st = round_store(a.out, a.out_dt, orow * a.out_stride + d, sm[d] * scale);
    ss2 = fmaf(st, st, ss2);
  }
  ss2 = block_sum(ss2, red);
  if (a.out_inv && tid == 0) a.out_inv[orow] = 1.f / fmaxf(sqrtf(ss2), (float)kCosEps);
}

// ---- helpers of the streaming kernel: mbarrier / bulk-copy PTX, shared-memory vector decode ------
constexpr int kPsStageBytes = 16 * 1024;  // upper bound of a ring stage (a chunk = up to 32 whole token rows)

__device__ __forceinline__ uint32_t ps_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ps_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void ps_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ps_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ps_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void ps_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ float warp_max_f32(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <int DT> struct SmemVec;   // 16 bytes of tokens in shared memory -> floats
template <> struct SmemVec<TSIM_F32> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void cvt(const uint4& v, float* o) {
    o[0] = __uint_as_float(v.x); o[1] = __uint_as_float(v.y); o[2] = __uint_as_float(v.z); o[3] = __uint_as_float(v.w);
  }
};
template <> struct SmemVec<TSIM_F16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void cvt(const uint4& v, float* o) {
    const __half2* h = (const __half2*)&v;
#pragma unroll
    for (int j = 0; j < 4; ++j) { float2 f = __half22float2(h[j]); o[2 * j] = f.x; o[2 * j + 1] = f.y; }
  }
};
template <> struct SmemVec<TSIM_BF16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void cvt(const uint4& v, float* o) {
    // bf16 -> fp32 is a 16-bit shift: two integer ops per pair
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { o[2 * j] = __uint_as_float(w[j] << 16); o[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u); }
  }
};

// ---- K1, streaming kernel -------------------------------------------------------------------------
// One persistent CTA per SM; every warp is a self-contained streaming reducer with its OWN two-stage
// shared-memory ring, so there is no hand-off between warps anywhere:
//   * it draws (sentence, token-split) items from a global counter and reads an item's mask weights one
//     item ahead;
//   * lane 0 moves each chunk of up to 32 whole token rows (one contiguous run of the token tensor,
//     fully masked chunks skipped, trailing padding never read) with ONE 1-D bulk async copy (TMA,
//     UBLKCP) that completes on the stage's mbarrier -- always two chunks in flight per warp, across
//     item boundaries;
//   * the warp accumulates weight * token from its stage into registers (lane = 16-byte columns
//     lane, lane + 32, ...), and at an item's last chunk finishes the row itself: token count, mean
//     (modules.py:168-170), L2 norm, cast, store at out_rows[b], inverse norm -- warp shuffles only.
// ~14 chunks (210 KB) in flight per SM; no block-level barrier after start-up.
constexpr int kPwStages = 2;          // ring stages per warp
constexpr int kPwMaxWarps = 8;
constexpr int kPwMaxTL = 256;         // tokens per item (the launcher raises the split count to keep this)
struct PwHdr { float w[32]; int ntok; int last; float cnt; int pad; long long item; long long pad2; };

template <int DT, int NV>
__global__ void __launch_bounds__(kPwMaxWarps * 32, 1)
pool_norm_warp_kernel(PoolArgs a, int64_t nitems, int CT, int stage_bytes, int* counter) {
  extern __shared__ __align__(128) unsigned char pw_raw[];
  constexpr int VEC = SmemVec<DT>::N;
  constexpr int ESZ = 16 / VEC;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  // per warp: [stages][stage_bytes] data | hdr[stages] | wbuf[kPwMaxTL] | full[stages]
  const size_t per_warp = (size_t)kPwStages * stage_bytes + kPwStages * sizeof(PwHdr) + kPwMaxTL * sizeof(float) + 64;
  unsigned char* base = pw_raw + (size_t)warp * per_warp;
  unsigned char* data = base;
  PwHdr* hdr = (PwHdr*)(base + (size_t)kPwStages * stage_bytes);
  float* wbuf = (float*)(hdr + kPwStages);
  uint64_t* full = (uint64_t*)(wbuf + kPwMaxTL);
  const int nvec = (int)(a.D / VEC);
  if (lane == 0) {
    for (int s = 0; s < kPwStages; ++s) ps_mbar_init(ps_smem(&full[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();

  const int64_t stride_items = (int64_t)gridDim.x * nwarps;
  auto next_item = [&]() -> int64_t {       // warp g starts with item g; further items come from the counter
    int v = 0;
    if (lane == 0) v = atomicAdd(counter, 1);
    return (int64_t)__shfl_sync(0xffffffffu, v, 0) + stride_items;
  };
  float wn[kPwMaxTL / 32];
  auto load_mask = [&](int64_t item) {
    const int64_t b = item / a.S; const int sp = (int)(item % a.S);
    const int l0 = sp * a.TL, tl = min((int)a.L, l0 + a.TL) - l0;
#pragma unroll
    for (int j = 0; j < kPwMaxTL / 32; ++j)
      wn[j] = (lane + 32 * j < tl) ? mask_value(a.mask, a.mask_dt, b * a.msb + l0 + lane + 32 * j) : 0.f;
  };

  // ---- issue side: the (item, chunk) cursor ----
  int64_t p_item = (int64_t)blockIdx.x * nwarps + warp, p_next = -1;
  int p_lend = 0, p_lc = 0, p_l0 = 0;
  float p_cnt = 0.f;
  bool p_open = false;                    // wbuf / p_lend describe p_item
  if (p_item < nitems) load_mask(p_item);
  uint32_t issued = 0, consumed = 0;
  auto open_item = [&]() {                // wn holds p_item's mask: publish it, prefetch the next one's
    const int sp = (int)(p_item % a.S);
    p_l0 = sp * a.TL;
    const int tl = min((int)a.L, p_l0 + a.TL) - p_l0;
    float cnt = 0.f;
    int lend = 0;
#pragma unroll
    for (int j = 0; j < kPwMaxTL / 32; ++j) {
      if (lane + 32 * j < tl) wbuf[lane + 32 * j] = wn[j];
      cnt += wn[j];
      if (wn[j] != 0.f) lend = lane + 32 * j + 1;
    }
    p_cnt = warp_sum_f32(cnt);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lend = max(lend, __shfl_xor_sync(0xffffffffu, lend, o));
    p_lend = lend; p_lc = 0; p_open = true;
    __syncwarp();
    p_next = next_item();
    if (p_next < nitems) load_mask(p_next);
  };
  // issue the next chunk into stage issued % kPwStages; false when there is nothing left to issue
  auto issue_one = [&]() -> bool {
    for (;;) {
      if (p_item >= nitems) return false;
      if (!p_open) open_item();
      // chunks up to the item's last real token; the chunk that holds it closes the item; an item
      // without any real token is a bare closing stage
      while (p_lc < p_lend || p_lc == 0) {
        const int lc = p_lc;
        p_lc += CT;
        const int nt = max(0, min(CT, p_lend - lc));
        const bool last = lc + CT >= p_lend;
        const float w = lane < nt ? wbuf[lc + lane] : 0.f;
        const unsigned nz = __ballot_sync(0xffffffffu, w != 0.f);
        if (!nz && !last) continue;                                     // all masked: not read at all
        const int st = (int)(issued % kPwStages);
        hdr[st].w[lane] = w;
        if (lane == 0) { hdr[st].ntok = nt; hdr[st].last = last ? 1 : 0; hdr[st].cnt = p_cnt; hdr[st].item = p_item; }
        __syncwarp();
        if (lane == 0) {
          const uint32_t bytes = (uint32_t)nt * (uint32_t)a.D * ESZ;
          const uint32_t fb = ps_smem(&full[st]);
          if (bytes) {
            // the stage was read by this warp's generic loads: order them before the async-proxy write
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            ps_mbar_expect_tx(fb, bytes);
            const int64_t b = p_item / a.S;
            ps_bulk_load(ps_smem(data + (size_t)st * stage_bytes),
                         (const unsigned char*)a.tok + (b * a.sb + (int64_t)(p_l0 + lc) * a.D) * ESZ, bytes, fb);
          } else {
            ps_mbar_arrive(fb);
          }
        }
        ++issued;
        if (last) { p_item = p_next; p_open = false; }
        return true;
      }
      p_item = p_next; p_open = false;      // not reached: the loop above always ends with a closing chunk
    }
  };

  float acc[NV][VEC];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[i][j] = 0.f;

  for (;;) {
    while (issued - consumed < (uint32_t)kPwStages && issue_one()) {}
    if (issued == consumed) break;
    const int st = (int)(consumed % kPwStages);
    ps_mbar_wait(ps_smem(&full[st]), (consumed / kPwStages) & 1u);
    const PwHdr* h = hdr + st;
    const int ntok = h->ntok;
    const uint4* src = (const uint4*)(data + (size_t)st * stage_bytes);
    if (!(TSIM_KNOB_DEV(a.dbg) & 1)) {
      for (int t = 0; t < ntok; t += 2) {
        // two tokens per step, loads first; a masked token's values must not reach the sum (they may be Inf/NaN)
        const bool ok1 = t + 1 < ntok;
        const float w0 = h->w[t], w1 = ok1 ? h->w[t + 1] : 0.f;
        uint4 r0[NV], r1[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int v = lane + 32 * i;
          if (v < nvec) {
            r0[i] = src[(size_t)t * nvec + v];
            r1[i] = src[(size_t)(ok1 ? t + 1 : t) * nvec + v];
          }
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          if (lane + 32 * i < nvec) {
            float x0[VEC], x1[VEC];
            SmemVec<DT>::cvt(r0[i], x0);
            SmemVec<DT>::cvt(r1[i], x1);
            if (w0 != 0.f) {
#pragma unroll
              for (int j = 0; j < VEC; ++j) acc[i][j] = fmaf(w0, x0[j], acc[i][j]);
            }
            if (w1 != 0.f) {
#pragma unroll
              for (int j = 0; j < VEC; ++j) acc[i][j] = fmaf(w1, x1[j], acc[i][j]);
            }
          }
        }
      }
    }
    const int last = h->last;
    float cnt = h->cnt;
    const int64_t item = h->item;
    __syncwarp();                 // every lane is done with the stage (and its header) before it is refilled
    ++consumed;
    if (!last) continue;

    // ---- the item is complete: finish the row in registers ----
    const int64_t b = item / a.S; const int sp = (int)(item % a.S);
    bool finish = true;
    if (a.S > 1) {
      float* pr = a.partial + ((int64_t)b * a.S + sp) * a.D;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int v = lane + 32 * i;
        if (v < nvec) {
#pragma unroll
          for (int j = 0; j < VEC; j += 4)
            *(float4*)(pr + (size_t)v * VEC + j) = make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
        }
      }
      // the last warp to finish a split of row b adds the partial sums in a fixed order
      __threadfence();
      __syncwarp();
      int lastw = 0;
      if (lane == 0) lastw = (atomicAdd(&a.ticket[b], 1) == a.S - 1);
      lastw = __shfl_sync(0xffffffffu, lastw, 0);
      finish = lastw != 0;
      if (finish) {
        __threadfence();
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int v = lane + 32 * i;
#pragma unroll
          for (int j = 0; j < VEC; ++j) acc[i][j] = 0.f;
          if (v < nvec) {
            for (int s2 = 0; s2 < a.S; ++s2) {
              const float* ps = a.partial + ((int64_t)b * a.S + s2) * a.D + (size_t)v * VEC;
#pragma unroll
              for (int j = 0; j < VEC; j += 4) {
                const float4 u = __ldcg((const float4*)(ps + j));
                acc[i][j] += u.x; acc[i][j + 1] += u.y; acc[i][j + 2] += u.z; acc[i][j + 3] += u.w;
              }
            }
          }
        }
        float c = 0.f;
        for (int l = lane; l < a.L; l += 32) c += mask_value(a.mask, a.mask_dt, b * a.msb + l);
        cnt = warp_sum_f32(c);
      }
    }
    if (finish) {
      const float denom = fmaxf(cnt, kPoolEps);  // modules.py:168
      float ss = 0.f, amax = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float m = acc[i][j] / denom;     // modules.py:170 (lanes past the row hold zeros)
          acc[i][j] = m;
          ss = fmaf(m, m, ss);
          amax = fmaxf(amax, fabsf(m));
        }
      float scale = 1.f;
      if (a.normalize) scale = 1.f / fmaxf(sqrtf(warp_sum_f32(ss)), (float)kCosEps);
      if (a.out_dt == TSIM_E4M3) {
        // per-row power-of-two scale: largest |element| lands in [64, 128) (e4m3 max is 448)
        const float bm = warp_max_f32(amax) * scale;
        if (bm > 0.f) {
          int e;
          frexpf(bm, &e);  // bm = f * 2^e, f in [0.5, 1)
          scale *= exp2f((float)(7 - e));
        }
      }
      const int64_t orow = a.out_rows ? a.out_rows[b] : b;
      float ss2 = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int v = lane + 32 * i;
        if (v < nvec) {
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            const float stv = round_store(a.out, a.out_dt, orow * a.out_stride + (int64_t)v * VEC + j, acc[i][j] * scale);
            ss2 = fmaf(stv, stv, ss2);
          }
        }
      }
      if (a.out_inv) {
        ss2 = warp_sum_f32(ss2);
        if (lane == 0) a.out_inv[orow] = 1.f / fmaxf(sqrtf(ss2), (float)kCosEps);
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[i][j] = 0.f;
  }
}

// Sum of squares of one 16-byte vector of a stored row.
template <int DT> struct SqVec;
template <> struct SqVec<TSIM_F32> {
  static constexpr int N = 4;
  static __device__ __forceinline__ float sq(const uint4& raw) {
    const float* f = (const float*)&raw;
    return f[0] * f[0] + f[1] * f[1] + f[2] * f[2] + f[3] * f[3];
  }
};
template <> struct SqVec<TSIM_F16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ float sq(const uint4& raw) {
    const __half2* h = (const __half2*)&raw;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) { float2 f = __half22float2(h[j]); s = fmaf(f.x, f.x, s); s = fmaf(f.y, f.y, s); }
    return s;
  }
};
template <> struct SqVec<TSIM_BF16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ float sq(const uint4& raw) {
    const __nv_bfloat162* h = (const __nv_bfloat162*)&raw;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) { float2 f = __bfloat1622float2(h[j]); s = fmaf(f.x, f.x, s); s = fmaf(f.y, f.y, s); }
    return s;
  }
};
template <> struct SqVec<TSIM_E4M3> {
  static constexpr int N = 16;
  static __device__ __forceinline__ float sq(const uint4& raw) {
    // two e4m3 values per conversion (cvt.rn.f16x2.e4m3x2); the squares are exact in fp32
    const __nv_fp8x2_storage_t* e = (const __nv_fp8x2_storage_t*)&raw;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const __half2_raw h = __nv_cvt_fp8x2_to_halfraw2(e[j], __NV_E4M3);
      const float2 f = __half22float2(*(const __half2*)&h);
      s = fmaf(f.x, f.x, s);
      s = fmaf(f.y, f.y, s);
    }
    return s;
  }
};

// One warp per row, four rows in flight per warp, grid-stride over rows; float accumulation (the
// value only scales the approximate scores of the candidate pass; final scores are recomputed in
// float64).  HBM-bound: N*D*e bytes in, N*4 out.
template <int DT>
__global__ void __launch_bounds__(256) row_inv_norm_kernel(const void* x, int64_t N, int64_t D,
                                                           int64_t stride, float* out, int vec_ok) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  constexpr int R = 4;
  constexpr int VN = SqVec<DT>::N;
  const int esz = 16 / VN;
  for (int64_t row0 = gw * R; row0 < N; row0 += nwarps * R) {
    float ss[R];
#pragma unroll
    for (int r = 0; r < R; ++r) ss[r] = 0.f;
    if (vec_ok) {
      for (int64_t d = (int64_t)lane * VN; d < D; d += 32 * VN) {
        uint4 raw[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int64_t row = row0 + r < N ? row0 + r : N - 1;
          raw[r] = __ldg((const uint4*)((const unsigned char*)x + (row * stride + d) * esz));
        }
#pragma unroll
        for (int r = 0; r < R; ++r) ss[r] += SqVec<DT>::sq(raw[r]);
      }
    } else {
      for (int64_t d = lane; d < D; d += 32) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int64_t row = row0 + r < N ? row0 + r : N - 1;
          const float v = Elem<DT>::ld(x, row * stride + d);
          ss[r] = fmaf(v, v, ss[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float t = warp_sum_f32(ss[r]);
      if (lane == 0 && row0 + r < N) out[row0 + r] = 1.f / fmaxf(sqrtf(t), (float)kCosEps);
    }
  }
}

// e4m3 rows: the sums of squares come off the tensor cores.  Converting an e4m3 byte to float and squaring it costs
// ~2.5 instructions per BYTE on the CUDA cores: the kernel above ran at 36 % of HBM on e4m3 rows (4 ms per 25M x 384).
// mma.sync.m16n8k32 (e4m3 x e4m3 -> f32) multiplies a 16 x 32 block of A by a 32 x 8 block of B; feeding 16 rows as A
// and THE SAME registers as B -- rows 0-7 for one MMA, rows 8-15 for a second -- gives blocks of X X^T whose diagonal
// is the rows' sums of squares (products exact, float accumulation).  Any assignment of a row's bytes to the k
// positions is fine as long as A and B agree, and they do by construction (thread (g, kq) holds bytes kq * 16 .. + 15
// of a 64-byte step of row g in the registers that the fragment layout calls row g / column g): two 16-byte loads and
// four MMAs per thread per 16 rows x 64 bytes, ~10 instructions per KILOBYTE.
__device__ __forceinline__ void mma_e4m3_16x8x32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                 uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.f32.e4m3.e4m3.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256) row_inv_norm_e4m3_mma_kernel(const unsigned char* x, int64_t N, int64_t D,
                                                                    int64_t stride, float* out) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, kq = lane & 3;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  for (int64_t row0 = gw * 16; row0 < N; row0 += nwarps * 16) {
    const int64_t ra = min(row0 + g, N - 1), rb = min(row0 + g + 8, N - 1);    // (rows past the end: clamped, not stored)
    const unsigned char* pa = x + ra * stride + kq * 16;
    const unsigned char* pb = x + rb * stride + kq * 16;
    float d1[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f};
    for (int64_t b0 = 0; b0 < D; b0 += 256) {                // four 64-byte steps per round: eight loads in flight
      uint4 w[4], v[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int64_t off = b0 + s * 64;
        const bool ok = off + kq * 16 < D;
        w[s] = ok ? __ldg(reinterpret_cast<const uint4*>(pa + off)) : zero;
        v[s] = ok ? __ldg(reinterpret_cast<const uint4*>(pb + off)) : zero;
      }
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        mma_e4m3_16x8x32(d1, w[s].x, v[s].x, w[s].y, v[s].y, w[s].x, w[s].y);   // columns = rows 0-7
        mma_e4m3_16x8x32(d2, w[s].x, v[s].x, w[s].y, v[s].y, v[s].x, v[s].y);   // columns = rows 8-15
        mma_e4m3_16x8x32(d1, w[s].z, v[s].z, w[s].w, v[s].w, w[s].z, w[s].w);
        mma_e4m3_16x8x32(d2, w[s].z, v[s].z, w[s].w, v[s].w, v[s].z, v[s].w);
      }
    }
    // diagonals: (row g, column g) sits in thread (g, kq = g / 2), element g % 2 of d1; (row g + 8, column g) in d2[2 + g % 2]
    if (kq == (g >> 1)) {
      const float sa = (g & 1) ? d1[1] : d1[0], sb = (g & 1) ? d2[3] : d2[2];
      if (row0 + g < N) out[row0 + g] = 1.f / fmaxf(sqrtf(sa), (float)kCosEps);
      if (row0 + g + 8 < N) out[row0 + g + 8] = 1.f / fmaxf(sqrtf(sb), (float)kCosEps);
    }
  }
}

template <int DT, int VEC>
int launch_pool(const PoolArgs& a, cudaStream_t st) {
  const int nvec = (int)(a.D / VEC);
  int threads;
  if (nvec >= 256) threads = 256;
  else {
    int rpi = 256 / nvec;
    threads = nvec * rpi;
    threads = ((threads + 31) / 32) * 32;
    if (threads > 256) threads = 256;
  }
  const int rpi = threads / nvec > 0 ? threads / nvec : 1;
  size_t smem = ((size_t)rpi * a.D + a.TL) * sizeof(float);
  if (smem > 200 * 1024) { set_error("pool_norm: D=%lld too large", (long long)a.D); return TSIM_ERR_UNSUPPORTED; }
  if (smem > 48 * 1024)
    TSIM_CUDA(cudaFuncSetAttribute(pool_norm_kernel<DT, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pool_norm_kernel<DT, VEC><<<(unsigned)(a.B * a.S), threads, smem, st>>>(a);
  TSIM_CUDA(cudaGetLastError());
  count_launch();
  return TSIM_OK;
}

int pool_splits(int64_t B, int64_t L) {
  int sms = device_sm_count();
  int64_t want = (2 * (int64_t)sms + B - 1) / B;       // >= 2 CTAs per SM overall
  int64_t maxs = (L + 7) / 8;                           // >= 8 tokens per CTA
  int64_t S = want < 1 ? 1 : want;
  if (S > maxs) S = maxs;
  const int64_t mins = (L + kPwMaxTL - 1) / kPwMaxTL;   // <= kPwMaxTL tokens per item (streaming kernel)
  if (S < mins) S = mins;
  if (S < 1) S = 1;
  return (int)S;
}

// The streaming kernel needs token rows that are contiguous within a sentence (one bulk copy per
// chunk) and 16-byte granularity.  TSIM_POOL_MODE=1 (experiment knob) forces the register-staged kernel.
int pool_mode() {
  return knob_int("TSIM_POOL_MODE", 0);
}
// warp-autonomous variant: at most 8 sixteen-byte columns per lane (row <= 4 KB), items of <= 256 tokens
bool pool_warp_ok(const PoolArgs& a, int esz, bool vec_ok) {
  const int64_t rowb = a.D * esz;
  return vec_ok && a.sl == a.D && rowb <= 4096 && a.TL <= kPwMaxTL;
}

template <int DT, int NV>
int launch_pool_warp_nv(const PoolArgs& a, int* counter, cudaStream_t st) {
  constexpr int ESZ = 16 / SmemVec<DT>::N;
  const int64_t rowb = a.D * ESZ;
  int CT = (int)(kPsStageBytes / rowb);
  if (CT > 32) CT = 32;
  const int stage_bytes = (int)((CT * rowb + 127) / 128 * 128);
  const size_t per_warp = (size_t)kPwStages * stage_bytes + kPwStages * sizeof(PwHdr) + kPwMaxTL * sizeof(float) + 64;
  int nw = (int)((size_t)(225 * 1024) / per_warp);
  if (nw > kPwMaxWarps) nw = kPwMaxWarps;
  if (nw < 1) nw = 1;
  const size_t smem = (size_t)nw * per_warp + 128;
  auto kern = pool_norm_warp_kernel<DT, NV>;
  TSIM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t nitems = a.B * a.S;
  int64_t ctas = (nitems + nw - 1) / nw;
  const int64_t cap = device_sm_count();
  if (ctas > cap) ctas = cap;
  kern<<<(unsigned)ctas, nw * 32, smem, st>>>(a, nitems, CT, stage_bytes, counter);
  TSIM_CUDA(cudaGetLastError());
  count_launch();
  return TSIM_OK;
}
template <int DT>
int launch_pool_warp(const PoolArgs& a, int* counter, cudaStream_t st) {
  const int nvec = (int)(a.D / SmemVec<DT>::N);
  const int nv = (nvec + 31) / 32;
  switch (nv) {
    case 1: return launch_pool_warp_nv<DT, 1>(a, counter, st);
    case 2: return launch_pool_warp_nv<DT, 2>(a, counter, st);
    case 3: return launch_pool_warp_nv<DT, 3>(a, counter, st);
    case 4: return launch_pool_warp_nv<DT, 4>(a, counter, st);
    case 5: case 6: return launch_pool_warp_nv<DT, 6>(a, counter, st);
    default: return launch_pool_warp_nv<DT, 8>(a, counter, st);
  }
}

}  // namespace

int launch_row_inv_norm(const void* x, int dt, int64_t N, int64_t D, int64_t stride, float* out,
                        cudaStream_t st) {
  if (N == 0) return TSIM_OK;
  const int esz = dtype_size(dt);
  if (!esz) { set_error("row_inv_norm: bad dtype %d", dt); return TSIM_ERR_INVALID_ARG; }
  const int per16 = 16 / esz;
  const int vec_ok = (D % per16 == 0) && (stride % per16 == 0) && (((uintptr_t)x & 15) == 0);
  const int wpb = 8;
  int64_t want = (N + 4 * wpb - 1) / (4 * wpb);
  const int64_t cap = (int64_t)device_sm_count() * 8;      // persistent: 8 CTAs per SM, grid-stride over rows
  unsigned grid = (unsigned)(want < cap ? want : cap);
  if (dt == TSIM_E4M3 && vec_ok && !knob_on("TSIM_NO_MMA_NORM")) {      // tensor-core sums of squares, 16 rows per warp
    int64_t want16 = (N + 16 * wpb - 1) / (16 * wpb);
    row_inv_norm_e4m3_mma_kernel<<<(unsigned)(want16 < cap ? want16 : cap), wpb * 32, 0, st>>>((const unsigned char*)x, N, D, stride, out);
    TSIM_CUDA(cudaGetLastError());
    count_launch();
    return TSIM_OK;
  }
  switch (dt) {
    case TSIM_F32: row_inv_norm_kernel<TSIM_F32><<<grid, wpb * 32, 0, st>>>(x, N, D, stride, out, vec_ok); break;
    case TSIM_F16: row_inv_norm_kernel<TSIM_F16><<<grid, wpb * 32, 0, st>>>(x, N, D, stride, out, vec_ok); break;
    case TSIM_BF16: row_inv_norm_kernel<TSIM_BF16><<<grid, wpb * 32, 0, st>>>(x, N, D, stride, out, vec_ok); break;
    default: row_inv_norm_kernel<TSIM_E4M3><<<grid, wpb * 32, 0, st>>>(x, N, D, stride, out, vec_ok); break;
  }
  TSIM_CUDA(cudaGetLastError());
  count_launch();
  return TSIM_OK;
}

}  // namespace tsim

using namespace tsim;

// workspace layout: [item counter, 256 B][ticket, B ints][partial sums, B * S * D floats]
extern "C" size_t tsim_pool_workspace_bytes(int64_t B, int64_t L, int64_t D) {
  if (B <= 0 || L <= 0 || D <= 0) return 512;
  int S = pool_splits(B, L);
  size_t partial = (S > 1) ? (size_t)B * S * D * sizeof(float) : 0;
  size_t ticket = (size_t)B * sizeof(int);
  return 256 + ((ticket + 255) / 256) * 256 + ((partial + 255) / 256) * 256 + 256;
}

extern "C" int tsim_pool_norm(const void* tok, int tok_dt, const void* mask, int mask_dt,
                              int64_t B, int64_t L, int64_t D, int64_t tok_stride_b,
                              int64_t tok_stride_l, int64_t mask_stride_b, void* out, int out_dt,
                              int64_t out_stride, const int64_t* out_rows, float* out_inv_norm,
                              int normalize, void* ws, size_t ws_bytes, void* stream) {
  TSIM_CHECK_ARG(B >= 0 && L >= 0 && D > 0, "pool_norm: bad shape B=%lld L=%lld D=%lld", (long long)B, (long long)L, (long long)D);
  if (B == 0) return TSIM_OK;
  TSIM_CHECK_ARG(tok && mask && out, "pool_norm: null pointer");
  TSIM_CHECK_ARG(L > 0, "pool_norm: L must be > 0");
  TSIM_CHECK_ARG(tok_dt == TSIM_F32 || tok_dt == TSIM_F16 || tok_dt == TSIM_BF16, "pool_norm: token dtype %d unsupported", tok_dt);
  TSIM_CHECK_ARG(mask_dt == TSIM_I64 || mask_dt == TSIM_I32 || mask_dt == TSIM_U8 || mask_dt == TSIM_F32, "pool_norm: mask dtype %d unsupported", mask_dt);
  TSIM_CHECK_ARG(out_dt == TSIM_F32 || out_dt == TSIM_BF16 || out_dt == TSIM_E4M3, "pool_norm: output dtype %d unsupported", out_dt);
  TSIM_CHECK_ARG(out_stride >= D, "pool_norm: out_stride < D");
  cudaStream_t st = (cudaStream_t)stream;
  PoolArgs a;
  a.tok = tok; a.mask = mask; a.mask_dt = mask_dt;
  a.B = B; a.L = L; a.D = D; a.sb = tok_stride_b; a.sl = tok_stride_l; a.msb = mask_stride_b;
  a.S = pool_splits(B, L);
  a.TL = (int)((L + a.S - 1) / a.S);
  a.out = out; a.out_dt = out_dt; a.out_stride = out_stride; a.out_rows = out_rows;
  a.out_inv = out_inv_norm; a.normalize = normalize;
  a.partial = nullptr; a.ticket = nullptr;
  a.dbg = knob_int("TSIM_POOL_DEBUG", 0);   // experiment knob, compiled out of the release library
  const size_t ticket_al = (((size_t)B * sizeof(int) + 255) / 256) * 256;
  const size_t need = 256 + (a.S > 1 ? ticket_al + (size_t)B * a.S * D * sizeof(float) : 0);
  if (!ws || ws_bytes < need) { set_error("pool_norm: workspace too small (%zu < %zu)", ws_bytes, need); return TSIM_ERR_WORKSPACE; }
  int* counter = (int*)ws;
  if (a.S > 1) {
    a.ticket = (int*)((char*)ws + 256);
    a.partial = (float*)((char*)ws + 256 + ticket_al);
  }
  // the item counter and the tickets are adjacent: one memset
  TSIM_CUDA(cudaMemsetAsync(ws, 0, 256 + (a.S > 1 ? (size_t)B * sizeof(int) : 0), st));
  const int esz = dtype_size(tok_dt);
  const int vec = 16 / esz;
  const bool vec_ok = (D % vec == 0) && (tok_stride_b % vec == 0) && (tok_stride_l % vec == 0) &&
                      (((uintptr_t)tok & 15) == 0);
  const int mode = pool_mode();
  if (mode != 1 && pool_warp_ok(a, esz, vec_ok)) {
    switch (tok_dt) {
      case TSIM_F32: return launch_pool_warp<TSIM_F32>(a, counter, st);
      case TSIM_F16: return launch_pool_warp<TSIM_F16>(a, counter, st);
      default: return launch_pool_warp<TSIM_BF16>(a, counter, st);
    }
  }
  switch (tok_dt) {
    case TSIM_F32: return vec_ok ? launch_pool<TSIM_F32, 4>(a, st) : launch_pool<TSIM_F32, 1>(a, st);
    case TSIM_F16: return vec_ok ? launch_pool<TSIM_F16, 8>(a, st) : launch_pool<TSIM_F16, 1>(a, st);
    default: return vec_ok ? launch_pool<TSIM_BF16, 8>(a, st) : launch_pool<TSIM_BF16, 1>(a, st);
  }
}

extern "C" int tsim_row_inv_norm(const void* x, int dt, int64_t N, int64_t D, int64_t stride,
                                 float* out, void* stream) {
  TSIM_CHECK_ARG(N >= 0 && D > 0 && stride >= D, "row_inv_norm: bad shape");
  if (N == 0) return TSIM_OK;
  TSIM_CHECK_ARG(x && out, "row_inv_norm: null pointer");
  return launch_row_inv_norm(x, dt, N, D, stride, out, (cudaStream_t)stream);
}
