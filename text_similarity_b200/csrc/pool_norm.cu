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
  int dbg;                // TSIM_POOL_DEBUG bits (diagnosis only): 1 = consumers skip the accumulation,
                          // 2 = static round-robin items, 4 = finisher skips its work
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

// ---- K1, streaming variant ---------------------------------------------------------------------
// Persistent CTAs, two per SM, three warp roles around a shared-memory ring of up to 8 stages (a stage =
// one chunk of whole token rows, <= 16 KB; ~90 KB per CTA):
//  * producer warp: takes (sentence, token-split) items from a global counter (CTAs that drew short
//    sentences simply take more), reads an item's mask weights one item AHEAD (the latency hides
//    behind the copies in flight), skips chunks whose tokens are all masked (trailing padding is
//    never read) and moves each remaining chunk of up to 32 tokens -- one contiguous run of the
//    token tensor -- into a ring stage with ONE 1-D bulk async copy (TMA, UBLKCP) that completes on
//    the stage's mbarrier;
//  * 8 consumer warps (thread = 16-byte column x token sub-row): wait on the barrier, accumulate
//    weight * token from shared memory in fp32, hand the stage back; at an item's END stage they
//    drop their partial sums into one of two fold buffers and go straight on to the next item;
//  * finisher warp: folds the sub-rows in a fixed order, divides by the token count, L2-normalises,
//    casts, stores the row and its inverse norm -- with warp shuffles only, off the streaming path.
// HBM-bound: ~180 KB of copies in flight per SM, independent of register pressure.
constexpr int kPsMaxStages = 8;          // ring depth: as many stages as fit in half an SM's shared memory
constexpr int kPsStageBytes = 16 * 1024;  // upper bound of a stage (a chunk = up to 32 whole token rows)
constexpr int kPsConsumers = 256;
constexpr int kPsThreads = kPsConsumers + 64;   // + producer warp + finisher warp
constexpr int kPsMaxTL = 512;    // tokens per item (the launcher raises the split count to keep this)

// kind 0: data chunk; 1: data chunk (ntok may be 0) that also ends item `item` (cnt = sum of its weights);
// 2: no more items
struct PsHdr { float w[32]; int ntok; int kind; float cnt; int pad; long long item; long long pad2; };
struct PsMail { long long item; float cnt; int pad; };   // consumers -> finisher, one per fold buffer

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

template <int DT>
__global__ void __launch_bounds__(kPsThreads, 2) pool_norm_stream_kernel(PoolArgs a, int64_t nitems, int CT, int nst, int stage_bytes, int* counter) {
  extern __shared__ __align__(128) unsigned char ps_raw[];
  constexpr int VEC = SmemVec<DT>::N;
  constexpr int ESZ = 16 / VEC;
  constexpr int NCW = kPsConsumers / 32;                               // consumer warps
  unsigned char* data = ps_raw;                                        // [nst][stage_bytes]
  PsHdr* hdr = (PsHdr*)(ps_raw + (size_t)nst * stage_bytes);           // [nst]
  float* wbuf = (float*)(hdr + nst);                             // [kPsMaxTL] producer-private weights
  const int nvec = (int)(a.D / VEC);
  const int rpi = kPsConsumers / nvec;                                 // token sub-rows per pass (>= 1)
  float* fold = wbuf + kPsMaxTL;                                       // [2][rpi * D] fold buffers
  PsMail* mail = (PsMail*)(fold + 2 * (size_t)rpi * a.D);              // [2]
  uint64_t* bars = (uint64_t*)(mail + 2);                              // full[nst], empty[nst], fold_full[2], fold_empty[2]
  uint64_t* fold_full = bars + 2 * nst;
  uint64_t* fold_empty = fold_full + 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < nst; ++s) {
      ps_mbar_init(ps_smem(&bars[s]), 1);
      ps_mbar_init(ps_smem(&bars[nst + s]), NCW);
    }
    for (int s = 0; s < 2; ++s) { ps_mbar_init(ps_smem(&fold_full[s]), NCW); ps_mbar_init(ps_smem(&fold_empty[s]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp == NCW) {
    // ===================== producer warp =====================
    int stage = 0; uint32_t phase = 0;
    float wn[kPsMaxTL / 32];
    auto next_item = [&]() -> int64_t {     // CTA c starts with item c; further items come from the counter
      int v = 0;
      if (lane == 0) v = atomicAdd(counter, 1);
      return (int64_t)__shfl_sync(0xffffffffu, v, 0) + gridDim.x;
    };
    int64_t static_next = (int64_t)blockIdx.x + gridDim.x;   // dbg bit 2: static round-robin items
    auto load_mask = [&](int64_t item) {
      const int64_t b = item / a.S; const int sp = (int)(item % a.S);
      const int l0 = sp * a.TL, tl = min((int)a.L, l0 + a.TL) - l0;
#pragma unroll
      for (int j = 0; j < kPsMaxTL / 32; ++j)
        wn[j] = (lane + 32 * j < tl) ? mask_value(a.mask, a.mask_dt, b * a.msb + l0 + lane + 32 * j) : 0.f;
    };
    int64_t item = blockIdx.x;
    if (item < nitems) load_mask(item);
    while (item < nitems) {
      const int64_t b = item / a.S; const int sp = (int)(item % a.S);
      const int l0 = sp * a.TL, tl = min((int)a.L, l0 + a.TL) - l0;
      float cnt = 0.f;
      int lend = 0;                                                   // one past the item's last real token
#pragma unroll
      for (int j = 0; j < kPsMaxTL / 32; ++j) {
        if (lane + 32 * j < tl) wbuf[lane + 32 * j] = wn[j];
        cnt += wn[j];
        if (wn[j] != 0.f) lend = lane + 32 * j + 1;
      }
      cnt = warp_sum_f32(cnt);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lend = max(lend, __shfl_xor_sync(0xffffffffu, lend, o));
      __syncwarp();
      int64_t nxt;
      if (a.dbg & 2) { nxt = static_next; static_next += gridDim.x; } else nxt = next_item();
      if (nxt < nitems) load_mask(nxt);                               // in flight during this item's copies
      // chunks up to the last real token (trailing padding is never read); the chunk that holds it
      // also closes the item (kind 1); an item without any real token is a bare END stage
      for (int lc = 0; lc < lend || lc == 0; lc += CT) {
        const int nt = max(0, min(CT, lend - lc));
        const bool last = lc + CT >= lend;
        const float w = lane < nt ? wbuf[lc + lane] : 0.f;
        const unsigned nz = __ballot_sync(0xffffffffu, w != 0.f);
        if (!nz && !last) continue;                                   // all masked: not read at all
        ps_mbar_wait(ps_smem(&bars[nst + stage]), phase ^ 1);
        hdr[stage].w[lane] = w;
        if (lane == 0) {
          hdr[stage].ntok = nt; hdr[stage].kind = last ? 1 : 0;
          hdr[stage].cnt = cnt; hdr[stage].item = item;
        }
        __syncwarp();
        if (lane == 0) {
          const uint32_t bytes = (uint32_t)nt * (uint32_t)a.D * ESZ;
          const uint32_t fb = ps_smem(&bars[stage]);
          if (bytes) {
            ps_mbar_expect_tx(fb, bytes);
            ps_bulk_load(ps_smem(data + (size_t)stage * stage_bytes),
                         (const unsigned char*)a.tok + (b * a.sb + (int64_t)(l0 + lc) * a.D) * ESZ, bytes, fb);
          } else {
            ps_mbar_arrive(fb);
          }
        }
        if (++stage == nst) { stage = 0; phase ^= 1; }
      }
      item = nxt;
    }
    ps_mbar_wait(ps_smem(&bars[nst + stage]), phase ^ 1);
    if (lane == 0) {
      hdr[stage].kind = 2;
      ps_mbar_arrive(ps_smem(&bars[stage]));
    }
    return;
  }

  if (warp == NCW + 1) {
    // ===================== finisher warp =====================
    int fb = 0; uint32_t fphase = 0;   // bit fb = phase of fold_full[fb]
    for (;;) {
      ps_mbar_wait(ps_smem(&fold_full[fb]), (fphase >> fb) & 1u);
      fphase ^= 1u << fb;
      const int64_t item = mail[fb].item;
      float cnt = mail[fb].cnt;
      if (item < 0) break;
      if (a.dbg & 4) { __syncwarp(); if (lane == 0) ps_mbar_arrive(ps_smem(&fold_empty[fb])); fb ^= 1; continue; }
      const int64_t b = item / a.S; const int sp = (int)(item % a.S);
      float* rb = fold + (size_t)fb * rpi * a.D;
      // fold the token sub-rows in a fixed order; pooled sums land in row 0 of the buffer
      for (int d = lane * 4; d < a.D; d += 128) {
        float4 t = *(const float4*)(rb + d);
        for (int rr = 1; rr < rpi; ++rr) {
          const float4 u = *(const float4*)(rb + (size_t)rr * a.D + d);
          t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
        }
        if (a.S > 1) *(float4*)(a.partial + ((int64_t)b * a.S + sp) * a.D + d) = t;
        else *(float4*)(rb + d) = t;
      }
      bool finish = true;
      if (a.S > 1) {
        // the last CTA to finish a split of row b adds the partial sums in a fixed order
        __threadfence();
        __syncwarp();
        int last = 0;
        if (lane == 0) last = (atomicAdd(&a.ticket[b], 1) == a.S - 1);
        last = __shfl_sync(0xffffffffu, last, 0);
        finish = last != 0;
        if (finish) {
          __threadfence();
          for (int d = lane * 4; d < a.D; d += 128) {
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int s2 = 0; s2 < a.S; ++s2) {
              const float4 u = __ldcg((const float4*)(a.partial + ((int64_t)b * a.S + s2) * a.D + d));
              t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
            }
            *(float4*)(rb + d) = t;
          }
          float c = 0.f;
          for (int l = lane; l < a.L; l += 32) c += mask_value(a.mask, a.mask_dt, b * a.msb + l);
          cnt = warp_sum_f32(c);
        }
      }
      if (finish) {
        __syncwarp();
        const float denom = fmaxf(cnt, kPoolEps);  // modules.py:168
        float ss = 0.f, amax = 0.f;
        for (int d = lane * 4; d < a.D; d += 128) {
          float4 t = *(const float4*)(rb + d);
          t.x /= denom; t.y /= denom; t.z /= denom; t.w /= denom;  // modules.py:170
          *(float4*)(rb + d) = t;
          ss = fmaf(t.x, t.x, ss); ss = fmaf(t.y, t.y, ss); ss = fmaf(t.z, t.z, ss); ss = fmaf(t.w, t.w, ss);
          amax = fmaxf(fmaxf(amax, fmaxf(fabsf(t.x), fabsf(t.y))), fmaxf(fabsf(t.z), fabsf(t.w)));
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
        for (int d = lane * 4; d < a.D; d += 128) {
          const float4 t = *(const float4*)(rb + d);
          const float v4[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float st = round_store(a.out, a.out_dt, orow * a.out_stride + d + j, v4[j] * scale);
            ss2 = fmaf(st, st, ss2);
          }
        }
        if (a.out_inv) {
          ss2 = warp_sum_f32(ss2);
          if (lane == 0) a.out_inv[orow] = 1.f / fmaxf(sqrtf(ss2), (float)kCosEps);
        }
      }
      __syncwarp();
      if (lane == 0) ps_mbar_arrive(ps_smem(&fold_empty[fb]));
      fb ^= 1;
    }
    return;
  }

  // ===================== consumers =====================
  const bool active = tid < rpi * nvec;
  const int r = tid / nvec, v = tid - r * nvec;
  int stage = 0; uint32_t phase = 0;
  int fb = 0; uint32_t ephase = 0;     // fold buffer to fill next; bit fb = phase of fold_empty[fb]
  for (;;) {
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
    float cnt = 0.f;
    int64_t item = -1;
    int kind;
    for (;;) {
      ps_mbar_wait(ps_smem(&bars[stage]), phase);
      const PsHdr* h = hdr + stage;
      kind = h->kind;
      if (kind != 2 && active && !(a.dbg & 1)) {
        const int ntok = h->ntok;
        const uint4* src = (const uint4*)(data + (size_t)stage * stage_bytes) + v;
        for (int t = r; t < ntok; t += 2 * rpi) {
          // two tokens per step, loads first; a masked token's values must not reach the sum (they may be Inf/NaN)
          const int t1 = t + rpi;
          const bool ok1 = t1 < ntok;
          const float w0 = h->w[t], w1 = ok1 ? h->w[t1] : 0.f;
          const uint4 raw0 = src[(size_t)t * nvec];
          const uint4 raw1 = src[(size_t)(ok1 ? t1 : t) * nvec];
          float x0[VEC], x1[VEC];
          SmemVec<DT>::cvt(raw0, x0);
          SmemVec<DT>::cvt(raw1, x1);
          if (w0 != 0.f) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] = fmaf(w0, x0[j], acc[j]);
          }
          if (w1 != 0.f) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] = fmaf(w1, x1[j], acc[j]);
          }
        }
      }
      if (kind == 1) { cnt = h->cnt; item = h->item; }
      __syncwarp();
      if (lane == 0) ps_mbar_arrive(ps_smem(&bars[nst + stage]));
      if (++stage == nst) { stage = 0; phase ^= 1; }
      if (kind != 0) break;
    }
    // ---- item done (or no more items: item = -1): hand the partial sums to the finisher ----
    // wait until the finisher has released this fold buffer (the first use of each is free)
    ps_mbar_wait(ps_smem(&fold_empty[fb]), ((ephase >> fb) & 1u) ^ 1u);
    ephase ^= 1u << fb;
    if (active && kind == 1) {
      float* rb = fold + (size_t)fb * rpi * a.D + (size_t)r * a.D + (size_t)v * VEC;
#pragma unroll
      for (int j = 0; j < VEC; j += 4) *(float4*)(rb + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
    }
    if (tid == 0) { mail[fb].item = item; mail[fb].cnt = cnt; }
    __syncwarp();
    if (lane == 0) ps_mbar_arrive(ps_smem(&fold_full[fb]));
    fb ^= 1;
    if (kind == 2) break;
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
    const __nv_fp8_e4m3* e = (const __nv_fp8_e4m3*)&raw;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) { const float v = float(e[j]); s = fmaf(v, v, s); }
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
  const int64_t mins = (L + kPsMaxTL - 1) / kPsMaxTL;   // <= kPsMaxTL tokens per item (streaming variant)
  if (S < mins) S = mins;
  if (S < 1) S = 1;
  return (int)S;
}

// The streaming variant needs token rows that are contiguous within a sentence (one bulk copy per
// chunk), 16-byte granularity, a row of at most one stage and at most 256 16-byte columns.
bool pool_stream_ok(const PoolArgs& a, int esz, bool vec_ok) {
  const char* v1 = getenv("TSIM_POOL_V1");   // experiment knob: force the register-staged kernel
  if (v1 && v1[0] == '1') return false;
  const int64_t rowb = a.D * esz;
  return vec_ok && a.sl == a.D && rowb <= kPsStageBytes && rowb / 16 <= kPsConsumers && a.TL <= kPsMaxTL;
}

template <int DT>
int launch_pool_stream(const PoolArgs& a, int* counter, cudaStream_t st) {
  constexpr int ESZ = 16 / SmemVec<DT>::N;
  const int nvec = (int)(a.D * ESZ / 16);
  const int rpi = kPsConsumers / nvec;
  int CT = (int)(kPsStageBytes / (a.D * ESZ));
  if (CT > 32) CT = 32;
  const int stage_bytes = (int)((CT * a.D * ESZ + 127) / 128 * 128);
  const size_t fixed = kPsMaxTL * sizeof(float) + 2 * (size_t)rpi * a.D * sizeof(float) + 2 * sizeof(PsMail) +
                       (2 * kPsMaxStages + 4) * sizeof(uint64_t) + 256;
  const size_t budget = 113 * 1024;        // two CTAs per SM
  int nst = (int)((budget - fixed) / (stage_bytes + sizeof(PsHdr)));
  if (nst > kPsMaxStages) nst = kPsMaxStages;
  if (nst < 2) nst = 2;
  const size_t smem = (size_t)nst * (stage_bytes + sizeof(PsHdr)) + fixed;
  TSIM_CUDA(cudaFuncSetAttribute(pool_norm_stream_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t nitems = a.B * a.S;
  const int64_t cap = 2 * (int64_t)device_sm_count();
  const unsigned grid = (unsigned)(nitems < cap ? nitems : cap);
  pool_norm_stream_kernel<DT><<<grid, kPsThreads, smem, st>>>(a, nitems, CT, nst, stage_bytes, counter);
  TSIM_CUDA(cudaGetLastError());
  count_launch();
  return TSIM_OK;
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
  { const char* d = getenv("TSIM_POOL_DEBUG"); a.dbg = d ? atoi(d) : 0; }
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
  if (pool_stream_ok(a, esz, vec_ok)) {
    switch (tok_dt) {
      case TSIM_F32: return launch_pool_stream<TSIM_F32>(a, counter, st);
      case TSIM_F16: return launch_pool_stream<TSIM_F16>(a, counter, st);
      default: return launch_pool_stream<TSIM_BF16>(a, counter, st);
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
