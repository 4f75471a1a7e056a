/*
 * tsim.h -- C ABI of the B200-native semantic-search hot path (libtsim.so).
 *
 * The reference (cr1m5onk1ng/text_similarity) is pure Python and has no FFI of its own;
 * its boundary for this path is the Python class surface.  Each entry point below names
 * the reference lines whose device work it replaces.  The Python mirror of that surface
 * (text_similarity_b200/, re-exported under the reference's module paths in src/) binds
 * these symbols with ctypes; INTEGRATION.md shows the stub a maintainer of the reference
 * would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless it says host;
 *   - the library never allocates or frees caller-visible memory: outputs and workspace
 *     are caller-allocated (workspace size from the *_workspace_bytes functions);
 *   - every call is asynchronous on the CUDA stream passed as `stream` (a cudaStream_t cast
 *     to void*; NULL = legacy default stream); no hidden synchronisation;
 *   - return value: TSIM_OK (0) or a negative TSIM_ERR_* code; tsim_last_error() returns a
 *     thread-local human-readable message for the last failing call on this thread;
 *   - there is no CPU fallback: an unsupported argument is an error.
 */
#ifndef TSIM_H_
#define TSIM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSIM_ABI_VERSION 2

/* element types */
#define TSIM_F32 0
#define TSIM_F16 1
#define TSIM_BF16 2
#define TSIM_E4M3 3 /* float8_e4m3fn */
/* attention-mask element types */
#define TSIM_I64 10
#define TSIM_I32 11
#define TSIM_U8 12 /* also torch.bool */

/* status codes */
#define TSIM_OK 0
#define TSIM_ERR_INVALID_ARG (-1)
#define TSIM_ERR_UNSUPPORTED (-2)
#define TSIM_ERR_MISALIGNED (-3)
#define TSIM_ERR_WORKSPACE (-4)
#define TSIM_ERR_CUDA (-5)

/* search modes */
#define TSIM_MODE_AUTO 0   /* tensor-core path where supported, exact scan otherwise */
#define TSIM_MODE_EXACT 1  /* force the float64 exact-scan path for every query */
#define TSIM_MODE_TENSOR 2 /* require the tcgen05 path (error if unsupported) */

int tsim_version(void);
const char* tsim_last_error(void);

/* ------------------------------------------------------------------------------------
 * K1: fused masked mean-pool + L2-normalise + cast.
 * Replaces AvgPoolingStrategy.forward, reference src/modules/modules.py:158-171 (same
 * arithmetic inlined at src/models/sentence_encoder.py:35-38):
 *     out[b,:] = sum_l tok[b,l,:] * mask[b,l] / max(sum_l mask[b,l], 1e-9)
 * and, when `normalize` != 0, the x / max(||x||, 1e-8) that F.cosine_similarity
 * (src/pipeline/search_pipeline.py:77) and cos_sim (src/utils/metrics.py:99-100) apply
 * inside the similarity, hoisted here so the corpus is stored unit-norm.
 *
 *   tok   [B, L, D]  TSIM_F32 / F16 / BF16, strides in ELEMENTS (innermost contiguous)
 *   mask  [B, L]     TSIM_I64 / I32 / U8 / F32, row stride in elements
 *   out   row b is written at out + out_rows[b] * out_stride (out_rows may be NULL = b),
 *         as TSIM_F32 / BF16 / E4M3.  E4M3 rows are stored times a per-row power of two
 *         chosen so the largest element is near 2^7; cosine is scale free and the factor
 *         is folded into out_inv_norm.
 *   out_inv_norm [B] float, nullable: 1 / max(||row as stored||, 1e-8), indexed like out rows.
 *   ws    tsim_pool_workspace_bytes(B, L, D) bytes.
 * ---------------------------------------------------------------------------------- */
size_t tsim_pool_workspace_bytes(int64_t B, int64_t L, int64_t D);
int tsim_pool_norm(const void* tok, int tok_dt, const void* mask, int mask_dt,
                   int64_t B, int64_t L, int64_t D,
                   int64_t tok_stride_b, int64_t tok_stride_l, int64_t mask_stride_b,
                   void* out, int out_dt, int64_t out_stride, const int64_t* out_rows,
                   float* out_inv_norm, int normalize,
                   void* ws, size_t ws_bytes, void* stream);

/* 1 / max(||x[i,:]||, 1e-8) for every row of a stored matrix (the norm pass of
 * F.cosine_similarity, search_pipeline.py:77, done once per corpus instead of once per
 * query).  x [N, D] TSIM_F32/F16/BF16/E4M3, row stride in elements. */
int tsim_row_inv_norm(const void* x, int dt, int64_t N, int64_t D, int64_t stride,
                      float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * K2 + K3: exact cosine top-k of Q queries against N corpus rows.
 * Replaces the per-query loop of SentenceMiningPipeline._search, reference
 * src/pipeline/search_pipeline.py:73-79 (F.cosine_similarity :77 + torch.topk :78), and
 * cos_sim + argmax, src/utils/metrics.py:99-101,477.
 *
 *   q      [Q, D] q_dt, row stride q_stride elements      (any norm; not modified)
 *   corpus [N, D] c_dt, row stride c_stride elements      (any norm; not modified)
 *   corpus_inv_norm [N] float = tsim_row_inv_norm(corpus) (from K1 or the call above);
 *          may be NULL, then it is computed into the workspace on every call.
 *   k      results per query, 1 <= k <= 1024 (k <= 100 on the tensor-core path).
 *   idx_base           added to corpus row numbers in out_idx (contiguous row shards).
 *   exclude_self_base  >= 0: corpus row (idx_base + j) == exclude_self_base + query number
 *                      is skipped (all-pairs mining); -1: off.
 *   out_score   [Q, k] float   cosine, best first
 *   out_score64 [Q, k] double  nullable; the float64 value out_score was rounded from
 *                      (carry it through shard merges so ranking stays exact)
 *   out_idx     [Q, k] int64   idx_base + row; ties by lower index; -1 past the last row
 *   out_flags   [Q]    int32   nullable; diagnostic: which stage answered the query -- 0 the first
 *                      tensor pass, 2 the wide (112-candidate) retry pass, 1 the float64 scan
 *
 * Result definition (what is "exact"): rows are ranked by the float64 cosine of the
 * STORED values, dot / (max(||q||,1e-8) * max(||c||,1e-8)), descending, ties by ascending
 * index.  The tensor-core pass only nominates candidates; a per-query safety check proves
 * no row outside the candidates can be in the top-k, otherwise the query is recomputed by
 * a float64 scan of the whole shard.  Supported: TSIM_MODE_TENSOR needs q_dt == c_dt ==
 * TSIM_BF16 (D % 8 == 0) or TSIM_E4M3 (D % 16 == 0), 16-byte aligned rows, k <= 100;
 * everything else runs the exact scan.
 * ---------------------------------------------------------------------------------- */
size_t tsim_search_workspace_bytes(int64_t Q, int64_t N, int64_t D, int k,
                                   int q_dt, int c_dt, int mode);
int tsim_search_topk(const void* q, int q_dt, int64_t q_stride,
                     const void* corpus, int c_dt, int64_t c_stride,
                     const float* corpus_inv_norm,
                     int64_t Q, int64_t N, int64_t D, int k,
                     int64_t idx_base, int64_t exclude_self_base, int mode,
                     float* out_score, double* out_score64, int64_t* out_idx,
                     int32_t* out_flags,
                     void* ws, size_t ws_bytes, void* stream);

/* The same search for fp32 / fp16 rows at tensor-core speed: the candidate pass reads bf16 SHADOWS of the
 * queries and the corpus (plain element-wise casts, made once per corpus by the caller; any norm),
 * 112 candidates per query are nominated, and the survivors are re-scored in float64 on the ORIGINAL rows,
 * so the result is the one tsim_search_topk defines on (q, corpus) -- same indices, same score bits.  The
 * completeness proof uses the rounding bound of the shadows (6e-3 instead of 5e-5); queries it cannot cover
 * (many rows within ~1e-2 of the k-th best) are answered by the float64 scan, as usual.  Needs D % 8 == 0,
 * 16-byte aligned shadow rows and k <= 24; otherwise the whole call runs the exact scan.
 *   q_shadow [Q, D], corpus_shadow [N, D]  shadow_dt = TSIM_BF16, row strides in elements
 *   shadow_inv_norm [N] float = tsim_row_inv_norm(corpus_shadow), may be NULL (computed per call)
 *   ws     tsim_search_shadow_workspace_bytes(Q, N, D, k, shadow_dt) bytes. */
size_t tsim_search_shadow_workspace_bytes(int64_t Q, int64_t N, int64_t D, int k, int shadow_dt);
int tsim_search_topk_shadow(const void* q, int q_dt, int64_t q_stride,
                            const void* corpus, int c_dt, int64_t c_stride,
                            const void* q_shadow, int64_t qs_stride,
                            const void* corpus_shadow, int64_t cs_stride, int shadow_dt,
                            const float* shadow_inv_norm,
                            int64_t Q, int64_t N, int64_t D, int k,
                            int64_t idx_base, int64_t exclude_self_base,
                            float* out_score, double* out_score64, int64_t* out_idx,
                            int32_t* out_flags,
                            void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Plan handles ("Ownership": the C side may cache TMA descriptors keyed by (ptr, shape) in
 * an opaque handle the Python side owns and destroys).  A handle fixes (Q, N, D, k, dtypes, mode) on the
 * CURRENT device: the launch plan is made once, and the cuTensorMapEncodeTiled descriptors of every array
 * it has seen (queries, corpus, shadows; up to 16, least recently used replaced) are kept, so a repeated
 * tsim_plan_search over the same arrays does no planning and encodes nothing.  Same result, same
 * workspace contract as tsim_search_topk / tsim_search_topk_shadow (reference lines replaced: the same,
 * src/pipeline/search_pipeline.py:73-79).  A handle is not thread-safe: one search at a time per handle.
 *   shadow_dt  TSIM_BF16: the plan of tsim_search_topk_shadow (q_dt / c_dt = dtypes of the ORIGINAL rows,
 *              corpus_inv_norm = the shadow's, q_shadow / corpus_shadow required); -1: no shadow (the shadow
 *              arguments of tsim_plan_search are ignored).
 *   tsim_plan_create returns NULL on error (tsim_last_error() says why).
 * ---------------------------------------------------------------------------------- */
typedef struct tsim_plan tsim_plan_t;
tsim_plan_t* tsim_plan_create(int64_t Q, int64_t N, int64_t D, int k, int q_dt, int c_dt, int mode,
                              int shadow_dt);
/* Plan for fp32 / fp16 rows with a SPLIT bf16 shadow (k <= 100): every element x is carried as hi = bf16(x) and
 * lo = bf16(x - hi).  With Dp = D rounded up to a multiple of 64 (zero padded): corpus_shadow rows are [hi | lo], 2 Dp
 * wide; q_shadow rows are 3 Dp wide and hold, for every 64-element block j, the triple (hi_j, lo_j, hi_j).  One bf16
 * tensor pass of 3 Dp / 64 k-blocks multiplies query block 3 j + r with corpus block hi_j (r < 2) or lo_j (r == 2):
 * qh.ch + ql.ch + qh.cl = q.c to ~1e-5 of ||q|| ||c|| -- tight enough for the ordinary candidate-list proof at
 * k = 100, where the rounded shadow of tsim_search_topk_shadow (error 6e-3) stops at k = 24.  The corpus shadow costs
 * as many bytes as fp32 rows do (the second read of hi_j is an L2 hit).
 * corpus_inv_norm must be 1 / ||row|| of the ORIGINAL rows (tsim_row_inv_norm on them; null: computed per call).
 * Results are defined on, and re-scored in float64 from, the original rows (reference lines replaced: the same,
 * src/pipeline/search_pipeline.py:73-79). */
tsim_plan_t* tsim_plan_create_split_shadow(int64_t Q, int64_t N, int64_t D, int k, int q_dt, int c_dt);
void tsim_plan_destroy(tsim_plan_t* plan);
size_t tsim_plan_workspace_bytes(const tsim_plan_t* plan);
int tsim_plan_search(tsim_plan_t* plan,
                     const void* q, int64_t q_stride, const void* corpus, int64_t c_stride,
                     const float* corpus_inv_norm,
                     const void* q_shadow, int64_t qs_stride,
                     const void* corpus_shadow, int64_t cs_stride,
                     int64_t idx_base, int64_t exclude_self_base,
                     float* out_score, double* out_score64, int64_t* out_idx, int32_t* out_flags,
                     void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * K3 (second pass): merge n_lists candidate lists per query into the top k_out, ranked by
 * (score descending, index ascending); entries with index < 0 are padding.
 * The reference has no merge: its chunk loop overwrites earlier chunks' results
 * (src/pipeline/search_pipeline.py:83,88); this is the repaired
 * intent, and the step after the NCCL all-gather of per-shard results.
 *   sc [Q, n_lists * k_in] double, ix [Q, n_lists * k_in] int64; n_lists * k_in <= 4096.
 * ---------------------------------------------------------------------------------- */
int tsim_merge_topk(const double* sc, const int64_t* ix, int64_t Q, int64_t n_lists,
                    int k_in, int k_out,
                    float* out_score, double* out_score64, int64_t* out_idx, void* stream);

/* The same merge reading its lists in place from any regular layout: entry j of list l of query q sits at
 * [l * list_stride + q * query_stride + j] (elements) of sc and of ix.  tsim_merge_topk is list_stride = k_in,
 * query_stride = n_lists * k_in; the receive buffer of the NCCL all-gather of per-shard results -- n_lists
 * messages of [Q, k] scores followed by [Q, k] rows -- is list_stride = 2 * Q * k, query_stride = k with
 * ix = sc + Q * k: no re-layout copy between the collective and the merge. */
int tsim_merge_topk_strided(const double* sc, const int64_t* ix, int64_t Q, int64_t n_lists, int k_in,
                            int64_t list_stride, int64_t query_stride, int k_out,
                            float* out_score, double* out_score64, int64_t* out_idx, void* stream);

/* Measurement hook (bench.py / profiling only): when both events are non-NULL, the calling
 * thread's following tsim_search_topk calls record `start` immediately before and `stop`
 * immediately after the candidate-pass kernel (tcgen05 search, or the exact scan when that is
 * the whole-call path) on the call's stream, so the dominant kernel can be timed on its own
 * launching stream.  Events are cudaEvent_t cast to void*; pass NULL, NULL to switch off. */
int tsim_set_timing_events(void* start, void* stop);

/* Number of CUDA kernels this library has launched in the process so far (bench.py reports the
 * difference over its timed region as `gpu_launches`). */
uint64_t tsim_launch_count(void);

/* Process-wide counters for tests: out[0] kernels launched, out[1] cuTensorMapEncodeTiled calls, out[2]
 * environment variables read (always 0 in the release library: it reads none), out[3] launch plans made. */
void tsim_debug_counters(uint64_t out[4]);
/* Test hooks for the error bound the completeness proof rests on (tests/test_gpu_eps.py).
 * tsim_debug_eps: the bound of |tensor-core cosine - exact cosine| (units of ||q|| ||c||) the library assumes for
 * rows of width D and element type dt (shadow != 0: for a bf16 shadow of fp32 / fp16 rows).
 * tsim_debug_tensor_pass: ONE tcgen05 candidate pass with cold 112-entry lists and nothing else -- no sample, no
 * threshold, no re-score.  q must hold 128 rows (rows Q..127 are read, their results unused), Q <= 128, both
 * arrays of type dt (TSIM_BF16 / TSIM_E4M3).  Writes *out_lists = L = min(ceil(N / 256), SM count) (host) and
 * out_keys [Q][L][112] packed keys, 0 = empty slot, else (ordered fp32 of dot(q, c) * corpus_inv_norm[c]) << 32 |
 * (0xffffffff - row); list j holds the 112 best rows of tiles j, j + L, ...  thr_scratch: [Q] uint32 scratch. */
float tsim_debug_eps(int64_t D, int dt, int shadow);
int tsim_debug_tensor_pass(const void* q, int64_t q_stride, const void* corpus, int64_t c_stride, int dt,
                           const float* corpus_inv_norm, int64_t Q, int64_t N, int64_t D,
                           uint64_t* out_keys, int64_t* out_lists, uint32_t* thr_scratch, void* stream);
/* Bit 0: built with -DTSIM_EXPERIMENT (environment knobs and in-kernel diagnosis switches compiled in;
 * never the library the package loads by default). */
int tsim_build_flags(void);

#ifdef __cplusplus
}
#endif
#endif /* TSIM_H_ */
