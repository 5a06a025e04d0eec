/*
 * hyptok_b200.h -- C ABI of the B200 (sm_100a) merge-loop hot path of HypTokenizer.
 *
 * The reference (sangaprabhav/HypTokenizer) is pure Python on torch tensors and has no
 * FFI of its own; the boundary it offers is the Python API of embedding/lorentz_model.py
 * and tokenizer/{hyperbolic_merge,fast_hyperbolic_merge,frequency_aware_hyperbolic_merge}.py.
 * This library is what sits UNDER that API: every entry point names the reference
 * function(s) it replaces.  INTEGRATION.md shows the ctypes stub a maintainer adds.
 *
 * Conventions
 *   - all pointers are DEVICE pointers owned by the caller unless marked `host`
 *   - no hidden allocation: workspaces are passed in, sizes come from *_workspace_bytes()
 *   - `stream` is a cudaStream_t passed as void*; every call is stream-ordered, none syncs
 *   - return value: HYP_OK or a negative HYP_ERR_*; hyp_last_error() gives the text
 *   - fp32 everywhere; rows are D = d+1 floats (time component first), `ld*` are in floats
 *   - `semantics`: HYP_SEM_REFERENCE = the shipped arithmetic (negated product, 0*NaN mask),
 *                  HYP_SEM_LORENTZ   = corrected geometry
 *   - arithmetic on the pre-clamp Minkowski product follows ATen's CPU summation order
 *     (SURVEY.md Appendix D), so it is bit-identical to the reference's torch CPU value
 */
#ifndef HYPTOK_B200_H
#define HYPTOK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HYP_ABI_VERSION 1

#define HYP_OK 0
#define HYP_ERR_ARG (-1)         /* bad argument (NULL, negative size, unsupported D)          */
#define HYP_ERR_CUDA (-2)        /* CUDA runtime / driver error, see hyp_last_error()           */
#define HYP_ERR_FULL (-3)        /* vocabulary table full: reference raises ValueError          */
#define HYP_ERR_UNSUPPORTED (-4) /* device is not sm_100 or a feature is unavailable            */
#define HYP_ERR_WORKSPACE (-5)   /* workspace too small                                         */

#define HYP_SEM_REFERENCE 0
#define HYP_SEM_LORENTZ 1

#define HYP_MAX_D 1025 /* largest supported row length D = d+1 */
#define HYP_MAX_PEERS 8 /* GPUs of one NVSwitch domain a hyp_ctx can span */

int hyp_abi_version(void);
const char *hyp_last_error(void);
/* 0 if the current device is compute capability 10.x, HYP_ERR_UNSUPPORTED otherwise. */
int hyp_check_device(void);

/* ---- K1: fused pointwise Lorentz ops (embedding/lorentz_model.py) ----------------------
 * n independent row pairs.  Operand r of x is at x + r*ldx (ldx == 0 broadcasts one row). */

/* minkowski_dot, lorentz_model.py:14-25.  out[n] */
int hyp_minkowski_dot(const float *x, int64_t ldx, const float *y, int64_t ldy, float *out,
                      int64_t n, int D, void *stream);
/* distance, lorentz_model.py:122-138: acosh(clamp(sgn*<x,y>, 1.0f)) / sqrt(c).  out[n] */
int hyp_distance(const float *x, int64_t ldx, const float *y, int64_t ldy, float *out,
                 int64_t n, int D, float c, int semantics, void *stream);
/* Backward of hyp_distance (what torch autograd derives for lorentz_model.py:122-138; used by
 * multimodal/contrastive_loss.py:63-95 and by training loops that differentiate `distance`):
 * grad_x[n][D], grad_y[n][D] (either may be NULL) from grad_out[n].  The clamp passes gradient where
 * sgn*<x,y> >= 1 (inclusive), so the shipped semantics yields zeros, as the reference does. */
int hyp_distance_backward(const float *x, int64_t ldx, const float *y, int64_t ldy,
                          const float *grad_out, float *grad_x, float *grad_y, int64_t n, int D,
                          float c, int semantics, void *stream);
/* log_map, lorentz_model.py:96-119.  out[n][D] (ldo) */
int hyp_log_map(const float *x, int64_t ldx, const float *y, int64_t ldy, float *out,
                int64_t ldo, int64_t n, int D, int semantics, void *stream);
/* exp_map, lorentz_model.py:73-93.  out[n][D] */
int hyp_exp_map(const float *x, int64_t ldx, const float *v, int64_t ldv, float *out,
                int64_t ldo, int64_t n, int D, void *stream);
/* project_to_hyperboloid, lorentz_model.py:41-56.  out[n][D]; out may alias x */
int hyp_project(const float *x, int64_t ldx, float *out, int64_t ldo, int64_t n, int D, float c,
                void *stream);
/* weighted midpoint of tokenizer/hyperbolic_merge.py:323-340 for n index pairs of table E:
 * project(exp_xi(w_j * log_xi(xj))), w_j = len_j/(len_i+len_j) (double, rounded to fp32).
 * `project`==0 gives the un-projected point of frequency_aware_hyperbolic_merge.py:139-141. */
int hyp_midpoint(const float *E, int64_t ldE, const int32_t *idx_i, const int32_t *idx_j,
                 const int32_t *len_i, const int32_t *len_j, float *out, int64_t ldo, int64_t n,
                 int D, float c, int semantics, int project, void *stream);

/* ---- K3: exact re-scoring of index pairs (hyperbolic_merge.py:230-241,
 * fast_hyperbolic_merge.py:320-324, :450-454).  u_out (pre-clamp, pre-sign product) may be NULL. */
int hyp_rescore_pairs(const float *E, int64_t ldE, const int32_t *idx_i, const int32_t *idx_j,
                      float *d_out, float *u_out, int64_t n, int D, float c, int semantics,
                      void *stream);

/* ---- K2 (exact fp32 path): all-pairs Lorentz distance, lorentz_model.py:141-178 ----------- */

/* batch_distance: full (n1, n2) matrix, out[n1][ldo]. */
int hyp_batch_distance(const float *x, int64_t ldx, int64_t n1, const float *y, int64_t ldy,
                       int64_t n2, float *out, int64_t ldo, int D, float c, int semantics,
                       void *stream);

/* Backward of hyp_batch_distance, first half: w[n1][ldw] = grad_out[i][j] * d(distance)/d<x_i,y_j>
 * (the B x B loop of multimodal/contrastive_loss.py:38-45 under autograd).  The row gradients are
 * then two plain GEMMs, grad_x = w @ (y0, -ys), grad_y = w^T @ (x0, -xs). */
int hyp_batch_distance_backward_coef(const float *x, int64_t ldx, int64_t n1, const float *y,
                                     int64_t ldy, int64_t n2, const float *grad_out, int64_t ldg,
                                     float *w, int64_t ldw, int D, float c, int semantics,
                                     void *stream);

/* Result record of the reductions below: the argmin over the key (d, i, j) -- what the
 * reference's stable sort on distance selects (hyperbolic_merge.py:378). i == -1: none. */
typedef struct hyp_best {
  float d;
  int32_t i;
  int32_t j;
  uint32_t count_lo; /* number of pairs with d < threshold, i < j (low / high word) */
  uint32_t count_hi;
  uint32_t pad[3];
} hyp_best; /* 32 bytes */

int64_t hyp_allpairs_workspace_bytes(int64_t n);
/* _find_merge_candidates + sort + [0] (hyperbolic_merge.py:247-269,:378) without materialising
 * the list: argmin over i<j<n of (d,i,j), plus the count of pairs with d < threshold.
 * NaN distances never qualify (`NaN < thr` is False). `best` is device memory. */
int hyp_allpairs_min(const float *E, int64_t ldE, int64_t n, int D, float c, int semantics,
                     float threshold, hyp_best *best, void *workspace, int64_t workspace_bytes,
                     void *stream);
/* The candidate list itself (hyperbolic_merge.py:259-269): writes up to `capacity` records
 * (i, j, d) with i<j, d<threshold in UNSPECIFIED order (the host sorts by (i,j) to get the
 * reference's row-major order) and the total count into *count (device, uint64). */
int hyp_allpairs_emit(const float *E, int64_t ldE, int64_t n, int D, float c, int semantics,
                      float threshold, int32_t *out_i, int32_t *out_j, float *out_d,
                      int64_t capacity, unsigned long long *count, void *stream);

/* ---- the first K candidates in (d, i, j) order without materialising the list: what
 * AdaptiveMergeCache.add_batch keeps (fast_hyperbolic_merge.py:91-95: concat, stable sort on d, truncate to
 * max_size = 10 000) when the list has 5*10^7 entries.  A radix select over the 31 significant bits of d >= 0,
 * driven by the host: histogram passes narrow the K-th distance down to its bit pattern, a per-row count of the
 * pairs AT that distance finds the row where the K-th pair (row-major ties) sits, and one emit with a cut writes
 * the <= K + n survivors, which the host sorts.  Every pass is one sweep of the exact all-pairs kernel.
 *   hyp_allpairs_hist      hist[1 << 12] (uint64, zeroed by the callee): candidates (i < j, d < threshold) whose
 *                          distance bits 30 .. 31 - prefix_len equal `prefix`, binned on the next `bin_bits` bits
 *   hyp_allpairs_row_ties  row_ties[n] (zeroed by the callee): per row i, candidates (i, j) with bits(d) == d_bits
 *   hyp_allpairs_emit_cut  hyp_allpairs_emit restricted to bits(d) < cut_bits, or == cut_bits in rows <= cut_row */
int hyp_allpairs_hist(const float *E, int64_t ldE, int64_t n, int D, float c, int semantics,
                      float threshold, uint32_t prefix, int prefix_len, int bin_bits,
                      unsigned long long *hist, void *stream);
int hyp_allpairs_row_ties(const float *E, int64_t ldE, int64_t n, int D, float c, int semantics,
                          float threshold, uint32_t d_bits, unsigned int *row_ties, void *stream);
int hyp_allpairs_emit_cut(const float *E, int64_t ldE, int64_t n, int D, float c, int semantics,
                          float threshold, uint32_t cut_bits, int64_t cut_row, int32_t *out_i,
                          int32_t *out_j, float *out_d, int64_t capacity, unsigned long long *count,
                          void *stream);

/* Exact per-row nearest neighbours (what the reference asks FAISS for, fast_hyperbolic_merge.py:
 * 301-304, hyperbolic_merge.py:217, but in the true Lorentz distance instead of Klein-L2 and
 * without sampling): for every row i of the shard [row0, row0+nrows) the k smallest (d(i,j), j)
 * over all j in [0,n), j != i, ascending; ties break on j.  out_idx/out_d are [nrows][k]
 * (-1 / +inf padded when n-1 < k).  k <= 64.  Rows are independent: one shard per GPU. */
int hyp_allpairs_topk(const float *E, int64_t ldE, int64_t n, int64_t row0, int64_t nrows, int D,
                      float c, int semantics, int k, int32_t *out_idx, float *out_d, void *stream);

/* ---- K2 (tensor-core path): the same per-row top-k through a tcgen05 TF32 Gram GEMM ----------
 * (TMA-fed tiles, TMEM accumulators; the signed Minkowski product, time-like term included as hi/lo
 * TF32 parts of x0, is ONE MMA chain, so the accumulator is the pre-clamp value itself), used as a
 * certified FILTER: two GEMM passes bound each row's k-th best and collect a provable superset of
 * the exact top-k, which is re-scored in fp32 (ATen order) and sorted.  Rows whose candidate buffers
 * overflow (massive ties, e.g. the shipped semantics where every distance is 0) are recomputed by the
 * exact kernel inside the same call (device-driven, no host round trip), so the output is always
 * bit-identical to hyp_allpairs_topk; row_flags[r] != 0 marks the rows that took that route.
 * k <= 32, d <= 124.  workspace 256-B aligned. */
int64_t hyp_gram_topk_workspace_bytes(int64_t n, int64_t nrows, int D);
int hyp_gram_topk(const float *E, int64_t ldE, int64_t n, int64_t row0, int64_t nrows, int D, float c,
                  int semantics, int k, int32_t *out_idx, float *out_d, int32_t *row_flags,
                  void *workspace, int64_t workspace_bytes, void *stream);

/* ---- K8: multi-GPU context and the sharded all-pairs / top-k (BASELINE configs[2]) --------
 * The reference has no multi-GPU path; this is the north star's "rows shard across the GPUs of one box,
 * per-shard top-k merged by an all-gather".  A hyp_ctx owns, on every rank, a double-buffered gather buffer
 * of `world` slots (cudaMalloc) that every peer maps through CUDA IPC, so a kernel on rank r stores its
 * results straight into all ranks' buffers over NVLink; the remaining collective is one barrier kernel
 * (release flag to every peer, acquire spin on the local flags, bounded: a missing peer sets the status
 * word instead of hanging).  One host thread per ctx, one process per GPU, all ranks call in the same order.
 *   hyp_ctx_create   on the current device; slot_bytes = the most one rank contributes to one gather,
 *                    a multiple of 256
 *   hyp_ctx_export   this rank's 64-byte IPC handle (host memory), to be exchanged by the caller
 *                    (torch.distributed / MPI / files: the library does no host-side communication)
 *   hyp_ctx_connect  handles of all ranks, rank-major (host memory, world * 64 bytes); world == 1 needs none
 *   hyp_ctx_status   0, or 1 after a barrier timed out (synchronous copy) */
typedef struct hyp_ctx hyp_ctx;
#define HYP_IPC_HANDLE_BYTES 64
int hyp_ctx_create(hyp_ctx **ctx, int rank, int world, int64_t slot_bytes);
int hyp_ctx_export(hyp_ctx *ctx, void *handle_out);
int hyp_ctx_connect(hyp_ctx *ctx, const void *handles);
int hyp_ctx_status(hyp_ctx *ctx, int *status_out);
int hyp_ctx_destroy(hyp_ctx *ctx);
/* All-gather of one slot per rank: `local` (device, 16-byte aligned, bytes % 16 == 0, bytes <= slot_bytes) is
 * stored into slot `rank` of every rank's buffer; *gathered = this rank's buffer (device pointer, slot g at
 * g * slot_bytes), complete for every rank once the call's work on `stream` is done.  It stays valid until
 * the second next gather on this context. */
int hyp_allgather_topk(hyp_ctx *ctx, const void *local, int64_t bytes, void **gathered, void *stream);
/* hyp_gram_topk over this rank's row shard [rank * per, ...), per = ceil(n / world), fused with the
 * all-gather: the finishing kernels write each row's k records {int32 idx, float d} into every rank's buffer.
 * *gathered: slot g (at g * slot_bytes) = [per][k] records of table rows g * per + r (rows >= n are
 * unspecified); with slot_bytes == per * k * 8 the buffer is the dense [world * per][k] table.  Needs
 * slot_bytes >= per * k * 8.  row_flags[per], workspace as for hyp_gram_topk(n, per, D). */
int hyp_gram_topk_allgather(hyp_ctx *ctx, const float *E, int64_t ldE, int64_t n, int D, float c,
                            int semantics, int k, int32_t *row_flags, void *workspace,
                            int64_t workspace_bytes, void **gathered, void *stream);

/* ---- K4/K5: incremental merge loop (hyperbolic_merge.py:309-412, the loop of
 * scripts/train_hyperbolic_tokenizer.py:236-283, fast_hyperbolic_merge.py:467-576) ------------
 * Device-resident state; the loop never leaves the GPU.  Because the reference never removes
 * a merged pair and never changes an existing row, argmin over all pairs after appending row n
 * equals min(previous argmin, argmin over pairs (i, n)) -- SURVEY.md section 0.3. */
typedef struct hyp_merge_state {
  double threshold;   /* merge_threshold: a Python float in the reference. Compared as the
                         reference does: `d < (float)thr` when n > 100 (tensor `<`,
                         hyperbolic_merge.py:262), `(double)d < thr` when n <= 100 (:288)   */
  int32_t n;          /* current vocabulary size                                   */
  int32_t capacity;   /* max_vocab_size (rows allocated in E)                      */
  float best_d;       /* running argmin over all i<j<n of (d, i, j); i == -1: none */
  int32_t best_i;
  int32_t best_j;
  int32_t steps_done; /* merges appended by the last hyp_merge_steps call          */
  int32_t stop;       /* 0 ran max_steps, 1 no candidate below threshold, 2 table full */
  int32_t pad;
} hyp_merge_state;    /* 40 bytes */

typedef struct hyp_merge_record {
  int32_t i;
  int32_t j;
  float d;
  int32_t n_new; /* row index the merged token received */
} hyp_merge_record;

int64_t hyp_merge_workspace_bytes(void);
/* Fill *state on the device from an argmin record produced by hyp_allpairs_min (no host round
 * trip between the initial all-pairs search and the loop). */
int hyp_merge_state_init(hyp_merge_state *state, const hyp_best *best, int32_t n, int32_t capacity,
                         double threshold, void *stream);
/* Runs up to `max_steps` merges on the device: pick state.best if best_d < threshold,
 * append midpoint row n (lengths in `len`, len[n] = len[i]+len[j]), score row n against rows
 * 0..n-1, fold into the running argmin.  Appends one record per merge to `log`
 * (capacity >= max_steps).  `threshold_every`/`threshold_mul`: every `threshold_every` steps
 * (step > 0, counted from `step0`) threshold *= threshold_mul, as the reference loops do
 * (x1.05 / x1.1 per 1000); pass 0 to disable.  Stops early when nothing is below threshold
 * (state->stop = 1) or the table is full (stop = 2, the reference's ValueError).
 * `capacity_hint` = an upper bound, known to the host, on the rows the table can reach during this
 * call (min(state->capacity, n + max_steps); the state itself lives on the device): when that many
 * rows fit in the shared memory of one persistent CTA per SM the table-resident kernel runs; up to
 * four times that capacity the same kernel keeps the first rows resident and scores the rest from
 * L2; beyond, the L2-streaming kernel runs.  All produce identical bits.
 * A call on a state whose `stop` is already non-zero does nothing (steps_done = 0, stop kept), so a
 * host may queue a long run as several calls (advancing `log` and `step0`) without a round trip in
 * between and read each call's log while the next one runs. */
int hyp_merge_steps(float *E, int64_t ldE, int32_t *len, int D, float c, int semantics,
                    hyp_merge_state *state, hyp_merge_record *log, int32_t max_steps,
                    int32_t step0, int32_t threshold_every, double threshold_mul,
                    int32_t capacity_hint, void *workspace, int64_t workspace_bytes, void *stream);
/* One row against the table, top-k form (K4): the k smallest (d(E[i], q), i) over i < n, i != exclude_row (-1: none),
 * ascending, ties on i; q = D floats in device memory (e.g. E + row * ldE).  What one query of the reference's FAISS
 * index returns (fast_hyperbolic_merge.py:301-304), in the true Lorentz distance; bit-identical to the row of
 * hyp_allpairs_topk.  out_idx / out_d [k], (-1, +inf) padded.  k <= 64.  workspace: 8 n bytes. */
int64_t hyp_gemv_topk_workspace_bytes(int64_t n);
int hyp_gemv_topk(const float *E, int64_t ldE, int64_t n, const float *q, int64_t exclude_row, int D, float c,
                  int semantics, int k, int32_t *out_idx, float *out_d, void *workspace,
                  int64_t workspace_bytes, void *stream);
/* One row against the table: argmin over i<n of (d(E[i], E[row]), i) and count below
 * threshold; the single-step building block (K4). */
int hyp_row_min(const float *E, int64_t ldE, int64_t n, int64_t row, int D, float c, int semantics,
                float threshold, hyp_best *best, void *workspace, int64_t workspace_bytes,
                void *stream);

/* ---- K7: frequency-aware scoring, inner loop (frequency_aware_hyperbolic_merge.py:114-166) ---
 * For each of C candidates: the UN-projected weighted midpoint of rows (idx_i, idx_j) (:139-141)
 * and its distance to the S rows sample[c*S .. c*S+S) (:149-153); out[C][S].  The sample indices
 * are drawn on the host with the reference's own torch.randperm calls. */
int hyp_coherence_distances(const float *E, int64_t ldE, const int32_t *idx_i, const int32_t *idx_j,
                            const int32_t *len_i, const int32_t *len_j, const int32_t *sample, int S,
                            float *out, int64_t C, int D, float c, int semantics, void *stream);

/* The whole candidate score of frequency_aware_hyperbolic_merge.py:114-199 on the device, float64 in the reference's
 * operation order (numpy's pairwise mean; exp from CUDA's libm, <= 1 ulp from numpy's):
 *   coherence[c] = 1 / (1 + exp(mean_s d(mid_c, E[sample[c][s]]) - threshold)) over the samples != i, j (0 if none),
 *   score[c]     = alpha / (1 + dist[c]) + beta * freq_score[c] + gamma * coherence[c].
 * dist[C]: the candidates' distances (fp32, as hyp_allpairs_emit / hyp_rescore_pairs give them); freq_score[C]:
 * log1p(f) / log1p(f_max) from the host (a lookup by token strings).  coherence may be NULL.  S <= 64. */
int hyp_score_candidates(const float *E, int64_t ldE, const int32_t *idx_i, const int32_t *idx_j,
                         const int32_t *len_i, const int32_t *len_j, const int32_t *sample, int S,
                         const float *dist, const double *freq_score, double alpha, double beta,
                         double gamma, double threshold, double *score, double *coherence, int64_t C,
                         int D, float c, int semantics, void *stream);

/* ---- K6: pair counting (frequency_aware_hyperbolic_merge.py:92-112) ----------------------- */
/* Adjacent code-point pairs inside each `line.strip()` of a UTF-8 byte stream (16-byte aligned),
 * lines split at '\n' / '\r' (text-mode universal newlines), never across lines; strip() removes
 * every str.isspace() code point.  ASCII pairs land in ascii_counts[a*128+b] (uint64[16384]); all
 * other pairs in an open-addressing table: hash_keys[h] = (cp_a << 32 | cp_b), 0xFFFF.. = empty,
 * hash_vals[h] = count; hash_capacity must be a power of two.  The callee zeroes/initialises all
 * outputs.  *overflow (device int) != 0 if the table filled up.  Input must be valid UTF-8 (the
 * reference would raise UnicodeDecodeError otherwise).  Two kernels with identical results serve the
 * call, chosen on the device by the stream's alphabet (no host round trip): a one-CTA kernel samples 16 KiB of the
 * stream, ranks the ASCII bytes by frequency (the counting kernels keep private counters for the 27 most frequent) and
 * writes its verdict and the rank table into a 16 KiB per-GPU scratch (64 slots for calls in flight) the library
 * allocates on first use -- the one hidden allocation of this ABI. */
int hyp_pair_count(const uint8_t *text, int64_t n_bytes, unsigned long long *ascii_counts,
                   unsigned long long *hash_keys, unsigned long long *hash_vals,
                   int64_t hash_capacity, int *overflow, void *stream);

/* General path of the same count (frequency_aware_hyperbolic_merge.py:92-112 over the output of a tokenize()
 * WITH merge rules, hyperbolic_merge.py:414-446): adjacent token-id pairs within each text, over the
 * [offsets[t], offsets[t] + n_tokens[t]) layout hyp_apply_merges writes (`tokens` has n_slots = offsets[n_texts]
 * int32 slots).  64-bit keys (first << 32 | second, ids as uint32), device radix sort, run-length encode:
 * out_keys[i] ascending (by the ids biased with 2^31, i.e. negative "character without id" tokens first),
 * out_counts[i]; *n_unique (device int64) = number of distinct pairs, which may exceed `capacity` (then only
 * the first `capacity` were written: call again with more room).  workspace 256-byte aligned. */
int64_t hyp_pair_count_sorted_workspace_bytes(int64_t n_slots);
int hyp_pair_count_sorted(const int32_t *tokens, const int64_t *offsets, const int32_t *n_tokens,
                          int64_t n_texts, int64_t n_slots, unsigned long long *out_keys,
                          unsigned long long *out_counts, int64_t capacity, int64_t *n_unique,
                          void *workspace, int64_t workspace_bytes, void *stream);

/* ---- batched tokenize (tokenizer/hyperbolic_merge.py:414-446; what scripts/benchmark_efficiency.py
 * :58-94 measures) -------------------------------------------------------------------------------
 * n_texts UTF-8 texts, text t = bytes [offsets[t], offsets[t+1]).  Tokens are host-assigned symbol ids
 * (one per distinct string in a rule or single-character vocabulary entry): ascii_sym[128] maps ASCII
 * code points to ids (-1: none), (cp_sorted, cp_sym)[n_cp] the other code points (sorted).  Rules live
 * in an open-addressing table: rule_keys[h] = (a << 32 | b), 0xFFFF.. = empty, rule_vals[h] = merged id;
 * capacity a power of two; duplicates already resolved last-wins.  Output: tokens[offsets[t] ..
 * offsets[t] + n_tokens[t]) = the text's final token ids (characters without an id as -(cp+1)).
 * `tokens` needs one int32 per input byte. */
int hyp_apply_merges(const uint8_t *text, const int64_t *offsets, int64_t n_texts,
                     const int32_t *ascii_sym, const uint32_t *cp_sorted, const int32_t *cp_sym,
                     int32_t n_cp, const unsigned long long *rule_keys, const int32_t *rule_vals,
                     int64_t rule_capacity, int32_t *tokens, int32_t *n_tokens, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HYPTOK_B200_H */
