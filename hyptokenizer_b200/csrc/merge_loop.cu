// merge_loop.cu -- K4 (one row against the table) and K5 (device-resident merge loop), sm_100a.
//
// Replaces the per-step "recompute all pairs, build a Python list, sort, take [0], merge" of
// tokenizer/hyperbolic_merge.py:357-412, scripts/train_hyperbolic_tokenizer.py:236-283 and
// tokenizer/fast_hyperbolic_merge.py:467-576 (whose HNSW index and AdaptiveMergeCache exist only
// to make that affordable).  The reference never removes a merged pair and never changes an
// existing row, so after appending row n:
//     argmin_{i<j<=n} (d,i,j) = min( argmin_{i<j<n} (d,i,j) , argmin_{i<n} (d(i,n), i, n) )
// i.e. one new row scored against the table (HBM/L2-bound GEMV + min) per merge.
//
// K5 is ONE cooperative launch for up to max_steps merges: every CTA recomputes the (tiny)
// midpoint row redundantly and bit-identically, CTA 0 appends it, all CTAs scan their share of
// rows, one grid barrier, every CTA reduces the per-CTA minima to the same new state.
// Compiled with -fmad=false.
#include <stdlib.h>

#include "common.cuh"

namespace hyp {

constexpr int kLoopThreads = 512;
constexpr int kLoopWarps = kLoopThreads / 32;
constexpr int kMaxLoopBlocks = 148 * 4;
constexpr int kRowsPerIter = 4;  // rows a warp keeps in flight (independent reduction chains)
constexpr int kXRing = 1024;     // exchange slots of the resident loop (step k uses slot k mod kXRing)

struct LoopWorkspace {
  unsigned int barrier;    // monotonically increasing arrival counter
  unsigned int ticket;     // hyp_row_min: last-block-done
  unsigned int published;  // resident loop: scans whose global appends CTA 0 has fenced (release sequence number)
  unsigned int pad[5];
  long long prof[8];               // resident loop: phase cycle counters of CTA 0 and CTA G-1
  long long mprof[8];              // debug: midpoint sub-phase cycles of CTA 0
  unsigned long long reserved[4];
  struct alignas(16) XSlot {       // resident loop: per-step exchange word, polled with ONE 128-bit load
    unsigned long long key;        //   RED.MIN target: d_bits << 32 | row (kNoKey when nobody improved on `best`)
    unsigned long long cw;         //   ONE 64-bit scalar: arrivals in the low word (never reset: use u of a slot
                                   //   completes at G * u), CTAs that posted a key during this use in the high word
  } xring[kXRing];
  Key slot[2][kMaxLoopBlocks];     // L2 loop: per-CTA minima, double-buffered by step parity
  unsigned long long below[kMaxLoopBlocks];
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void ld_acquire_v2_u64(const void *p, unsigned long long &a, unsigned long long &b) {
  asm volatile("ld.acquire.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void red_release_add_u32(unsigned int *p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void ld_relaxed_v2_u64(const void *p, unsigned long long &a, unsigned long long &b) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(unsigned int *p, unsigned int v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_relaxed_add_u32(unsigned int *p, unsigned int v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void red_relaxed_add_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// Every wait on another CTA is bounded: ~2^24 polls (each an L2 round trip, some 10 s in all) and the kernel traps --
// the launch fails with an error instead of hanging the GPU if the replicated state of the CTAs ever diverged.
constexpr unsigned int kSpinLimit = 1u << 24;
__device__ __forceinline__ void spin_guard(unsigned int &spins) {
  if (++spins > kSpinLimit) __trap();
}

// All CTAs of a cooperative launch (co-resident by construction).
__device__ __forceinline__ void grid_barrier(unsigned int *ctr, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    red_release_add_u32(ctr, 1u);
    unsigned int spins = 0;
    while (ld_acquire_u32(ctr) < target) spin_guard(spins);
  }
  __syncthreads();
}

// Score table rows against `q` (D floats in shared memory): R rows per warp iteration.
// Table reads bypass L1 (ld.global.cg): rows appended by another CTA share cache lines with
// rows this SM has already read.
template <int R>
__device__ __forceinline__ void warp_mdot_rows(const float *__restrict__ E, int64_t ldE, const int64_t *rows,
                                               int nrows, const float *__restrict__ q, int D, int lane,
                                               float *m_out) {
  const int N = D - 1;
  const float *qs = q + 1;
  if (N < 8) {
    for (int a = 0; a < R; ++a) {
      if (a >= nrows) break;
      const float *xs = E + rows[a] * ldE + 1;
      float s = thread_sum_aten([&](int e) { return __fmul_rn(__ldcg(xs + e), qs[e]); }, N);
      m_out[a] = __fsub_rn(__fmul_rn(__ldcg(xs - 1), q[0]), s);
    }
    return;
  }
  const int vs = N >> 3, full = vs >> 2;
  const int c = lane >> 3, l = lane & 7;
  float P[R], acc[R], x0[R];
  const float *xs[R];
#pragma unroll
  for (int a = 0; a < R; ++a) {
    xs[a] = E + rows[a < nrows ? a : 0] * ldE + 1;
    P[a] = 0.f;
    acc[a] = 0.f;
  }
  for (int r = 0; r < full; ++r) {
    const int e = 32 * r + lane;
    const float qe = qs[e];
#pragma unroll
    for (int a = 0; a < R; ++a) P[a] = __fadd_rn(P[a], __fmul_rn(__ldcg(xs[a] + e), qe));
  }
  if (c == 0)
    for (int k = 4 * full; k < vs; ++k) {
      const int e = 8 * k + l;
      const float qe = qs[e];
#pragma unroll
      for (int a = 0; a < R; ++a) P[a] = __fadd_rn(P[a], __fmul_rn(__ldcg(xs[a] + e), qe));
    }
  for (int k = 8 * vs; k < N; ++k) {
    const float qe = qs[k];
#pragma unroll
    for (int a = 0; a < R; ++a) acc[a] = __fadd_rn(acc[a], __fmul_rn(__ldcg(xs[a] + k), qe));
  }
#pragma unroll
  for (int a = 0; a < R; ++a) x0[a] = __ldcg(xs[a] - 1);
#pragma unroll
  for (int a = 0; a < R; ++a) {
    float p1 = __shfl_down_sync(HYP_FULL_MASK, P[a], 8);
    float p2 = __shfl_down_sync(HYP_FULL_MASK, P[a], 16);
    float p3 = __shfl_down_sync(HYP_FULL_MASK, P[a], 24);
    P[a] = __fadd_rn(__fadd_rn(__fadd_rn(P[a], p1), p2), p3);
  }
#pragma unroll
  for (int qn = 0; qn < 8; ++qn)
#pragma unroll
    for (int a = 0; a < R; ++a) acc[a] = __fadd_rn(acc[a], __shfl_sync(HYP_FULL_MASK, P[a], qn));
#pragma unroll
  for (int a = 0; a < R; ++a) m_out[a] = __fsub_rn(__fmul_rn(x0[a], q[0]), acc[a]);
}

// Scan rows [0, n) strided over all warps of the grid; returns this CTA's min key (thread 0).
__device__ __forceinline__ Key block_scan_rows(const float *__restrict__ E, int64_t ldE, int n, int row_new,
                                               const float *q, int D, float sqrt_c, float sgn, float thr_f,
                                               double thr_d, bool cmp_double, unsigned long long *below_out,
                                               Key *s_keys, unsigned long long *s_cnt) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t total_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  Key best = key_none();
  unsigned long long below = 0;
  for (int64_t base = gw * kRowsPerIter; base < n; base += total_warps * kRowsPerIter) {
    int64_t rows[kRowsPerIter];
    int nrows = 0;
#pragma unroll
    for (int a = 0; a < kRowsPerIter; ++a) {
      rows[a] = base + a;
      if (base + a < n) nrows = a + 1;
    }
    float m[kRowsPerIter];
    warp_mdot_rows<kRowsPerIter>(E, ldE, rows, nrows, q, D, lane, m);
#pragma unroll
    for (int a = 0; a < kRowsPerIter; ++a) {
      if (a < nrows && rows[a] != row_new) {
        float d = dist_from_mdot(m[a], sgn, sqrt_c);
        if (d == d) {
          bool lt = cmp_double ? ((double)d < thr_d) : (d < thr_f);
          if (lt) ++below;
          int lo = rows[a] < row_new ? (int)rows[a] : row_new;
          int hi = rows[a] < row_new ? row_new : (int)rows[a];
          Key k{d, lo, hi};
          if (key_less(k, best)) best = k;
        }
      }
    }
  }
  // every lane of a warp holds the same `best`; combine across warps
  if (lane == 0) {
    s_keys[warp] = best;
    s_cnt[warp] = below;
  }
  __syncthreads();
  Key b = key_none();
  if (threadIdx.x == 0) {
    unsigned long long cnt = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      if (key_less(s_keys[w], b)) b = s_keys[w];
      cnt += s_cnt[w];
    }
    *below_out = cnt;
  }
  __syncthreads();
  return b;
}

__global__ void __launch_bounds__(kLoopThreads)
row_min_kernel(const float *__restrict__ E, int64_t ldE, int n, int row, int D, float sqrt_c, float sgn,
               float thr, LoopWorkspace *ws, hyp_best *out) {
  extern __shared__ float smem[];
  __shared__ Key s_keys[kLoopWarps];
  __shared__ unsigned long long s_cnt[kLoopWarps];
  __shared__ int s_last;
  float *q = smem;
  for (int k = threadIdx.x; k < D; k += blockDim.x) q[k] = __ldcg(E + (int64_t)row * ldE + k);
  __syncthreads();
  unsigned long long below = 0;
  Key b = block_scan_rows(E, ldE, n, row, q, D, sqrt_c, sgn, thr, (double)thr, false, &below, s_keys, s_cnt);
  if (threadIdx.x == 0) {
    ws->slot[0][blockIdx.x] = b;
    ws->below[blockIdx.x] = below;
    __threadfence();
    s_last = (atomicAdd(&ws->ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    Key r = key_none();
    unsigned long long cnt = 0;
    for (int q2 = 0; q2 < (int)gridDim.x; ++q2) {
      Key k;
      k.d = __ldcg(&ws->slot[0][q2].d);
      k.i = __ldcg(&ws->slot[0][q2].i);
      k.j = __ldcg(&ws->slot[0][q2].j);
      if (key_less(k, r)) r = k;
      cnt += __ldcg(&ws->below[q2]);
    }
    hyp_best o;
    o.d = r.i < 0 ? __int_as_float(0x7f800000) : r.d;
    o.i = r.i;
    o.j = r.j;
    o.count_lo = (uint32_t)(cnt & 0xffffffffu);
    o.count_hi = (uint32_t)(cnt >> 32);
    o.pad[0] = o.pad[1] = o.pad[2] = 0;
    *out = o;
    ws->ticket = 0;
  }
}

struct LoopParams {
  float *E;
  int64_t ldE;
  int32_t *len;
  int D;
  float c, sqrt_c, sgn;
  int semantics;
  hyp_merge_state *state;
  hyp_merge_record *log;
  int max_steps, step0, thr_every;
  double thr_mul;
  LoopWorkspace *ws;
};

__global__ void __launch_bounds__(kLoopThreads) merge_loop_l2_kernel(const LoopParams p) {
  extern __shared__ float smem[];
  const int D = p.D;
  float *q = smem;           // [D] new row
  float *xi = q + D;         // [D]
  float *xj = xi + D;        // [D]
  float *scratch = xj + D;   // [2D]
  __shared__ Key s_keys[kLoopWarps];
  __shared__ unsigned long long s_cnt[kLoopWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // a launch on a state that has already stopped does nothing (lets a caller queue several launches back to back)
  if (p.state->stop != 0) {
    if (blockIdx.x == 0 && threadIdx.x == 0) p.state->steps_done = 0;
    return;
  }
  // replicated loop state (identical in every CTA)
  int n = p.state->n;
  const int cap = p.state->capacity;
  Key best{p.state->best_d, p.state->best_i, p.state->best_j};
  double thr = p.state->threshold;
  int done = 0, stop = 0;
  unsigned int arrivals = 0;

  for (int k = 0; k < p.max_steps; ++k) {
    const bool cmp_double = n <= 100;  // hyperbolic_merge.py:247 vs :270 (tensor `<` vs Python float `<`)
    const bool have = best.i >= 0 && (cmp_double ? ((double)best.d < thr) : (best.d < (float)thr));
    if (!have) { stop = 1; break; }
    if (n >= cap) { stop = 2; break; }

    // ---- midpoint row (hyperbolic_merge.py:317-340), redundantly in every CTA -------------
    for (int e = threadIdx.x; e < D; e += blockDim.x) {
      xi[e] = __ldcg(p.E + (int64_t)best.i * p.ldE + e);
      xj[e] = __ldcg(p.E + (int64_t)best.j * p.ldE + e);
    }
    __syncthreads();
    if (warp == 0) {
      const int li = __ldcg(p.len + best.i), lj = __ldcg(p.len + best.j);
      warp_midpoint(xi, xj, li, lj, D, p.c, p.semantics, true, scratch, lane,
                    [&](int e, float v) { q[e] = v; });
      if (blockIdx.x == 0) {
        __syncwarp();
        for (int e = lane; e < D; e += 32) p.E[(int64_t)n * p.ldE + e] = q[e];
        if (lane == 0) {
          p.len[n] = li + lj;
          p.log[k] = hyp_merge_record{best.i, best.j, best.d, n};
        }
      }
    }
    __syncthreads();

    // ---- score row n against rows 0..n-1 ------------------------------------------------------
    unsigned long long below = 0;
    const float thr_f = (float)thr;
    Key b = block_scan_rows(p.E, p.ldE, n, n, q, D, p.sqrt_c, p.sgn, thr_f, thr, cmp_double, &below, s_keys, s_cnt);
    const int par = k & 1;
    if (threadIdx.x == 0) p.ws->slot[par][blockIdx.x] = b;
    arrivals += gridDim.x;
    grid_barrier(&p.ws->barrier, arrivals);

    // ---- every CTA folds the per-CTA minima into the same new state ------------------------------
    Key r = key_none();
    for (int t = threadIdx.x; t < (int)gridDim.x; t += blockDim.x) {
      Key kk;
      const Key *src = &p.ws->slot[par][t];
      kk.d = __ldcg(&src->d);
      kk.i = __ldcg(&src->i);
      kk.j = __ldcg(&src->j);
      if (key_less(kk, r)) r = kk;
    }
    r = warp_key_min(r);
    if (lane == 0) s_keys[warp] = r;
    __syncthreads();
    r = s_keys[0];
    for (int w = 1; w < kLoopWarps; ++w)
      if (key_less(s_keys[w], r)) r = s_keys[w];
    __syncthreads();
    if (key_less(r, best)) best = r;
    ++n;
    ++done;
    const int step = p.step0 + k;
    if (p.thr_every > 0 && step > 0 && step % p.thr_every == 0) thr *= p.thr_mul;
  }

  if (blockIdx.x == 0 && threadIdx.x == 0) {
    p.state->n = n;
    p.state->best_d = best.d;
    p.state->best_i = best.i;
    p.state->best_j = best.j;
    p.state->threshold = thr;
    p.state->steps_done = done;
    p.state->stop = stop;
  }
}


// ---------------------------------------------------------------------------------------------
// K5, table-resident variant.  The whole table lives in the shared memory of the persistent grid:
// CTA b owns rows r = b (mod G) in slot r / G.  Spatial parts are stored as float4 groups,
// T4[g][slot] = elements 4g..4g+3 of the row in `slot` (time parts in T0[slot]), so that
//  (a) one THREAD scores one row with conflict-free LDS.128 (a warp reads 512 contiguous bytes) and
//      ~3 instructions per element -- the warp-per-row L2 loop is issue-bound at ~10x that;
//  (b) the ATen summation order falls out of streaming the row in natural order into 32 register
//      accumulators P[e mod 32] (partial c = (e/8) mod 4, lane l = e mod 8  <=>  8c + l = e mod 32);
//  (c) appending a row is 25 strided float4 stores (odd slot stride).
// 148 SMs x ~220 KB hold ~80 k rows of D=101: the table of BASELINE configs[1] (50 k x 101 fp32 =
// 20.2 MB) never leaves the SMs between merges.  Per merge: one 64-bit atomicMin per CTA on
// (d_bits << 32 | row) and one grid barrier.
// ---------------------------------------------------------------------------------------------
constexpr int kResThreads = 384;
constexpr int kResWarps = kResThreads / 32;
constexpr unsigned long long kNoKey = 0xffffffffffffffffULL;

struct ResidentParams {
  LoopParams lp;
  int slots;     // slot capacity per CTA (odd)
};

// Two fp32 products in one instruction (FMUL2): each half is an IEEE round-to-nearest product, bit-identical to
// __fmul_rn.  Only the PRODUCTS are packed: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (single
// rounding) even under --fmad=false, which would break the torch summation order, so the adds stay scalar FADDs.
__device__ __forceinline__ void fmul2_rn(unsigned long long a, unsigned long long b, float &lo, float &hi) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(d));
}

// exact-order Minkowski product of the row in `slot` with q; N known at compile time.
template <int N>
__device__ __forceinline__ float resident_mdot_static(const float4 *__restrict__ T4, const float *__restrict__ T0, int S,
                                                      int slot, const float4 *__restrict__ q4, float q0) {
  constexpr int vs = N / 8, full = vs / 4, G4 = (N + 3) / 4;
  const ulonglong2 *T2 = reinterpret_cast<const ulonglong2 *>(T4);
  const ulonglong2 *q2 = reinterpret_cast<const ulonglong2 *>(q4);
  float P[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) P[k] = 0.f;
  float tail = 0.f;
#pragma unroll
  for (int g = 0; g < G4; ++g) {
    const ulonglong2 a = T2[(size_t)g * S + slot];
    const ulonglong2 b = q2[g];
    float pr[4];
    fmul2_rn(a.x, b.x, pr[0], pr[1]);
    if (4 * g + 2 < N) fmul2_rn(a.y, b.y, pr[2], pr[3]);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int e = 4 * g + c;
      if (e < N) {
        if (e < 32 * full) P[e & 31] = __fadd_rn(P[e & 31], pr[c]);
        else if (e < 8 * vs) P[e & 7] = __fadd_rn(P[e & 7], pr[c]);   // left-over lane vectors join partial 0
        else tail = __fadd_rn(tail, pr[c]);
      }
    }
  }
  float acc = tail;
#pragma unroll
  for (int l = 0; l < 8; ++l) {
    const float L = __fadd_rn(__fadd_rn(__fadd_rn(P[l], P[8 + l]), P[16 + l]), P[24 + l]);
    acc = __fadd_rn(acc, L);
  }
  return __fsub_rn(__fmul_rn(T0[slot], q0), acc);
}

// any N (including N < 8): scalar reads of the same layout
__device__ __forceinline__ float resident_mdot_dynamic(const float4 *__restrict__ T4, const float *__restrict__ T0, int S,
                                                       int slot, const float4 *__restrict__ q4, float q0, int N) {
  const float *Tf = reinterpret_cast<const float *>(T4);
  const float *qf = reinterpret_cast<const float *>(q4);
  float s = thread_sum_aten(
      [&](int e) { return __fmul_rn(Tf[((size_t)(e >> 2) * S + slot) * 4 + (e & 3)], qf[e]); }, N);
  return __fsub_rn(__fmul_rn(T0[slot], q0), s);
}

// Rows [row0, n) of the table in global memory against q, a warp per row (4 in flight), this warp taking every
// `stride`-th group of 4 from `first`.  Lane 0 folds (distance bits, row) into (best_d, best_r).  Out of line: it runs
// only for tables larger than the grid's shared memory and must not cost the resident path registers.
__device__ __noinline__ void overflow_scan(const float *E, int64_t ldE, int64_t row0, int n, const float *q, int D,
                                           float sqrt_c, float sgn, int64_t first, int64_t stride, int lane,
                                           unsigned int &best_d, unsigned int &best_r) {
  for (int64_t base = row0 + first * kRowsPerIter; base < n; base += stride * kRowsPerIter) {
    int64_t rows[kRowsPerIter];
    int nr = 0;
#pragma unroll
    for (int a = 0; a < kRowsPerIter; ++a) {
      rows[a] = base + a;
      if (base + a < n) nr = a + 1;
    }
    float mm[kRowsPerIter];
    warp_mdot_rows<kRowsPerIter>(E, ldE, rows, nr, q, D, lane, mm);
#pragma unroll
    for (int a = 0; a < kRowsPerIter; ++a) {
      if (a < nr) {
        const float d = dist_from_mdot(mm[a], sgn, sqrt_c);            // identical on every lane
        if (d == d && lane == 0 && __float_as_uint(d) < best_d) {      // rows ascend: strict < keeps the smallest
          best_d = __float_as_uint(d);
          best_r = (unsigned int)rows[a];
        }
      }
    }
  }
}

// The merged row of (xi, xj), out of line: for the call sites that run once per launch or once per failed guess, so
// that only the steady-state copy (the speculating warp's) is inlined into the loop.
template <int NS>
__device__ __noinline__ void midpoint_cold(const float *xi, const float *xj, int len_i, int len_j, int D, float c,
                                           int semantics, float *scr, int lane, float *qr, float *qf) {
  warp_midpoint<NS>(xi, xj, len_i, len_j, D, c, semantics, true, scr, lane, [&](int e, float v) {
    qr[e] = v;
    if (e) qf[e - 1] = v;
  });
}

// named barrier over the first `nthreads` threads of the CTA (the speculating warp stays out of it)
__device__ __forceinline__ void bar_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// -DHYP_LOOP_PROF (make EXTRA=-DHYP_LOOP_PROF): thread 0 of CTA 0 and of CTA G-1 accumulate the cycles of the
// three phases of an iteration into ws->prof (read by bench.py); -DHYP_MID_PROF: the speculating warp of CTA 0
// accumulates the sub-phases of its midpoint into ws->mprof.  The product build reads no clocks.
#ifdef HYP_LOOP_PROF
#define HYP_PHASE(acc)                                                        \
  do {                                                                        \
    const long long t_ = clock64();                                           \
    acc += t_ - t_mark;                                                       \
    t_mark = t_;                                                              \
  } while (0)
#else
#define HYP_PHASE(acc) do { } while (0)
#endif

// Software pipeline of the loop.  Merge k appends q_k = midpoint(best_k), scans it against the rows below it
// (scan_k), and the grid exchange of the per-CTA minima (exchange_k) gives best_{k+1}.  The exchange is a round
// trip through L2 that nothing can shorten, so it is overlapped with the NEXT merge's work under the guess that
// exchange_k will not beat best_k (the reference's loop settles on one pair after a few merges):
//   iteration k:   scan warps:        scan_{k+1} with q'_{k+1} = midpoint(best_k), then thread 0 polls exchange_k
//                  speculating warp:  q''_{k+2} = midpoint(best_k)
//   decide:        unchanged -> commit merge k, post the arrival of exchange_{k+1}, rotate the three q buffers
//                  changed   -> commit merge k with the new best, recompute q_{k+1}, redo scan_{k+1} and q'_{k+2}
// Every scan and every midpoint is computed in full each merge; a wrong guess only costs the redo.  (Guessing
// only after a guess that held was tried: the extra branch cost 4 % of the steady-state rate, and wrong guesses are
// confined to the first few dozen merges of a run -- each merge halves the best distance until it reaches 0 and
// index order then pins one pair for good.)  What a
// speculative scan writes (row n+1 of the table in shared and global memory, its len and log entry) lies beyond the
// committed state and is rewritten by the redo; a guess is only made when merge k+1 is due under it, and a changed
// best (smaller distance) cannot make it undue, so nothing speculative survives the launch.
// OVER: the table may outgrow the grid's shared memory (rows >= slots * G are scored from L2).  A separate
// instantiation: the extra code cost the all-resident kernel 12 % (registers, code layout) when it was merely present.
template <int NS, bool OVER>
__global__ void __launch_bounds__(kResThreads, 1) merge_loop_resident_kernel(const ResidentParams rp) {
  extern __shared__ __align__(16) float smem[];
  const LoopParams &p = rp.lp;
  const int D = p.D, S = rp.slots, N = D - 1;
  const int G4 = (N + 3) / 4;
  const int G = gridDim.x, b = blockIdx.x;
  constexpr int kWork = kResThreads - 32;      // threads that scan; the last warp speculates
  constexpr int kWorkWarps = kWork / 32;
  // layout (floats): qq4[3][4*G4] | T4[4*G4*S] | T0[S] | qrow[3][D] | xi[D] xj[D] scratch[2][2D]
  float4 *qq4 = reinterpret_cast<float4 *>(smem);
  float4 *T4 = qq4 + 3 * G4;
  float *T0 = reinterpret_cast<float *>(T4 + (size_t)G4 * S);
  float *qrow = T0 + S;
  float *xi = qrow + 3 * D;
  float *xj = xi + D;
  float *scratch = xj + D;
  float *Tf = reinterpret_cast<float *>(T4);
  __shared__ unsigned long long s_key[kResWarps];
  __shared__ unsigned long long s_win[2];
  __shared__ int s_len[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool spec_warp = warp == kResWarps - 1;

  // a launch on a state that has already stopped does nothing (lets a caller queue several launches back to back)
  if (p.state->stop != 0) {
    if (b == 0 && threadIdx.x == 0) p.state->steps_done = 0;
    return;
  }
  int n = p.state->n;
  const int cap = p.state->capacity;
  Key best{p.state->best_d, p.state->best_i, p.state->best_j};
  double thr = p.state->threshold;
  int done = 0, stop = 0;
  int cur = 0, nx1 = 1, nx2 = 2;               // q buffers: q_k | q'_{k+1} | q''_{k+2}
  unsigned int arrivals = (unsigned int)G;
#ifdef HYP_MID_PROF
  long long m_prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
#ifdef HYP_LOOP_PROF
  long long t_mid = 0, t_scan = 0, t_bar = 0, t_mark = 0;
#endif

  auto midpoint_rare = [&](int which, float *scr) {
    midpoint_cold<NS>(xi, xj, s_len[0], s_len[1], D, p.c, p.semantics, scr, lane, qrow + which * D,
                      reinterpret_cast<float *>(qq4 + which * G4));
  };
  auto midpoint_into = [&](int which, float *scr, long long *tp = nullptr) {
    float *qr = qrow + which * D;
    float *qf = reinterpret_cast<float *>(qq4 + which * G4);
    warp_midpoint<NS>(xi, xj, s_len[0], s_len[1], D, p.c, p.semantics, true, scr, lane, [&](int e, float v) {
      qr[e] = v;
      if (e) qf[e - 1] = v;
    }, tp);
  };

  // ---- load the rows this CTA owns: a warp reads one row (coalesced), scatters it conflict-free --
  {
    const int owned0 = (n > b) ? (n - b + G - 1) / G : 0;
    for (int sl = warp; sl < owned0 && (!OVER || sl < S); sl += kResWarps) {   // OVER: rows beyond S slots stay in global memory
      const float *row = p.E + ((int64_t)sl * G + b) * p.ldE;
      for (int k = lane; k < D; k += 32) {
        const float v = __ldcg(row + k);
        if (k == 0) T0[sl] = v;
        else Tf[((size_t)((k - 1) >> 2) * S + sl) * 4 + ((k - 1) & 3)] = v;
      }
    }
    float *qf = reinterpret_cast<float *>(qq4);
    for (int k = threadIdx.x; k < 12 * G4; k += blockDim.x) qf[k] = 0.f;   // padding lanes of the q4 buffers stay zero
    // operand rows of the current best pair (kept in shared memory until the best pair changes)
    if (best.i >= 0) {
      for (int e = threadIdx.x; e < D; e += blockDim.x) {
        xi[e] = __ldcg(p.E + (int64_t)best.i * p.ldE + e);
        xj[e] = __ldcg(p.E + (int64_t)best.j * p.ldE + e);
      }
      if (threadIdx.x == 0) s_len[0] = __ldcg(p.len + best.i);
      if (threadIdx.x == 1) s_len[1] = __ldcg(p.len + best.j);
    }
  }
  if (b == 0) {
    for (int sl = threadIdx.x; sl < kXRing; sl += blockDim.x) {
      p.ws->xring[sl].key = kNoKey;
      p.ws->xring[sl].cw = 0ull;
    }
  }
  grid_barrier(&p.ws->barrier, arrivals);

  // loop-carried bookkeeping without integer division: rows this CTA owns below n, and whether/where row n lands here
  int owned = (n > b) ? (n - b + G - 1) / G : 0;
  int n_mod = n % G, n_div = n / G;
  float thr_f = (float)thr;
  long long next_thr = -1;                   // first k after which the threshold is multiplied
  if (p.thr_every > 0) {
    next_thr = (p.thr_every - p.step0 % p.thr_every) % p.thr_every;
    if (p.step0 + next_thr == 0) next_thr = p.thr_every;
  }

  // is merge kk due?  0: yes; -1: max_steps reached; 1: no pair under the threshold; 2: vocabulary full
  auto due = [&](int kk, int nn, const Key &bb, double th, float th_f) -> int {
    if (kk >= p.max_steps) return -1;
    bool have = bb.i >= 0;
    if (have) have = (nn <= 100) ? ((double)bb.d < th) : (bb.d < th_f);     // hyperbolic_merge.py:288 vs :262
    if (!have) return 1;
    if (nn >= cap) return 2;
    return 0;
  };

  // scan_kk: score row nn (= the q buffer `qb`) against the resident rows below it, append it, and leave the CTA's
  // (distance bits, row) minimum in (c_d, c_r) on warp 0.  Called by the scanning warps only.
  unsigned int c_d = 0xffffffffu, c_r = 0xffffffffu;
  unsigned long long e_key = 0, e_cnt = 0;     // thread 0: the early poll of the previous exchange (see scan)
  bool e_issued = false;
  unsigned int nscan = 0;                      // scans started so far; identical in every CTA
  auto scan = [&](int qb, int nn, int own, int nmod, int ndiv, int kk, const LoopWorkspace::XSlot *early) {
    ++nscan;
    const float *q = qrow + qb * D;
    const float4 *q4 = qq4 + qb * G4;
    const float q0 = q[0];
    unsigned int my_d = 0xffffffffu, my_r = 0xffffffffu;   // d >= 0, so the bits order like d
    if (b == 0 && warp == kWorkWarps - 1) {
      // append row nn to the table in global memory (the caller's `embeddings`) and log the merge.  Done by the
      // last scanning warp (it has the fewest rows), which then fences and publishes the scan's sequence number:
      // the arrivals of the exchange stay free of fences (a fence next to the arrival cost 13 % of the merge
      // rate).  Readers (the failed-guess path, and every CTA once per kXRing/4 merges for the recycled slots)
      // wait for `published` before they touch what was written here.
      for (int e = lane; e < D; e += 32) p.E[(int64_t)nn * p.ldE + e] = q[e];
      if (lane == 0) {
        p.len[nn] = s_len[0] + s_len[1];
        p.log[kk] = hyp_merge_record{best.i, best.j, best.d, nn};
        // recycle the exchange slot used kXRing/2 merges ago (every CTA is past it) for its use kXRing/2 merges on
        LoopWorkspace::XSlot *rs = &p.ws->xring[(kk + kXRing / 2) & (kXRing - 1)];
        rs->key = kNoKey;
        reinterpret_cast<unsigned int *>(&rs->cw)[1] = 0u;        // its "posted" word; the arrivals keep counting
      }
      __threadfence();
      __syncwarp();
      if (lane == 0) st_relaxed_u32(&p.ws->published, nscan);
    }
    // Rows that did not fit the grid's shared memory (index >= S * G) are scored from L2, a warp per row, coalesced.
    // They were written by CTA 0 in earlier scans: its appending warp publishes a scan's number behind a fence, and
    // the load below (issued now, consumed after the resident rows) acquires it.
    const int64_t over0 = (int64_t)S * G;
    const bool has_over = OVER && (int64_t)nn > over0;
    unsigned int pub = 0;
    if (OVER && has_over && lane == 0) pub = ld_acquire_u32(&p.ws->published);
    const int own_res = (OVER && own > S) ? S : own;
    for (int t = threadIdx.x; t < own_res; t += kWork) {
      const float m = (NS > 0) ? resident_mdot_static<(NS > 0 ? NS : 8)>(T4, T0, S, t, q4, q0)
                               : resident_mdot_dynamic(T4, T0, S, t, q4, q0, N);
      if (early && t == 0) {
        // thread 0, half way through its scan: every CTA posted its arrival for the PREVIOUS exchange before it
        // started this scan, so by now the slot is almost surely complete.  The load's round trip through L2
        // hides behind the rest of the scan; poll() takes the value if the count is full.
        ld_relaxed_v2_u64(early, e_key, e_cnt);
        e_issued = true;
      }
      const float d = dist_from_mdot(m, p.sgn, p.sqrt_c);
      if (d == d && __float_as_uint(d) < my_d) {           // rows ascend with t: strict < keeps the smallest row
        my_d = __float_as_uint(d);
        my_r = (unsigned int)(t * G + b);
      }
    }
    if (OVER && has_over) {
      if (lane == 0) {
        unsigned int spins = 0;
        while (pub + 1u < nscan) {                                         // every earlier scan's append is visible
          pub = ld_acquire_u32(&p.ws->published);
          spin_guard(spins);
        }
      }
      __syncwarp();
      overflow_scan(p.E, p.ldE, over0, nn, q, D, p.sqrt_c, p.sgn, (int64_t)b * kWorkWarps + warp, (int64_t)G * kWorkWarps,
                    lane, my_d, my_r);
    }
    // the owner of row nn appends it (its slot is beyond `own`, so nobody reads it during this scan)
    if (nmod == b && (!OVER || ndiv < S)) {
      for (int e = threadIdx.x; e < D; e += kWork) {
        if (e == 0) T0[ndiv] = q[0];
        else Tf[((size_t)((e - 1) >> 2) * S + ndiv) * 4 + ((e - 1) & 3)] = q[e];
      }
    }
    // lexicographic (d, row) minimum of the warp with two REDUX instructions, then of the CTA on warp 0
    const unsigned int w_d = __reduce_min_sync(HYP_FULL_MASK, my_d);
    const unsigned int w_r = __reduce_min_sync(HYP_FULL_MASK, my_d == w_d ? my_r : 0xffffffffu);
    if (lane == 0) s_key[warp] = ((unsigned long long)w_d << 32) | w_r;
    bar_named(1, kWork);
    if (warp == 0) {
      const unsigned long long wk = lane < kWorkWarps ? s_key[lane] : kNoKey;
      const unsigned int k_d = (unsigned int)(wk >> 32), k_r = (unsigned int)(wk & 0xffffffffu);
      c_d = __reduce_min_sync(HYP_FULL_MASK, k_d);
      c_r = __reduce_min_sync(HYP_FULL_MASK, k_d == c_d ? k_r : 0xffffffffu);
    }
  };

  // ---- exchange on ONE 16-byte slot per merge (slot kk mod kXRing; its arrival counter is never reset, use u completes
  // at G * u).  A CTA posts its minimum (64-bit RED.MIN) only when it beats `best` -- every CTA holds the same `best`,
  // so a key that does not beat it cannot change the outcome -- and then arrives with ONE 64-bit RED.ADD on `cw`:
  // +1 arrival, and +1 in the high word if it posted (a release, ordered behind its RED.MIN; everything else arrives
  // relaxed: CTA 0's appends are published separately, see scan).  Every CTA reads the slot with a single 128-bit
  // load.  Memory-model argument: `cw` is one 64-bit scalar and all updates of it are RMWs, so the load that sees the
  // full arrival count sees the exact number of posts of this use.  No post (the steady state): no RED.MIN on the slot
  // exists since its reset, so the key element reads kNoKey whenever the hardware fetched it.  A post (a failed guess, rare):
  // the count is re-read with an ACQUIRE, which synchronises with the posters' release arrivals, and only then the key
  // is loaded.  So nothing relies on a vector load being one atomic access, and there is no second round trip on the
  // path that matters.  Thread 0 only.
  auto arrive = [&](int kk, int nn) {
    LoopWorkspace::XSlot *xr = &p.ws->xring[kk & (kXRing - 1)];
    const bool post = c_r != 0xffffffffu && key_less(Key{__uint_as_float(c_d), (int)c_r, nn}, best);
    if (post) atomicMin(&xr->key, ((unsigned long long)c_d << 32) | c_r);
    if (post) red_release_add_u64(&xr->cw, 0x100000001ull);
    else red_relaxed_add_u64(&xr->cw, 1ull);
  };
  auto poll = [&](int kk) {
    const LoopWorkspace::XSlot *xr = &p.ws->xring[kk & (kXRing - 1)];
    const unsigned int target = (unsigned int)G * (unsigned int)((kk / kXRing) + 1);
    unsigned long long key = e_key, cc = e_cnt;
    if (!e_issued) ld_relaxed_v2_u64(xr, key, cc);
    unsigned int spins = 0;
    while ((unsigned int)(cc & 0xffffffffu) < target) {
      ld_relaxed_v2_u64(xr, key, cc);
      spin_guard(spins);
    }
    e_issued = false;
    if ((unsigned int)(cc >> 32) != 0u) {
      (void)ld_acquire_u64(&xr->cw);           // synchronises with the release arrival of every CTA that posted
      key = ld_relaxed_u64(&xr->key);
    }
    // (no post: no RED.MIN on this slot exists since CTA 0 reset it -- every CTA's arrival is already in `cw` and none
    //  of them posted -- so the key that came with the vector load IS kNoKey, whenever its element was read.  Using it
    //  also keeps the register of the early load live until here: a register of an in-flight load that the compiler
    //  reuses stalls the scan for the whole round trip.)
    s_win[kk & 1] = key;
  };
  // Thread 0: wait until CTA 0 has fenced the appends of every scan before the one in flight, then fence: what those
  // scans wrote (rows / len below n, recycled exchange slots) may be read after the next CTA barrier.
  auto wait_published = [&]() {
    unsigned int spins = 0;
    while (ld_relaxed_u32(&p.ws->published) + 1u < nscan) spin_guard(spins);
    __threadfence();
  };

  // ---- the loop, as phases with ONE copy of the scan and of each midpoint (the kernel is sensitive to its code size).
  // A phase = [warp 0 computes the real merged row into nx1] ; { scan of nx1 || speculating warp: the row after it
  // into nx2 } ; CTA barrier.  Two kinds:
  //   real phase (prologue, and after a failed guess): scan_k of the real q_k; then post its arrival and rotate.
  //   guess phase (steady state): scan_{k+1} of q'_{k+1} while exchange_k is in flight, then read exchange_k:
  //     unchanged -> commit merge k, post the arrival of scan_{k+1}, rotate;
  //     changed   -> commit merge k with the new best, fetch its operands, next phase is a real one.
  // invariant before a guess phase: merge k is due; `best`, q_k (buffer cur) are final; scan_k is done and its arrival
  // posted; q'_{k+1} (buffer nx1) is the merged row for `best`, computed from xi, xj, s_len
  int code = due(0, n, best, thr, thr_f);
  int k = 0;                                   // next merge to commit
  bool guess = false;                          // kind of the next phase
#ifdef HYP_LOOP_PROF
  if (threadIdx.x == 0) t_mark = clock64();
#endif
  while (code == 0) {
    // what this phase scans: merge ks, row ns, and the owner bookkeeping that goes with row ns
    const bool bump = guess && (long long)k == next_thr;   // step0 + k is a positive multiple of thr_every
    const double thr_n = bump ? thr * p.thr_mul : thr;
    const float thr_nf = bump ? (float)thr_n : thr_f;
    const bool wrap = n_mod + 1 == G;
    const int own1 = guess ? owned + (n_mod == b ? 1 : 0) : owned;
    const int nmod1 = guess ? (wrap ? 0 : n_mod + 1) : n_mod, ndiv1 = guess ? n_div + (wrap ? 1 : 0) : n_div;
    const int ks = guess ? k + 1 : k, ns = guess ? n + 1 : n;
    int code_s = guess ? due(k + 1, n + 1, best, thr_n, thr_nf) : 0;   // is merge k+1 due if `best` stays?

    if (!guess) {
      if (warp == 0) midpoint_rare(nx1, scratch);          // the real q_k
      __syncthreads();
    }
    if (spec_warp) {
#ifdef HYP_MID_PROF
      if (code_s == 0) midpoint_into(nx2, scratch + 2 * D, m_prof);
#else
      if (code_s == 0) midpoint_into(nx2, scratch + 2 * D);
#endif
    } else {
      if (code_s == 0) scan(nx1, ns, own1, nmod1, ndiv1, ks, guess ? &p.ws->xring[k & (kXRing - 1)] : nullptr);
      if (guess && threadIdx.x == 0) {
        HYP_PHASE(t_scan);
        poll(k);
        HYP_PHASE(t_bar);
      }
    }
    __syncthreads();

    if (!guess) {
      // scan_k is done for real: post it, q_k becomes `cur`, the row after it is the next guess
      if (threadIdx.x == 0) arrive(k, n);
      const int t = cur; cur = nx1; nx1 = nx2; nx2 = t;
      guess = true;
      continue;
    }

    const unsigned long long win = s_win[k & 1];
    bool changed = false;
    if (win != kNoKey) {
      Key r{__uint_as_float((unsigned int)(win >> 32)), (int)(win & 0xffffffffu), n};
      if (key_less(r, best)) { best = r; changed = true; }
    }
    // commit merge k
    ++done;
    ++n;
    owned = own1; n_mod = nmod1; n_div = ndiv1;
    thr = thr_n; thr_f = thr_nf;
    if (bump) next_thr += p.thr_every;

    if (changed) {
      // new best pair = (i_win, n-1): row n-1 is q_k (still in shared memory), row i_win comes from L2
      const float *qk = qrow + cur * D;
      const int ln = s_len[0] + s_len[1];
      if (threadIdx.x == 0) wait_published();
      __syncthreads();
      for (int e = threadIdx.x; e < D; e += blockDim.x) {
        xj[e] = qk[e];
        xi[e] = __ldcg(p.E + (int64_t)best.i * p.ldE + e);
      }
      if (threadIdx.x == 0) s_len[0] = __ldcg(p.len + best.i);
      if (threadIdx.x == 1) s_len[1] = ln;
      code = due(k + 1, n, best, thr, thr_f);
      __syncthreads();
      ++k;
      guess = false;                           // the guess failed: the next phase recomputes q_{k+1} and scans it
      continue;
    }
    code = code_s;
    if (code != 0) break;
    if (threadIdx.x == 0) {
      if ((k & (kXRing / 4 - 1)) == kXRing / 4 - 1) wait_published();   // recycled slots: see scan
      arrive(k + 1, n);
    }
    const int t = cur; cur = nx1; nx1 = nx2; nx2 = t;
    ++k;
    if (threadIdx.x == 0) HYP_PHASE(t_mid);
  }
  stop = code > 0 ? code : 0;

#ifdef HYP_LOOP_PROF
  if (threadIdx.x == 0 && (b == 0 || b == G - 1)) {
    // [0] decide + commit (+ redo on a failed guess), [1] scan, [2] poll
    long long *prof = p.ws->prof + (b == 0 ? 0 : 4);
    prof[0] = t_mid; prof[1] = t_scan; prof[2] = t_bar; prof[3] = done;
  }
#endif
#ifdef HYP_MID_PROF
  if (b == 0 && spec_warp && lane == 0) {
    for (int q_ = 0; q_ < 8; ++q_) p.ws->mprof[q_] = m_prof[q_];
    p.ws->prof[3] = done;
  }
#endif
  if (b == 0 && threadIdx.x == 0) {
    p.state->n = n;
    p.state->best_d = best.d;
    p.state->best_i = best.i;
    p.state->best_j = best.j;
    p.state->threshold = thr;
    p.state->steps_done = done;
    p.state->stop = stop;
  }
}

__global__ void merge_state_init_kernel(hyp_merge_state *st, const hyp_best *best, int n, int capacity, double thr) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    st->threshold = thr;
    st->n = n;
    st->capacity = capacity;
    st->best_d = best->d;
    st->best_i = best->i;
    st->best_j = best->j;
    st->steps_done = 0;
    st->stop = 0;
    st->pad = 0;
  }
}

static int loop_grid(size_t smem, int *blocks_out) {
  int dev = 0, sms = 148, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, merge_loop_l2_kernel, kLoopThreads, smem);
  if (per_sm < 1) {
    set_error("merge loop: kernel does not fit on an SM (smem=%zu)", smem);
    return HYP_ERR_CUDA;
  }
  if (per_sm > 2) per_sm = 2;
  int g = sms * per_sm;
  if (g > kMaxLoopBlocks) g = kMaxLoopBlocks;
  *blocks_out = g;
  return HYP_OK;
}

}  // namespace hyp

using namespace hyp;

extern "C" int64_t hyp_merge_workspace_bytes(void) { return (int64_t)sizeof(LoopWorkspace); }

extern "C" int hyp_merge_state_init(hyp_merge_state *state, const hyp_best *best, int32_t n, int32_t capacity,
                                    double threshold, void *stream) {
  if (!state || !best || n < 0 || capacity < n) {
    set_error("hyp_merge_state_init: bad arguments");
    return HYP_ERR_ARG;
  }
  merge_state_init_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(state, best, n, capacity, threshold);
  return check_launch("hyp_merge_state_init");
}

extern "C" int hyp_row_min(const float *E, int64_t ldE, int64_t n, int64_t row, int D, float c, int semantics,
                           float threshold, hyp_best *best, void *workspace, int64_t workspace_bytes,
                           void *stream) {
  if (!E || !best || !workspace || n < 0 || row < 0 || D < 2 || D > HYP_MAX_D || !(c > 0.f)) {
    set_error("hyp_row_min: bad arguments");
    return HYP_ERR_ARG;
  }
  if (workspace_bytes < (int64_t)sizeof(LoopWorkspace)) {
    set_error("hyp_row_min: workspace %lld < %zu bytes", (long long)workspace_bytes, sizeof(LoopWorkspace));
    return HYP_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(workspace, 0, 192, st);
  int64_t want = (n + (int64_t)kLoopWarps * kRowsPerIter - 1) / ((int64_t)kLoopWarps * kRowsPerIter);
  int grid = (int)(want < 1 ? 1 : (want > kMaxLoopBlocks ? kMaxLoopBlocks : want));
  row_min_kernel<<<grid, kLoopThreads, (size_t)D * sizeof(float), st>>>(
      E, ldE, (int)n, (int)row, D, sqrtf(c), semantics == HYP_SEM_REFERENCE ? -1.f : 1.f, threshold,
      (LoopWorkspace *)workspace, best);
  return check_launch("hyp_row_min");
}

extern "C" int hyp_merge_steps(float *E, int64_t ldE, int32_t *len, int D, float c, int semantics,
                               hyp_merge_state *state, hyp_merge_record *log, int32_t max_steps, int32_t step0,
                               int32_t threshold_every, double threshold_mul, int32_t capacity_hint,
                               void *workspace, int64_t workspace_bytes, void *stream) {
  if (!E || !len || !state || !log || !workspace || D < 2 || D > HYP_MAX_D || max_steps < 0 || !(c > 0.f)) {
    set_error("hyp_merge_steps: bad arguments");
    return HYP_ERR_ARG;
  }
  if (workspace_bytes < (int64_t)sizeof(LoopWorkspace)) {
    set_error("hyp_merge_steps: workspace %lld < %zu bytes", (long long)workspace_bytes, sizeof(LoopWorkspace));
    return HYP_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  LoopParams p;
  p.E = E; p.ldE = ldE; p.len = len; p.D = D; p.c = c; p.sqrt_c = sqrtf(c);
  p.sgn = semantics == HYP_SEM_REFERENCE ? -1.f : 1.f;
  p.semantics = semantics;
  p.state = state; p.log = log; p.max_steps = max_steps; p.step0 = step0;
  p.thr_every = threshold_every; p.thr_mul = threshold_mul;
  p.ws = (LoopWorkspace *)workspace;
  int dev = 0, sms = 148, max_smem = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaMemsetAsync(workspace, 0, 192, st);

  // table-resident variant when every row up to `capacity_hint` fits in the grid's shared memory
  const int Nsp = D - 1, G4 = (Nsp + 3) / 4;
  const int64_t fixed = ((int64_t)12 * G4 + (int64_t)9 * D) * (int64_t)sizeof(float) + 1024;
  int slots = (int)(((int64_t)max_smem - fixed) / (((int64_t)4 * G4 + 1) * (int64_t)sizeof(float)));
  if ((slots & 1) == 0) --slots;  // odd stride: conflict-free scattered stores
  const char *force = getenv("HYP_MERGE_LOOP");
  const bool want_l2 = force && force[0] == 'l';
  const int64_t rows_max = capacity_hint > 0 ? capacity_hint : 0;
  if (const char *e = getenv("HYP_RESIDENT_SLOTS")) {       // tests: shrink the resident part to exercise the overflow path
    const int want = atoi(e);
    if (want >= 1 && want < slots) slots = want | 1;
  }
  // Rows beyond slots * sms stay in global memory and are scored from L2 by the same kernel (a warp per row); past
  // four times the resident capacity the plain L2 loop is as good.
  if (!want_l2 && slots >= 1 && rows_max > 0 && (rows_max + sms - 1) / sms <= (int64_t)4 * slots) {
    ResidentParams rp;
    rp.lp = p;
    rp.slots = slots;
    const size_t smem = ((size_t)12 * G4 + (size_t)4 * G4 * slots + slots + (size_t)9 * D) * sizeof(float);
    const bool over = (rows_max + sms - 1) / sms > slots;
    const void *fn = over ? (Nsp == 100  ? (const void *)merge_loop_resident_kernel<100, true>
                             : Nsp == 50 ? (const void *)merge_loop_resident_kernel<50, true>
                                         : (const void *)merge_loop_resident_kernel<0, true>)
                          : (Nsp == 100  ? (const void *)merge_loop_resident_kernel<100, false>
                             : Nsp == 50 ? (const void *)merge_loop_resident_kernel<50, false>
                                         : (const void *)merge_loop_resident_kernel<0, false>);
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("hyp_merge_steps: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
      return HYP_ERR_CUDA;
    }
    void *args[] = {(void *)&rp};
    e = cudaLaunchCooperativeKernel(fn, dim3(sms), dim3(kResThreads), args, smem, st);
    if (e != cudaSuccess) {
      set_error("hyp_merge_steps: cooperative launch (resident) failed: %s", cudaGetErrorString(e));
      return HYP_ERR_CUDA;
    }
    return check_launch("hyp_merge_steps");
  }

  const size_t smem = (size_t)5 * D * sizeof(float);
  int grid = 0;
  int rc = loop_grid(smem, &grid);
  if (rc) return rc;
  void *args[] = {(void *)&p};
  cudaError_t e = cudaLaunchCooperativeKernel((const void *)merge_loop_l2_kernel, dim3(grid), dim3(kLoopThreads), args,
                                              smem, st);
  if (e != cudaSuccess) {
    set_error("hyp_merge_steps: cooperative launch failed: %s", cudaGetErrorString(e));
    return HYP_ERR_CUDA;
  }
  return check_launch("hyp_merge_steps");
}
