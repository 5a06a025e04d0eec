// common.cuh -- shared device helpers for the hyptok_b200 kernels (sm_100a).
//
// Everything that feeds the pre-clamp Minkowski product is written with explicit
// round-to-nearest intrinsics (no FMA contraction) and follows ATen's CPU summation
// order, so `u` is bit-identical to the reference's value.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hyptok_b200.h"

#define HYP_FULL_MASK 0xffffffffu

namespace hyp {

void set_error(const char *fmt, ...);
int check_launch(const char *what);

// ---------------------------------------------------------------------------------------------
// scalar math
// ---------------------------------------------------------------------------------------------

// torch.clamp(x, min=lo): NaN propagates (fmaxf would swallow it).
__device__ __forceinline__ float clamp_min(float x, float lo) { return (x < lo) ? lo : x; }
__device__ __forceinline__ float clamp_max(float x, float hi) { return (x > hi) ? hi : x; }

// acosh for u >= 1 in the form log1p(t + sqrt(t (t+2))), t = u-1: <= 2.5 ulp of torch's CPU
// acosh and bit-equal to it ~85 % of the time. acosh(1) == 0 exactly,
// NaN -> NaN.
// (out of line: it is the rare branch, and the merge loop's speed depends on the size of its inlined code)
static __device__ __noinline__ float acosh_huge(float u) { return __fadd_rn(logf(u), 0.69314718f); }
__device__ __forceinline__ float acosh_ge1(float u) {
  if (u > 1e18f) return acosh_huge(u);  // t*(t+2) would overflow
  float t = __fsub_rn(u, 1.0f);
  float s = __fsqrt_rn(__fmul_rn(t, __fadd_rn(t, 2.0f)));
  return log1pf(__fadd_rn(t, s));
}

// distance from the pre-sign Minkowski product m (lorentz_model.py:134-138):
//   u = clamp(sgn * m, 1.0f);  d = acosh(u) / sqrt(c)
// `1.0 + 1e-8` is a Python double and rounds to 1.0f when torch clamps an fp32 tensor.
__device__ __forceinline__ float dist_from_mdot(float m, float sgn, float sqrt_c) {
  float u = clamp_min(sgn < 0.f ? -m : m, 1.0f);
  return __fdiv_rn(acosh_ge1(u), sqrt_c);
}

// d(distance)/d<x,y> times the incoming gradient g, as autograd derives it through
// acosh(clamp(sgn*m, min=1.0f)) / sqrt(c): division first (g / sqrt_c), acosh' = rsqrt(u*u - 1), the clamp passes
// the gradient where its input >= min (inclusive; NaN fails the test) and d(sgn*m)/dm = sgn.  u == 1 gives
// g * inf exactly as torch does.
__device__ __forceinline__ float dist_grad_from_mdot(float m, float sgn, float sqrt_c, float g) {
  const float s = sgn < 0.f ? -m : m;
  if (!(s >= 1.0f)) return 0.f;
  const float r = __fdiv_rn(1.0f, __fsqrt_rn(__fsub_rn(__fmul_rn(s, s), 1.0f)));
  const float gm = __fmul_rn(__fdiv_rn(g, sqrt_c), r);
  return sgn < 0.f ? -gm : gm;
}

// ---------------------------------------------------------------------------------------------
// ATen-order reductions, one WARP per vector (coalesced: lane t owns elements t, t+32, ...)
//
// sum_fp32(p[0..N)):  N >= 8: vs = N/8 lane-vectors; full = vs/4 rounds. part[c][l] (c = k mod 4)
// accumulates vector k = 4r+c in order r = 0..full-1; left-over vectors k >= 4*full all go to
// part[0]; lanes[l] = ((part0+part1)+part2)+part3; acc = scalar tail (from 0) then += lanes[0..7].
// Thread t = 8c + l therefore owns exactly the elements t + 32r  -- a coalesced stride.
// ---------------------------------------------------------------------------------------------
template <int NS = 0, typename ProdFn>
__device__ __forceinline__ float warp_sum_aten(ProdFn prod, int Nrt, int lane) {
  const int N = NS > 0 ? NS : Nrt;   // NS > 0: trip counts are compile-time constants, every loop unrolls
  if (N < 8) {
    // scalar row_sum with 4 interleaved partials; every lane computes it redundantly.
    float part[4] = {0.f, 0.f, 0.f, 0.f};
    int full = N >> 2;
    for (int r = 0; r < full; ++r)
#pragma unroll
      for (int k = 0; k < 4; ++k) part[k] = __fadd_rn(part[k], prod(4 * r + k));
    for (int i = 4 * full; i < N; ++i) part[0] = __fadd_rn(part[0], prod(i));
    float a = __fadd_rn(part[0], part[1]);
    a = __fadd_rn(a, part[2]);
    return __fadd_rn(a, part[3]);
  }
  const int vs = N >> 3;
  const int full = vs >> 2;
  const int c = lane >> 3, l = lane & 7;
  float P = 0.f;
  if (NS > 0) {
#pragma unroll
    for (int r = 0; r < (NS > 0 ? (NS >> 3) >> 2 : 1); ++r) P = __fadd_rn(P, prod(32 * r + lane));
    if (c == 0) {
#pragma unroll
      for (int k = 4 * ((NS >> 3) >> 2); k < (NS >> 3); ++k) P = __fadd_rn(P, prod(8 * k + l));
    }
  } else {
#pragma unroll 4
    for (int r = 0; r < full; ++r) P = __fadd_rn(P, prod(32 * r + lane));
    if (c == 0) {
#pragma unroll 3
      for (int k = 4 * full; k < vs; ++k) P = __fadd_rn(P, prod(8 * k + l));
    }
  }
  float p1 = __shfl_down_sync(HYP_FULL_MASK, P, 8);
  float p2 = __shfl_down_sync(HYP_FULL_MASK, P, 16);
  float p3 = __shfl_down_sync(HYP_FULL_MASK, P, 24);
  float L = __fadd_rn(__fadd_rn(__fadd_rn(P, p1), p2), p3);  // valid on lanes 0..7
  float acc = 0.f;
  if (NS > 0) {
#pragma unroll
    for (int k = 8 * (NS >> 3); k < NS; ++k) acc = __fadd_rn(acc, prod(k));
  } else {
#pragma unroll 7
    for (int k = 8 * vs; k < N; ++k) acc = __fadd_rn(acc, prod(k));
  }
  float Lq[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) Lq[q] = __shfl_sync(HYP_FULL_MASK, L, q);   // independent shuffles first
#pragma unroll
  for (int q = 0; q < 8; ++q) acc = __fadd_rn(acc, Lq[q]);
  return acc;  // identical on every lane
}

// <x,y> with signature (+,-,...,-), lorentz_model.py:25:  fl(fl(x0*y0) - sum_fp32(fl(xs*ys))).
template <int NS = 0>
__device__ __forceinline__ float warp_mdot(const float *__restrict__ x, const float *__restrict__ y,
                                           int D, int lane) {
  const float *xs = x + 1, *ys = y + 1;
  float s = warp_sum_aten<NS>([&](int e) { return __fmul_rn(xs[e], ys[e]); }, D - 1, lane);
  return __fsub_rn(__fmul_rn(x[0], y[0]), s);
}

// torch.norm(v, dim=-1) of N contiguous floats (lorentz_model.py:53): 8 lane accumulators
// (rounded products, no FMA) over k in order, lanes folded 0..7, tail in groups of 4 with rounded
// products, then a <= 3 element remainder with FMA, then sqrt.
template <int NS = 0, typename ElemFn>
__device__ __forceinline__ float warp_norm_aten(ElemFn elem, int Nrt, int lane) {
  const int N = NS > 0 ? NS : Nrt;
  const int vs = N >> 3;
  float a = 0.f;
  if (lane < 8) {
    if (NS > 0) {
      float sq[NS > 0 ? (NS >> 3) + 1 : 1];             // all loads and squares first, then the ordered adds
#pragma unroll
      for (int k = 0; k < (NS >> 3); ++k) {
        const float v = elem(8 * k + lane);
        sq[k] = __fmul_rn(v, v);
      }
#pragma unroll
      for (int k = 0; k < (NS >> 3); ++k) a = __fadd_rn(a, sq[k]);
    } else {
#pragma unroll 4
      for (int k = 0; k < vs; ++k) {
        float v = elem(8 * k + lane);
        a = __fadd_rn(a, __fmul_rn(v, v));
      }
    }
  }
  float b = __shfl_sync(HYP_FULL_MASK, a, 0);
#pragma unroll
  for (int q = 1; q < 8; ++q) b = __fadd_rn(b, __shfl_sync(HYP_FULL_MASK, a, q));
  int k = 8 * vs;
  while (N - k >= 4) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v = elem(k + q);
      b = __fadd_rn(b, __fmul_rn(v, v));
    }
    k += 4;
  }
  for (; k < N; ++k) {
    float v = elem(k);
    b = __fmaf_rn(v, v, b);
  }
  return __fsqrt_rn(b);
}

// ---------------------------------------------------------------------------------------------
// thread-level ATen-order sum over products given by prod(e), e in [0, N).
// Visits lane l = 0..7, partial c = 0..3 so only three running values live per sum
// (used by the register-blocked all-pairs tile kernel and by thread-per-row scans).
// ---------------------------------------------------------------------------------------------
template <typename ProdFn>
__device__ __forceinline__ float thread_sum_aten(ProdFn prod, int N) {
  if (N < 8) {
    float part[4] = {0.f, 0.f, 0.f, 0.f};
    int full = N >> 2;
    for (int r = 0; r < full; ++r)
#pragma unroll
      for (int k = 0; k < 4; ++k) part[k] = __fadd_rn(part[k], prod(4 * r + k));
    for (int i = 4 * full; i < N; ++i) part[0] = __fadd_rn(part[0], prod(i));
    float a = __fadd_rn(part[0], part[1]);
    a = __fadd_rn(a, part[2]);
    return __fadd_rn(a, part[3]);
  }
  const int vs = N >> 3, full = vs >> 2;
  float acc = 0.f;
  for (int k = 8 * vs; k < N; ++k) acc = __fadd_rn(acc, prod(k));
  for (int l = 0; l < 8; ++l) {
    float L = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float P = 0.f;
      for (int r = 0; r < full; ++r) P = __fadd_rn(P, prod(32 * r + 8 * c + l));
      if (c == 0)
        for (int k = 4 * full; k < vs; ++k) P = __fadd_rn(P, prod(8 * k + l));
      L = (c == 0) ? P : __fadd_rn(L, P);
    }
    acc = __fadd_rn(acc, L);
  }
  return acc;
}

// ---------------------------------------------------------------------------------------------
// (d, i, j) lexicographic key helpers.  d >= 0 (or NaN, which never qualifies), so the fp32 bit
// pattern orders like the value.
// ---------------------------------------------------------------------------------------------
struct Key {
  float d;
  int i, j;
};
__device__ __forceinline__ Key key_none() { return Key{__int_as_float(0x7f800000), -1, -1}; }
__device__ __forceinline__ bool key_less(const Key &a, const Key &b) {
  // a valid (i >= 0) key always beats an empty one; NaN is filtered before keys are formed
  if (a.i < 0) return false;
  if (b.i < 0) return true;
  if (a.d < b.d) return true;
  if (a.d > b.d) return false;
  if (a.i != b.i) return a.i < b.i;
  return a.j < b.j;
}
__device__ __forceinline__ Key warp_key_min(Key k) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Key other;
    other.d = __shfl_xor_sync(HYP_FULL_MASK, k.d, o);
    other.i = __shfl_xor_sync(HYP_FULL_MASK, k.i, o);
    other.j = __shfl_xor_sync(HYP_FULL_MASK, k.j, o);
    if (key_less(other, k)) k = other;
  }
  return k;
}

// ---------------------------------------------------------------------------------------------
// Where a per-row top-k list goes.  Separate arrays (idx[r][k], d[r][k]) for single-GPU callers, and/or
// interleaved {idx, d bits} records written straight into the gather buffers of every rank of a
// hyp_ctx (peer memory over NVLink): pairs[g] is rank g's buffer, already offset to this shard's first
// row, so the kernel that produces a row also performs its part of the all-gather.
// ---------------------------------------------------------------------------------------------
struct TopkSink {
  int32_t *idx;
  float *d;
  int2 *pairs[HYP_MAX_PEERS];
  int n_pairs;
};
__device__ __forceinline__ void sink_write(const TopkSink &s, int64_t r, int k, int q, int32_t j, float dv) {
  if (s.idx) s.idx[r * k + q] = j;
  if (s.d) s.d[r * k + q] = dv;
  const int2 rec = make_int2(j, __float_as_int(dv));
#pragma unroll
  for (int g = 0; g < HYP_MAX_PEERS; ++g)      // (static indices: a dynamic one would copy the parameter struct to local memory)
    if (g < s.n_pairs) s.pairs[g][r * k + q] = rec;
}

// allpairs.cu: exact per-row top-k of rows [row0, row0+nrows) -- or, with row_list / row_count (device memory), of the
// listed shard-relative rows only -- into `sink`.
int launch_allpairs_topk(const float *E, int64_t ldE, int64_t n, int64_t row0, int64_t nrows, int D, float c,
                         int semantics, int k, const TopkSink &sink, const int32_t *row_list, const int32_t *row_count,
                         cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// the midpoint chain of hyperbolic_merge.py:323-340, one warp, rows in shared or global memory.
// `v` is scratch for D floats (shared).  Writes D floats to out (any address space).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float logmap_coef(float m, int semantics, float *m_signed) {
  // lorentz_model.py:108-117 (reference) / Appendix B (lorentz)
  float coef;
  if (semantics == HYP_SEM_REFERENCE) {
    float u = clamp_min(-m, 1.0f);
    coef = __fdiv_rn(acosh_ge1(u), __fsqrt_rn(__fsub_rn(__fmul_rn(u, u), 1.0f)));
    coef = clamp_max(coef, 1e4f);
    float bad = ((coef != coef) || (coef > 1e4f)) ? 1.f : 0.f;
    // mask*1 + (1-mask)*coef : 0*NaN stays NaN, exactly as shipped
    coef = __fadd_rn(__fmul_rn(bad, 1.0f), __fmul_rn(__fsub_rn(1.0f, bad), coef));
    *m_signed = m;  // y + <x,y> x
  } else {
    float u = clamp_min(m, 1.0f);
    coef = __fdiv_rn(acosh_ge1(u), __fsqrt_rn(__fsub_rn(__fmul_rn(u, u), 1.0f)));
    coef = clamp_max(coef, 1e4f);
    if ((coef != coef) || (coef > 1e4f)) coef = 1.0f;
    *m_signed = -m;  // y - <x,y> x
  }
  return coef;
}

// f(k) for k = lane, lane + 32, ... < D.  With NS > 0 (D == NS + 1 known at compile time) the trip count is a
// constant and the iterations unroll, so their loads and divisions overlap instead of running back to back.
template <int NS, typename F>
__device__ __forceinline__ void warp_for_elems(int D, int lane, F f) {
  if (NS > 0) {
#pragma unroll
    for (int it = 0; it < (NS > 0 ? (NS + 1 + 31) / 32 : 1); ++it) {
      const int k = lane + 32 * it;
      if (k < NS + 1) f(k);
    }
  } else {
    for (int k = lane; k < D; k += 32) f(k);
  }
}

// out = exp_x(v) (lorentz_model.py:85-93); v holds D floats readable by the whole warp.
template <int NS = 0, typename OutFn>
__device__ __forceinline__ void warp_expmap(const float *__restrict__ x, const float *v, int D,
                                            int lane, OutFn out) {
  const float *vs = v + 1;
  float sq = warp_sum_aten<NS>([&](int e) { return __fmul_rn(vs[e], vs[e]); }, D - 1, lane);
  float vn = __fsqrt_rn(clamp_min(sq, 1e-8f));
  float small = (vn < 1e-6f) ? 1.f : 0.f;
  float den = __fadd_rn(vn, small);
  float ch = coshf(vn), sh = sinhf(vn);
  float keep = __fsub_rn(1.0f, small);
  warp_for_elems<NS>(D, lane, [&](int k) {
    float dir = __fdiv_rn(v[k], den);
    dir = __fadd_rn(__fmul_rn(small, 0.0f), __fmul_rn(keep, dir));
    out(k, __fadd_rn(__fmul_rn(ch, x[k]), __fmul_rn(sh, dir)));
  });
}

// Full chain; `buf` = 2*D floats of shared scratch private to the warp. Result row left in
// buf[0..D) and returned through out(k, value).
template <int NS = 0, typename OutFn>
__device__ __forceinline__ void warp_midpoint(const float *__restrict__ xi, const float *__restrict__ xj,
                                              int len_i, int len_j, int D, float c, int semantics,
                                              bool project, float *buf, int lane, OutFn out,
                                              long long *tp = nullptr) {
  float *v = buf, *m_row = buf + D;
  long long tc0 = tp ? clock64() : 0;
  float m = warp_mdot<NS>(xi, xj, D, lane);
  if (tp) { long long t = clock64(); tp[0] += t - tc0; tc0 = t; }
  float ms;
  float coef = logmap_coef(m, semantics, &ms);
  if (tp) { long long t = clock64(); tp[1] += t - tc0; tc0 = t; }
  const float w = (float)((double)len_j / (double)(len_i + len_j));  // Python float, cast to fp32 by `*`
  if (tp) { long long t = clock64(); tp[2] += t - tc0; tc0 = t; }
  warp_for_elems<NS>(D, lane, [&](int k) {
    float lg = __fmul_rn(coef, __fadd_rn(xj[k], __fmul_rn(ms, xi[k])));
    v[k] = __fmul_rn(lg, w);
  });
  __syncwarp();
  if (tp) { long long t = clock64(); tp[3] += t - tc0; tc0 = t; }
  warp_expmap<NS>(xi, v, D, lane, [&](int k, float val) { m_row[k] = val; });
  __syncwarp();
  if (tp) { long long t = clock64(); tp[4] += t - tc0; tc0 = t; }
  if (project) {
    // lorentz_model.py:52-55: x0 = sqrt(1 + (c*r)*r)
    float r = warp_norm_aten<NS>([&](int e) { return m_row[1 + e]; }, D - 1, lane);
    float x0 = __fsqrt_rn(__fadd_rn(1.0f, __fmul_rn(__fmul_rn(c, r), r)));
    if (tp) { long long t = clock64(); tp[5] += t - tc0; tc0 = t; }
    warp_for_elems<NS>(D, lane, [&](int k) { out(k, k == 0 ? x0 : m_row[k]); });
    if (tp) { long long t = clock64(); tp[6] += t - tc0; tc0 = t; }
  } else {
    warp_for_elems<NS>(D, lane, [&](int k) { out(k, m_row[k]); });
  }
  __syncwarp();
}

}  // namespace hyp
