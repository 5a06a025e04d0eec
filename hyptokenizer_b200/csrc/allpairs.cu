// allpairs.cu -- K2, exact fp32 path: all-pairs Lorentz distance on CUDA cores (sm_100a).
//
// Replaces embedding/lorentz_model.py:141-178 (`batch_distance`, which materialises an
// (n, n, d) temporary) and the candidate scan of tokenizer/hyperbolic_merge.py:247-269.
// The Minkowski product of every pair is accumulated in ATen's CPU order (SURVEY.md
// Appendix D) with separately rounded products, so it is bit-identical to the reference;
// that is why this path runs on the FP32 pipe and not on tensor cores (the tcgen05 Gram
// kernel in gram_tc.cu trades that for speed and re-scores its finalists here).
//
// Tiling: 64x64 pairs per CTA, 256 threads, 4x4 pairs per thread; both operand tiles are
// staged TRANSPOSED in shared memory ([k][row], 16-byte aligned rows) so each k costs two
// LDS.128 per 16 products.  Persistent CTAs walk the (upper-triangular) tile list.
// Compiled with -fmad=false.
#include <math.h>

#include "common.cuh"

namespace hyp {

constexpr int TM = 64, TN = 64, TPAD = 68, NTHREADS = 256;
constexpr int kMaxBlocks = 148 * 8;

enum Mode { kDense = 0, kMin = 1, kEmit = 2, kGradCoef = 3, kHist = 4, kRowTies = 5 };
constexpr int kHistBits = 12;              // bins per radix-select pass over the 31 significant bits of d >= 0

struct MinWorkspace {
  unsigned int ticket;
  unsigned int pad0;
  unsigned long long count;
  Key block_best[kMaxBlocks];
};

struct Params {
  const float *x;
  int64_t ldx;
  int64_t n1;
  const float *y;
  int64_t ldy;
  int64_t n2;
  int D;
  float sqrt_c, sgn, thr;
  // dense / grad-coef (out = W)
  float *out;
  int64_t ldo;
  const float *gout;   // grad-coef: incoming gradient of the distance matrix
  int64_t ldg;
  // min
  MinWorkspace *ws;
  hyp_best *best;
  // emit
  int32_t *out_i, *out_j;
  float *out_d;
  int64_t capacity;
  unsigned long long *emit_count;
  // emit with a cut (global top-K): only pairs with bits(d) < cut_bits, or == cut_bits in rows <= cut_row
  unsigned int cut_bits;
  int64_t cut_row;
  // hist: pairs whose distance bits start with `prefix` (prefix_len bits below the sign) are binned on the next
  // kHistBits (or fewer) bits; row ties: pairs with bits(d) == cut_bits counted per row
  unsigned int prefix, prefix_shift, bin_shift, bin_mask;
  unsigned long long *hist;
  unsigned int *row_ties;
  // tiles
  int64_t tiles_m, tiles_n, n_tiles;
  int triangular;
};

__device__ __forceinline__ void tile_coords(const Params &p, int64_t t, int64_t &tm, int64_t &tn) {
  if (!p.triangular) {
    tm = t / p.tiles_n;
    tn = t - tm * p.tiles_n;
    return;
  }
  // t enumerates (tm, tn) with tn >= tm, row-major
  const double T = (double)p.tiles_m;
  double b = 2.0 * T + 1.0;
  int64_t r = (int64_t)floor((b - sqrt(b * b - 8.0 * (double)t)) * 0.5);
  if (r < 0) r = 0;
  if (r >= p.tiles_m) r = p.tiles_m - 1;
  auto start = [&](int64_t q) { return q * p.tiles_m - q * (q - 1) / 2; };
  while (r > 0 && start(r) > t) --r;
  while (r + 1 < p.tiles_m && start(r + 1) <= t) ++r;
  tm = r;
  tn = r + (t - start(r));
}

// Load rows [r0, r0+64) of a row-major table into a transposed tile: T[k][r], k = spatial index.
__device__ __forceinline__ void load_tile_T(const float *__restrict__ src, int64_t ld, int64_t n, int64_t r0,
                                            int d, float *__restrict__ T, float *__restrict__ t0) {
  // a warp reads one row's contiguous floats (coalesced); 8 warps -> 8 rows per pass
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int r = w; r < TM; r += NTHREADS / 32) {
    const int64_t gr = r0 + r;
    const bool ok = gr < n;
    const float *row = src + (ok ? gr : 0) * ld;
    for (int k = lane; k < d; k += 32) T[k * TPAD + r] = ok ? row[1 + k] : 0.f;
    if (lane == 0) t0[r] = ok ? row[0] : 0.f;
  }
}

// Exact-order spatial dot products of a 4x4 block of pairs from two transposed tiles.
__device__ __forceinline__ void tile_dots(const float *__restrict__ As, const float *__restrict__ Bs, int d, int ty,
                                          int tx, float (&S)[4][4]) {
  if (d >= 8) {
    const int vs = d >> 3, full = vs >> 2;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) S[r][c] = 0.f;
    // scalar tail first (from zero), then lanes 0..7 in order
    for (int e = 8 * vs; e < d; ++e) {
      const float4 a = *reinterpret_cast<const float4 *>(As + e * TPAD + 4 * ty);
      const float4 b = *reinterpret_cast<const float4 *>(Bs + e * TPAD + 4 * tx);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) S[r][c] = __fadd_rn(S[r][c], __fmul_rn(av[r], bv[c]));
    }
    for (int l = 0; l < 8; ++l) {
      float L[4][4];
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        float P[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) P[r][c] = 0.f;
        auto step = [&](int e) {
          const float4 a = *reinterpret_cast<const float4 *>(As + e * TPAD + 4 * ty);
          const float4 b = *reinterpret_cast<const float4 *>(Bs + e * TPAD + 4 * tx);
          const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) P[r][c] = __fadd_rn(P[r][c], __fmul_rn(av[r], bv[c]));
        };
        for (int r = 0; r < full; ++r) step(32 * r + 8 * c4 + l);
        if (c4 == 0)
          for (int k = 4 * full; k < vs; ++k) step(8 * k + l);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) L[r][c] = (c4 == 0) ? P[r][c] : __fadd_rn(L[r][c], P[r][c]);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) S[r][c] = __fadd_rn(S[r][c], L[r][c]);
    }
  } else {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        S[r][c] = thread_sum_aten(
            [&](int e) { return __fmul_rn(As[e * TPAD + 4 * ty + r], Bs[e * TPAD + 4 * tx + c]); }, d);
  }
}

template <int MODE>
__global__ void __launch_bounds__(NTHREADS) allpairs_tile_kernel(const Params p) {
  extern __shared__ __align__(16) float smem[];
  const int d = p.D - 1;
  float *As = smem;                 // [d][TPAD]
  float *Bs = As + (size_t)d * TPAD;  // [d][TPAD]
  float *a0 = Bs + (size_t)d * TPAD;  // [64]
  float *b0 = a0 + TM;              // [64]
  __shared__ Key s_keys[NTHREADS / 32];
  __shared__ unsigned long long s_cnt[NTHREADS / 32];
  __shared__ int s_last;
  __shared__ unsigned int s_hist[MODE == kHist ? (1 << kHistBits) : 1];
  if (MODE == kHist) {
    for (int q = threadIdx.x; q < (1 << kHistBits); q += NTHREADS) s_hist[q] = 0;
    __syncthreads();
  }

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Key best = key_none();
  unsigned long long below = 0;

  for (int64_t t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
    int64_t tm, tn;
    tile_coords(p, t, tm, tn);
    const int64_t i0 = tm * TM, j0 = tn * TN;
    __syncthreads();  // previous tile fully consumed
    load_tile_T(p.x, p.ldx, p.n1, i0, d, As, a0);
    load_tile_T(p.y, p.ldy, p.n2, j0, d, Bs, b0);
    __syncthreads();

    float S[4][4];
    tile_dots(As, Bs, d, ty, tx, S);

    // epilogue: m = fl(fl(x0*y0) - S), distance, then the mode's consumer
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int64_t gi = i0 + 4 * ty + r;
      float dist[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float m = __fsub_rn(__fmul_rn(a0[4 * ty + r], b0[4 * tx + c]), S[r][c]);
        dist[c] = dist_from_mdot(m, p.sgn, p.sqrt_c);
      }
      if (MODE == kGradCoef) {
        // backward of the dense mode: W[i][j] = g[i][j] * d(distance)/d<x_i, y_j>; the caller turns W into the
        // row gradients with two plain GEMMs (W @ Y', W^T @ X')
        if (gi < p.n1) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int64_t gj = j0 + 4 * tx + c;
            if (gj < p.n2) {
              const float m = __fsub_rn(__fmul_rn(a0[4 * ty + r], b0[4 * tx + c]), S[r][c]);
              p.out[gi * p.ldo + gj] = dist_grad_from_mdot(m, p.sgn, p.sqrt_c, p.gout[gi * p.ldg + gj]);
            }
          }
        }
      } else if (MODE == kDense) {
        if (gi < p.n1) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int64_t gj = j0 + 4 * tx + c;
            if (gj < p.n2) p.out[gi * p.ldo + gj] = dist[c];
          }
        }
      } else if (MODE == kHist) {
        // radix-select pass: candidates (i < j, d < thr) whose distance bits carry the prefix, binned on the next bits.
        // Equal bins of a thread's row are merged before they touch shared memory (in the shipped semantics every
        // distance is 0: one bin).
        unsigned int prev = 0xffffffffu, run = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int64_t gj = j0 + 4 * tx + c;
          if (gi < gj && gj < p.n2 && gi < p.n1 && dist[c] == dist[c] && dist[c] < p.thr) {
            const unsigned int bits = __float_as_uint(dist[c]);
            if ((bits >> p.prefix_shift) == p.prefix || p.prefix_shift >= 32u) {
              const unsigned int bin = (bits >> p.bin_shift) & p.bin_mask;
              if (bin == prev) ++run;
              else {
                if (run) atomicAdd(&s_hist[prev], run);
                prev = bin;
                run = 1;
              }
            }
          }
        }
        if (run) atomicAdd(&s_hist[prev], run);
      } else if (MODE == kRowTies) {
        unsigned int ties = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int64_t gj = j0 + 4 * tx + c;
          if (gi < gj && gj < p.n2 && gi < p.n1 && dist[c] < p.thr && __float_as_uint(dist[c]) == p.cut_bits) ++ties;
        }
        // the 16 threads of a half warp share their four rows: one atomic per row and half warp
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) ties += __shfl_xor_sync(HYP_FULL_MASK, ties, o);
        if (tx == 0 && ties && gi < p.n1) atomicAdd(p.row_ties + gi, ties);
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int64_t gj = j0 + 4 * tx + c;
          if (gi < gj && gj < p.n2 && gi < p.n1 && dist[c] == dist[c]) {
            if (MODE == kMin) {
              if (dist[c] < p.thr) ++below;
              Key k{dist[c], (int)gi, (int)gj};
              if (key_less(k, best)) best = k;
            } else if (dist[c] < p.thr) {
              const unsigned int bits = __float_as_uint(dist[c]);
              if (bits < p.cut_bits || (bits == p.cut_bits && gi <= p.cut_row)) {
                unsigned long long pos = atomicAdd(p.emit_count, 1ULL);
                if ((int64_t)pos < p.capacity) {
                  p.out_i[pos] = (int)gi;
                  p.out_j[pos] = (int)gj;
                  p.out_d[pos] = dist[c];
                }
              }
            }
          }
        }
      }
    }
  }

  if (MODE == kHist) {
    __syncthreads();
    for (int q = threadIdx.x; q < (1 << kHistBits); q += NTHREADS) {
      const unsigned int v = s_hist[q];
      if (v) atomicAdd(p.hist + q, (unsigned long long)v);
    }
  }
  if (MODE == kMin) {
    best = warp_key_min(best);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(HYP_FULL_MASK, below, o);
    if (lane == 0) {
      s_keys[warp] = best;
      s_cnt[warp] = below;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      Key b = s_keys[0];
      unsigned long long cnt = s_cnt[0];
      for (int w = 1; w < NTHREADS / 32; ++w) {
        if (key_less(s_keys[w], b)) b = s_keys[w];
        cnt += s_cnt[w];
      }
      p.ws->block_best[blockIdx.x] = b;
      if (cnt) atomicAdd(&p.ws->count, cnt);
      __threadfence();
      unsigned int tk = atomicAdd(&p.ws->ticket, 1u);
      s_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      Key b = key_none();
      for (int q = threadIdx.x; q < (int)gridDim.x; q += NTHREADS) {
        Key k = p.ws->block_best[q];
        if (key_less(k, b)) b = k;
      }
      b = warp_key_min(b);
      if (lane == 0) s_keys[warp] = b;
      __syncthreads();
      if (threadIdx.x == 0) {
        Key r = s_keys[0];
        for (int w = 1; w < NTHREADS / 32; ++w)
          if (key_less(s_keys[w], r)) r = s_keys[w];
        unsigned long long cnt = *((volatile unsigned long long *)&p.ws->count);
        hyp_best o;
        o.d = r.i < 0 ? __int_as_float(0x7f800000) : r.d;
        o.i = r.i;
        o.j = r.j;
        o.count_lo = (uint32_t)(cnt & 0xffffffffu);
        o.count_hi = (uint32_t)(cnt >> 32);
        o.pad[0] = o.pad[1] = o.pad[2] = 0;
        *p.best = o;
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Exact per-row k nearest neighbours: out[i][0..k) = the k smallest (d(i,j), j), j != i, ascending.
// One CTA owns a 64-row tile and streams every column tile past it, so the per-row candidate lists
// live in shared memory for the CTA's lifetime (no cross-CTA merging).  After each 64x64 tile the
// distances go through shared memory to 64 "row owner" threads that run a threshold test against
// the row's current k-th best and replace-max on the rare hit (expected k ln(V/k) inserts per row).
// Key = d_bits << 32 | j  (d >= 0, so the bit pattern orders like the value; ties break on j).
// ---------------------------------------------------------------------------------------------
constexpr int kMaxTopK = 64;

// With `row_list` (shard-relative row numbers, *row_count of them, both in device memory) the kernel serves only
// those rows, 64 list entries per tile: the device-driven redo of the rows the tensor-core filter flagged
// (gram_tc.cu), launched without the host knowing how many there are.
__global__ void __launch_bounds__(NTHREADS)
allpairs_topk_kernel(const float *__restrict__ E, int64_t ldE, int64_t n, int64_t row0, int64_t nrows, int D,
                     float sqrt_c, float sgn, int k, const TopkSink sink, const int32_t *__restrict__ row_list,
                     const int32_t *__restrict__ row_count) {
  extern __shared__ __align__(16) float smem[];
  const int d = D - 1;
  float *As = smem;
  float *Bs = As + (size_t)d * TPAD;
  float *a0 = Bs + (size_t)d * TPAD;
  float *b0 = a0 + TM;
  float *dist = b0 + TN;                                        // [64][65]
  // [64][k] keys, 8-byte aligned.  The offset is rounded in floats, not through an integer cast of the pointer: a
  // pointer that has been through uintptr_t loses its shared address space and every access becomes a generic LD/ST.
  const size_t lists_off = ((size_t)(2 * d) * TPAD + TM + TN + (size_t)TM * (TN + 1) + 1) & ~(size_t)1;
  unsigned long long *lists = reinterpret_cast<unsigned long long *>(smem + lists_off);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t col_tiles = (n + TN - 1) / TN;
  const int64_t n_listed = row_list ? (int64_t)*row_count : nrows;
  const int64_t row_tiles = (n_listed + TM - 1) / TM;
  const unsigned long long kEmpty = 0xffffffffffffffffULL;
  const int64_t i_end = (row0 + nrows < n) ? row0 + nrows : n;
  // table row served by slot r of tile rt (-1: none)
  auto row_of = [&](int64_t rt, int r) -> int64_t {
    const int64_t pos = rt * TM + r;
    if (pos >= n_listed) return -1;
    const int64_t gr = row_list ? row0 + row_list[pos] : row0 + pos;
    return gr < i_end ? gr : -1;
  };

  for (int64_t rt = blockIdx.x; rt < row_tiles; rt += gridDim.x) {
    __syncthreads();
    // rows of this tile: [i0, i0+64) clipped to the shard, or 64 entries of the list
    {
      const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
      for (int r = w; r < TM; r += NTHREADS / 32) {
        const int64_t gr = row_of(rt, r);
        const bool ok = gr >= 0;
        const float *row = E + (ok ? gr : 0) * ldE;
        for (int kk = lane; kk < d; kk += 32) As[kk * TPAD + r] = ok ? row[1 + kk] : 0.f;
        if (lane == 0) a0[r] = ok ? row[0] : 0.f;
      }
    }
    for (int q = threadIdx.x; q < TM * k; q += NTHREADS) lists[q] = kEmpty;
    // per-row running threshold (largest kept key) and its position: owner threads 0..63
    unsigned long long worst = kEmpty;
    int worst_pos = 0, filled = 0;

    for (int64_t ct = 0; ct < col_tiles; ++ct) {
      const int64_t j0 = ct * TN;
      __syncthreads();
      load_tile_T(E, ldE, n, j0, d, Bs, b0);
      __syncthreads();
      float S[4][4];
      tile_dots(As, Bs, d, ty, tx, S);
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float m = __fsub_rn(__fmul_rn(a0[4 * ty + r], b0[4 * tx + c]), S[r][c]);
          dist[(4 * ty + r) * (TN + 1) + 4 * tx + c] = dist_from_mdot(m, sgn, sqrt_c);
        }
      __syncthreads();
      if (threadIdx.x < TM) {
        const int r = threadIdx.x;
        const int64_t gi = row_of(rt, r);
        if (gi >= 0) {
          unsigned long long *mine = lists + (size_t)r * k;
          const int64_t jn = (j0 + TN < n) ? TN : n - j0;
          for (int c = 0; c < (int)jn; ++c) {
            const float dv = dist[r * (TN + 1) + c];
            const int64_t gj = j0 + c;
            if (!(dv == dv) || gj == gi) continue;
            const unsigned long long key = ((unsigned long long)__float_as_uint(dv) << 32) | (unsigned int)gj;
            if (filled < k) {
              mine[filled++] = key;
              if (filled == k) {
                worst = 0;
                for (int q = 0; q < k; ++q)
                  if (mine[q] >= worst) { worst = mine[q]; worst_pos = q; }
              }
            } else if (key < worst) {
              mine[worst_pos] = key;
              worst = 0;
              for (int q = 0; q < k; ++q)
                if (mine[q] >= worst) { worst = mine[q]; worst_pos = q; }
            }
          }
        }
      }
    }
    __syncthreads();
    // sort each row's list (insertion sort by its owner; k <= 64) and write out
    if (threadIdx.x < TM) {
      const int r = threadIdx.x;
      const int64_t gi = row_of(rt, r);
      if (gi >= 0) {
        unsigned long long *mine = lists + (size_t)r * k;
        for (int a = 1; a < k; ++a) {
          unsigned long long v = mine[a];
          int bpos = a - 1;
          while (bpos >= 0 && mine[bpos] > v) { mine[bpos + 1] = mine[bpos]; --bpos; }
          mine[bpos + 1] = v;
        }
        const int64_t orow = gi - row0;
        for (int q = 0; q < k; ++q) {
          const unsigned long long v = mine[q];
          const bool empty = v == kEmpty;
          sink_write(sink, orow, k, q, empty ? -1 : (int32_t)(v & 0xffffffffu),
                     empty ? __int_as_float(0x7f800000) : __uint_as_float((unsigned int)(v >> 32)));
        }
      }
    }
  }
}

static size_t tile_smem_bytes(int D) { return ((size_t)2 * (D - 1) * TPAD + TM + TN) * sizeof(float); }

template <int MODE>
static int launch_tiles(Params &p, cudaStream_t st, const char *what) {
  const size_t smem = tile_smem_bytes(p.D);
  if (smem > 200 * 1024) {
    set_error("%s: D=%d needs %zu bytes of shared memory per CTA (max 200 KiB => D <= 368)", what, p.D, smem);
    return HYP_ERR_ARG;
  }
  cudaError_t e = cudaFuncSetAttribute(allpairs_tile_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    return HYP_ERR_CUDA;
  }
  int per_sm = 0, dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, allpairs_tile_kernel<MODE>, NTHREADS, smem);
  if (per_sm < 1) per_sm = 1;
  int64_t grid = (int64_t)sms * per_sm;
  if (grid > kMaxBlocks) grid = kMaxBlocks;
  if (grid > p.n_tiles) grid = p.n_tiles;
  if (grid < 1) grid = 1;
  allpairs_tile_kernel<MODE><<<(int)grid, NTHREADS, smem, st>>>(p);
  return check_launch(what);
}

static int fill_params(Params &p, const float *x, int64_t ldx, int64_t n1, const float *y, int64_t ldy, int64_t n2,
                       int D, float c, int semantics, int triangular) {
  if (((!x || !y) && n1 > 0 && n2 > 0) || n1 < 0 || n2 < 0 || D < 2 || D > HYP_MAX_D || !(c > 0.f)) {
    set_error("all-pairs: bad arguments (n1=%lld n2=%lld D=%d c=%g)", (long long)n1, (long long)n2, D, c);
    return HYP_ERR_ARG;
  }
  p = Params{};
  p.x = x; p.ldx = ldx; p.n1 = n1; p.y = y; p.ldy = ldy; p.n2 = n2; p.D = D;
  p.sqrt_c = sqrtf(c);
  p.sgn = semantics == HYP_SEM_REFERENCE ? -1.f : 1.f;
  p.tiles_m = (n1 + TM - 1) / TM;
  p.tiles_n = (n2 + TN - 1) / TN;
  p.triangular = triangular;
  p.n_tiles = triangular ? p.tiles_m * (p.tiles_m + 1) / 2 : p.tiles_m * p.tiles_n;
  return HYP_OK;
}

}  // namespace hyp

using namespace hyp;

extern "C" int hyp_batch_distance(const float *x, int64_t ldx, int64_t n1, const float *y, int64_t ldy, int64_t n2,
                                  float *out, int64_t ldo, int D, float c, int semantics, void *stream) {
  Params p;
  int rc = fill_params(p, x, ldx, n1, y, ldy, n2, D, c, semantics, 0);
  if (rc) return rc;
  if (n1 == 0 || n2 == 0) return HYP_OK;
  if (!out) return HYP_ERR_ARG;
  p.out = out;
  p.ldo = ldo;
  return launch_tiles<kDense>(p, (cudaStream_t)stream, "hyp_batch_distance");
}

extern "C" int hyp_batch_distance_backward_coef(const float *x, int64_t ldx, int64_t n1, const float *y, int64_t ldy,
                                                int64_t n2, const float *grad_out, int64_t ldg, float *w,
                                                int64_t ldw, int D, float c, int semantics, void *stream) {
  Params p;
  int rc = fill_params(p, x, ldx, n1, y, ldy, n2, D, c, semantics, 0);
  if (rc) return rc;
  if (n1 == 0 || n2 == 0) return HYP_OK;
  if (!grad_out || !w) return HYP_ERR_ARG;
  p.out = w;
  p.ldo = ldw;
  p.gout = grad_out;
  p.ldg = ldg;
  return launch_tiles<kGradCoef>(p, (cudaStream_t)stream, "hyp_batch_distance_backward_coef");
}

extern "C" int64_t hyp_allpairs_workspace_bytes(int64_t) { return (int64_t)sizeof(MinWorkspace); }

extern "C" int hyp_allpairs_min(const float *E, int64_t ldE, int64_t n, int D, float c, int semantics, float threshold,
                                hyp_best *best, void *workspace, int64_t workspace_bytes, void *stream) {
  Params p;
  int rc = fill_params(p, E, ldE, n, E, ldE, n, D, c, semantics, 1);
  if (rc) return rc;
  if (!best || !workspace) return HYP_ERR_ARG;
  if (workspace_bytes < (int64_t)sizeof(MinWorkspace)) {
    set_error("hyp_allpairs_min: workspace %lld < %zu bytes", (long long)workspace_bytes, sizeof(MinWorkspace));
    return HYP_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(workspace, 0, 16, st);
  p.ws = (MinWorkspace *)workspace;
  p.best = best;
  p.thr = threshold;
  return launch_tiles<kMin>(p, st, "hyp_allpairs_min");
}

extern "C" int hyp_allpairs_emit(const float *E, int64_t ldE, int64_t n, int D, float c, int semantics, float threshold,
                                 int32_t *out_i, int32_t *out_j, float *out_d, int64_t capacity,
                                 unsigned long long *count, void *stream) {
  Params p;
  int rc = fill_params(p, E, ldE, n, E, ldE, n, D, c, semantics, 1);
  if (rc) return rc;
  if (!count || capacity < 0 || (capacity > 0 && (!out_i || !out_j || !out_d))) return HYP_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(count, 0, sizeof(unsigned long long), st);
  if (n < 2) return HYP_OK;
  p.thr = threshold;
  p.out_i = out_i; p.out_j = out_j; p.out_d = out_d;
  p.capacity = capacity;
  p.emit_count = count;
  p.cut_bits = 0xffffffffu;                  // no cut: every candidate
  p.cut_row = 0;
  return launch_tiles<kEmit>(p, st, "hyp_allpairs_emit");
}

extern "C" int hyp_allpairs_emit_cut(const float *E, int64_t ldE, int64_t n, int D, float c, int semantics,
                                     float threshold, uint32_t cut_bits, int64_t cut_row, int32_t *out_i, int32_t *out_j,
                                     float *out_d, int64_t capacity, unsigned long long *count, void *stream) {
  Params p;
  int rc = fill_params(p, E, ldE, n, E, ldE, n, D, c, semantics, 1);
  if (rc) return rc;
  if (!count || capacity < 0 || (capacity > 0 && (!out_i || !out_j || !out_d))) return HYP_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(count, 0, sizeof(unsigned long long), st);
  if (n < 2) return HYP_OK;
  p.thr = threshold;
  p.out_i = out_i; p.out_j = out_j; p.out_d = out_d;
  p.capacity = capacity;
  p.emit_count = count;
  p.cut_bits = cut_bits;
  p.cut_row = cut_row;
  return launch_tiles<kEmit>(p, st, "hyp_allpairs_emit_cut");
}

extern "C" int hyp_allpairs_hist(const float *E, int64_t ldE, int64_t n, int D, float c, int semantics, float threshold,
                                 uint32_t prefix, int prefix_len, int bin_bits, unsigned long long *hist, void *stream) {
  Params p;
  int rc = fill_params(p, E, ldE, n, E, ldE, n, D, c, semantics, 1);
  if (rc) return rc;
  if (!hist || prefix_len < 0 || bin_bits < 1 || bin_bits > kHistBits || prefix_len + bin_bits > 31) {
    set_error("hyp_allpairs_hist: bad arguments (prefix_len=%d bin_bits=%d: 1 <= bin_bits <= %d, prefix_len + bin_bits <= 31)",
              prefix_len, bin_bits, kHistBits);
    return HYP_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(hist, 0, sizeof(unsigned long long) << kHistBits, st);
  if (n < 2) return HYP_OK;
  p.thr = threshold;
  p.prefix = prefix;
  p.prefix_shift = prefix_len == 0 ? 32u : (unsigned int)(31 - prefix_len);
  p.bin_shift = (unsigned int)(31 - prefix_len - bin_bits);
  p.bin_mask = (1u << bin_bits) - 1u;
  p.hist = hist;
  return launch_tiles<kHist>(p, st, "hyp_allpairs_hist");
}

extern "C" int hyp_allpairs_row_ties(const float *E, int64_t ldE, int64_t n, int D, float c, int semantics,
                                     float threshold, uint32_t d_bits, unsigned int *row_ties, void *stream) {
  Params p;
  int rc = fill_params(p, E, ldE, n, E, ldE, n, D, c, semantics, 1);
  if (rc) return rc;
  if (!row_ties && n > 0) return HYP_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (n > 0) cudaMemsetAsync(row_ties, 0, sizeof(unsigned int) * (size_t)n, st);
  if (n < 2) return HYP_OK;
  p.thr = threshold;
  p.cut_bits = d_bits;
  p.row_ties = row_ties;
  return launch_tiles<kRowTies>(p, st, "hyp_allpairs_row_ties");
}

namespace hyp {
// Shared by hyp_allpairs_topk and the tensor-core path's redo of flagged rows (row_list / row_count in device memory).
int launch_allpairs_topk(const float *E, int64_t ldE, int64_t n, int64_t row0, int64_t nrows, int D, float c,
                         int semantics, int k, const TopkSink &sink, const int32_t *row_list, const int32_t *row_count,
                         cudaStream_t st) {
  const size_t smem = tile_smem_bytes(D) + ((size_t)TM * (TN + 1) + 4) * sizeof(float) + (size_t)TM * k * 8 + 16;
  if (smem > 200 * 1024) {
    set_error("hyp_allpairs_topk: D=%d k=%d needs %zu bytes of shared memory per CTA", D, k, smem);
    return HYP_ERR_ARG;
  }
  cudaError_t e = cudaFuncSetAttribute(allpairs_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("hyp_allpairs_topk: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return HYP_ERR_CUDA;
  }
  int per_sm = 0, dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, allpairs_topk_kernel, NTHREADS, smem);
  if (per_sm < 1) per_sm = 1;
  int64_t row_tiles = (nrows + TM - 1) / TM;
  int64_t grid = (int64_t)sms * per_sm;
  if (grid > row_tiles) grid = row_tiles;
  allpairs_topk_kernel<<<(int)grid, NTHREADS, smem, st>>>(E, ldE, n, row0, nrows, D, sqrtf(c),
                                                          semantics == HYP_SEM_REFERENCE ? -1.f : 1.f, k, sink, row_list,
                                                          row_count);
  return check_launch("hyp_allpairs_topk");
}
}  // namespace hyp

extern "C" int hyp_allpairs_topk(const float *E, int64_t ldE, int64_t n, int64_t row0, int64_t nrows, int D, float c,
                                 int semantics, int k, int32_t *out_idx, float *out_d, void *stream) {
  if (n < 0 || row0 < 0 || nrows < 0 || row0 + nrows > n || D < 2 || D > HYP_MAX_D || !(c > 0.f) || k < 1 ||
      k > kMaxTopK) {
    set_error("hyp_allpairs_topk: bad arguments (n=%lld row0=%lld nrows=%lld D=%d k=%d, k <= %d)", (long long)n,
              (long long)row0, (long long)nrows, D, k, kMaxTopK);
    return HYP_ERR_ARG;
  }
  if (nrows == 0) return HYP_OK;
  if (!E || !out_idx || !out_d) return HYP_ERR_ARG;
  TopkSink sink{};
  sink.idx = out_idx;
  sink.d = out_d;
  return launch_allpairs_topk(E, ldE, n, row0, nrows, D, c, semantics, k, sink, nullptr, nullptr, (cudaStream_t)stream);
}
