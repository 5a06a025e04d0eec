// gemv_topk.cu -- K4, top-k form: ONE row against the table, the k nearest rows (sm_100a).
//
// The per-merge incremental update of the north star: the merged token's new row is scored against all n rows (the
// reference recomputes the whole n x n matrix instead, tokenizer/hyperbolic_merge.py:247-269) and the k smallest
// (distance, index) keys are kept -- what a query of the reference's FAISS index returns for one vector
// (fast_hyperbolic_merge.py:301-304), in the true Lorentz distance.  hyp_row_min (merge_loop.cu) is the k = 1 form the
// device-resident loop uses; this entry serves host-driven steps and single-row neighbour queries.
//   row_keys_kernel   a warp per table row: exact ATen-order Minkowski product with the query, distance,
//                     key = d bits << 32 | row (NaN distances and the excluded row: all ones).  HBM-bound: 4 n D bytes.
//   key_select_kernel one CTA: radix select of the k-th smallest 64-bit key (8 passes of 8 bits over the key array,
//                     which is L2-resident), then the <= k keys up to it, rank-sorted.
// Compiled with -fmad=false: the products round exactly like the reference's.
#include "common.cuh"

namespace hyp {

constexpr unsigned long long kNoKey64 = 0xffffffffffffffffULL;
constexpr int kSelThreadsK = 1024;
constexpr int kMaxGemvK = 64;

__global__ void __launch_bounds__(256)
row_keys_kernel(const float *__restrict__ E, int64_t ldE, int64_t n, const float *__restrict__ q, int64_t exclude, int D,
                float sqrt_c, float sgn, unsigned long long *__restrict__ keys) {
  extern __shared__ float qs[];                        // the query row, staged once per block
  for (int e = threadIdx.x; e < D; e += blockDim.x) qs[e] = q[e];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n; r += nwarps) {
    // operand order as the all-pairs kernels: <table row, query> (the product is commutative term by term)
    const float m = warp_mdot(E + r * ldE, qs, D, lane);
    if (lane == 0) {
      const float d = dist_from_mdot(m, sgn, sqrt_c);
      keys[r] = (r == exclude || !(d == d)) ? kNoKey64 : (((unsigned long long)__float_as_uint(d) << 32) | (unsigned int)r);
    }
  }
}

__global__ void __launch_bounds__(kSelThreadsK)
key_select_kernel(const unsigned long long *__restrict__ keys, int64_t n, int k, int32_t *__restrict__ out_idx,
                  float *__restrict__ out_d) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_want;
  __shared__ unsigned long long list[kMaxGemvK];
  __shared__ int s_count;
  const int tid = threadIdx.x;
  if (tid == 0) { s_prefix = 0; s_want = k; s_count = 0; }
  // the k-th smallest key, one byte per pass from the top: keys are distinct (the row index is part of them), except
  // for the "no key" value, which sorts last
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    const unsigned long long prefix = s_prefix;
    unsigned int prev = 0xffffffffu, run = 0;           // equal bins of a thread's consecutive keys are merged
    for (int64_t i = tid; i < n; i += kSelThreadsK) {
      const unsigned long long key = keys[i];
      if (pass == 0 || (key >> (shift + 8)) == (prefix >> (shift + 8))) {
        const unsigned int bin = (unsigned int)(key >> shift) & 0xffu;
        if (bin == prev) ++run;
        else {
          if (run) atomicAdd(&hist[prev], run);
          prev = bin;
          run = 1;
        }
      }
    }
    if (run) atomicAdd(&hist[prev], run);
    __syncthreads();
    if (tid == 0) {
      int want = s_want;
      unsigned int b = 0;
      for (; b < 255u; ++b) {
        if ((unsigned int)want <= hist[b]) break;
        want -= (int)hist[b];
      }
      s_want = want;
      s_prefix = prefix | ((unsigned long long)b << shift);
    }
    __syncthreads();
  }
  const unsigned long long kth = s_prefix;               // (fewer than k real keys: the descent ends on "no key")
  for (int64_t i = tid; i < n; i += kSelThreadsK) {
    const unsigned long long key = keys[i];
    if (key <= kth && key != kNoKey64) {
      const int pos = atomicAdd(&s_count, 1);
      if (pos < kMaxGemvK) list[pos] = key;
    }
  }
  __syncthreads();
  const int m = s_count < k ? s_count : k;
  if (tid < m) {
    const unsigned long long mine = list[tid];
    int rank = 0;
    for (int t = 0; t < m; ++t) rank += list[t] < mine;
    out_idx[rank] = (int32_t)(mine & 0xffffffffu);
    out_d[rank] = __uint_as_float((unsigned int)(mine >> 32));
  }
  if (tid >= m && tid < k) {
    out_idx[tid] = -1;
    out_d[tid] = __int_as_float(0x7f800000);
  }
}

}  // namespace hyp

using namespace hyp;

extern "C" int64_t hyp_gemv_topk_workspace_bytes(int64_t n) { return n < 0 ? -1 : (n > 0 ? n : 1) * 8; }

extern "C" int hyp_gemv_topk(const float *E, int64_t ldE, int64_t n, const float *q, int64_t exclude_row, int D, float c,
                             int semantics, int k, int32_t *out_idx, float *out_d, void *workspace,
                             int64_t workspace_bytes, void *stream) {
  if (n < 0 || D < 2 || D > HYP_MAX_D || !(c > 0.f) || k < 1 || k > kMaxGemvK || !out_idx || !out_d ||
      (n > 0 && (!E || !q))) {
    set_error("hyp_gemv_topk: bad arguments (n=%lld D=%d k=%d, k <= %d)", (long long)n, D, k, kMaxGemvK);
    return HYP_ERR_ARG;
  }
  if (n > 0 && (!workspace || workspace_bytes < n * 8 || ((uintptr_t)workspace & 7) != 0)) {
    set_error("hyp_gemv_topk: workspace must be 8-byte aligned and hold %lld bytes", (long long)(n * 8));
    return HYP_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long *keys = (unsigned long long *)workspace;
  if (n > 0) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t blocks = (n + 7) / 8;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    row_keys_kernel<<<(int)blocks, 256, (size_t)D * sizeof(float), st>>>(E, ldE, n, q, exclude_row, D, sqrtf(c),
                                                                        semantics == HYP_SEM_REFERENCE ? -1.f : 1.f, keys);
    int rc = check_launch("hyp_gemv_topk(keys)");
    if (rc) return rc;
  }
  key_select_kernel<<<1, kSelThreadsK, 0, st>>>(keys, n, k, out_idx, out_d);
  return check_launch("hyp_gemv_topk(select)");
}
