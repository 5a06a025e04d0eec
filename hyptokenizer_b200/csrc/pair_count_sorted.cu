// pair_count_sorted.cu -- K6, general path: adjacent-token pair counts over a token-id stream (sm_100a).
//
// FrequencyAwareHyperbolicTokenizer._compute_pair_frequencies (tokenizer/frequency_aware_hyperbolic_merge.py:92-112)
// counts the adjacent pairs of `self.tokenize(line.strip())` per line.  With empty merge rules that is the character
// bigram count of pair_count.cu (the case the reference actually reaches inside __init__); with rules
// (a tokenizer that was load()ed, tokenize() first called after merges: hyperbolic_merge.py:414-446) the tokens are
// multi-character and the alphabet is the vocabulary -- up to 10^5 symbols, so a dense table is out.  This file counts
// over the token ids hyp_apply_merges produces: 64-bit keys (first << 32 | second), one per token slot that has a
// successor in its text, a device radix sort, a run-length encode.  HBM-bound integer work: the stream is read once,
// the keys are written once and pass through the sort's digit passes (CUB DeviceRadixSort / DeviceRunLengthEncode from
// the CUDA toolkit do the sort and the segmented reduction; key generation and the finish are ours).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>

#include "common.cuh"

namespace hyp {

// Token ids are int32: symbol ids >= 0, characters without an id -(code point + 1) >= -0x110000.  Biased by 2^31 they
// order as unsigned, and the key of "no pair here" -- (INT_MIN, INT_MIN), not a token -- becomes 0 and sorts first.
__device__ __forceinline__ unsigned long long biased_key(int32_t a, int32_t b) {
  return ((unsigned long long)((uint32_t)a ^ 0x80000000u) << 32) | ((uint32_t)b ^ 0x80000000u);
}

// one thread per token slot of the [offsets[t], offsets[t + 1]) layout hyp_apply_merges writes: slot s of text t holds a
// token iff s - offsets[t] < n_tokens[t]; it forms a pair with its successor iff that one holds a token too
__global__ void __launch_bounds__(256)
token_pair_keys_kernel(const int32_t *__restrict__ tokens, const int64_t *__restrict__ offsets,
                       const int32_t *__restrict__ n_tokens, int64_t n_texts, int64_t n_slots,
                       unsigned long long *__restrict__ keys) {
  // a warp per text: texts are a few hundred bytes, lanes stride over the slots of their text
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t t = warp; t < n_texts; t += nwarps) {
    const int64_t b0 = offsets[t], b1 = offsets[t + 1];
    const int64_t nt = n_tokens[t];
    for (int64_t s = b0 + lane; s < b1; s += 32) {
      const int64_t k = s - b0;
      keys[s] = (k + 1 < nt) ? biased_key(tokens[s], tokens[s + 1]) : 0ull;
    }
  }
  (void)n_slots;
}

__global__ void __launch_bounds__(256)
token_pair_finish_kernel(const unsigned long long *__restrict__ uk, const int32_t *__restrict__ uc,
                         const int32_t *__restrict__ num_runs, unsigned long long *__restrict__ out_keys,
                         unsigned long long *__restrict__ out_counts, int64_t capacity, int64_t *__restrict__ n_unique) {
  const int64_t runs = *num_runs;
  const int64_t skip = (runs > 0 && uk[0] == 0ull) ? 1 : 0;          // the "no pair here" run sorts first
  const int64_t n = runs - skip;
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_unique = n;             // may exceed `capacity`: the caller checks
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n && i < capacity;
       i += (int64_t)gridDim.x * blockDim.x) {
    out_keys[i] = uk[i + skip] ^ 0x8000000080000000ull;
    out_counts[i] = (unsigned long long)uc[i + skip];
  }
}

struct SortedLayout {
  size_t off_a, off_b, off_uc, off_runs, off_tmp, tmp_bytes, total;
};

static SortedLayout sorted_layout(int64_t n_slots) {
  SortedLayout L{};
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 255) & ~(size_t)255; return at; };
  const size_t N = (size_t)(n_slots > 0 ? n_slots : 1);
  L.off_a = take(N * 8);
  L.off_b = take(N * 8);
  L.off_uc = take(N * 4);
  L.off_runs = take(256);
  size_t t1 = 0, t2 = 0;
  cub::DoubleBuffer<unsigned long long> db(nullptr, nullptr);
  cub::DeviceRadixSort::SortKeys(nullptr, t1, db, (int64_t)N, 0, 64, (cudaStream_t)0);
  cub::DeviceRunLengthEncode::Encode(nullptr, t2, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                     (int32_t *)nullptr, (int32_t *)nullptr, (int64_t)N, (cudaStream_t)0);
  L.tmp_bytes = t1 > t2 ? t1 : t2;
  L.off_tmp = take(L.tmp_bytes);
  L.total = o;
  return L;
}

}  // namespace hyp

using namespace hyp;

extern "C" int64_t hyp_pair_count_sorted_workspace_bytes(int64_t n_slots) {
  if (n_slots < 0) return -1;
  return (int64_t)sorted_layout(n_slots).total;
}

extern "C" int hyp_pair_count_sorted(const int32_t *tokens, const int64_t *offsets, const int32_t *n_tokens,
                                     int64_t n_texts, int64_t n_slots, unsigned long long *out_keys,
                                     unsigned long long *out_counts, int64_t capacity, int64_t *n_unique,
                                     void *workspace, int64_t workspace_bytes, void *stream) {
  if (n_texts < 0 || n_slots < 0 || capacity < 0 || !n_unique || (capacity > 0 && (!out_keys || !out_counts)) ||
      (n_texts > 0 && (!offsets || !n_tokens)) || (n_slots > 0 && !tokens) || n_slots >= (1LL << 31) * 2) {
    set_error("hyp_pair_count_sorted: bad arguments (n_texts=%lld n_slots=%lld capacity=%lld)", (long long)n_texts,
              (long long)n_slots, (long long)capacity);
    return HYP_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (n_slots == 0 || n_texts == 0) {
    cudaMemsetAsync(n_unique, 0, sizeof(int64_t), st);
    return HYP_OK;
  }
  const SortedLayout L = sorted_layout(n_slots);
  if (!workspace || workspace_bytes < (int64_t)L.total || ((uintptr_t)workspace & 255) != 0) {
    set_error("hyp_pair_count_sorted: workspace must be 256-byte aligned and hold %zu bytes (got %lld)", L.total,
              (long long)workspace_bytes);
    return HYP_ERR_WORKSPACE;
  }
  uint8_t *ws = (uint8_t *)workspace;
  unsigned long long *ka = (unsigned long long *)(ws + L.off_a), *kb = (unsigned long long *)(ws + L.off_b);
  int32_t *uc = (int32_t *)(ws + L.off_uc), *runs = (int32_t *)(ws + L.off_runs);
  void *tmp = ws + L.off_tmp;
  // slots behind the last text (or between texts) never get a key: start from "no pair" everywhere
  cudaMemsetAsync(ka, 0, (size_t)n_slots * 8, st);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t blocks = (n_texts + 7) / 8;
  if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
  token_pair_keys_kernel<<<(int)blocks, 256, 0, st>>>(tokens, offsets, n_tokens, n_texts, n_slots, ka);
  int rc = check_launch("hyp_pair_count_sorted(keys)");
  if (rc) return rc;
  cub::DoubleBuffer<unsigned long long> db(ka, kb);
  size_t tb = L.tmp_bytes;
  cudaError_t e = cub::DeviceRadixSort::SortKeys(tmp, tb, db, n_slots, 0, 64, st);
  if (e == cudaSuccess) {
    tb = L.tmp_bytes;
    unsigned long long *sorted = db.Current();
    unsigned long long *uk = sorted == ka ? kb : ka;                  // the buffer the sort left free
    e = cub::DeviceRunLengthEncode::Encode(tmp, tb, (const unsigned long long *)sorted, uk, uc, runs, n_slots, st);
    if (e == cudaSuccess) {
      int64_t fb = (capacity + 255) / 256;
      if (fb < 1) fb = 1;
      if (fb > (int64_t)sms * 8) fb = (int64_t)sms * 8;
      token_pair_finish_kernel<<<(int)fb, 256, 0, st>>>(uk, uc, runs, out_keys, out_counts, capacity, n_unique);
      return check_launch("hyp_pair_count_sorted(finish)");
    }
  }
  set_error("hyp_pair_count_sorted: CUB: %s", cudaGetErrorString(e));
  return HYP_ERR_CUDA;
}
