// apply_merges.cu -- batched tokenize(): apply the merge rules to many texts at once (sm_100a).
//
// Replaces, for a whole corpus, the per-string Python loop of HyperbolicTokenizer.tokenize
// (tokenizer/hyperbolic_merge.py:414-446), the one thing the reference's own benchmark measures
// (scripts/benchmark_efficiency.py:58-94).  Semantics reproduced exactly: tokens = list(text); repeat
// left-to-right passes until a pass changes nothing; at position i, if (tokens[i], tokens[i+1]) is a rule key,
// tokens[i] becomes the merged token, tokens[i+1] is deleted and the scan STAYS at i (chain merge); no rule
// priority; later duplicates of a key overwrite earlier ones (that is resolved on the host when the table is
// built, as the reference's dict does).
//
// Tokens are symbol ids (host-assigned: one id per distinct string that occurs in a rule or as a
// single-character vocabulary entry); characters without an id travel as -(code point + 1) and never match.
// One thread per text: each text is an independent sequential rewrite; its token array lives in the
// caller-provided scratch at the text's own byte offset (a text never has more tokens than bytes).
#include "common.cuh"

namespace hyp {

__device__ __forceinline__ int rule_lookup(const unsigned long long *__restrict__ keys, const int32_t *__restrict__ vals,
                                           uint32_t mask, int a, int b) {
  if (a < 0 || b < 0) return -1;
  const unsigned long long key = ((unsigned long long)(uint32_t)a << 32) | (uint32_t)b;
  uint32_t h = (uint32_t)((key * 0x9E3779B97F4A7C15ULL) >> 32) & mask;
  for (;;) {
    const unsigned long long k = __ldg(keys + h);
    if (k == key) return __ldg(vals + h);
    if (k == 0xffffffffffffffffULL) return -1;
    h = (h + 1) & mask;
  }
}

__global__ void __launch_bounds__(128)
apply_merges_kernel(const uint8_t *__restrict__ text, const int64_t *__restrict__ offsets, int64_t n_texts,
                    const int32_t *__restrict__ ascii_sym, const uint32_t *__restrict__ cp_sorted,
                    const int32_t *__restrict__ cp_sym, int n_cp, const unsigned long long *__restrict__ rkeys,
                    const int32_t *__restrict__ rvals, uint32_t rmask, int32_t *__restrict__ tok,
                    int32_t *__restrict__ n_tok) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_texts) return;
  const int64_t b0 = offsets[t], b1 = offsets[t + 1];
  int32_t *out = tok + b0;
  // ---- list(text): decode UTF-8 into symbol ids ------------------------------------------------------
  int m = 0;
  for (int64_t p = b0; p < b1;) {
    uint32_t c = text[p];
    int need = 1;
    if (c >= 0x80) {
      need = (c >= 0xF0) ? 4 : (c >= 0xE0) ? 3 : 2;
      c &= 0xFF >> (need + 1);
      for (int k = 1; k < need && p + k < b1; ++k) c = (c << 6) | (text[p + k] & 0x3F);
    }
    p += need;
    int sym;
    if (c < 128) {
      sym = __ldg(ascii_sym + c);
    } else {
      int lo = 0, hi = n_cp;          // binary search among the non-ASCII code points that have an id
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(cp_sorted + mid) < c) lo = mid + 1; else hi = mid;
      }
      sym = (lo < n_cp && __ldg(cp_sorted + lo) == c) ? __ldg(cp_sym + lo) : -1;
    }
    out[m++] = sym >= 0 ? sym : -(int32_t)(c + 1);
  }
  // ---- passes until nothing changes (hyperbolic_merge.py:434-444) ---------------------------------------
  bool changed = m > 1;
  while (changed) {
    changed = false;
    int w = 0, r = 1;
    int cur = out[0];
    for (;;) {
      if (r < m) {
        const int merged = rule_lookup(rkeys, rvals, rmask, cur, out[r]);
        if (merged >= 0) {          // tokens[i] = merged; tokens.pop(i+1); stay at i
          cur = merged;
          ++r;
          changed = true;
          continue;
        }
      }
      out[w++] = cur;               // i += 1
      if (r >= m) break;
      cur = out[r++];
    }
    m = w;
    if (m < 2) break;
  }
  n_tok[t] = m;
}

}  // namespace hyp

using namespace hyp;

extern "C" int hyp_apply_merges(const uint8_t *text, const int64_t *offsets, int64_t n_texts, const int32_t *ascii_sym,
                                const uint32_t *cp_sorted, const int32_t *cp_sym, int32_t n_cp,
                                const unsigned long long *rule_keys, const int32_t *rule_vals, int64_t rule_capacity,
                                int32_t *tokens, int32_t *n_tokens, void *stream) {
  if (n_texts < 0 || n_cp < 0 || rule_capacity < 2 || (rule_capacity & (rule_capacity - 1)) != 0 ||
      rule_capacity > (1LL << 31)) {
    set_error("hyp_apply_merges: bad arguments (rule_capacity must be a power of two)");
    return HYP_ERR_ARG;
  }
  if (n_texts == 0) return HYP_OK;
  if (!offsets || !ascii_sym || !rule_keys || !rule_vals || !tokens || !n_tokens || (n_cp > 0 && (!cp_sorted || !cp_sym)))
    return HYP_ERR_ARG;
  const int64_t blocks = (n_texts + 127) / 128;
  apply_merges_kernel<<<(unsigned int)blocks, 128, 0, (cudaStream_t)stream>>>(
      text, offsets, n_texts, ascii_sym, cp_sorted, cp_sym, n_cp, rule_keys, rule_vals, (uint32_t)(rule_capacity - 1),
      tokens, n_tokens);
  return check_launch("hyp_apply_merges");
}
