// pair_count.cu -- K6: adjacent code-point pair counting over a UTF-8 corpus (sm_100a).
//
// Replaces FrequencyAwareHyperbolicTokenizer._compute_pair_frequencies
// (tokenizer/frequency_aware_hyperbolic_merge.py:92-112).  Inside __init__ the tokenizer's merge
// rules are empty, so `self.tokenize(line.strip())` is `list(line.strip())` and the
// reference counts, for every line of the text-mode file (universal newlines: '\n', '\r', '\r\n'),
// the adjacent code-point pairs of the line with leading/trailing `str.isspace()` characters
// removed.  A pair (a, b) at consecutive positions of one line is therefore counted iff some
// non-space character sits at or before `a` in that line and some non-space sits at or after `b`.
//
// HBM-bound integer/byte work: the stream is read once in 16-byte vectors into shared memory,
// ASCII pairs go to a CTA-private 128x128 shared-memory histogram (flushed once), everything else
// to a global open-addressing table keyed by (cp_a << 32 | cp_b).
#include "common.cuh"

namespace hyp {

constexpr int kPcThreads = 512;
constexpr int kChunk = 16 * 1024;      // bytes of text per CTA iteration
constexpr int kHalo = 64;              // bytes staged on both sides of a chunk for neighbour lookups
constexpr unsigned long long kEmptyKey = 0xffffffffffffffffULL;

__device__ __forceinline__ bool is_nl(uint32_t cp) { return cp == 0x0a || cp == 0x0d; }

// str.isspace() over all of Unicode (enumerated from CPython 3.12)
__device__ __forceinline__ bool is_space(uint32_t cp) {
  if (cp < 0x80) return (cp >= 0x09 && cp <= 0x0d) || (cp >= 0x1c && cp <= 0x20);
  return cp == 0x85 || cp == 0xa0 || cp == 0x1680 || (cp >= 0x2000 && cp <= 0x200a) || cp == 0x2028 ||
         cp == 0x2029 || cp == 0x202f || cp == 0x205f || cp == 0x3000;
}

// Byte accessor: shared-memory window [w0, w1) of the text, global memory outside it.
struct Text {
  const uint8_t *g;
  int64_t n;
  const uint8_t *s;
  int64_t w0, w1;
  __device__ __forceinline__ uint32_t at(int64_t p) const {
    return (p >= w0 && p < w1) ? s[p - w0] : __ldg(g + p);
  }
  __device__ __forceinline__ bool is_start(int64_t p) const { return (at(p) & 0xC0) != 0x80; }
  // decode the code point starting at p (valid UTF-8 assumed; truncated tail -> what is there)
  __device__ __forceinline__ uint32_t decode(int64_t p, int &len) const {
    uint32_t b0 = at(p);
    if (b0 < 0x80) { len = 1; return b0; }
    int need = (b0 >= 0xF0) ? 4 : (b0 >= 0xE0) ? 3 : 2;
    uint32_t cp = b0 & (0xFF >> (need + 1));
    int got = 1;
    for (; got < need && p + got < n; ++got) cp = (cp << 6) | (at(p + got) & 0x3F);
    len = got;
    return cp;
  }
  __device__ __forceinline__ int64_t prev_start(int64_t p) const {  // start of the code point before p, or -1
    int64_t q = p - 1;
    while (q >= 0 && !is_start(q)) --q;
    return q;
  }
};

__device__ __forceinline__ void hash_add(unsigned long long *keys, unsigned long long *vals, uint32_t cap_mask,
                                         unsigned long long key, int *overflow) {
  uint32_t h = (uint32_t)((key * 0x9E3779B97F4A7C15ULL) >> 32) & cap_mask;
  for (uint32_t probe = 0; probe <= cap_mask; ++probe) {
    unsigned long long cur = keys[h];
    if (cur == kEmptyKey) {
      unsigned long long old = atomicCAS(keys + h, kEmptyKey, key);
      cur = (old == kEmptyKey) ? key : old;
    }
    if (cur == key) {
      atomicAdd(vals + h, 1ULL);
      return;
    }
    h = (h + 1) & cap_mask;
  }
  atomicExch(overflow, 1);
}

__global__ void __launch_bounds__(kPcThreads)
pair_count_kernel(const uint8_t *__restrict__ text, int64_t n, unsigned long long *__restrict__ ascii_counts,
                  unsigned long long *hkeys, unsigned long long *hvals, uint32_t cap_mask, int *overflow) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint32_t *hist = reinterpret_cast<uint32_t *>(smem_raw);                 // [128*128]
  uint8_t *win = smem_raw + 128 * 128 * sizeof(uint32_t);                  // [kHalo + kChunk + kHalo]
  for (int k = threadIdx.x; k < 128 * 128; k += blockDim.x) hist[k] = 0;

  const int64_t n_chunks = (n + kChunk - 1) / kChunk;
  for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const int64_t c0 = ch * kChunk;
    const int64_t c1 = (c0 + kChunk < n) ? c0 + kChunk : n;
    const int64_t w0 = (c0 - kHalo > 0) ? c0 - kHalo : 0;      // c0, kHalo multiples of 16 -> w0 16-aligned
    const int64_t w1 = (c1 + kHalo < n) ? c1 + kHalo : n;
    __syncthreads();
    // stage the window with 16-byte loads (text base is 16-byte aligned: torch allocations are)
    const int64_t nvec = (w1 - w0) >> 4;
    const uint4 *src = reinterpret_cast<const uint4 *>(text + w0);
    for (int64_t v = threadIdx.x; v < nvec; v += blockDim.x) reinterpret_cast<uint4 *>(win)[v] = __ldg(src + v);
    for (int64_t b = (nvec << 4) + threadIdx.x; b < w1 - w0; b += blockDim.x) win[b] = __ldg(text + w0 + b);
    __syncthreads();
    Text T{text, n, win, w0, w1};

    // One thread takes 16 consecutive bytes.  ASCII fast path: the 16 bytes plus 4 bytes of context on either
    // side sit in six registers, and a pair (a, b) is counted when neither is a line break and
    //   left : a is not a space, or the byte before a is neither space nor break        (else: scan back)
    //   right: b is not a space, or the byte after  b is neither space nor break        (else: scan ahead)
    // Anything else in the 24-byte neighbourhood (a non-ASCII byte, a run of two or more spaces next to the
    // pair) takes the general per-position path below.
    for (int64_t base = c0 + 16 * (int64_t)threadIdx.x; base < c1; base += 16 * (int64_t)blockDim.x) {
      uint32_t W[6];
      {
        const uint4 mid = *reinterpret_cast<const uint4 *>(win + (base - w0));
        W[1] = mid.x; W[2] = mid.y; W[3] = mid.z; W[4] = mid.w;
        W[0] = (base - 4 >= w0) ? *reinterpret_cast<const uint32_t *>(win + (base - 4 - w0)) : 0x0a0a0a0au;
        W[5] = (base + 20 <= w1) ? *reinterpret_cast<const uint32_t *>(win + (base + 16 - w0)) : 0x0a0a0a0au;
      }
      const int64_t valid = n - base;                      // bytes of this 16-byte group that exist
      bool fast = valid >= 20 && base - 4 >= w0;           // groups touching either end of the text go slow
      fast = fast && (((W[0] | W[1] | W[2] | W[3] | W[4] | W[5]) & 0x80808080u) == 0);
      if (fast) {
        auto byte_at = [&](int j) -> uint32_t { return (W[(j + 4) >> 2] >> (8 * ((j + 4) & 3))) & 0xffu; };
        const unsigned long long space_mask = (0x1fULL << 9) | (0x1fULL << 28);   // 9..13, 28..32
        auto sp = [&](uint32_t c) -> bool { return c < 64 && ((space_mask >> c) & 1ULL); };
        auto nl = [&](uint32_t c) -> bool { return c == 0x0a || c == 0x0d; };
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t a = byte_at(i), bch = byte_at(i + 1);
          if (nl(a) || nl(bch)) continue;
          bool ok = true;
          bool slow = false;
          if (sp(a)) {
            const uint32_t pa = byte_at(i - 1);
            if (nl(pa)) ok = false;
            else if (sp(pa)) slow = true;
          }
          if (ok && sp(bch)) {
            const uint32_t nb = byte_at(i + 2);
            if (nl(nb)) ok = false;
            else if (sp(nb)) slow = true;
          }
          if (!ok) continue;
          if (slow) {
            // a run of spaces next to the pair: resolve with the general scans
            const int64_t pp = base + i;
            bool left = !sp(a), right = !sp(bch);
            if (!left) {
              int64_t q = pp - 1;
              while (q >= 0) {
                const uint32_t cq = T.at(q);
                if (cq >= 0x80) { int l2; const int64_t qs = T.is_start(q) ? q : T.prev_start(q); const uint32_t cp = T.decode(qs, l2);
                                  if (!is_space(cp)) { left = true; break; } q = qs - 1; continue; }
                if (nl(cq)) break;
                if (!sp(cq)) { left = true; break; }
                --q;
              }
            }
            if (left && !right) {
              int64_t q = pp + 2;
              while (q < n) {
                int l2;
                const uint32_t cq = T.decode(q, l2);
                if (is_nl(cq)) break;
                if (!is_space(cq)) { right = true; break; }
                q += l2;
              }
            }
            if (!(left && right)) continue;
          }
          atomicAdd(&hist[a * 128 + bch], 1u);
        }
        continue;
      }
      const int64_t p_end = (base + 16 < c1) ? base + 16 : c1;
      for (int64_t p = base; p < p_end; ++p) {
      if (!T.is_start(p)) continue;
      int la, lb;
      const uint32_t a = T.decode(p, la);
      if (is_nl(a)) continue;
      const int64_t pb = p + la;
      if (pb >= n) continue;
      const uint32_t b = T.decode(pb, lb);
      if (is_nl(b)) continue;
      // left condition: a non-space at or before `a` on this line
      bool left = !is_space(a);
      if (!left) {
        int64_t q = T.prev_start(p);
        while (q >= 0) {
          int l2;
          uint32_t cq = T.decode(q, l2);
          if (is_nl(cq)) break;
          if (!is_space(cq)) { left = true; break; }
          q = T.prev_start(q);
        }
      }
      if (!left) continue;
      bool right = !is_space(b);
      if (!right) {
        int64_t q = pb + lb;
        while (q < n) {
          int l2;
          uint32_t cq = T.decode(q, l2);
          if (is_nl(cq)) break;
          if (!is_space(cq)) { right = true; break; }
          q += l2;
        }
      }
      if (!right) continue;
      if (a < 128 && b < 128) {
        atomicAdd(&hist[a * 128 + b], 1u);
      } else {
        hash_add(hkeys, hvals, cap_mask, ((unsigned long long)a << 32) | b, overflow);
      }
      }
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < 128 * 128; k += blockDim.x) {
    uint32_t v = hist[k];
    if (v) atomicAdd(ascii_counts + k, (unsigned long long)v);
  }
}

}  // namespace hyp

using namespace hyp;

extern "C" int hyp_pair_count(const uint8_t *text, int64_t n_bytes, unsigned long long *ascii_counts,
                              unsigned long long *hash_keys, unsigned long long *hash_vals, int64_t hash_capacity,
                              int *overflow, void *stream) {
  if (n_bytes < 0 || !ascii_counts || !hash_keys || !hash_vals || !overflow || hash_capacity < 2 ||
      (hash_capacity & (hash_capacity - 1)) != 0 || hash_capacity > (1LL << 31) || (n_bytes > 0 && !text)) {
    set_error("hyp_pair_count: bad arguments (hash_capacity must be a power of two)");
    return HYP_ERR_ARG;
  }
  if (((uintptr_t)text & 15) != 0) {
    set_error("hyp_pair_count: text must be 16-byte aligned");
    return HYP_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(ascii_counts, 0, 128 * 128 * sizeof(unsigned long long), st);
  cudaMemsetAsync(hash_keys, 0xff, hash_capacity * sizeof(unsigned long long), st);
  cudaMemsetAsync(hash_vals, 0, hash_capacity * sizeof(unsigned long long), st);
  cudaMemsetAsync(overflow, 0, sizeof(int), st);
  if (n_bytes == 0) return HYP_OK;
  const size_t smem = 128 * 128 * sizeof(uint32_t) + kChunk + 2 * kHalo;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(pair_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t chunks = (n_bytes + kChunk - 1) / kChunk;
  int grid = (int)(chunks < (int64_t)sms * 2 ? chunks : (int64_t)sms * 2);
  pair_count_kernel<<<grid, kPcThreads, smem, st>>>(text, n_bytes, ascii_counts, hash_keys, hash_vals,
                                                    (uint32_t)(hash_capacity - 1), overflow);
  return check_launch("hyp_pair_count");
}
