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
// HBM-bound integer/byte work, two kernels with identical results.  v1 (kept, HYP_PAIR_COUNT=v1): the stream is read
// once in 16-byte vectors into shared memory, ASCII pairs go to a CTA-private 128x128 shared-memory histogram with
// atomics (flushed once).  v2 (default, further down): private non-atomic counters for the frequent symbols.  In both,
// pairs with a non-ASCII code point go to a global open-addressing table keyed by (cp_a << 32 | cp_b).
#include <cstdlib>

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


// ---------------------------------------------------------------------------------------------------------------
// v2: lane-private counters.
//
// v1 above retires one shared-memory atomic per input byte, and ATOMS runs at ~1 lane per clock per SM on this part:
// 148 SMs x 1.97 GHz = the ~310 GB/s it measures, with 80 thread instructions per byte of per-position branching on top.
// v2 removes both (680 GB/s on the same stream):
//  * the ASCII bytes are ranked by frequency (one ranking per stream, from the sample pair_count_select_kernel takes;
//    until late round 2 every CTA ranked its own first chunk).  Pairs of the 27 most frequent symbols are
//    counted in PRIVATE one-byte counters, one column of 28 x 28 bins per PAIR of lanes, word-interleaved so that column
//    c of a warp lives in bank c (+16 in odd rows): an update is a plain LDS.U8 / +1 / STS.U8, no atomics and no bank conflicts, in two
//    predicated phases (even lanes, then odd lanes) because two lanes share a column.  A counter that wraps carries
//    256 into the global table.  Pairs within the 80 most frequent symbols go to a CTA histogram with ATOMS, anything
//    rarer straight to the global table.  Counts stay exact for any input; the alphabet only decides how fast.
//  * the hot loop has no data-dependent branch: a byte outside the private alphabet has the last private rank, whose row
//    and column of the table are junk bins; the byte offset of a bin is one add of two table entries,
//    A[first byte] + B[second byte]; pairs that are not counted are predicated off;
//  * classification is by table as well: C[byte] carries "space", "line break" and "outside the private alphabet" in
//    three 10-bit planes, and acc = 2 acc + C[byte] over the thread's 19-byte neighbourhood yields the three position
//    masks with one IMAD per byte; the count / resolve-slowly decisions of all 16 pairs are mask algebra;
//  * no barrier in the main loop: a thread reads its 16 bytes (+ 4 of context on either side) straight from global
//    memory, a warp 512 contiguous bytes, the next step's words in flight while this step is counted.  One CTA of 16
//    warps per SM (the private counters take 16 x 12.25 KB of its shared memory).
// What was measured on the way (B200, 256 MiB of the config-4 stream): lane-private columns with 6 warps 240 GB/s
// (latency-bound, every pair a chain of dependent branches) -> branch-free 398 -> two groups per thread side by side
// 383 (no gain: not the chain) -> lane-pair columns + 12 warps, still staged through shared-memory windows 361 (one
// barrier per 512 bytes per warp) -> barrier-free 438 -> predicated phases 454 -> table classification 487 (569 on
// 1 GiB) -> a 28 x 28 table, which fits 16 warps instead of 12: 680 on 1 GiB.  ncu before that last step: 533 warp
// instructions per 512 bytes, issue slots 53 % busy, shared-memory pipe 54 %, a third of the stalls on LDS -> +1.
#ifndef HYP_PC_SYMS
#define HYP_PC_SYMS 28      // 27 private symbols, 16 warps per SM: 680 GB/s; 32 (31 symbols, 12 warps): 569 GB/s
#endif
constexpr int kPrivSyms = HYP_PC_SYMS;                     // private table dimension: kPrivSyms - 1 symbols + the junk rank
constexpr int kPrivDiv = (kPrivSyms + 3) / 4;             // ranks sharing a byte lane of the word: ra / kPrivDiv
constexpr int kPrivRows = kPrivDiv * kPrivSyms;           // word rows per column
constexpr uint32_t kJunkRank = kPrivSyms - 1;
constexpr int kV2Threads = HYP_PC_SYMS <= 28 ? 512 : 384;
constexpr int kV2Warps = kV2Threads / 32;
constexpr int kV2Chunk = kV2Threads * 16;                 // bytes of text per CTA step (8 KiB), one 16-byte group per thread
constexpr int kPrivPerWarp = kPrivRows * 64;              // rows x 16 lane pairs x 4 bytes
// CTA histogram (shared-memory atomics) for the pairs outside the private alphabet: the 80 most frequent symbols.  It
// was 64 x 64 until late round 2; a 75-symbol stream then sent 7 % of its pairs to the GLOBAL table, where the atomics
// of all SMs on a few hundred hot addresses serialise in L2 (most of the 6.1 ms per GiB of that stream).
constexpr int kV2Hist = HYP_PC_SYMS <= 28 ? 80 : 64;
constexpr size_t kV2Smem = (size_t)kV2Warps * kPrivPerWarp + kV2Hist * kV2Hist * 4 + 512 + 512 + 1024 + 256 + 128;
static_assert(kV2Smem <= 232448, "v2 shared memory");

// byte offset of private bin (ra, rb) inside a lane pair's column: word row (ra % kPrivDiv) * kPrivSyms + rb (16 words
// per row), byte ra / kPrivDiv of the word
__device__ __forceinline__ uint32_t priv_off_a(uint32_t ra) { return (ra % kPrivDiv) * (kPrivSyms * 64u) + ra / kPrivDiv; }
__device__ __forceinline__ uint32_t priv_off_b(uint32_t rb) { return rb * 64u; }

struct V2Ctx {
  uint32_t *hist;                // [kV2Hist * kV2Hist]
  const uint8_t *sym;            // byte -> frequency rank 0 .. kV2Hist - 1, 0xff = none
  unsigned long long *ascii_counts;
  // cold path: pairs resolved by the general scans, pairs outside the private alphabet.  It never touches the private
  // counters (a column is shared by two lanes and only the converged hot loop keeps them apart).
  __device__ __forceinline__ void add(uint32_t a, uint32_t b) const {
    const uint32_t ra = sym[a], rb = sym[b];
    if (ra < (uint32_t)kV2Hist && rb < (uint32_t)kV2Hist) atomicAdd(&hist[ra * kV2Hist + rb], 1u);
    else atomicAdd(ascii_counts + a * 128u + b, 1ULL);
  }
};

// One 16-byte group of a thread, classified: which pairs are counted outright, which need the general scans, which
// of the counted ones lie outside the private alphabet, and the two table entries of each of its 17 bytes.
struct V2Group {
  int64_t base;
  uint32_t counted, slow, other;
  bool general;                  // a non-ASCII byte or an end of the text in the neighbourhood: per-position path
  uint32_t A[16], B[17];         // B[0] is unused
};

// the 24-byte neighbourhood of the group at `base` (a multiple of 16): bytes base-4 .. base+19 as six words, straight
// from global memory (a warp reads 512 contiguous bytes; the context words hit the lines its neighbours fetch).
// Anything that is not there -- before the text, or within 20 bytes of its end -- reads as 0x80 and sends the group
// down the per-position path.
__device__ __forceinline__ void v2_load(uint32_t (&W)[6], const uint8_t *__restrict__ text, int64_t n, int64_t base) {
#pragma unroll
  for (int k = 0; k < 6; ++k) W[k] = 0x80808080u;
  if (base + 20 <= n) {
    const uint4 mid = __ldg(reinterpret_cast<const uint4 *>(text + base));
    W[1] = mid.x; W[2] = mid.y; W[3] = mid.z; W[4] = mid.w;
    W[5] = __ldg(reinterpret_cast<const uint32_t *>(text + base + 16));
    if (base >= 4) W[0] = __ldg(reinterpret_cast<const uint32_t *>(text + base - 4));
  }
}

__device__ __forceinline__ void v2_prepare(V2Group &g, int64_t base, int64_t n, const uint32_t (&W)[6],
                                           const uint16_t *tabA, const uint16_t *tabB, const uint32_t *tabC) {
  g.base = base;
  g.counted = g.slow = g.other = 0;
  g.general = false;
#pragma unroll
  for (int j = 0; j < 16; ++j) g.A[j] = g.B[j + 1] = 0;
  if (base >= n) return;
  if (((W[0] | W[1] | W[2] | W[3] | W[4] | W[5]) & 0x80808080u) != 0) {
    g.general = true;
    return;
  }
  // Classification by table: tabC[byte] = space | line break << 10 | outside the private alphabet << 20, and
  // acc = 2 acc + tabC[byte] over positions 8 .. -1 (and 17 .. 9) leaves three 10-bit planes, position -1 (9) at bit 0.
  auto byte_at = [&](int j) -> uint32_t { return (W[(j + 4) >> 2] >> (8 * ((j + 4) & 3))) & 0xffu; };
  uint32_t lo = 0, hi = 0;
#pragma unroll
  for (int j = 17; j >= -1; --j) {
    const uint32_t c = byte_at(j);
    if (j >= 0 && j < 16) g.A[j] = tabA[c];
    if (j >= 1 && j < 17) g.B[j] = tabB[c];
    if (j >= 9) hi = hi * 2u + tabC[c];
    else lo = lo * 2u + tabC[c];
  }
  // bit k of a plane = position k - 1; pair i = (byte i, byte i + 1): prev at bit i, a at i + 1, b at i + 2, next at i + 3
  const uint32_t S = (lo & 0x3ffu) | ((hi & 0x1ffu) << 10);
  const uint32_t N = ((lo >> 10) & 0x3ffu) | (((hi >> 10) & 0x1ffu) << 10);
  const uint32_t P = ((lo >> 20) & 0x3ffu) | (((hi >> 20) & 0x1ffu) << 10);
  const uint32_t valid = ~((N >> 1) | (N >> 2)) & 0xffffu;
  const uint32_t left_unres = (S >> 1) & S & ~N, right_unres = (S >> 2) & (S >> 3) & ~(N >> 3);
  g.counted = valid & (~(S >> 1) | ~S) & (~(S >> 2) | ~(S >> 3));
  g.slow = valid & (left_unres | right_unres);
  g.other = g.counted & ((P >> 1) | (P >> 2));
}

__global__ void __launch_bounds__(kV2Threads, 1)
pair_count_v2_kernel(const uint8_t *__restrict__ text, int64_t n, unsigned long long *__restrict__ ascii_counts,
                     unsigned long long *hkeys, unsigned long long *hvals, uint32_t cap_mask, int *overflow,
                     const int *__restrict__ select) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  if (*select != 2) return;                    // the stream's alphabet chose the other kernel (pair_count_select_kernel)
  uint8_t *priv = smem_raw;                                                    // [warps][256 word rows][32 lanes][4]
  uint32_t *hist64 = reinterpret_cast<uint32_t *>(smem_raw + kV2Warps * kPrivPerWarp);   // [kV2Hist * kV2Hist]
  uint16_t *tabA = reinterpret_cast<uint16_t *>(hist64 + kV2Hist * kV2Hist);    // byte -> priv_off_a(private rank)
  uint16_t *tabB = tabA + 256;                                                  // byte -> priv_off_b(private rank)
  uint32_t *tabC = reinterpret_cast<uint32_t *>(tabB + 256);                    // byte -> class bits, see v2_prepare
  uint8_t *sym = reinterpret_cast<uint8_t *>(tabC + 256);                                                    // byte -> frequency rank, 0xff = none
  uint8_t *inv = sym + 256;                                                     // rank -> byte
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n_chunks = (n + kV2Chunk - 1) / kV2Chunk;

  int64_t ch = blockIdx.x;
  if (ch >= n_chunks) return;
  {  // frequency ranks of the ASCII bytes: one ranking for the whole stream (pair_count_select_kernel)
    uint32_t rank = tid < 128 ? reinterpret_cast<const uint8_t *>(select + 1)[tid] : 0xffu;
    if (rank >= (uint32_t)kV2Hist) rank = 0xffu;
    if (tid < 128) {
      const uint32_t pr = rank < kJunkRank ? rank : kJunkRank;
      sym[tid] = (uint8_t)rank;
      sym[128 + tid] = 0xff;
      tabA[tid] = (uint16_t)priv_off_a(pr);
      tabB[tid] = (uint16_t)priv_off_b(pr);
      const uint32_t t = tid;
      const uint32_t is_sp = (t >= 0x09 && t <= 0x0d) || (t >= 0x1c && t <= 0x20), is_br = t == 0x0a || t == 0x0d;
      tabC[tid] = is_sp | (is_br << 10) | ((pr == kJunkRank ? 1u : 0u) << 20);
      tabA[128 + tid] = (uint16_t)priv_off_a(kJunkRank);
      tabB[128 + tid] = (uint16_t)priv_off_b(kJunkRank);
      tabC[128 + tid] = 1u << 20;
      if (rank != 0xffu) inv[rank] = (uint8_t)tid;
    }
    uint4 *z = reinterpret_cast<uint4 *>(smem_raw);
    const int nz = (kV2Warps * kPrivPerWarp + kV2Hist * kV2Hist * 4) / 16;
    for (int k = tid; k < nz; k += kV2Threads) z[k] = make_uint4(0, 0, 0, 0);
    __syncthreads();
  }
  uint8_t *const priv_col = priv + warp * kPrivPerWarp + (lane >> 1) * 4;   // shared by lanes 2t and 2t + 1
  const bool odd = lane & 1;
  const V2Ctx ctx{hist64, sym, ascii_counts};
  const Text T{text, n, nullptr, 0, 0};        // the cold paths read the text from global memory (L1)

  // No barrier in the main loop: every warp streams its own 512 bytes per step, the next step's words in flight
  // while this step is counted.
  const int64_t step = (int64_t)gridDim.x * kV2Chunk;
  int64_t base = ch * kV2Chunk + 16 * (int64_t)tid;
  uint32_t Wn[6];
  v2_load(Wn, text, n, base);
  for (; ch < n_chunks; ch += gridDim.x, base += step) {
    uint32_t W[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) W[k] = Wn[k];
    if (ch + gridDim.x < n_chunks) v2_load(Wn, text, n, base + step);

    V2Group g;
    v2_prepare(g, base, n, W, tabA, tabB, tabC);

    // Hot loop, executed by the whole (converged) warp: one LDS.U8 -> +1 -> STS.U8 per pair, in two phases because a
    // column of counters belongs to a PAIR of lanes -- even lanes update, then odd lanes.  A lane that has nothing to
    // add in a phase issues nothing: the RMW is predicated.  __syncwarp orders the phases for the compiler too.
    const uint32_t mine_e = odd ? 0u : g.counted, mine_o = odd ? g.counted : 0u;
    uint32_t wacc[4] = {0, 0, 0, 0};           // byte 3 - (i & 3) of wacc[i >> 2]: 1 if the counter of pair i passed 255
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      uint8_t *p = priv_col + (g.A[i] + g.B[i + 1]);
      uint32_t l = 0;                          // the counter before this lane's +1
      if ((mine_e >> i) & 1u) { l = *p; *p = (uint8_t)(l + 1u); }
      __syncwarp();
      if ((mine_o >> i) & 1u) { l = *p; *p = (uint8_t)(l + 1u); }
      __syncwarp();
      wacc[i >> 2] = __byte_perm(wacc[i >> 2], l + 1u, 0x2105);
    }
    uint32_t wrap = 0;                         // bit i: the counter of pair i passed 255
    if (wacc[0] | wacc[1] | wacc[2] | wacc[3]) {
#pragma unroll
      for (int i = 0; i < 16; ++i) wrap |= ((wacc[i >> 2] >> (8 * (3 - (i & 3)))) & 1u) << i;
    }

    // Cold paths.
    {
      uint32_t wrapped = wrap & ~g.other;
      while (wrapped) {                          // a private counter passed 255: carry into the global table
        const int i = __ffs(wrapped) - 1;
        wrapped &= wrapped - 1;
        atomicAdd(ascii_counts + T.at(base + i) * 128u + T.at(base + i + 1), 256ULL);
      }
      uint32_t other = g.other;
      while (other) {                            // counted pairs outside the private alphabet
        const int i = __ffs(other) - 1;
        other &= other - 1;
        ctx.add(T.at(base + i), T.at(base + i + 1));
      }
      uint32_t slow = g.slow;
      while (slow) {
        // a run of spaces next to the pair: resolve with the general scans
        const int i = __ffs(slow) - 1;
        slow &= slow - 1;
        const int64_t pp = base + i;
        const uint32_t a = T.at(pp), bch = T.at(pp + 1);
        auto sp = [&](uint32_t c) -> bool { return (c >= 0x09 && c <= 0x0d) || (c >= 0x1c && c <= 0x20); };
        bool left = !sp(a), right = !sp(bch);
        if (!left) {
          int64_t q = pp - 1;
          while (q >= 0) {
            const uint32_t cq = T.at(q);
            if (cq >= 0x80) {
              int l2;
              const int64_t qs = T.is_start(q) ? q : T.prev_start(q);
              const uint32_t cp = T.decode(qs, l2);
              if (!is_space(cp)) { left = true; break; }
              q = qs - 1;
              continue;
            }
            if (is_nl(cq)) break;
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
        if (left && right) ctx.add(a, bch);
      }
      if (g.general) {
      const int64_t p_end = (base + 16 < n) ? base + 16 : n;
      for (int64_t p = base; p < p_end; ++p) {
        if (!T.is_start(p)) continue;
        int la, lb;
        const uint32_t a = T.decode(p, la);
        if (is_nl(a)) continue;
        const int64_t pb = p + la;
        if (pb >= n) continue;
        const uint32_t b = T.decode(pb, lb);
        if (is_nl(b)) continue;
        bool left = !is_space(a);
        if (!left) {
          int64_t q = T.prev_start(p);
          while (q >= 0) {
            int l2;
            const uint32_t cq = T.decode(q, l2);
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
            const uint32_t cq = T.decode(q, l2);
            if (is_nl(cq)) break;
            if (!is_space(cq)) { right = true; break; }
            q += l2;
          }
        }
        if (!right) continue;
        if (a < 128 && b < 128) ctx.add(a, b);
        else hash_add(hkeys, hvals, cap_mask, ((unsigned long long)a << 32) | b, overflow);
      }
      }
    }
  }
  __syncthreads();

  // flush: private counters (sum over the warp's 16 columns per bin), then the CTA histogram
  const uint32_t *rows = reinterpret_cast<const uint32_t *>(priv + warp * kPrivPerWarp);
  for (int g = 0; g < kPrivRows; ++g) {
    const uint32_t w = lane < 16 ? rows[g * 16 + lane] : 0u;
    const uint32_t s0 = __reduce_add_sync(HYP_FULL_MASK, w & 0xffu), s1 = __reduce_add_sync(HYP_FULL_MASK, (w >> 8) & 0xffu);
    const uint32_t s2 = __reduce_add_sync(HYP_FULL_MASK, (w >> 16) & 0xffu), s3 = __reduce_add_sync(HYP_FULL_MASK, w >> 24);
    if (lane < 4) {
      const uint32_t s = lane == 0 ? s0 : lane == 1 ? s1 : lane == 2 ? s2 : s3;
      const uint32_t ra = (uint32_t)(g / kPrivSyms) + kPrivDiv * lane, rb = g % kPrivSyms;   // inverse of priv_off_a/b
      if (s && ra < kJunkRank && rb != kJunkRank)
        atomicAdd(ascii_counts + (uint32_t)inv[ra] * 128u + inv[rb], (unsigned long long)s);
    }
  }
  for (int k = tid; k < kV2Hist * kV2Hist; k += kV2Threads) {
    const uint32_t v = hist64[k];
    if (v) atomicAdd(ascii_counts + (uint32_t)inv[k / kV2Hist] * 128u + inv[k % kV2Hist], (unsigned long long)v);
  }
}


// ---------------------------------------------------------------------------------------------------------------
// v3 (default): count first, correct later.
//
// v2 decides per pair whether it counts (strip() and line breaks) BEFORE touching a counter: three table lookups and a
// mask computation per byte, and a two-phase update because two lanes share a column -- 33 thread instructions and 7
// shared-memory operations per byte (ncu, round 1), 10 % of the HBM roofline.  v3 turns it around:
//  * every adjacent byte pair of an all-ASCII 16-byte group is counted UNCONDITIONALLY in the private table.  A line
//    break has the junk rank, so pairs that touch one land in junk bins by themselves; what strip() excludes -- pairs
//    inside the leading / trailing white space of a line -- always contains two ADJACENT bytes <= 0x20, which one AND /
//    OR per byte over the table entries detects.  Only such a group (and one with a byte outside the private alphabet, a
//    non-ASCII byte or an end of the text) enters the cold path, which classifies it exactly like v2 and CORRECTS: -1 in
//    the global table for a private pair that did not count, +1 for a counted pair outside the private alphabet;
//  * one table lookup per byte: tab[c] = row offset << 16 | flags | column offset, a pair's counter is at
//    row(first) + column(second) + the lane's base, one IADD3;
//  * one column of 28 x 28 one-byte bins per LANE, word-interleaved (lane L owns bank L): plain LDS.U8 / +1 / STS.U8 in
//    ONE phase, conflict-free for any data.  Two consecutive pairs are updated together (both loads, then both stores);
//    when they hit the same bin the second store carries +2.  8 warps per SM (8 x 24.5 KB of counters), so the stream
//    is prefetched four 4 KiB steps ahead into registers to keep ~16 KB per SM in flight.
// Result bits are identical to v1 / v2 for any input (tests/test_gpu_paircount.py runs all three).
constexpr int kV3Syms = 28;                                // 27 private symbols + the junk rank
constexpr uint32_t kV3Junk = kV3Syms - 1;
constexpr int kV3Threads = 256, kV3Warps = kV3Threads / 32;
constexpr int kV3Chunk = kV3Threads * 16;                  // 4 KiB of text per CTA step, one 16-byte group per thread
constexpr int kV3RowBytes = (kV3Syms / 4) * 128;           // a row of 28 bins = 7 words per lane x 32 lanes
constexpr int kV3PrivPerWarp = 25600;                      // 28 rows = 25 088 bytes, padded so that the low 10 bits of a
                                                           // warp's base are zero (it is OR-ed into the offsets)
static_assert(kV3PrivPerWarp >= kV3Syms * kV3RowBytes && (kV3PrivPerWarp & 1023) == 0, "v3 counter block");
constexpr int kV3Depth = 4;                                // steps in flight per thread: 16 KB per SM
constexpr int kV3Bins = kV3Syms * kV3Syms;                 // 784
constexpr size_t kV3Smem = (size_t)kV3Warps * kV3PrivPerWarp + 64 * 64 * 4 + kV3Bins * 4 + 256 * 4 + 256 * 4 + 256 + 64 + 64;
constexpr uint32_t kV3FlagJ = 1u << 12;                    // ASCII byte outside the private alphabet (not a line break)
constexpr uint32_t kV3FlagW = 1u << 13;                    // byte <= 0x20: every str.isspace() ASCII byte and line break

__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v)); }

struct V3Cold {
  uint32_t *hist64;
  const uint8_t *sym;
  const uint32_t *tabC;
  unsigned long long *ascii_counts, *hkeys, *hvals;
  uint32_t cap_mask;
  int *overflow;
  const uint8_t *text;
  int64_t n;
};

// does the pair at (pp, pp + 1), neither byte a line break, survive strip()?  (general scans over the text)
__device__ __noinline__ bool v3_pair_counts(const Text &T, int64_t pp, int64_t n) {
  auto sp = [&](uint32_t c) -> bool { return (c >= 0x09 && c <= 0x0d) || (c >= 0x1c && c <= 0x20); };
  const uint32_t a = T.at(pp), bch = T.at(pp + 1);
  bool left = !sp(a), right = !sp(bch);
  if (!left) {
    int64_t q = pp - 1;
    while (q >= 0) {
      const uint32_t cq = T.at(q);
      if (cq >= 0x80) {
        int l2;
        const int64_t qs = T.is_start(q) ? q : T.prev_start(q);
        const uint32_t cp = T.decode(qs, l2);
        if (!is_space(cp)) { left = true; break; }
        q = qs - 1;
        continue;
      }
      if (is_nl(cq)) break;
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
  return left && right;
}

// Everything the hot loop does not decide, for ONE 16-byte group (called by the lanes that need it, diverged).
//   classify the group may hold pairs that strip() excludes or bytes outside the private alphabet
//   general  the group was not counted at all (non-ASCII byte or an end of the text nearby): per-position path
__device__ __noinline__ void v3_cold(const V3Cold k, int64_t base, bool classify, bool general,
                                     uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4, uint32_t w5) {
  const Text T{k.text, k.n, nullptr, 0, 0};
  auto add = [&](uint32_t a, uint32_t b) {
    const uint32_t ra = k.sym[a], rb = k.sym[b];
    if ((ra | rb) < 64u) atomicAdd(&k.hist64[ra * 64u + rb], 1u);
    else atomicAdd(k.ascii_counts + a * 128u + b, 1ULL);
  };
  if (classify) {
    const uint32_t W[6] = {w0, w1, w2, w3, w4, w5};
    auto byte_at = [&](int j) -> uint32_t { return (W[(j + 4) >> 2] >> (8 * ((j + 4) & 3))) & 0xffu; };
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int j = 17; j >= -1; --j) {
      const uint32_t c = byte_at(j);
      if (j >= 9) hi = hi * 2u + k.tabC[c];
      else lo = lo * 2u + k.tabC[c];
    }
    // planes as in v2_prepare: bit q = position q - 1; pair i: prev at bit i, a at i + 1, b at i + 2, next at i + 3
    const uint32_t S = (lo & 0x3ffu) | ((hi & 0x1ffu) << 10);
    const uint32_t N = ((lo >> 10) & 0x3ffu) | (((hi >> 10) & 0x1ffu) << 10);
    const uint32_t P = ((lo >> 20) & 0x3ffu) | (((hi >> 20) & 0x1ffu) << 10);
    const uint32_t valid = ~((N >> 1) | (N >> 2)) & 0xffffu;
    const uint32_t left_unres = (S >> 1) & S & ~N, right_unres = (S >> 2) & (S >> 3) & ~(N >> 3);
    const uint32_t counted = valid & (~(S >> 1) | ~S) & (~(S >> 2) | ~(S >> 3));
    const uint32_t slow = valid & (left_unres | right_unres);
    const uint32_t priv = valid & ~((P >> 1) | (P >> 2));      // both bytes in the private alphabet: the hot loop counted it
    uint32_t minus = priv & ~counted & ~slow;                  // ... but strip() excludes it
    while (minus) {
      const int i = __ffs(minus) - 1;
      minus &= minus - 1;
      atomicAdd(k.ascii_counts + byte_at(i) * 128u + byte_at(i + 1), ~0ULL);
    }
    uint32_t other = counted & ~priv;                          // counts, and the hot loop put it into a junk bin
    while (other) {
      const int i = __ffs(other) - 1;
      other &= other - 1;
      add(byte_at(i), byte_at(i + 1));
    }
    uint32_t rest = slow;
    while (rest) {
      const int i = __ffs(rest) - 1;
      rest &= rest - 1;
      const bool counts = v3_pair_counts(T, base + i, k.n);
      const bool is_priv = (priv >> i) & 1u;
      if (is_priv && !counts) atomicAdd(k.ascii_counts + byte_at(i) * 128u + byte_at(i + 1), ~0ULL);
      if (!is_priv && counts) add(byte_at(i), byte_at(i + 1));
    }
  }
  if (general) {
    const int64_t p_end = (base + 16 < k.n) ? base + 16 : k.n;
    for (int64_t p = base; p < p_end; ++p) {
      if (!T.is_start(p)) continue;
      int la, lb;
      const uint32_t a = T.decode(p, la);
      if (is_nl(a)) continue;
      const int64_t pb = p + la;
      if (pb >= k.n) continue;
      const uint32_t b = T.decode(pb, lb);
      if (is_nl(b)) continue;
      bool left = !is_space(a);
      if (!left) {
        int64_t q = T.prev_start(p);
        while (q >= 0) {
          int l2;
          const uint32_t cq = T.decode(q, l2);
          if (is_nl(cq)) break;
          if (!is_space(cq)) { left = true; break; }
          q = T.prev_start(q);
        }
      }
      if (!left) continue;
      bool right = !is_space(b);
      if (!right) {
        int64_t q = pb + lb;
        while (q < k.n) {
          int l2;
          const uint32_t cq = T.decode(q, l2);
          if (is_nl(cq)) break;
          if (!is_space(cq)) { right = true; break; }
          q += l2;
        }
      }
      if (!right) continue;
      if (a < 128 && b < 128) add(a, b);
      else hash_add(k.hkeys, k.hvals, k.cap_mask, ((unsigned long long)a << 32) | b, k.overflow);
    }
  }
}

__global__ void __launch_bounds__(kV3Threads, 1)
pair_count_v3_kernel(const uint8_t *__restrict__ text, int64_t n, unsigned long long *__restrict__ ascii_counts,
                     unsigned long long *hkeys, unsigned long long *hvals, uint32_t cap_mask, int *overflow,
                     const int *__restrict__ select) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  if (*select != 3) return;                    // the stream's alphabet chose the other kernel (pair_count_select_kernel)
  uint8_t *priv = smem_raw;                                                     // [warps][28 rows][7 words][32 lanes][4]
  uint32_t *hist64 = reinterpret_cast<uint32_t *>(smem_raw + kV3Warps * kV3PrivPerWarp);   // [64*64]
  uint32_t *carry = hist64 + 64 * 64;                                           // [28*28]: wraps of the one-byte counters
  uint32_t *tab = carry + kV3Bins;                                              // byte -> row << 16 | flags | column
  uint32_t *tabC = tab + 256;                                                   // byte -> class planes (cold path)
  uint8_t *sym = reinterpret_cast<uint8_t *>(tabC + 256);                       // byte -> frequency rank, 0xff = none
  uint8_t *inv = sym + 256;                                                     // rank -> byte
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n_chunks = (n + kV3Chunk - 1) / kV3Chunk;

  int64_t ch = blockIdx.x;
  if (ch >= n_chunks) return;
  {  // frequency ranks of the ASCII bytes: one ranking for the whole stream (pair_count_select_kernel)
    uint32_t rank = tid < 128 ? reinterpret_cast<const uint8_t *>(select + 1)[tid] : 0xffu;
    if (rank >= 64u) rank = 0xffu;
    if (tid < 128) {
      const uint32_t pr = rank < kV3Junk ? rank : kV3Junk;
      const uint32_t t = tid;
      const uint32_t is_sp = (t >= 0x09 && t <= 0x0d) || (t >= 0x1c && t <= 0x20), is_br = t == 0x0a || t == 0x0d;
      const uint32_t junk_off = (kV3Junk * kV3RowBytes) << 16 | ((kV3Junk >> 2) * 128u + (kV3Junk & 3u));
      sym[tid] = (uint8_t)rank;
      sym[128 + tid] = 0xff;
      tab[tid] = ((pr * kV3RowBytes) << 16) | ((pr >> 2) * 128u + (pr & 3u)) |
                 ((pr == kV3Junk && !is_br) ? kV3FlagJ : 0u) | (t <= 0x20 ? kV3FlagW : 0u);
      tab[128 + tid] = junk_off;
      tabC[tid] = is_sp | (is_br << 10) | ((pr == kV3Junk ? 1u : 0u) << 20);
      tabC[128 + tid] = 1u << 20;
      if (rank != 0xffu) inv[rank] = (uint8_t)tid;
    }
    uint4 *z = reinterpret_cast<uint4 *>(smem_raw);
    const int nz = (kV3Warps * kV3PrivPerWarp + 64 * 64 * 4 + kV3Bins * 4) / 16;
    for (int q = tid; q < nz; q += kV3Threads) z[q] = make_uint4(0, 0, 0, 0);
    __syncthreads();
  }
  // A counter's byte offset inside the warp's block is row(first) + column(second) + 4 * lane.  The column offsets use
  // bits 0-1 and 7-9, 4 * lane bits 2-6: `(t & 0x383) | lane4` is one LOP3, the row is added by one LEA.HI, and the warp's
  // base rides in the address mode of LDS / STS.  volatile: the read-modify-writes of one thread stay in program order.
  volatile uint8_t *const col = priv;
  const uint32_t lane4 = 4u * lane | (uint32_t)warp * kV3PrivPerWarp;   // (+ the warp's block: bits 10 and up)
  const V3Cold cold{hist64, sym, tabC, ascii_counts, hkeys, hvals, cap_mask, overflow, text, n};

  const int64_t step = (int64_t)gridDim.x * kV3Chunk;
  int64_t base = ch * kV3Chunk + 16 * (int64_t)tid;
  uint32_t Wq[kV3Depth][6];
#pragma unroll
  for (int sidx = 0; sidx < kV3Depth; ++sidx) v2_load(Wq[sidx], text, n, base + sidx * step);
  // chunks whose groups all have their full neighbourhood inside the text are loaded without per-thread checks
  const int64_t interior_hi = (n - 20) / kV3Chunk;         // chunk c is interior iff 1 <= c < interior_hi
  // (Tried: two 16-byte groups per thread and step, to give the scheduler two independent streams -- 970 GB/s against
  //  1061 for this form on the config-4 stream; 16 warps racing on 8 warps' counters, as an upper bound for "more
  //  warps": 1335.  Carrying a wrapped counter where it happens -- a branch per two-pair step, the carry bin taken from
  //  the counter's own address, no positions collected: 790 GB/s with a divergent detour, 768 with a warp vote and
  //  predicated atomics; a branch between the steps stops ptxas from overlapping one step's loads with the previous
  //  step's stores.  Software-pipelining the table lookups of group g + 1 in front of the updates of group g (ptxas
  //  cannot hoist a table load over a counter store, same array): 929.  Cheaper wrap bookkeeping (the form below against
  //  one PRMT per pair and a position loop): +1 %.  The next step's loads issued before this step's stores, repaired by
  //  a software store-to-load forward (4 compares + 4 selects + 2 masks per step): 977 -> 904 on English letters.  The
  //  rate follows the instruction count (ALU pipe 62 % busy at 2 warps per scheduler), see DESIGN.md.)

  for (;;) {
#pragma unroll
    for (int sidx = 0; sidx < kV3Depth; ++sidx) {
      if (ch >= n_chunks) goto done;           // (uniform over the CTA)
      uint32_t W[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) W[q] = Wq[sidx][q];
      {                                        // the step four ahead
        const int64_t chn = ch + (int64_t)kV3Depth * gridDim.x;
        const uint8_t *pn = text + base + kV3Depth * step;
        if (chn >= 1 && chn < interior_hi) {
          const uint4 mid = __ldg(reinterpret_cast<const uint4 *>(pn));
          Wq[sidx][1] = mid.x; Wq[sidx][2] = mid.y; Wq[sidx][3] = mid.z; Wq[sidx][4] = mid.w;
          Wq[sidx][0] = __ldg(reinterpret_cast<const uint32_t *>(pn - 4));
          Wq[sidx][5] = __ldg(reinterpret_cast<const uint32_t *>(pn + 16));
        } else {
          v2_load(Wq[sidx], text, n, base + kV3Depth * step);   // (reads as 0x80.. past the end)
        }
      }

      const bool exists = base < n;
      const bool plain = exists && (((W[0] | W[1] | W[2] | W[3] | W[4] | W[5]) & 0x80808080u) == 0);
      // hot view of positions 0..16: a group that is not plain counts line breaks (junk bins) and is redone below
      uint32_t H[5];
#pragma unroll
      for (int q = 0; q < 5; ++q) H[q] = plain ? W[q + 1] : 0x0a0a0a0au;
      uint32_t t[17];
#pragma unroll
      for (int j = 0; j < 17; ++j) t[j] = tab[(H[j >> 2] >> (8 * (j & 3))) & 0xffu];
      uint32_t adj = 0, any = t[16];
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        adj |= t[j] & t[j + 1];
        adj |= t[j + 1] & t[j + 2];
        any |= t[j] | t[j + 1];
      }
      // Which counters passed 255?  Per two-pair step one IMAD (c1 << 16 | c0: bits 8 and 24 are the wrap flags; the
      // FMA pipe, not the ALU the rest of the loop saturates) and a four-instruction OR tree at the end of the group.
      uint32_t e[8];
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const uint32_t a0 = (t[i] >> 16) + ((t[i + 1] & 0x383u) | lane4);
        const uint32_t a1 = (t[i + 1] >> 16) + ((t[i + 2] & 0x383u) | lane4);
        uint32_t c0 = col[a0];
        uint32_t c1 = col[a1];
        // same bin: the first store writes the old value back, the second carries both increments (so that a wrap
        // is seen once, in c1)
        const bool same = a0 == a1;
        c0 += same ? 0u : 1u;
        c1 += same ? 2u : 1u;
        col[a0] = (uint8_t)c0;
        col[a1] = (uint8_t)c1;
        e[i >> 1] = c1 * 65536u + c0;
      }
      // A one-byte counter passed 255 (some lane of the warp in 87 % of the groups): +1 in the CTA's carry table, whose
      // bins are worth 256 at the flush.  The step index is static here, so the bin row * 28 + column comes straight
      // from the step's table entries (the counter's address without the lane bits: (a >> 7) << 2 | a & 3).
      if (((e[0] | e[1] | e[2]) | (e[3] | e[4] | e[5]) | (e[6] | e[7])) & 0x01000100u) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (e[k] & 0x01000100u) {
            const uint32_t r0 = (t[2 * k] >> 16) + (t[2 * k + 1] & 0x383u), r1 = (t[2 * k + 1] >> 16) + (t[2 * k + 2] & 0x383u);
            if (e[k] & 0x00000100u) atomicAdd(&carry[((r0 >> 7) << 2) | (r0 & 3u)], 1u);
            if (e[k] & 0x01000000u) atomicAdd(&carry[((r1 >> 7) << 2) | (r1 & 3u)], 1u);
          }
        }
      }
      // white space next to white space or a line break (positions -1 .. 17), a byte outside the private alphabet
      const uint32_t cm1 = W[0] >> 24, c17 = (W[5] >> 8) & 0xffu;
      const bool edge = ((t[0] & kV3FlagW) && cm1 <= 0x20u) || ((t[16] & kV3FlagW) && c17 <= 0x20u);
      const bool classify = plain && (((adj & kV3FlagW) | (any & kV3FlagJ)) != 0 || edge);
      const bool general = exists && !plain;
      if (classify | general) v3_cold(cold, base, classify, general, W[0], W[1], W[2], W[3], W[4], W[5]);
      ch += gridDim.x;
      base += step;
    }
  }
done:
  __syncthreads();

  // flush: private counters (sum over the warp's 32 columns per bin), then the CTA histogram
  const uint32_t *cols = reinterpret_cast<const uint32_t *>(priv + warp * kV3PrivPerWarp);
  for (int g = 0; g < (int)kV3Junk * (kV3Syms / 4); ++g) {      // word rows of ranks 0 .. 26 (the junk row is dropped)
    const uint32_t w = cols[g * 32 + lane];
    const uint32_t s0 = __reduce_add_sync(HYP_FULL_MASK, w & 0xffu), s1 = __reduce_add_sync(HYP_FULL_MASK, (w >> 8) & 0xffu);
    const uint32_t s2 = __reduce_add_sync(HYP_FULL_MASK, (w >> 16) & 0xffu), s3 = __reduce_add_sync(HYP_FULL_MASK, w >> 24);
    if (lane < 4) {
      const uint32_t s = lane == 0 ? s0 : lane == 1 ? s1 : lane == 2 ? s2 : s3;
      const uint32_t ra = (uint32_t)g / (kV3Syms / 4), rb = 4u * ((uint32_t)g % (kV3Syms / 4)) + lane;
      if (s && rb != kV3Junk) atomicAdd(ascii_counts + (uint32_t)inv[ra] * 128u + inv[rb], (unsigned long long)s);
    }
  }
  for (int q = tid; q < 64 * 64; q += kV3Threads) {
    const uint32_t v = hist64[q];
    if (v) atomicAdd(ascii_counts + (uint32_t)inv[q >> 6] * 128u + inv[q & 63], (unsigned long long)v);
  }
  for (int q = tid; q < kV3Bins; q += kV3Threads) {
    const uint32_t v = (q / kV3Syms == (int)kV3Junk || q % kV3Syms == (int)kV3Junk) ? 0u : carry[q];   // (junk bins wrap too)
    if (v) atomicAdd(ascii_counts + (uint32_t)inv[q / kV3Syms] * 128u + inv[q % kV3Syms], 256ULL * v);
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Which kernel for this stream?  v3 counts every pair of the 27 most frequent symbols unconditionally and repairs the
// rest in a cold path; that pays while the cold path is rare.  One CTA looks at 8 windows of 2 KiB spread over the
// stream: the share of bytes outside the 27 most frequent symbols of the sample (line breaks aside) and of adjacent
// white-space bytes.  Above 0.4 % of the positions v2 -- which decides every pair before it counts, at a fixed cost --
// is the faster one (wide alphabets, indented text).  The verdict is a device word both kernels read: the one that is
// not chosen returns at once, so there is no host round trip.  The same kernel ranks the ASCII bytes of its sample by
// frequency and publishes the ranks behind the verdict: both counting kernels build their lookup tables from them.
// Results are identical either way (and for any rank table: the ranks decide how fast, never what is counted).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kSelWindows = 8, kSelWindow = 2048, kSelThreads = 1024;
constexpr int kSelSlotInts = 64;               // one verdict slot: [0] = 2 | 3, bytes 4 .. 131 = frequency rank of ASCII byte c
__global__ void __launch_bounds__(kSelThreads)
pair_count_select_kernel(const uint8_t *__restrict__ text, int64_t n, int *__restrict__ select, int force) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned int member[4];           // bit c: ASCII byte c is one of the 27 most frequent (or a line break)
  __shared__ unsigned int bad;
  const int tid = threadIdx.x;
  if (tid < 256) hist[tid] = 0;
  if (tid < 4) member[tid] = 0;
  if (tid == 0) bad = 0;
  __syncthreads();
  const int64_t span = n < (int64_t)kSelWindows * kSelWindow ? n : (int64_t)kSelWindows * kSelWindow;
  const int64_t stride = n > span ? ((n - kSelWindow) / (kSelWindows - 1)) & ~(int64_t)15 : kSelWindow;
  auto pos_of = [&](int64_t q) -> int64_t {    // q-th sampled byte -> position in the text
    return n > span ? (q / kSelWindow) * stride + (q % kSelWindow) : q;
  };
  // Frequency ranks of the ASCII bytes over the whole sample, ONE ranking for every counting CTA.  (Round 2 ranked per
  // CTA from its first 4 KiB: a letter of 0.1 % of the text is absent from a third of such chunks, loses its private
  // bin there, and every 16-byte group that holds it takes the cold path -- half the throughput on English letter
  // frequencies.)  Ties, absent bytes above all: lower-case letters and the space first, then by byte value.
  for (int64_t q = tid; q < span; q += kSelThreads) atomicAdd(&hist[__ldg(text + pos_of(q))], 1u);
  __syncthreads();
  if (tid < 128) {
    bool in = tid == 0x0a || tid == 0x0d;
    unsigned int rank = 0xffu;
    if (!in) {
      auto key = [&](int c) -> unsigned int {
        const unsigned int likely = ((c >= 'a' && c <= 'z') || c == ' ') ? 1u : 0u;
        return hist[c] * 2u + likely;
      };
      const unsigned int mine = key(tid);
      rank = 0;
      for (int w = 0; w < 128; ++w) {
        const unsigned int c = key(w);
        rank += (w != 0x0a && w != 0x0d && (c > mine || (c == mine && w < tid))) ? 1u : 0u;
      }
      in = rank < 27u;                         // (published as it is, 0 .. 125: each kernel keeps the ranks it has bins for)
    }
    if (in) atomicOr(&member[tid >> 5], 1u << (tid & 31));
    reinterpret_cast<uint8_t *>(select + 1)[tid] = (uint8_t)rank;
  }
  __syncthreads();
  unsigned int mine_bad = 0;
  for (int64_t q = tid; q < span; q += kSelThreads) {
    const int64_t p = pos_of(q);
    const unsigned int c = __ldg(text + p);
    const unsigned int nx = p + 1 < n ? __ldg(text + p + 1) : 0x41u;
    const bool outside = c >= 128u || !((member[c >> 5] >> (c & 31)) & 1u);
    const bool ws_pair = c <= 0x20u && nx <= 0x20u;
    mine_bad += (outside || ws_pair) ? 1u : 0u;
  }
  mine_bad = __reduce_add_sync(HYP_FULL_MASK, mine_bad);
  if ((tid & 31) == 0 && mine_bad) atomicAdd(&bad, mine_bad);
  __syncthreads();
  if (tid == 0) *select = force ? force : ((unsigned long long)bad * 250ull > (unsigned long long)span) ? 2 : 3;   // > 0.4 %
}

// 16 KiB per GPU for the selection verdicts (and rank tables) of calls in flight (allocated on first use)
static int *select_slot() {
  constexpr int kDevs = 64, kSlots = 64;
  static int *base[kDevs] = {nullptr};
  static unsigned int next[kDevs] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kDevs) return nullptr;
  if (!base[dev] && cudaMalloc((void **)&base[dev], kSlots * kSelSlotInts * sizeof(int)) != cudaSuccess) {
    cudaGetLastError();
    base[dev] = nullptr;
    return nullptr;
  }
  return base[dev] + (size_t)(next[dev]++ % kSlots) * kSelSlotInts;
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
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // HYP_PAIR_COUNT=v1 | v2 | v3 forces one kernel (same results; A/B runs, and the tests run every one of them);
  // default: v3 or v2 by the stream's alphabet, decided on the device (pair_count_select_kernel)
  const char *ev = getenv("HYP_PAIR_COUNT");
  const int variant = (ev && ev[0] == 'v' && ev[1] >= '1' && ev[1] <= '3') ? ev[1] - '0' : 0;
  if (variant != 1) {
    int *select = select_slot();
    if (!select) {
      set_error("hyp_pair_count: no device memory for the kernel selection slot");
      return HYP_ERR_CUDA;
    }
    pair_count_select_kernel<<<1, kSelThreads, 0, st>>>(text, n_bytes, select, variant);
    {
      int rc = check_launch("hyp_pair_count(select)");
      if (rc) return rc;
    }
    // (per device, so set before every launch: a process-wide "already set" flag would skip it on a second GPU)
    if (cudaFuncSetAttribute(pair_count_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kV3Smem) != cudaSuccess ||
        cudaFuncSetAttribute(pair_count_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kV2Smem) != cudaSuccess) {
      set_error("hyp_pair_count: cudaFuncSetAttribute: %s", cudaGetErrorString(cudaGetLastError()));
      return HYP_ERR_CUDA;
    }
    if (variant == 0 || variant == 3) {
      const int64_t chunks3 = (n_bytes + kV3Chunk - 1) / kV3Chunk;
      const int grid3 = (int)(chunks3 < (int64_t)sms ? chunks3 : (int64_t)sms);
      pair_count_v3_kernel<<<grid3, kV3Threads, kV3Smem, st>>>(text, n_bytes, ascii_counts, hash_keys, hash_vals,
                                                               (uint32_t)(hash_capacity - 1), overflow, select);
      int rc = check_launch("hyp_pair_count(v3)");
      if (rc) return rc;
    }
    if (variant == 0 || variant == 2) {
      const int64_t chunks2 = (n_bytes + kV2Chunk - 1) / kV2Chunk;
      const int grid2 = (int)(chunks2 < (int64_t)sms ? chunks2 : (int64_t)sms);
      pair_count_v2_kernel<<<grid2, kV2Threads, kV2Smem, st>>>(text, n_bytes, ascii_counts, hash_keys, hash_vals,
                                                               (uint32_t)(hash_capacity - 1), overflow, select);
      int rc = check_launch("hyp_pair_count(v2)");
      if (rc) return rc;
    }
    return HYP_OK;
  }
  const size_t smem = 128 * 128 * sizeof(uint32_t) + kChunk + 2 * kHalo;
  if (cudaFuncSetAttribute(pair_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_error("hyp_pair_count: cudaFuncSetAttribute: %s", cudaGetErrorString(cudaGetLastError()));
    return HYP_ERR_CUDA;
  }
  int64_t chunks = (n_bytes + kChunk - 1) / kChunk;
  int grid = (int)(chunks < (int64_t)sms * 2 ? chunks : (int64_t)sms * 2);
  pair_count_kernel<<<grid, kPcThreads, smem, st>>>(text, n_bytes, ascii_counts, hash_keys, hash_vals,
                                                    (uint32_t)(hash_capacity - 1), overflow);
  return check_launch("hyp_pair_count");
}
