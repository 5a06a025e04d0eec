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
// 148 SMs x 1.97 GHz = the 312 GB/s it measures, with 80 thread instructions per byte of per-position branching on top.
// v2 removes both:
//  * classification is SIMD-in-register over the thread's 24-byte neighbourhood: two 24-bit masks (space, line break)
//    from carry-free byte arithmetic, and the count / resolve-slowly decisions of all 16 pairs as mask algebra;
//  * each CTA ranks the ASCII bytes of its first window by frequency.  Pairs of the 32 most frequent symbols are
//    counted in LANE-PRIVATE one-byte counters (word-interleaved: lane t owns bank t, so a warp's 32 updates are 32
//    plain LDS.U8 / STS.U8 with no conflicts and no atomics; a counter that wraps carries 256 into the global table),
//    pairs within the 64 most frequent go to a CTA histogram with ATOMS, anything rarer straight to the global table.
//    Counts stay exact for any input; the alphabet only decides how fast.
//  * the stream is staged through two shared-memory windows with cp.async, the next chunk in flight while this one
//    is counted (one CTA of 6 warps per SM: the private counters take 6 x 32 KB).
constexpr int kV2Threads = 192;
constexpr int kV2Warps = kV2Threads / 32;
constexpr int kV2Iters = 2;
constexpr int kV2Chunk = kV2Threads * 16 * kV2Iters;      // 6144 bytes of text per CTA iteration
constexpr int kV2Win = kHalo + kV2Chunk + kHalo;          // 6272
constexpr int kPrivPerWarp = 32 * 32 * 32;                // 1024 bins x 32 lanes x 1 byte
constexpr size_t kV2Smem = (size_t)kV2Warps * kPrivPerWarp + 64 * 64 * 4 + 2 * kV2Win + 256 + 256 + 64;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

// four bytes (all < 0x80) -> bit 7 of each byte set where the byte is an ASCII space (9..13, 28..32) / a line break
__device__ __forceinline__ void classify4(uint32_t w, uint32_t &sp, uint32_t &nl) {
  const uint32_t ge9 = w + 0x77777777u, ge14 = w + 0x72727272u, ge28 = w + 0x64646464u, ge33 = w + 0x5f5f5f5fu;
  sp = ((ge9 & ~ge14) | (ge28 & ~ge33)) & 0x80808080u;
  const uint32_t ne10 = (w ^ 0x0a0a0a0au) + 0x7f7f7f7fu, ne13 = (w ^ 0x0d0d0d0du) + 0x7f7f7f7fu;
  nl = ~(ne10 & ne13) & 0x80808080u;
}
// bits 7, 15, 23, 31 -> bits 0..3
__device__ __forceinline__ uint32_t gather4(uint32_t m) { return (((m >> 7) * 0x00204081u) >> 21) & 0xfu; }

struct V2Ctx {
  uint8_t *priv_lane;            // this lane's byte column of its warp's private counters
  uint32_t *hist64;
  const uint8_t *sym;            // byte -> frequency rank 0..63, 0xff = none
  unsigned long long *ascii_counts;
  // cold path: pairs resolved by the general scans, pairs outside the private alphabet
  __device__ __forceinline__ void add(uint32_t a, uint32_t b) const {
    const uint32_t ra = sym[a], rb = sym[b], r = ra | rb;
    if (ra < 31u && rb < 31u) {               // rank 31 is the junk row / column of the private table (see the hot loop)
      const uint32_t idx = ra * 32u + rb;
      uint8_t *p = priv_lane + (idx >> 2) * 128u + (idx & 3u);
      const uint32_t v = (uint32_t)*p + 1u;
      *p = (uint8_t)v;
      if ((v & 0xffu) == 0u) atomicAdd(ascii_counts + a * 128u + b, 256ULL);
    } else if (r < 64u) {
      atomicAdd(&hist64[ra * 64u + rb], 1u);
    } else {
      atomicAdd(ascii_counts + a * 128u + b, 1ULL);
    }
  }
};

__global__ void __launch_bounds__(kV2Threads, 1)
pair_count_v2_kernel(const uint8_t *__restrict__ text, int64_t n, unsigned long long *__restrict__ ascii_counts,
                     unsigned long long *hkeys, unsigned long long *hvals, uint32_t cap_mask, int *overflow) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t *priv = smem_raw;                                                    // [warps][256 word rows][32 lanes][4]
  uint32_t *hist64 = reinterpret_cast<uint32_t *>(smem_raw + kV2Warps * kPrivPerWarp);   // [64*64]
  uint8_t *win_base = reinterpret_cast<uint8_t *>(hist64 + 64 * 64);            // two windows
  uint8_t *sym = win_base + 2 * kV2Win;                                         // byte -> frequency rank, 0xff = none
  uint8_t *psym = sym + 256;                                                    // byte -> min(rank, 31)
  uint8_t *inv = psym + 256;                                                    // rank -> byte
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n_chunks = (n + kV2Chunk - 1) / kV2Chunk;

  auto window = [&](int64_t ch, int64_t &c0, int64_t &c1, int64_t &w0, int64_t &w1) {
    c0 = ch * kV2Chunk;
    c1 = (c0 + kV2Chunk < n) ? c0 + kV2Chunk : n;
    w0 = (c0 - kHalo > 0) ? c0 - kHalo : 0;
    w1 = (c1 + kHalo < n) ? c1 + kHalo : n;
  };
  auto load_window = [&](uint8_t *win, int64_t ch) {
    int64_t c0, c1, w0, w1;
    window(ch, c0, c1, w0, w1);
    const int nvec = (int)((w1 - w0) >> 4);
    for (int v = tid; v < nvec; v += kV2Threads) cp_async16(win + 16 * v, text + w0 + 16 * (int64_t)v);
    for (int b = (nvec << 4) + tid; b < (int)(w1 - w0); b += kV2Threads) win[b] = __ldg(text + w0 + b);
  };

  int64_t ch = blockIdx.x;
  if (ch >= n_chunks) return;
  load_window(win_base, ch);
  cp_async_wait_all();
  for (int k = tid; k < 256; k += kV2Threads) hist64[k] = 0;
  __syncthreads();
  {  // frequency ranks of the ASCII bytes of this CTA's first window
    int64_t c0, c1, w0, w1;
    window(ch, c0, c1, w0, w1);
    for (int b = tid; b < (int)(w1 - w0); b += kV2Threads) atomicAdd(&hist64[win_base[b]], 1u);
    __syncthreads();
    uint32_t rank = 0xffu;
    if (tid < 128 && tid != 0x0a && tid != 0x0d) {
      const uint32_t mine = hist64[tid];
      rank = 0;
      for (int w = 0; w < 128; ++w) {
        const uint32_t c = hist64[w];
        rank += (w != 0x0a && w != 0x0d && (c > mine || (c == mine && w < tid))) ? 1u : 0u;
      }
      if (rank >= 64u) rank = 0xffu;
    }
    __syncthreads();
    if (tid < 128) {
      sym[tid] = (uint8_t)rank;
      sym[128 + tid] = 0xff;
      psym[tid] = (uint8_t)(rank < 31u ? rank : 31u);
      psym[128 + tid] = 31;
      if (rank != 0xffu) inv[rank] = (uint8_t)tid;
    }
    uint4 *z = reinterpret_cast<uint4 *>(smem_raw);
    const int nz = (kV2Warps * kPrivPerWarp + 64 * 64 * 4) / 16;
    for (int k = tid; k < nz; k += kV2Threads) z[k] = make_uint4(0, 0, 0, 0);
    __syncthreads();
  }
  const V2Ctx ctx{priv + warp * kPrivPerWarp + lane * 4, hist64, sym, ascii_counts};

  int buf = 0;
  for (; ch < n_chunks; ch += gridDim.x) {
    const uint8_t *win = win_base + buf * kV2Win;
    if (ch + gridDim.x < n_chunks) load_window(win_base + (buf ^ 1) * kV2Win, ch + gridDim.x);
    int64_t c0, c1, w0, w1;
    window(ch, c0, c1, w0, w1);
    Text T{text, n, win, w0, w1};
#pragma unroll 1
    for (int it = 0; it < kV2Iters; ++it) {
      const int64_t base = c0 + 16 * (int64_t)(tid + it * kV2Threads);
      if (base >= c1) break;
      uint32_t W[6];
      {
        const uint4 mid = *reinterpret_cast<const uint4 *>(win + (base - w0));
        W[1] = mid.x; W[2] = mid.y; W[3] = mid.z; W[4] = mid.w;
        W[0] = (base - 4 >= w0) ? *reinterpret_cast<const uint32_t *>(win + (base - 4 - w0)) : 0x80808080u;
        W[5] = (base + 20 <= w1) ? *reinterpret_cast<const uint32_t *>(win + (base + 16 - w0)) : 0x80808080u;
      }
      // groups at either end of the text, or with a non-ASCII byte in their 24-byte neighbourhood, go slow
      if (((W[0] | W[1] | W[2] | W[3] | W[4] | W[5]) & 0x80808080u) == 0) {
        uint32_t S = 0, N = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          uint32_t sp, nl;
          classify4(W[k], sp, nl);
          S |= gather4(sp) << (4 * k);
          N |= gather4(nl) << (4 * k);
        }
        // bit q = position + 4; pair i = (byte i, byte i+1) sits at bit i + 4
        const uint32_t valid = ~N & ~(N >> 1) & 0x000ffff0u;
        const uint32_t left_unres = S & (S << 1) & ~(N << 1), right_unres = (S >> 1) & (S >> 2) & ~(N >> 2);
        const uint32_t fast = valid & (~S | ~(S << 1)) & (~(S >> 1) | ~(S >> 2));
        uint32_t slow = (valid & (left_unres | right_unres)) >> 4;
        auto byte_at = [&](int j) -> uint32_t { return (W[(j + 4) >> 2] >> (8 * ((j + 4) & 3))) & 0xffu; };
        // Hot loop, branch-free: pr = private rank 0..30, or 31 for every byte outside the private alphabet, so that
        // pr_a * 32 + pr_b lands in a junk bin (row or column 31) by itself; pairs that are not counted are sent to
        // junk bin 1023.  One dependent LDS.U8 -> +1 -> STS.U8 per pair; everything else is independent of that chain.
        uint32_t pr[17];
#pragma unroll
        for (int j = 0; j < 17; ++j) pr[j] = psym[byte_at(j)];
        uint32_t nb = 0;                           // bit j: byte j is outside the private alphabet
#pragma unroll
        for (int j = 0; j < 17; ++j) nb |= ((pr[j] + 1u) >> 5) << j;
        const uint32_t counted = fast >> 4;
        uint32_t other = counted & (nb | (nb >> 1));
        uint32_t wrapped = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t idx = ((counted >> i) & 1u) ? pr[i] * 32u + pr[i + 1] : 1023u;
          uint8_t *p = ctx.priv_lane + (idx >> 2) * 128u + (idx & 3u);
          const uint32_t v = (uint32_t)*p + 1u;
          *p = (uint8_t)v;
          wrapped |= (v >> 8) << i;
        }
        wrapped &= counted & ~other;
        while (wrapped) {                          // a private counter passed 255: carry into the global table
          const int i = __ffs(wrapped) - 1;
          wrapped &= wrapped - 1;
          atomicAdd(ascii_counts + (uint32_t)win[base + i - w0] * 128u + win[base + i + 1 - w0], 256ULL);
        }
        while (other) {                            // counted pairs outside the private alphabet
          const int i = __ffs(other) - 1;
          other &= other - 1;
          ctx.add(win[base + i - w0], win[base + i + 1 - w0]);
        }
        while (slow) {
          // a run of spaces next to the pair: resolve with the general scans
          const int i = __ffs(slow) - 1;
          slow &= slow - 1;
          const int64_t pp = base + i;
          const uint32_t a = win[pp - w0], bch = win[pp + 1 - w0];
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
    cp_async_wait_all();
    __syncthreads();
    buf ^= 1;
  }

  // flush: private counters (sum over the warp's 32 lanes per bin), then the CTA histogram
  const uint32_t *rows = reinterpret_cast<const uint32_t *>(priv + warp * kPrivPerWarp);
  for (int g = 0; g < 256; ++g) {
    const uint32_t w = rows[g * 32 + lane];
    const uint32_t s0 = __reduce_add_sync(HYP_FULL_MASK, w & 0xffu), s1 = __reduce_add_sync(HYP_FULL_MASK, (w >> 8) & 0xffu);
    const uint32_t s2 = __reduce_add_sync(HYP_FULL_MASK, (w >> 16) & 0xffu), s3 = __reduce_add_sync(HYP_FULL_MASK, w >> 24);
    if (lane < 4) {
      const uint32_t s = lane == 0 ? s0 : lane == 1 ? s1 : lane == 2 ? s2 : s3;
      const uint32_t idx = g * 4 + lane;
      if (s && (idx >> 5) != 31u && (idx & 31u) != 31u)
        atomicAdd(ascii_counts + (uint32_t)inv[idx >> 5] * 128u + inv[idx & 31u], (unsigned long long)s);
    }
  }
  for (int k = tid; k < 64 * 64; k += kV2Threads) {
    const uint32_t v = hist64[k];
    if (v) atomicAdd(ascii_counts + (uint32_t)inv[k >> 6] * 128u + inv[k & 63], (unsigned long long)v);
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
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  static const int variant = [] {
    const char *e = getenv("HYP_PAIR_COUNT");
    return (e && e[0] == 'v' && e[1] == '2') ? 2 : 1;
  }();
  if (variant == 2) {
    static bool attr2 = false;
    if (!attr2) {
      cudaFuncSetAttribute(pair_count_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kV2Smem);
      attr2 = true;
    }
    const int64_t chunks2 = (n_bytes + kV2Chunk - 1) / kV2Chunk;
    const int grid2 = (int)(chunks2 < (int64_t)sms ? chunks2 : (int64_t)sms);
    pair_count_v2_kernel<<<grid2, kV2Threads, kV2Smem, st>>>(text, n_bytes, ascii_counts, hash_keys, hash_vals,
                                                             (uint32_t)(hash_capacity - 1), overflow);
    return check_launch("hyp_pair_count");
  }
  const size_t smem = 128 * 128 * sizeof(uint32_t) + kChunk + 2 * kHalo;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(pair_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  int64_t chunks = (n_bytes + kChunk - 1) / kChunk;
  int grid = (int)(chunks < (int64_t)sms * 2 ? chunks : (int64_t)sms * 2);
  pair_count_kernel<<<grid, kPcThreads, smem, st>>>(text, n_bytes, ascii_counts, hash_keys, hash_vals,
                                                    (uint32_t)(hash_capacity - 1), overflow);
  return check_launch("hyp_pair_count");
}
