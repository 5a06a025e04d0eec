// gram_tc.cu -- K2, tensor-core path: all-pairs Lorentz distance as a Gram GEMM on tcgen05 (sm_100a).
//
// Replaces the O(n^2 d) broadcast of embedding/lorentz_model.py:141-178 + the neighbour search the
// reference delegates to FAISS (fast_hyperbolic_merge.py:301-304) for per-row top-k at large V.
//
//   u'[i][j] = sgn * (x0_i x0_j - sum_k xs_i[k] xs_j[k])
//                                                 ONE tcgen05.mma kind::tf32 chain per tile: A/B tiles fed by
//                                                 TMA (SWIZZLE_128B, K-major), fp32 accumulators in TMEM; the
//                                                 time-like term rides inside the MMA as hi/lo TF32 parts of
//                                                 x0 in the K padding (tc_pack_kernel), so the accumulator IS u'
//   key      = max(u', 1)                         clamped pre-acosh value; acosh is monotone, so the
//                                                 order of keys is the order of distances
//
// TF32 keeps 10 mantissa bits of each operand, so |u' - u| <= eps_i = 2^-9 |xs_i| max_j|xs_j| (+ fp32
// accumulation slack).  The kernel therefore never DECIDES anything; it bounds and filters:
//   pass 1  per row, the minimum key of every `step`-th 128-column tile  -> tilemin[tile][row]
//           (k distinct tiles' minima are k distinct elements, so the k-th smallest tile minimum is an
//            upper bound tau_i of the row's true k-th smallest key)
//   select  tau_i = k-th smallest of the row's tile minima              (kth_select_kernel)
//   pass 2  collect every column with key <= tau_i + 2 eps_i            -> cand[row][segment][<= cap]
//           a certified superset of the exact top-k
//   finish  exact fp32 re-score of the candidates in ATen order (allpairs.cu arithmetic), sort by
//           (d, j), emit k -- into the caller's arrays and/or straight into the gather buffers of every
//           rank of a hyp_ctx (peer memory over NVLink: the producing kernel performs the all-gather)
//   redo    rows whose candidate buffers overflowed (massive ties) are listed on the device and
//           recomputed by the exact CUDA-core kernel, so the result is ALWAYS the exact one.
//
// Work item = (pair of 128-row blocks, column segment): a shard of a few dozen row blocks (V/8 rows per
// GPU) still fills all 148 SMs, and the last wave of a big shard is short.
// Warp roles per CTA (352 threads): warp 0 TMA producer, warps 1 and 10 one MMA issuer thread per row
// block of the pair (warp 1 also owns the TMEM allocation), warps 2..9 epilogue, four per row block
// (thread <-> accumulator row = TMEM lane).  The A tiles of the pair stay in shared memory for the whole
// item, the column tiles stream through a 2-stage B ring shared by both row blocks; accumulators are
// double-buffered per row block in TMEM (4 x 128 columns) so the epilogue of tile t overlaps the MMA of t+1.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include <stdio.h>

#include "common.cuh"

namespace hyp {

constexpr int TC_M = 128;          // rows per CTA block (UMMA M)
constexpr int TC_N = 128;          // columns per tile (UMMA N)
constexpr int TC_KSLAB = 32;       // fp32 elements per 128-byte swizzle row
constexpr int TC_SLAB_BYTES = TC_M * 128;   // 16 KiB: 128 rows x 128 B
constexpr int TC_MAX_SLABS = 4;    // K padded up to 128
constexpr int TC_MAX_STAGES = 4;
constexpr int TC_ACC = 4;           // accumulator buffers in TMEM (4 x 128 columns = all 512)
constexpr int TC_EPI_WARPS = 8;    // 2 per TMEM lane quadrant: each handles one 64-column half of the tile
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS + 32;   // TMA, MMA issuer 0, 8 epilogue warps, MMA issuer 1
constexpr int TC_CAND_ROW = 512;   // candidate slots per row in the workspace, split evenly over the column segments
constexpr int TC_CAP = 384;        // most candidates of one row the finish kernel re-scores (more: exact redo)
constexpr int TC_MAX_SEG = 16;     // column segments per row block pair
constexpr bool kDefaultTs = false; // engine when HYP_TC_ENGINE is unset: false = both operands in shared memory

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, tf32 inputs, fp32 accumulate, M=128, N=128, K=8 per instruction
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// One lane of a CONVERGED warp.  The MMA warps run their loops with all 32 lanes (every operand of tcgen05.mma is then
// provably warp-uniform and lives in uniform registers) and only the instruction itself is issued by the elected lane.
// With the whole loop inside `if (lane == 0)` the compiler wrapped EVERY tcgen05.mma in an ELECT / 4 x R2UR.BROADCAST /
// BRA.U.ANY waterfall: ~84 cycles per MMA from one thread, above the tensor pipe's 64.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives row (lane base + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
}

// Tail slab (K not a multiple of 32): rows of 32 B (8 floats, SWIZZLE_32B = 6, 8-row group 256 B) or 64 B
// (16 floats, SWIZZLE_64B = 4, 8-row group 512 B); same canonical K-major form.
__device__ __forceinline__ uint64_t smem_desc_tail(uint32_t saddr, int row_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * row_bytes) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(row_bytes == 32 ? 6 : 4) << 61;
  return d;
}

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
}
// 16-byte load from shared memory by 32-bit shared address.  (A float4 load through a generic pointer that has passed
// through a lambda compiles to LD.E.128, a generic load: slower, and it shows up as long-scoreboard stalls.)
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(saddr));
  return r;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// tcgen05.ld writes its destination registers asynchronously; the compiler only knows that the asm statement "wrote"
// them, so nothing stops it from scheduling their use before the wait.  This empty asm makes the values "change" at a
// point behind the wait (volatile asms keep their order), which pins every use after it.
__device__ __forceinline__ void tmem_pin(float (&v)[32]) {
  asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]));
  asm volatile("" : "+f"(v[16]), "+f"(v[17]), "+f"(v[18]), "+f"(v[19]), "+f"(v[20]), "+f"(v[21]), "+f"(v[22]), "+f"(v[23]), "+f"(v[24]), "+f"(v[25]), "+f"(v[26]), "+f"(v[27]), "+f"(v[28]), "+f"(v[29]), "+f"(v[30]), "+f"(v[31]));
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 in
// [0,14), LBO >> 4 in [16,30) (unused for swizzled K-major, 1), SBO >> 4 in [32,46) = 1024 B between
// 8-row groups, version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// cute::UMMA::InstrDescriptor for kind::tf32: D=F32 (1<<4), A=B=TF32 (2<<7, 2<<10), K-major both,
// N>>3 in [17,23), M>>4 in [24,29).
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) |
                                ((uint32_t)(TC_M >> 4) << 24);

// ---------------------------------------------------------------------------------------------
// operand preparation.  The signed Minkowski product rides INSIDE the MMA, time-like term included:
//   XA[r] = [ xs_0 .. xs_{d-1},  hi,  lo,  hi,  lo, 0.. ]      x0 = hi + lo, hi = x0 cut to TF32's 10 mantissa bits
//   XB[r] = [-s xs_0 .. -s xs_{d-1}, s hi, s hi, s lo, s lo, 0.. ]                       s = sgn (+1 lorentz, -1 reference)
// so that  XA[i] . XB[j] = s (x0_i x0_j - xs_i . xs_j) = u'  directly.  hi x hi is exact in the tensor core (10-bit
// operands, fp32 accumulate), the three cross terms carry |lo| <= 2^-10 x0 rounded to TF32: ~2^-19 x0_i x0_j in
// total, inside the fp32 slack of the bound.  (One TF32 slot for x0 would cost 2^-11 x0_i x0_j ~ 5e-4, far more than
// the spread of u' near the k-th neighbour: SURVEY 7.)  d = 100 uses the four pad columns of K = 104: no extra
// MMA, and the epilogue needs neither the column time components nor an FMA per element.
// Also nrm[r] = |xs_r| and its maximum, for the error bound.
// ---------------------------------------------------------------------------------------------
// One warp per row, one float4 of each packed table per lane (Kp / 4 <= 32 of them), two rows in flight.  XB and
// nrm cover the whole table (every rank scores its rows against ALL columns), XA only the shard's rows.
__global__ void __launch_bounds__(256)
tc_pack_kernel(const float *__restrict__ E, int64_t ldE, int64_t n, int64_t row0, int64_t nrows, int D, int Kp,
               float sgn, float *__restrict__ XA, float *__restrict__ XB, float *__restrict__ nrm,
               unsigned int *__restrict__ max_nrm_bits) {
  const int lane = threadIdx.x & 31;
  const int d = D - 1;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool active = 4 * lane < Kp;
  float wmax = 0.f;
  for (int64_t r0 = 2 * warp; r0 < n; r0 += 2 * nwarps) {
    float a[2][4], x0[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t r = r0 + h < n ? r0 + h : r0;
      const float *row = E + r * ldE;
      x0[h] = __ldg(row);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int k = 4 * lane + t;
        a[h][t] = k < d ? __ldg(row + 1 + k) : 0.f;
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t r = r0 + h;
      if (r >= n) break;
      const float hi = __uint_as_float(__float_as_uint(x0[h]) & 0xffffe000u);
      const float lo = x0[h] - hi;                 // exact; NaN / inf propagate into the accumulator as they should
      float va[4], vb[4], ss = 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int k = 4 * lane + t;
        if (k < d) {
          va[t] = a[h][t];
          vb[t] = -sgn * a[h][t];
          ss = fmaf(a[h][t], a[h][t], ss);
        } else if (k < d + 4) {
          const int q = k - d;
          va[t] = (q & 1) ? lo : hi;               // hi, lo, hi, lo
          vb[t] = sgn * (q < 2 ? hi : lo);         // hi, hi, lo, lo
        } else {
          va[t] = 0.f;
          vb[t] = 0.f;
        }
      }
      if (active) {
        *reinterpret_cast<float4 *>(XB + r * Kp + 4 * lane) = make_float4(vb[0], vb[1], vb[2], vb[3]);
        if (r >= row0 && r < row0 + nrows)
          *reinterpret_cast<float4 *>(XA + (r - row0) * Kp + 4 * lane) = make_float4(va[0], va[1], va[2], va[3]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(HYP_FULL_MASK, ss, o);
      const float nr = sqrtf(ss);
      if (lane == 0) nrm[r] = nr;
      if (nr == nr) wmax = fmaxf(wmax, nr);
    }
  }
  if (lane == 0 && wmax > 0.f) atomicMax(max_nrm_bits, __float_as_uint(wmax));
}

struct TcParams {
  int64_t n;          // table rows (columns of the Gram matrix)
  int64_t row0;       // shard start (XA and its tensor map start at this row)
  int64_t nrows;      // shard rows
  int n_slabs;        // full 128-byte-row slabs (32 floats each)
  int tail_row_bytes; // 0, 32 or 64: one more slab with narrow rows for K mod 32 in {8, 16}
  int stage_bytes;    // bytes of one operand tile in shared memory (1024-aligned)
  int n_stages;       // depth of the B ring
  int n_ksteps;       // ceil((d + 4) / 8) MMAs per tile
  int debug;          // HYP_TC_DEBUG bit 0: epilogue skips TMEM loads + math, bit 1: no MMAs issued, bit 2: no TMA,
                      //              bit 3: epilogue loads TMEM but skips the math
  int rb_per_cta;     // 2: two row blocks share each B tile; 1: one row block per item
  int64_t n_ct;       // column tiles this pass visits: ct = t * ct_step, t in [0, n_ct)
  int ct_step;        // 1 = every tile; pass 1 samples (see hyp_gram_topk)
  int n_seg;          // column segments: work item w = (row pair w / n_seg, segment w % n_seg)
  int64_t seg_tiles;  // visited tiles per segment: t in [s * seg_tiles, min((s + 1) * seg_tiles, n_ct))
  // pass 1
  float *tilemin;     // [n_ct][ld_tm]
  int64_t ld_tm;
  // pass 2
  const float *thr;   // [nrows] tau + 2 eps
  int32_t *cand;      // [nrows][TC_CAND_ROW]: segment s of a row owns slots [s * seg_cap, (s + 1) * seg_cap)
  int32_t *cand_cnt;  // [nrows][TC_MAX_SEG]: hits per segment (> seg_cap == overflow)
  int seg_cap;
  // TS engine (A operands in tensor memory): the packed A table of the shard and its row pitch in floats
  const float *xa;
  int kp;
};

// NS / TB: the operand layout (full slabs, tail row bytes) as compile-time constants for the common dimensions
// (d = 100: 3 slabs + 32-byte tail; d = 50: 2 slabs), -1 = read it from the parameters.  With constants the k-step
// offsets of the descriptors are immediates of the uniform datapath and the MMA loop has no predicates.
template <int PASS, int NS = -1, int TB = -1>
__global__ void __launch_bounds__(TC_THREADS, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapA_tail,
               const __grid_constant__ CUtensorMap tmapB, const __grid_constant__ CUtensorMap tmapB_tail,
               const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // layout: A tile of row block 2rp | A tile of row block 2rp+1 | B stages | barriers | tmem ptr
  // Two row blocks share every B tile: the kernel is bound by the L2 -> SM traffic of the B stream (ablation:
  // TMA + barriers alone took 44 % of a pass, 6 TB/s), and this halves it.
  uint8_t *base = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *sA = base;
  uint8_t *sB = sA + 2 * (size_t)p.stage_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sB + (size_t)p.n_stages * p.stage_bytes);
  uint64_t *a_full = bars + 0, *a_empty = bars + 1;
  uint64_t *acc_full = bars + 2, *acc_empty = bars + 2 + TC_ACC;                        // [TC_ACC] each
  uint64_t *b_full = bars + 2 + 2 * TC_ACC, *b_empty = bars + 2 + 2 * TC_ACC + TC_MAX_STAGES;   // [n_stages] each
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 2 + 2 * TC_ACC + 2 * TC_MAX_STAGES);
  const uint32_t tile_tx = p.n_slabs * TC_SLAB_BYTES + TC_M * p.tail_row_bytes;
  const int tail_elem0 = p.n_slabs * TC_KSLAB;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row_blocks = (p.nrows + TC_M - 1) / TC_M;
  const int64_t row_pairs = (row_blocks + p.rb_per_cta - 1) / p.rb_per_cta;   // pairs (or single blocks)
  const int64_t items = row_pairs * p.n_seg;       // work items: (row pair, column segment), segment fastest

  if (threadIdx.x == 0) {
    mbar_init(a_full, 1);
    mbar_init(a_empty, 2);                          // one commit per MMA issuer
    for (int s = 0; s < TC_ACC; ++s) {
      mbar_init(acc_full + s, 1);
      mbar_init(acc_empty + s, TC_EPI_WARPS / 2);   // one arrival per warp of the draining group
    }
    for (int s = 0; s < p.n_stages; ++s) {
      mbar_init(b_full + s, 1);
      mbar_init(b_empty + s, 2);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, TC_ACC * TC_N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t bstage = 0, bphase = 0, aphase = 0;
      for (int64_t w = blockIdx.x; w < items; w += gridDim.x) {
        const int64_t rp = w / p.n_seg;
        const int64_t t0 = (w - rp * p.n_seg) * p.seg_tiles;
        const int64_t t1 = t0 + p.seg_tiles < p.n_ct ? t0 + p.seg_tiles : p.n_ct;
        mbar_wait(a_empty, aphase ^ 1);
        mbar_expect_tx(a_full, p.rb_per_cta * tile_tx);
        for (int h = 0; h < p.rb_per_cta; ++h) {
          // (the A map covers the shard's rows only; rows past it are zero-filled by TMA, their results never written)
          uint8_t *dstA = sA + (size_t)h * p.stage_bytes;
          const int arow = (int)((p.rb_per_cta * rp + h) * TC_M);
          for (int s = 0; s < p.n_slabs; ++s) tma_load_2d(&tmapA, a_full, dstA + s * TC_SLAB_BYTES, s * TC_KSLAB, arow);
          if (p.tail_row_bytes) tma_load_2d(&tmapA_tail, a_full, dstA + p.n_slabs * TC_SLAB_BYTES, tail_elem0, arow);
        }
        aphase ^= 1;
        for (int64_t t = t0; t < t1; ++t) {
          const int64_t ct = t * p.ct_step;
          mbar_wait(b_empty + bstage, bphase ^ 1);
          if (p.debug & 4) { mbar_arrive(b_full + bstage); if (++bstage == (uint32_t)p.n_stages) { bstage = 0; bphase ^= 1; } continue; }
          mbar_expect_tx(b_full + bstage, tile_tx);
          uint8_t *dst = sB + (size_t)bstage * p.stage_bytes;
          for (int s = 0; s < p.n_slabs; ++s)
            tma_load_2d(&tmapB, b_full + bstage, dst + s * TC_SLAB_BYTES, s * TC_KSLAB, (int)(ct * TC_N));
          if (p.tail_row_bytes)
            tma_load_2d(&tmapB_tail, b_full + bstage, dst + p.n_slabs * TC_SLAB_BYTES, tail_elem0, (int)(ct * TC_N));
          if (++bstage == (uint32_t)p.n_stages) { bstage = 0; bphase ^= 1; }
        }
      }
    }
  } else if (warp == 1 || warp == 2 + TC_EPI_WARPS) {
    // ================= MMA issuers: warp 1 for row block 2rp, the last warp for row block 2rp + 1 =================
    // A single thread issuing every tcgen05.mma ran at ~150 cycles per M128 x N128 x K8 instruction with TMA and
    // the epilogue switched off (the tensor pipe needs 64): the issue path, not shared memory, was the limit.  The
    // two row blocks of a pair have independent accumulators, so each gets its own issuing thread.
    // One thread issues every tcgen05.mma of the CTA, so its own instruction stream is the limit: the
    // descriptors of all k-steps are prepared once (low word = (address >> 4) + per-k-step offset, high word
    // constant per slab kind) and the tile loop only adds a base and issues.
    {
      constexpr int kMaxK = TC_MAX_SLABS * 4;
      const int n_slabs = NS >= 0 ? NS : p.n_slabs;
      const int tail_b = TB >= 0 ? TB : p.tail_row_bytes;
      const uint32_t hi128 = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
      const uint32_t hitail = (uint32_t)((8 * tail_b) >> 4) | (1u << 14) | ((tail_b == 32 ? 6u : 4u) << 29);
      auto koff = [&](int ks) -> uint32_t {
        const int slab = ks >> 2, within = ks & 3;
        return (uint32_t)((((slab >= n_slabs) ? n_slabs : slab) * TC_SLAB_BYTES + within * 32) >> 4);
      };
      auto khi = [&](int ks) -> uint32_t { return (ks >> 2) >= n_slabs ? hitail : hi128; };
      const int h = warp == 1 ? 0 : 1;
      const uint32_t a_lo = ((smem_u32(sA + (size_t)h * p.stage_bytes) >> 4) & 0x3fff) | (1u << 16);
      const bool idle = h >= p.rb_per_cta;              // single-block mode: the second issuer only keeps the barriers moving
      const int nk_rt = (p.debug & 2) ? 0 : p.n_ksteps;
      const int nk = (NS >= 0 && TB >= 0) ? (p.debug & 2 ? 0 : (NS * TC_KSLAB + TB / 4) / 8) : nk_rt;
      uint32_t bstage = 0, bphase = 0, aphase = 0;
      uint32_t tt = 0;                      // tiles issued by this CTA: accumulators 2*(tt&1)+h, use number tt>>1
      for (int64_t w = blockIdx.x; w < items; w += gridDim.x) {
        const int64_t t0 = (w % p.n_seg) * p.seg_tiles;
        const int64_t t1 = t0 + p.seg_tiles < p.n_ct ? t0 + p.seg_tiles : p.n_ct;
        mbar_wait(a_full, aphase);
        aphase ^= 1;
        for (int64_t t = t0; t < t1; ++t, ++tt) {
          mbar_wait(b_full + bstage, bphase);
          const uint32_t b_lo = ((smem_u32(sB + (size_t)bstage * p.stage_bytes) >> 4) & 0x3fff) | (1u << 16);
          if (!idle) {
            const uint32_t abuf = 2 * (tt & 1) + h;
            mbar_wait(acc_empty + abuf, ((tt >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_addr = tmem_base + abuf * TC_N;
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < kMaxK; ++ks) {
                if (ks < nk) {
                  const uint64_t ad = ((uint64_t)khi(ks) << 32) | (uint64_t)(a_lo + koff(ks));
                  const uint64_t bd = ((uint64_t)khi(ks) << 32) | (uint64_t)(b_lo + koff(ks));
                  umma_tf32(d_addr, ad, bd, kIdescTf32, ks > 0 ? 1u : 0u);
                }
              }
              umma_commit(acc_full + abuf);    // this row block's accumulator is ready for its epilogue group
            }
          }
          if (elect_one()) umma_commit(b_empty + bstage);     // B stage reusable once both issuers' MMAs have read it
          __syncwarp();
          if (++bstage == (uint32_t)p.n_stages) { bstage = 0; bphase ^= 1; }
        }
        if (elect_one()) umma_commit(a_empty);                // this issuer is done with its A tile
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue: warps 2..9 = two groups of four warps =================
    // Group h drains the accumulators of row block 2rp + h: both groups follow the same column tiles, so two
    // accumulators per group are in flight and one drain overlaps the MMAs of the next tile.  Within a group,
    // thread <-> accumulator row (TMEM lane), all 128 columns, and a row belongs to ONE thread for the whole pass
    // (its candidate list is a single stream).  The accumulator IS u' (see tc_pack_kernel): pass 1 keeps min(u') per
    // tile (fminf drops NaN operands, and min_j max(u',1) == max(min_j u', 1), so the clamp is applied once per tile);
    // pass 2 tests u' <= thr_i (thr_i >= 1, so the clamped region always passes) into a 32-bit hit mask and only
    // walks set bits.  Tiles that contain out-of-range columns (zero rows: u' = 0) or the row block's own diagonal
    // take the checked path.
    const int quad = warp & 3;                       // TMEM lanes [32*quad, +32) are readable by this warp
    const int grp = (warp - 2) >> 2;                 // which row block of the pair
    const int lane_base = 32 * quad;
    const int r_in_block = lane_base + lane;         // row of the tile
    const float inf = __int_as_float(0x7f800000);
    uint32_t T = 0;                                  // tiles drained by this group so far (accumulator ring position)
    for (int64_t w = blockIdx.x; w < items && grp < p.rb_per_cta; w += gridDim.x) {
      const int64_t rp = w / p.n_seg;
      const int seg = (int)(w - rp * p.n_seg);
      const int64_t t0 = seg * p.seg_tiles;
      const int64_t t1 = t0 + p.seg_tiles < p.n_ct ? t0 + p.seg_tiles : p.n_ct;
      const int64_t blk0 = p.row0 + (rp * p.rb_per_cta + grp) * TC_M;
      const int64_t gi = blk0 + r_in_block;
      const bool row_ok = gi < p.row0 + p.nrows;
      const float thr = (PASS == 2 && row_ok) ? __ldg(p.thr + (gi - p.row0)) : -inf;
      int cnt = 0;
      int32_t *const my_cand = p.cand + (gi - p.row0) * TC_CAND_ROW + (int64_t)seg * p.seg_cap;
      for (int64_t tix = t0; tix < t1; ++tix, ++T) {
      const int64_t ct = tix * p.ct_step;
      const uint32_t abuf = 2 * (uint32_t)(T & 1) + grp, accphase = (uint32_t)((T >> 1) & 1);
      const int64_t j0 = ct * TC_N;
      mbar_wait(acc_full + abuf, accphase);
      tc_fence_after();
      const bool checked = (j0 + TC_N > p.n) || (j0 < blk0 + TC_M && j0 + TC_N > blk0);
      float tmin = inf;
      const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + abuf * TC_N;
      // One 32-column chunk of the tile.
      auto consume = [&](int chunk, const float (&v)[32]) {
        if (PASS == 1) {
          if (!checked) {
            float m4[4] = {inf, inf, inf, inf};      // four independent min chains
#pragma unroll
            for (int c = 0; c < 32; ++c) m4[c & 3] = fminf(m4[c & 3], v[c]);
            tmin = fminf(tmin, fminf(fminf(m4[0], m4[1]), fminf(m4[2], m4[3])));
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const int64_t gj = j0 + chunk * 32 + c;
              if (gj < p.n && gj != gi) tmin = fminf(tmin, v[c]);
            }
          }
        } else {
          // hit mask: FSETP + SEL of an immediate bit per column, summed (disjoint bits) in a tree
          uint32_t bit[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) bit[c] = v[c] <= thr ? (1u << c) : 0u;
#pragma unroll
          for (int w = 16; w >= 1; w >>= 1)
#pragma unroll
            for (int c = 0; c < w; ++c) bit[c] |= bit[c + w];
          uint32_t hits = bit[0];
          while (hits) {
            const int c = __ffs(hits) - 1;
            hits &= hits - 1;
            const int64_t gj = j0 + chunk * 32 + c;
            if (!checked || (gj < p.n && gj != gi)) {
              if (cnt < p.seg_cap) my_cand[cnt] = (int32_t)gj;
              ++cnt;
            }
          }
        }
      };
      {
        // Three register buffers rotate over the four chunks so that only the first TMEM round trip of a tile is
        // exposed: chunks 2 and 3 are requested while chunks 0 and 1 are consumed.
        float va[32], vb[32], vc[32];
        const bool skip_math = (p.debug & 9) != 0;
        if (!(p.debug & 1)) {
          tmem_ld32_nowait(taddr, va);
          tmem_ld32_nowait(taddr + 32, vb);
          tmem_ld_wait();
          tmem_pin(va);
          tmem_pin(vb);
          tmem_ld32_nowait(taddr + 64, vc);
          if (!skip_math) consume(0, va);
          tmem_ld32_nowait(taddr + 96, va);
          if (!skip_math) consume(1, vb);
          tmem_ld_wait();
          tmem_pin(vc);
          tmem_pin(va);
          if (!skip_math) consume(2, vc);
          if (!skip_math) consume(3, va);
        }
      }
      // release the accumulator buffer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + abuf);
      if (PASS == 1 && row_ok) p.tilemin[tix * p.ld_tm + (gi - p.row0)] = fmaxf(tmin, 1.0f);
      }
      // hits of this (row, segment); every segment of every shard row is written by exactly one item
      if (PASS == 2 && row_ok) p.cand_cnt[(gi - p.row0) * TC_MAX_SEG + seg] = cnt;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, TC_ACC * TC_N);
  }
}

// ---------------------------------------------------------------------------------------------
// TS engine (HYP_TC_ENGINE=ts; an alternative, not the default): the same two passes with the A operands in TENSOR MEMORY.
//
// The kernel above is bound by shared-memory bandwidth: every M128 x N128 x K8 MMA reads 4 KB of A and 4 KB of B from
// shared memory, two chains per B tile, plus the TMA writes: with the epilogue switched off a full pass takes 2.18 ms =
// 2076 cycles per B tile against 1664 at the tensor pipe's own pace and ~2040 at 128 B/clk.
// tcgen05.mma can take A from tensor memory instead (the form
// `[d_tmem], [a_tmem], b_desc`: row m of A on lane m, element k in column a_col + k; checked on B200 by
// tools/ts_probe.cu): the A rows of a work item are written there ONCE per item, and an MMA then reads only B from
// shared memory -- half the bytes.  Tensor memory has 512 columns: two row blocks' A operands take 2 x 128 (K <= 128),
// which leaves 256 for the accumulators, so the column tile shrinks to N = 64 to keep them double-buffered for both
// row blocks (2 blocks x 2 buffers x 64 columns).  Shared memory now holds nothing but the B ring (64 x 416 B tiles).
// Roles as above, except that the epilogue warps of a row block also LOAD its A operand at the start of an item:
// thread <-> row reads its packed row from global memory (26 16-byte loads, once per ~260 tiles) and stores it with
// tcgen05.st; a_full / a_empty barriers order that against the block's MMA chain.
// Measured (B200, V = 100 k, d = 100): results bit-identical to the SS engine on every test; a full pass takes 2.77 ms
// against 2.57 ms -- 43 cycles per M128 x N64 x K8 MMA where the pipe's pace is 32: with half-width tiles every MMA
// still fetches its whole 4 KB A operand, now from tensor memory, and that read rate is the new bound.  N = 128 would
// halve it but needs all 512 columns for the accumulators alone.  Kept as a tested alternative and as the evidence for
// that bound; the way past BOTH limits is cta_group::2 (one B tile feeding two SMs).
// ---------------------------------------------------------------------------------------------
constexpr int TS_N = 64;
constexpr int TS_SLAB_BYTES = TS_N * 128;          // 8 KiB: 64 rows x 128 B
constexpr int TS_A_BASE = 256;                     // accumulator (parity, block h) at column (2 parity + h) * 64
constexpr int TS_A_COLS = 128;                     // columns reserved per row block's A operand
constexpr int TS_MAX_STAGES = 6;
constexpr uint32_t kIdescTf32N64 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TS_N >> 3) << 17) |
                                   ((uint32_t)(TC_M >> 4) << 24);

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
      "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
      "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

template <int PASS, int NS = -1, int TB = -1>
__global__ void __launch_bounds__(TC_THREADS, 1)
gram_ts_kernel(const __grid_constant__ CUtensorMap tmapB, const __grid_constant__ CUtensorMap tmapB_tail,
               const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *base = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *sB = base;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sB + (size_t)p.n_stages * p.stage_bytes);
  uint64_t *a_full = bars + 0, *a_empty = bars + 2;                                      // [2] each: per row block
  uint64_t *acc_full = bars + 4, *acc_empty = bars + 4 + TC_ACC;                         // [TC_ACC] each
  uint64_t *b_full = bars + 4 + 2 * TC_ACC, *b_empty = bars + 4 + 2 * TC_ACC + TS_MAX_STAGES;   // [n_stages] each
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 4 + 2 * TC_ACC + 2 * TS_MAX_STAGES);
  const uint32_t tile_tx = p.n_slabs * TS_SLAB_BYTES + TS_N * p.tail_row_bytes;
  const int tail_elem0 = p.n_slabs * TC_KSLAB;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row_blocks = (p.nrows + TC_M - 1) / TC_M;
  const int64_t row_pairs = (row_blocks + p.rb_per_cta - 1) / p.rb_per_cta;
  const int64_t items = row_pairs * p.n_seg;

  if (threadIdx.x == 0) {
    for (int h = 0; h < 2; ++h) {
      mbar_init(a_full + h, TC_EPI_WARPS / 2);      // one arrival per warp of the block's epilogue group
      mbar_init(a_empty + h, 1);                    // the block's MMA issuer
    }
    for (int s = 0; s < TC_ACC; ++s) {
      mbar_init(acc_full + s, 1);
      mbar_init(acc_empty + s, TC_EPI_WARPS / 2);
    }
    for (int s = 0; s < p.n_stages; ++s) {
      mbar_init(b_full + s, 1);
      mbar_init(b_empty + s, 2);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ================= TMA producer: the column tiles of the item's segment, nothing else =================
    if (lane == 0) {
      uint32_t bstage = 0, bphase = 0;
      for (int64_t w = blockIdx.x; w < items; w += gridDim.x) {
        const int64_t t0 = (w % p.n_seg) * p.seg_tiles;
        const int64_t t1 = t0 + p.seg_tiles < p.n_ct ? t0 + p.seg_tiles : p.n_ct;
        for (int64_t t = t0; t < t1; ++t) {
          const int64_t ct = t * p.ct_step;
          mbar_wait(b_empty + bstage, bphase ^ 1);
          mbar_expect_tx(b_full + bstage, tile_tx);
          uint8_t *dst = sB + (size_t)bstage * p.stage_bytes;
          for (int s = 0; s < p.n_slabs; ++s)
            tma_load_2d(&tmapB, b_full + bstage, dst + s * TS_SLAB_BYTES, s * TC_KSLAB, (int)(ct * TS_N));
          if (p.tail_row_bytes)
            tma_load_2d(&tmapB_tail, b_full + bstage, dst + p.n_slabs * TS_SLAB_BYTES, tail_elem0, (int)(ct * TS_N));
          if (++bstage == (uint32_t)p.n_stages) { bstage = 0; bphase ^= 1; }
        }
      }
    }
  } else if (warp == 1 || warp == 2 + TC_EPI_WARPS) {
    // ================= MMA issuers: warp 1 for row block 2rp, the last warp for row block 2rp + 1 =================
    {
      constexpr int kMaxK = TC_MAX_SLABS * 4;
      const int n_slabs = NS >= 0 ? NS : p.n_slabs;
      const int tail_b = TB >= 0 ? TB : p.tail_row_bytes;
      const uint32_t hi128 = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
      const uint32_t hitail = (uint32_t)((8 * tail_b) >> 4) | (1u << 14) | ((tail_b == 32 ? 6u : 4u) << 29);
      auto koff = [&](int ks) -> uint32_t {
        const int slab = ks >> 2, within = ks & 3;
        return (uint32_t)((((slab >= n_slabs) ? n_slabs : slab) * TS_SLAB_BYTES + within * 32) >> 4);
      };
      auto khi = [&](int ks) -> uint32_t { return (ks >> 2) >= n_slabs ? hitail : hi128; };
      const int h = warp == 1 ? 0 : 1;
      const bool idle = h >= p.rb_per_cta;              // single-block mode: the second issuer only keeps the B ring moving
      const uint32_t a_addr = tmem_base + TS_A_BASE + h * TS_A_COLS;
      const int nk = (NS >= 0 && TB >= 0) ? (NS * TC_KSLAB + TB / 4) / 8 : p.n_ksteps;
      uint32_t bstage = 0, bphase = 0, aphase = 0;
      uint32_t tt = 0;                      // tiles issued by this CTA: accumulator 2*(tt&1)+h, use number tt>>1
      for (int64_t w = blockIdx.x; w < items; w += gridDim.x) {
        const int64_t t0 = (w % p.n_seg) * p.seg_tiles;
        const int64_t t1 = t0 + p.seg_tiles < p.n_ct ? t0 + p.seg_tiles : p.n_ct;
        if (!idle) {
          mbar_wait(a_full + h, aphase);     // the block's A operand is in tensor memory
          aphase ^= 1;
          tc_fence_after();
        }
        for (int64_t t = t0; t < t1; ++t, ++tt) {
          mbar_wait(b_full + bstage, bphase);
          const uint32_t b_lo = ((smem_u32(sB + (size_t)bstage * p.stage_bytes) >> 4) & 0x3fff) | (1u << 16);
          if (!idle) {
            const uint32_t abuf = 2 * (tt & 1) + h;
            mbar_wait(acc_empty + abuf, ((tt >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_addr = tmem_base + abuf * TS_N;
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < kMaxK; ++ks) {
                if (ks < nk) {
                  const uint64_t bd = ((uint64_t)khi(ks) << 32) | (uint64_t)(b_lo + koff(ks));
                  umma_tf32_ts(d_addr, a_addr + 8 * ks, bd, kIdescTf32N64, ks > 0 ? 1u : 0u);
                }
              }
              umma_commit(acc_full + abuf);
            }
          }
          if (elect_one()) umma_commit(b_empty + bstage);
          __syncwarp();
          if (++bstage == (uint32_t)p.n_stages) { bstage = 0; bphase ^= 1; }
        }
        if (!idle) {
          if (elect_one()) umma_commit(a_empty + h);  // every MMA that reads this A operand has completed
          __syncwarp();
        }
      }
    }
  } else {
    // ================= epilogue (+ A loader): warps 2..9 = two groups of four warps, one per row block =================
    const int quad = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int lane_base = 32 * quad;
    const int r_in_block = lane_base + lane;
    const float inf = __int_as_float(0x7f800000);
    const uint32_t lane_addr = tmem_base + ((uint32_t)lane_base << 16);
    const int a_chunks = (p.n_ksteps * 8 + 31) / 32;
    uint32_t T = 0, ephase = 0;
    for (int64_t w = blockIdx.x; w < items && grp < p.rb_per_cta; w += gridDim.x) {
      const int64_t rp = w / p.n_seg;
      const int seg = (int)(w - rp * p.n_seg);
      const int64_t t0 = seg * p.seg_tiles;
      const int64_t t1 = t0 + p.seg_tiles < p.n_ct ? t0 + p.seg_tiles : p.n_ct;
      const int64_t rel = (rp * p.rb_per_cta + grp) * TC_M + r_in_block;      // row within the shard
      const int64_t blk0 = p.row0 + (rp * p.rb_per_cta + grp) * TC_M;
      const int64_t gi = p.row0 + rel;
      const bool row_ok = rel < p.nrows;
      // ---- this thread's row of the A operand -> tensor memory (the previous item's MMAs must be done with it)
      mbar_wait(a_empty + grp, ephase ^ 1);
      ephase ^= 1;
      tc_fence_after();
      {
        const float4 *src = reinterpret_cast<const float4 *>(p.xa + (row_ok ? rel : 0) * p.kp);
        for (int c = 0; c < a_chunks; ++c) {
          uint32_t r[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 v = row_ok ? __ldg(src + c * 8 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            r[4 * q] = __float_as_uint(v.x); r[4 * q + 1] = __float_as_uint(v.y);
            r[4 * q + 2] = __float_as_uint(v.z); r[4 * q + 3] = __float_as_uint(v.w);
          }
          tmem_st32(lane_addr + TS_A_BASE + grp * TS_A_COLS + c * 32, r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full + grp);
      }
      const float thr = (PASS == 2 && row_ok) ? __ldg(p.thr + rel) : -inf;
      int cnt = 0;
      int32_t *const my_cand = p.cand + rel * TC_CAND_ROW + (int64_t)seg * p.seg_cap;
      for (int64_t tix = t0; tix < t1; ++tix, ++T) {
        const int64_t ct = tix * p.ct_step;
        const uint32_t abuf = 2 * (uint32_t)(T & 1) + grp, accphase = (uint32_t)((T >> 1) & 1);
        const int64_t j0 = ct * TS_N;
        mbar_wait(acc_full + abuf, accphase);
        tc_fence_after();
        const bool checked = (j0 + TS_N > p.n) || (j0 < blk0 + TC_M && j0 + TS_N > blk0);
        float tmin = inf;
        const uint32_t taddr = lane_addr + abuf * TS_N;
        auto consume = [&](int chunk, const float (&v)[32]) {
          if (PASS == 1) {
            if (!checked) {
              float m4[4] = {inf, inf, inf, inf};
#pragma unroll
              for (int c = 0; c < 32; ++c) m4[c & 3] = fminf(m4[c & 3], v[c]);
              tmin = fminf(tmin, fminf(fminf(m4[0], m4[1]), fminf(m4[2], m4[3])));
            } else {
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                const int64_t gj = j0 + chunk * 32 + c;
                if (gj < p.n && gj != gi) tmin = fminf(tmin, v[c]);
              }
            }
          } else {
            uint32_t bit[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) bit[c] = v[c] <= thr ? (1u << c) : 0u;
#pragma unroll
            for (int wd = 16; wd >= 1; wd >>= 1)
#pragma unroll
              for (int c = 0; c < wd; ++c) bit[c] |= bit[c + wd];
            uint32_t hits = bit[0];
            while (hits) {
              const int c = __ffs(hits) - 1;
              hits &= hits - 1;
              const int64_t gj = j0 + chunk * 32 + c;
              if (!checked || (gj < p.n && gj != gi)) {
                if (cnt < p.seg_cap) my_cand[cnt] = (int32_t)gj;
                ++cnt;
              }
            }
          }
        };
        {
          float va[32], vb[32];
          tmem_ld32_nowait(taddr, va);
          tmem_ld32_nowait(taddr + 32, vb);
          tmem_ld_wait();
          tmem_pin(va);
          tmem_pin(vb);
          consume(0, va);
          consume(1, vb);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty + abuf);
        if (PASS == 1 && row_ok) p.tilemin[tix * p.ld_tm + rel] = fmaxf(tmin, 1.0f);
      }
      if (PASS == 2 && row_ok) p.cand_cnt[rel * TC_MAX_SEG + seg] = cnt;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// tau_i = k-th smallest tile minimum of row i, then
// thr_i = tau_i + 2 eps_i with eps_i = 2^-9 * 1.05 * |xs_i| * max|xs| + 8e-6 * max(1, tau_i).
// One WARP per row: the row's <= 1024 tile minima sit in registers (32 per lane), and the k-th smallest is found by a
// radix descent over the bit patterns (all values >= 1.0, so the bits order like the values), one REDUX per bit,
// starting at the highest bit in which the row's minimum and maximum differ (the minima of one row span ~2 % of
// their value: ~18 bits instead of 31).  A block takes 32 rows: their slice of the [tile][row] panel is read once,
// line by line, and transposed through shared memory.  Replaces a thread-per-row insertion select whose data-dependent
// inserts diverged and whose 100 blocks per 12 500-row shard left most SMs idle (0.13 ms of a 0.87 ms call at 8 GPUs).
constexpr int KS_PER_LANE = 32;            // most values per lane (n_ct <= 1024); the kernel is instantiated for 8 / 16 / 32
constexpr int KS_ROWS = 32;                // rows per block: one coalesced 128-byte line of the panel per tile

// k-th smallest (1-based rank k) of the warp's values v[q] = value number lane + 32 q, as a bit pattern
template <int Q>
__device__ __forceinline__ unsigned int warp_kth_smallest(const unsigned int (&v)[Q], int64_t n_vals, int k, int lane) {
  constexpr int KS_PER_LANE = Q;
  const unsigned int kInf = 0x7f800000u;
  unsigned int lo = kInf, hi = 0;
#pragma unroll
  for (int q = 0; q < KS_PER_LANE; ++q) {
    lo = min(lo, v[q]);
    if (lane + 32 * q < n_vals) hi = max(hi, v[q]);            // (the +inf padding does not widen the span)
  }
  lo = __reduce_min_sync(HYP_FULL_MASK, lo);
  hi = __reduce_max_sync(HYP_FULL_MASK, hi);
  // bits above `top` are common to every value of the row
  const int top = lo == hi ? -1 : 31 - __clz(lo ^ hi);
  unsigned int prefix = top >= 31 ? 0u : (lo >> (top + 1)) << (top + 1);
  int want = k;                                       // rank still to find among the values matching `prefix`
  for (int bit = top; bit >= 0; --bit) {
    unsigned int c0 = 0;
#pragma unroll
    for (int q = 0; q < KS_PER_LANE; ++q)
      c0 += (((v[q] ^ prefix) >> bit) == 0u) ? 1u : 0u;          // matches the prefix above `bit` and has bit == 0
    c0 = __reduce_add_sync(HYP_FULL_MASK, c0);
    if ((unsigned int)want > c0) {
      want -= (int)c0;
      prefix |= 1u << bit;
    }
  }
  // (fewer than k finite minima: the descent ends on +inf, thr = +inf, every column is a candidate and the row
  //  overflows into the exact redo)
  return n_vals < k ? kInf : prefix;
}

template <int Q>
__global__ void __launch_bounds__(256)
kth_select_warp_kernel(const float *__restrict__ tilemin, int64_t ld_tm, int64_t col_tiles, int64_t nrows,
                       int64_t row0, int k, const float *__restrict__ nrm,
                       const unsigned int *__restrict__ max_nrm_bits, float *__restrict__ thr) {
  extern __shared__ float panel[];                    // [col_tiles][KS_ROWS + 1]: the block's rows, transposed on the way in
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned int kInf = 0x7f800000u;
  for (int64_t r0 = (int64_t)blockIdx.x * KS_ROWS; r0 < nrows; r0 += (int64_t)gridDim.x * KS_ROWS) {
    __syncthreads();
    // tile t of the panel = 32 consecutive rows = one 128-byte line: every byte of the panel is read once, coalesced
    for (int64_t t = warp; t < col_tiles; t += 8)
      panel[t * (KS_ROWS + 1) + lane] = (r0 + lane < nrows) ? __ldcs(tilemin + t * ld_tm + r0 + lane) : __uint_as_float(kInf);
    __syncthreads();
    for (int rr = warp; rr < KS_ROWS; rr += 8) {
      const int64_t r = r0 + rr;
      if (r >= nrows) break;
      unsigned int v[Q];
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const int64_t t = lane + 32 * q;
        v[q] = t < col_tiles ? __float_as_uint(panel[t * (KS_ROWS + 1) + rr]) : kInf;   // bank (t + rr) mod 32: no conflicts
      }
      const float tau = __uint_as_float(warp_kth_smallest<Q>(v, col_tiles, k, lane));
      if (lane == 0) {
        // spatial TF32 rounding (2^-9 |xs_i| |xs_j|, 5 % margin) + fp32 accumulation and the split time-like term
        const float eps = 0.001953125f * 1.05f * nrm[row0 + r] * __uint_as_float(*max_nrm_bits) + 8e-6f * fmaxf(1.f, tau);
        thr[r] = tau + 2.f * eps;
      }
    }
  }
}

// Tables with more than 32 * 32 visited column tiles per row (n > 131 072 with every tile visited): thread per row,
// running k smallest with replace-max.
__global__ void kth_select_kernel(const float *__restrict__ tilemin, int64_t ld_tm, int64_t col_tiles, int64_t nrows,
                                  int64_t row0, int k, const float *__restrict__ nrm,
                                  const unsigned int *__restrict__ max_nrm_bits, float *__restrict__ thr) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  float best[32];
  const float inf = __int_as_float(0x7f800000);
  for (int q = 0; q < 32; ++q) best[q] = inf;
  float worst = inf;
  int wpos = 0;
  // 8 independent loads in flight per thread (the tile minima stream from HBM once), then the rare inserts
  for (int64_t t0 = 0; t0 < col_tiles; t0 += 8) {
    float vals[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) vals[u] = (t0 + u < col_tiles) ? __ldcs(tilemin + (t0 + u) * ld_tm + r) : inf;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float v = vals[u];
      if (v < worst) {
        best[wpos] = v;
        worst = -1.f;
        for (int q = 0; q < k; ++q)
          if (best[q] >= worst) { worst = best[q]; wpos = q; }
      }
    }
  }
  const float tau = worst;     // +inf when fewer than k tiles hold a finite minimum
  // spatial TF32 rounding (2^-9 |xs_i| |xs_j|, 5 % margin) + fp32 accumulation and the split time-like term
  const float eps = 0.001953125f * 1.05f * nrm[row0 + r] * __uint_as_float(*max_nrm_bits) + 8e-6f * fmaxf(1.f, tau);
  thr[r] = tau + 2.f * eps;
}

// exact fp32 re-score of each row's candidates (ATen order), sort by (d, j), write the first k into `sink`.
// One warp per row.  A row whose candidate buffers overflowed (massive ties) is not written: flags[r] = 1 and the
// row is appended to flist for the exact redo that follows on the same stream.
// NS > 0: the spatial dimension d as a compile-time constant (100, 50): the ATen-order loops unroll without predicates
// (the kernel is instruction-bound: 0.85e9 warp instructions per call at V = 100 k with the generic loops).
template <int NS>
__global__ void __launch_bounds__(128)
tc_finish_kernel(const float *__restrict__ E, int64_t ldE, int D, int64_t row0, int64_t nrows, float sqrt_c, float sgn,
                 int k, const int32_t *__restrict__ cand, const int32_t *__restrict__ cand_cnt, int n_seg, int seg_cap,
                 const TopkSink sink, int32_t *__restrict__ flags, int32_t *__restrict__ flist,
                 int32_t *__restrict__ nflag) {
  __shared__ unsigned long long keys[4][TC_CAP];
  __shared__ int32_t cj[4][TC_CAP];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned long long kEmpty = 0xffffffffffffffffULL;
  for (int64_t r = (int64_t)blockIdx.x * 4 + w; r < nrows; r += (int64_t)gridDim.x * 4) {
    // the row's candidates: n_seg lists (one per column segment) -> one contiguous list in shared memory
    const int c_mine = lane < n_seg ? cand_cnt[r * TC_MAX_SEG + lane] : 0;
    int incl = c_mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(HYP_FULL_MASK, incl, o);
      if (lane >= o) incl += up;
    }
    const int mtot_raw = __shfl_sync(HYP_FULL_MASK, incl, 31);
    const bool overflow = __any_sync(HYP_FULL_MASK, c_mine < 0 || c_mine > seg_cap) || mtot_raw > TC_CAP;
    const int mtot = overflow ? 0 : mtot_raw;
    if (!overflow) {
      for (int sg = 0; sg < n_seg; ++sg) {
        const int cs = __shfl_sync(HYP_FULL_MASK, c_mine, sg);
        const int os = __shfl_sync(HYP_FULL_MASK, incl, sg) - cs;
        const int32_t *src = cand + r * TC_CAND_ROW + (int64_t)sg * seg_cap;
        for (int t = lane; t < cs; t += 32) cj[w][os + t] = src[t];
      }
    }
    for (int q = lane; q < mtot; q += 32) keys[w][q] = kEmpty;
    __syncwarp();
    // exact re-score, four candidates per warp pass: a group of 8 lanes IS ATen's 8 summation lanes
    // (lane l owns elements 8k+l, partial k mod 4), so the order is reproduced with 4 registers per lane
    const float *xi = E + (row0 + r) * ldE;
    const int N = NS > 0 ? NS : D - 1, vs = N >> 3, full = vs >> 2;
    constexpr int KV = NS > 0 ? (NS >> 3) : 16;      // lane vectors held in registers (d <= 128)
    const int grp = lane >> 3, l8 = lane & 7;
    if (N >= 8) {
      // the query row's elements of this lane stay in registers for all candidates
#ifndef HYP_FIN_INFLIGHT
#define HYP_FIN_INFLIGHT 2
#endif
      constexpr int kInFlight = HYP_FIN_INFLIGHT;   // (generic loops: 4 in flight was slower -- 128 registers, lower occupancy)
      float xr[KV];
#pragma unroll
      for (int kk = 0; kk < KV; ++kk) xr[kk] = kk < vs ? __ldg(xi + 1 + 8 * kk + l8) : 0.f;
      const float xi0 = __ldg(xi);
      float xt[7];                    // the <= 7 tail elements of the query row (N mod 8), loaded once
#pragma unroll
      for (int e = 0; e < 7; ++e) xt[e] = (8 * vs + e < N) ? __ldg(xi + 1 + 8 * vs + e) : 0.f;
      for (int q0 = 0; q0 < mtot; q0 += 4 * kInFlight) {
        // kInFlight candidates per 8-lane group and pass: their loads are issued together
        float Ll[kInFlight], tailv[kInFlight], xj0[kInFlight];
        int jv[kInFlight];
        bool livev[kInFlight];
#pragma unroll
        for (int h = 0; h < kInFlight; ++h) {
          const int q = q0 + 4 * h + grp;
          livev[h] = q < mtot;
          jv[h] = cj[w][livev[h] ? q : 0];
        }
        float yv[kInFlight][KV], yt[kInFlight][7];
#pragma unroll
        for (int h = 0; h < kInFlight; ++h) {
          const float *xj = E + (int64_t)jv[h] * ldE;
#pragma unroll
          for (int kk = 0; kk < KV; ++kk) yv[h][kk] = kk < vs ? __ldg(xj + 1 + 8 * kk + l8) : 0.f;
#pragma unroll
          for (int e = 0; e < 7; ++e) yt[h][e] = (8 * vs + e < N) ? __ldg(xj + 1 + 8 * vs + e) : 0.f;
          xj0[h] = __ldg(xj);
        }
#pragma unroll
        for (int h = 0; h < kInFlight; ++h) {
          float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int kk = 0; kk < KV; ++kk) {
            if (kk < 4 * full) part[kk & 3] = __fadd_rn(part[kk & 3], __fmul_rn(xr[kk], yv[h][kk]));
            else if (kk < vs) part[0] = __fadd_rn(part[0], __fmul_rn(xr[kk], yv[h][kk]));
          }
          Ll[h] = __fadd_rn(__fadd_rn(__fadd_rn(part[0], part[1]), part[2]), part[3]);
          float acc = 0.f;
#pragma unroll
          for (int e = 0; e < 7; ++e)
            if (8 * vs + e < N) acc = __fadd_rn(acc, __fmul_rn(xt[e], yt[h][e]));
          tailv[h] = acc;
        }
#pragma unroll
        for (int h = 0; h < kInFlight; ++h) {
          float acc = tailv[h];
#pragma unroll
          for (int t = 0; t < 8; ++t) acc = __fadd_rn(acc, __shfl_sync(HYP_FULL_MASK, Ll[h], (lane & 24) + t));
          if (livev[h] && l8 == 0) {
            const float mm = __fsub_rn(__fmul_rn(xi0, xj0[h]), acc);
            const float dv = dist_from_mdot(mm, sgn, sqrt_c);
            if (dv == dv)
              keys[w][q0 + 4 * h + grp] = ((unsigned long long)__float_as_uint(dv) << 32) | (unsigned int)jv[h];
          }
        }
      }
    } else {
      for (int q = 0; q < mtot; ++q) {
        const int j = cj[w][q];
        const float mm = warp_mdot(xi, E + (int64_t)j * ldE, D, lane);
        if (lane == 0) {
          const float dv = dist_from_mdot(mm, sgn, sqrt_c);
          if (dv == dv) keys[w][q] = ((unsigned long long)__float_as_uint(dv) << 32) | (unsigned int)j;
        }
      }
    }
    __syncwarp();
    if (!overflow) {
      // rank sort of the keys, all distinct (distinct j)
      for (int q = lane; q < mtot; q += 32) {
        const unsigned long long mine = keys[w][q];
        if (mine == kEmpty) continue;
        int rank = 0;
        for (int t = 0; t < mtot; ++t) rank += keys[w][t] < mine;
        if (rank < k)
          sink_write(sink, r, k, rank, (int32_t)(mine & 0xffffffffu), __uint_as_float((unsigned int)(mine >> 32)));
      }
      // pad (fewer than k valid candidates can only happen with < k finite distances in the row)
      int valid = 0;
      for (int t = lane; t < mtot; t += 32) valid += keys[w][t] != kEmpty;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) valid += __shfl_xor_sync(HYP_FULL_MASK, valid, o);
      for (int q = valid + lane; q < k; q += 32) sink_write(sink, r, k, q, -1, __int_as_float(0x7f800000));
    }
    if (lane == 0) {
      flags[r] = overflow ? 1 : 0;
      if (overflow) flist[atomicAdd(nflag, 1)] = (int32_t)r;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (EncodeTiledFn)sym;
  }
  return fn;
}

struct TcLayout {
  int Kp, n_slabs, n_ksteps, tail_row_bytes, stage_bytes, n_stages;
  int stage_bytes_ts, n_stages_ts;     // TS engine: 64-column B tiles, nothing else in shared memory
  int64_t ld_tm, col_tiles, col_tiles_ts;
  size_t off_xa, off_xb, off_nrm, off_max, off_tilemin, off_thr, off_cand, off_cnt, off_flist, total;
};

static TcLayout tc_layout(int64_t n, int64_t nrows, int D) {
  TcLayout L;
  const int d = D - 1;
  const int kuse = d + 4;                    // spatial columns + the four time-like slots (tc_pack_kernel)
  L.Kp = ((kuse + TC_KSLAB - 1) / TC_KSLAB) * TC_KSLAB;
  L.n_ksteps = (kuse + 7) / 8;
  const int k8 = L.n_ksteps * 8;
  L.n_slabs = k8 / TC_KSLAB;
  int rem = k8 % TC_KSLAB;                 // 0, 8, 16 or 24 floats
  if (rem == 24) { L.n_slabs += 1; rem = 0; }   // 96-byte rows have no swizzle mode: take a full slab
  L.tail_row_bytes = rem * 4;
  L.stage_bytes = ((L.n_slabs * TC_SLAB_BYTES + TC_M * L.tail_row_bytes + 1023) / 1024) * 1024;
  const int budget = 226 * 1024 - 4096;     // two A tiles + the B ring (+ 1 KB alignment, barriers)
  L.n_stages = (budget - 2 * L.stage_bytes) / L.stage_bytes;
  if (L.n_stages > 2) L.n_stages = 2;      // measured: a deeper B ring does not help (the MMA issue rate and the
                                            // TMEM drain, not TMA latency, bound the tile loop); HYP_TC_STAGES overrides
  if (const char *e = getenv("HYP_TC_STAGES")) {
    const int want = atoi(e);
    if (want >= 1 && want <= TC_MAX_STAGES && (2 + want) * L.stage_bytes <= budget) L.n_stages = want;
  }
  if (L.n_stages < 1) L.n_stages = 1;
  L.col_tiles = (n + TC_N - 1) / TC_N;
  L.col_tiles_ts = (n + TS_N - 1) / TS_N;
  L.stage_bytes_ts = ((L.n_slabs * TS_SLAB_BYTES + TS_N * L.tail_row_bytes + 1023) / 1024) * 1024;
  L.n_stages_ts = budget / L.stage_bytes_ts;
  if (L.n_stages_ts > 4) L.n_stages_ts = 4;
  if (const char *e = getenv("HYP_TC_STAGES")) {
    const int want = atoi(e);
    if (want >= 1 && want <= TS_MAX_STAGES && want * L.stage_bytes_ts <= budget) L.n_stages_ts = want;
  }
  if (L.n_stages_ts < 1) L.n_stages_ts = 1;
  L.ld_tm = ((nrows + 31) / 32) * 32;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 255) & ~(size_t)255; return at; };
  L.off_xa = take((size_t)nrows * L.Kp * 4);   // XA: the shard's rows as A operands
  L.off_xb = take((size_t)n * L.Kp * 4);       // XB: every row as a B operand
  L.off_nrm = take((size_t)n * 4);
  L.off_max = take(256);                       // [0] max norm bits, [1] number of flagged rows
  L.off_tilemin = take((size_t)L.col_tiles_ts * L.ld_tm * 4);      // (the TS engine's 64-column tiles: the larger panel)
  L.off_thr = take((size_t)nrows * 4);
  L.off_cand = take((size_t)nrows * TC_CAND_ROW * 4);
  L.off_cnt = take((size_t)nrows * TC_MAX_SEG * 4);
  L.off_flist = take((size_t)nrows * 4);
  L.total = o;
  return L;
}

// Work decomposition of one pass: `pairs` row block pairs (or single blocks) x n_seg column segments over `sms`
// persistent CTAs.  Cost model in tile-times: waves * (tiles per segment + 2 for the A tiles of the item); the
// smallest segment count that reaches the minimum wins (fewer A loads, longer candidate lists per segment).
static void tc_segments(int64_t pairs, int64_t n_ct, int sms, int &n_seg, int64_t &seg_tiles, int a_cost = 2) {
  if (n_ct < 1) n_ct = 1;
  int lo = 1, hi = TC_MAX_SEG;
  if (const char *e = getenv("HYP_TC_SEG")) {          // force a segment count (tests, tuning)
    const int f = atoi(e);
    if (f >= 1 && f <= TC_MAX_SEG) lo = hi = f;
  }
  double best = 1e300;
  n_seg = 1;
  seg_tiles = n_ct;
  for (int S = lo; S <= hi; ++S) {
    const int64_t st = (n_ct + S - 1) / S;
    if (lo != hi && S > 1 && st < 6 * a_cost) break;   // segments too short to amortise their A operands
    const int64_t seff = (n_ct + st - 1) / st;         // no empty segments
    const int64_t waves = (pairs * seff + sms - 1) / sms;
    const double cost = (double)waves * (double)(st + a_cost);
    if (cost < best * 0.995) {
      best = cost;
      n_seg = (int)seff;
      seg_tiles = st;
    }
  }
}

// The whole pipeline on `st`: pack, bound pass, select, collect pass, finish, exact redo of flagged rows.
int gram_topk_run(const float *E, int64_t ldE, int64_t n, int64_t row0, int64_t nrows, int D, float c, int semantics,
                  int k, const TopkSink &sink, int32_t *row_flags, void *workspace, int64_t workspace_bytes,
                  cudaStream_t st) {
  const TcLayout L = tc_layout(n, nrows, D);
  if (workspace_bytes < (int64_t)L.total) {
    set_error("hyp_gram_topk: workspace %lld < %zu bytes", (long long)workspace_bytes, L.total);
    return HYP_ERR_WORKSPACE;
  }
  if (((uintptr_t)workspace & 255) != 0) {
    set_error("hyp_gram_topk: workspace must be 256-byte aligned");
    return HYP_ERR_ARG;
  }
  EncodeTiledFn encode = get_encode();
  if (!encode) {
    set_error("hyp_gram_topk: cuTensorMapEncodeTiled is unavailable in this driver");
    return HYP_ERR_UNSUPPORTED;
  }
  uint8_t *ws = (uint8_t *)workspace;
  float *XA = (float *)(ws + L.off_xa), *XB = (float *)(ws + L.off_xb), *nrm = (float *)(ws + L.off_nrm);
  const float sgn = semantics == HYP_SEM_REFERENCE ? -1.f : 1.f;
  unsigned int *maxn = (unsigned int *)(ws + L.off_max);
  int32_t *nflag = (int32_t *)(ws + L.off_max) + 1;
  float *tilemin = (float *)(ws + L.off_tilemin), *thr = (float *)(ws + L.off_thr);
  int32_t *cand = (int32_t *)(ws + L.off_cand), *cnt = (int32_t *)(ws + L.off_cnt);
  int32_t *flist = (int32_t *)(ws + L.off_flist);

  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);

  // HYP_TC_TIMING=1: per-stage device times on stderr (synchronises; for tuning only)
  const bool timing = getenv("HYP_TC_TIMING") != nullptr;
  cudaEvent_t tev[7];
  if (timing) {
    for (auto &e : tev) cudaEventCreate(&e);
    cudaEventRecord(tev[0], st);
  }
  cudaMemsetAsync(maxn, 0, 8, st);           // max norm and the flagged-row counter
  {
    int64_t blocks = (n + 15) / 16;            // 8 warps x 2 rows per block and sweep
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    tc_pack_kernel<<<(int)blocks, 256, 0, st>>>(E, ldE, n, row0, nrows, D, L.Kp, sgn, XA, XB, nrm, maxn);
  }
  int rc = check_launch("hyp_gram_topk(pack)");
  if (rc) return rc;

  if (timing) cudaEventRecord(tev[1], st);

  const cuuint64_t gdimA[2] = {(cuuint64_t)L.Kp, (cuuint64_t)nrows};
  const cuuint64_t gdimB[2] = {(cuuint64_t)L.Kp, (cuuint64_t)n};
  const cuuint64_t gstride[1] = {(cuuint64_t)L.Kp * 4};
  const cuuint32_t box[2] = {(cuuint32_t)TC_KSLAB, (cuuint32_t)TC_M};
  const cuuint32_t tbox[2] = {(cuuint32_t)(L.tail_row_bytes ? L.tail_row_bytes / 4 : TC_KSLAB), (cuuint32_t)TC_M};
  const cuuint32_t estride[2] = {1, 1};
  CUtensorMap maps[4];                       // A, A tail, B, B tail
  for (int m = 0; m < 4; ++m) {
    const bool tail = (m & 1) != 0;
    void *base = (m < 2) ? (void *)XA : (void *)XB;
    const CUtensorMapSwizzle sw = !tail || !L.tail_row_bytes ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : L.tail_row_bytes == 32   ? CU_TENSOR_MAP_SWIZZLE_32B
                                                             : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult cr = encode(&maps[m], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, m < 2 ? gdimA : gdimB, gstride,
                         tail ? tbox : box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      set_error("hyp_gram_topk: cuTensorMapEncodeTiled failed (%d) for map %d", (int)cr, m);
      return HYP_ERR_CUDA;
    }
  }

  // HYP_TC_ENGINE=ts: A operands in tensor memory, 64-column tiles (gram_ts_kernel); ss: both operands in shared memory
  const char *eng = getenv("HYP_TC_ENGINE");
  const bool ts = eng ? (eng[0] == 't') : kDefaultTs;
  CUtensorMap tsmaps[2];                     // TS engine: B, B tail with 64-row boxes
  if (ts) {
    const cuuint32_t box64[2] = {(cuuint32_t)TC_KSLAB, (cuuint32_t)TS_N};
    const cuuint32_t tbox64[2] = {(cuuint32_t)(L.tail_row_bytes ? L.tail_row_bytes / 4 : TC_KSLAB), (cuuint32_t)TS_N};
    for (int m = 0; m < 2; ++m) {
      const bool tail = m == 1;
      const CUtensorMapSwizzle sw = !tail || !L.tail_row_bytes ? CU_TENSOR_MAP_SWIZZLE_128B
                                    : L.tail_row_bytes == 32   ? CU_TENSOR_MAP_SWIZZLE_32B
                                                               : CU_TENSOR_MAP_SWIZZLE_64B;
      CUresult cr = encode(&tsmaps[m], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)XB, gdimB, gstride,
                           tail ? tbox64 : box64, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (cr != CUDA_SUCCESS) {
        set_error("hyp_gram_topk: cuTensorMapEncodeTiled failed (%d) for TS map %d", (int)cr, m);
        return HYP_ERR_CUDA;
      }
    }
  }

  TcParams p{};
  p.n = n; p.row0 = row0; p.nrows = nrows; p.n_slabs = L.n_slabs; p.n_ksteps = L.n_ksteps;
  p.tail_row_bytes = L.tail_row_bytes; p.stage_bytes = ts ? L.stage_bytes_ts : L.stage_bytes;
  p.n_stages = ts ? L.n_stages_ts : L.n_stages;
  p.xa = XA; p.kp = L.Kp;
  p.debug = getenv("HYP_TC_DEBUG") ? atoi(getenv("HYP_TC_DEBUG")) : 0;
  p.tilemin = tilemin; p.ld_tm = L.ld_tm; p.thr = thr; p.cand = cand; p.cand_cnt = cnt;
  const size_t smem = ts ? 1024 + (size_t)L.n_stages_ts * L.stage_bytes_ts + (6 + 2 * TC_ACC + 2 * TS_MAX_STAGES) * 8
                         : 1024 + (size_t)(2 + L.n_stages) * L.stage_bytes + (4 + 2 * TC_ACC + 2 * TC_MAX_STAGES) * 8;
  auto k1 = gram_tc_kernel<1>;
  auto k2 = gram_tc_kernel<2>;
  auto t1 = gram_ts_kernel<1>;
  auto t2 = gram_ts_kernel<2>;
  if (L.n_slabs == 3 && L.tail_row_bytes == 32) {          // d in 93 .. 100
    k1 = gram_tc_kernel<1, 3, 32>; k2 = gram_tc_kernel<2, 3, 32>; t1 = gram_ts_kernel<1, 3, 32>; t2 = gram_ts_kernel<2, 3, 32>;
  } else if (L.n_slabs == 2 && L.tail_row_bytes == 0) {    // d in 45 .. 60
    k1 = gram_tc_kernel<1, 2, 0>; k2 = gram_tc_kernel<2, 2, 0>; t1 = gram_ts_kernel<1, 2, 0>; t2 = gram_ts_kernel<2, 2, 0>;
  }
  if (ts) {
    cudaFuncSetAttribute(t1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(t2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  } else {
    cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  const int64_t col_tiles = ts ? L.col_tiles_ts : L.col_tiles;
  const int a_cost = ts ? 8 : 2;             // an item's A operands, in tile times
  const int64_t row_blocks = (nrows + TC_M - 1) / TC_M;
  // Two row blocks per item halve the L2 -> SM traffic of the column stream (the first bound this kernel hit); the
  // column segments provide the parallelism a small shard lacks, so pairing needs only two row blocks.
  p.rb_per_cta = row_blocks >= 2 ? 2 : 1;
  if (const char *e = getenv("HYP_TC_PAIR")) p.rb_per_cta = atoi(e) == 1 ? 1 : 2;
  const int64_t row_pairs = (row_blocks + p.rb_per_cta - 1) / p.rb_per_cta;

  // Pass 1 only has to BOUND each row's k-th best from above, and the k-th smallest minimum over ANY >= k distinct
  // column tiles does that: it visits every `step`-th tile.  The bound sits near rank k*step instead of k, so pass 2
  // collects ~step times as many candidates for the exact re-score; the two costs balance at step 2-3 at V=100k
  // (HYP_TC_SUB overrides).  Small tables keep every tile (at least 4k sampled tiles are required).
  int step = ts ? 4 : 2;                     // (64-column tiles: every 4th keeps the panel at the same number of minima)
  if (const char *e = getenv("HYP_TC_SUB")) step = atoi(e);
  if (step < 1) step = 1;
  while (step > 1 && (col_tiles + step - 1) / step < 4 * (int64_t)k) --step;
  p.ct_step = step;
  p.n_ct = (col_tiles + step - 1) / step;
  tc_segments(row_pairs, p.n_ct, sms, p.n_seg, p.seg_tiles, a_cost);
  p.seg_cap = 0;
  int64_t items = row_pairs * p.n_seg;
  if (ts) t1<<<(int)(items < sms ? items : sms), TC_THREADS, smem, st>>>(tsmaps[0], tsmaps[1], p);
  else k1<<<(int)(items < sms ? items : sms), TC_THREADS, smem, st>>>(maps[0], maps[1], maps[2], maps[3], p);
  rc = check_launch("hyp_gram_topk(pass 1)");
  if (rc) return rc;
  if (timing) cudaEventRecord(tev[2], st);
  if (p.n_ct <= 32 * KS_PER_LANE) {
    const size_t psm = (size_t)p.n_ct * (KS_ROWS + 1) * sizeof(float);
    int64_t sb = (nrows + KS_ROWS - 1) / KS_ROWS;
    if (sb > (int64_t)sms * 8) sb = (int64_t)sms * 8;
    auto launch_select = [&](auto kern) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm);
      kern<<<(int)sb, 256, psm, st>>>(tilemin, L.ld_tm, p.n_ct, nrows, row0, k, nrm, maxn, thr);
    };
    if (p.n_ct <= 32 * 8) launch_select(kth_select_warp_kernel<8>);
    else if (p.n_ct <= 32 * 16) launch_select(kth_select_warp_kernel<16>);
    else launch_select(kth_select_warp_kernel<32>);
  } else {
    kth_select_kernel<<<(int)((nrows + 127) / 128), 128, 0, st>>>(tilemin, L.ld_tm, p.n_ct, nrows, row0, k, nrm, maxn,
                                                                  thr);
  }
  rc = check_launch("hyp_gram_topk(select)");
  if (rc) return rc;
  if (timing) cudaEventRecord(tev[3], st);
  p.ct_step = 1;
  p.n_ct = col_tiles;
  tc_segments(row_pairs, p.n_ct, sms, p.n_seg, p.seg_tiles, a_cost);
  p.seg_cap = (TC_CAND_ROW / p.n_seg) & ~7;
  if (p.seg_cap > TC_CAP) p.seg_cap = TC_CAP;
  items = row_pairs * p.n_seg;
  if (ts) t2<<<(int)(items < sms ? items : sms), TC_THREADS, smem, st>>>(tsmaps[0], tsmaps[1], p);
  else k2<<<(int)(items < sms ? items : sms), TC_THREADS, smem, st>>>(maps[0], maps[1], maps[2], maps[3], p);
  rc = check_launch("hyp_gram_topk(pass 2)");
  if (rc) return rc;
  int64_t fb = (nrows + 3) / 4;
  if (fb > sms * 16) fb = sms * 16;
  if (timing) cudaEventRecord(tev[4], st);
  auto fin = tc_finish_kernel<0>;
  if (D - 1 == 100) fin = tc_finish_kernel<100>;
  else if (D - 1 == 50) fin = tc_finish_kernel<50>;
  fin<<<(int)fb, 128, 0, st>>>(E, ldE, D, row0, nrows, sqrtf(c), sgn, k, cand, cnt, p.n_seg, p.seg_cap, sink, row_flags,
                              flist, nflag);
  rc = check_launch("hyp_gram_topk(finish)");
  if (rc) return rc;
  if (timing) cudaEventRecord(tev[5], st);
  // flagged rows (none on non-degenerate data): the exact kernel reads their number from the device and returns at once
  // when there is nothing to do
  rc = launch_allpairs_topk(E, ldE, n, row0, nrows, D, c, semantics, k, sink, flist, nflag, st);
  if (timing) {
    cudaEventRecord(tev[6], st);
    cudaEventSynchronize(tev[6]);
    float ms[6];
    for (int q = 0; q < 6; ++q) cudaEventElapsedTime(&ms[q], tev[q], tev[q + 1]);
    fprintf(stderr, "[hyp_gram_topk] %s step=%d seg=%d/%d rb=%d pack %.3f  pass1 %.3f  select %.3f  pass2 %.3f  finish %.3f  "
                    "redo %.3f ms\n", ts ? "ts" : "ss", step, p.n_seg, (int)p.seg_tiles, p.rb_per_cta, ms[0], ms[1], ms[2], ms[3], ms[4], ms[5]);
    for (auto &e : tev) cudaEventDestroy(e);
  }
  return rc;
}

}  // namespace hyp

using namespace hyp;

extern "C" int64_t hyp_gram_topk_workspace_bytes(int64_t n, int64_t nrows, int D) {
  if (n < 0 || nrows < 0 || D < 2 || D - 1 + 4 > TC_MAX_SLABS * TC_KSLAB) return -1;
  return (int64_t)tc_layout(n, nrows, D).total;
}

namespace hyp {
int gram_topk_check_args(int64_t n, int64_t row0, int64_t nrows, int D, float c, int k) {
  if (n < 0 || row0 < 0 || nrows < 0 || row0 + nrows > n || D < 2 || !(c > 0.f) || k < 1 || k > 32) {
    set_error("hyp_gram_topk: bad arguments (n=%lld row0=%lld nrows=%lld D=%d k=%d, k <= %d)", (long long)n,
              (long long)row0, (long long)nrows, D, k, 32);
    return HYP_ERR_ARG;
  }
  if (D - 1 + 4 > TC_MAX_SLABS * TC_KSLAB) {
    set_error("hyp_gram_topk: d=%d (+4 time-like columns) exceeds the %d columns one shared-memory tile holds", D - 1,
              TC_MAX_SLABS * TC_KSLAB);
    return HYP_ERR_UNSUPPORTED;
  }
  return HYP_OK;
}
}  // namespace hyp

extern "C" int hyp_gram_topk(const float *E, int64_t ldE, int64_t n, int64_t row0, int64_t nrows, int D, float c,
                             int semantics, int k, int32_t *out_idx, float *out_d, int32_t *row_flags, void *workspace,
                             int64_t workspace_bytes, void *stream) {
  int rc = gram_topk_check_args(n, row0, nrows, D, c, k);
  if (rc) return rc;
  if (nrows == 0) return HYP_OK;
  if (!E || !out_idx || !out_d || !row_flags || !workspace) return HYP_ERR_ARG;
  TopkSink sink{};
  sink.idx = out_idx;
  sink.d = out_d;
  return gram_topk_run(E, ldE, n, row0, nrows, D, c, semantics, k, sink, row_flags, workspace, workspace_bytes,
                       (cudaStream_t)stream);
}
