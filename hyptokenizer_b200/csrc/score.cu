// score.cu -- K7: the inner loop of frequency-aware candidate scoring (sm_100a).
//
// Replaces, for a whole candidate list at once, the per-candidate Python loop of
// FrequencyAwareHyperbolicTokenizer._compute_semantic_coherence
// (tokenizer/frequency_aware_hyperbolic_merge.py:114-166): the UN-projected weighted midpoint of the
// candidate's two rows (:139-141) and its Lorentz distance to up to 50 sampled rows (:149-153).
// The sample indices come from the host (torch.randperm on the reference's global CPU generator, so
// the RNG stream is consumed exactly as the reference does).  hyp_coherence_distances returns the distances and leaves
// the float64 mean / sigmoid / weighted score to the host (numpy, as in the reference :160-199: what the tokenizer
// class uses, for bit parity of the scores); hyp_score_candidates does the whole score on the device.
// One warp per candidate; exact ATen-order products (compiled with -fmad=false).
#include "common.cuh"

namespace hyp {

constexpr int kScoreWarps = 4;

__global__ void __launch_bounds__(kScoreWarps * 32)
coherence_kernel(const float *__restrict__ E, int64_t ldE, const int32_t *__restrict__ ii,
                 const int32_t *__restrict__ jj, const int32_t *__restrict__ li, const int32_t *__restrict__ lj,
                 const int32_t *__restrict__ sample, int S, float *__restrict__ out, int64_t C, int D, float c,
                 int semantics, float sqrt_c, float sgn) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float *buf = smem + (size_t)w * 3 * D;   // 2D scratch + D merged point
  float *merged = buf + 2 * D;
  for (int64_t r = (int64_t)blockIdx.x * kScoreWarps + w; r < C; r += (int64_t)gridDim.x * kScoreWarps) {
    warp_midpoint(E + (int64_t)ii[r] * ldE, E + (int64_t)jj[r] * ldE, li[r], lj[r], D, c, semantics, false, buf, lane,
                  [&](int k, float v) { merged[k] = v; });
    __syncwarp();
    for (int s = 0; s < S; ++s) {
      const int32_t row = sample[r * S + s];
      float m = warp_mdot(merged, E + (int64_t)row * ldE, D, lane);
      if (lane == 0) out[r * S + s] = dist_from_mdot(m, sgn, sqrt_c);
    }
    __syncwarp();
  }
}

// The whole score of reference :114-199 for C candidates in one launch: coherence distances as above, then -- in
// float64, in the reference's own operation order -- the mean of the kept distances (numpy's pairwise sum: eight
// running sums for n >= 8, combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), the remainder added one by one),
// coh = 1 / (1 + exp(mean - threshold)) (0 when nothing is kept), dist_score = 1 / (1 + d) and
// score = alpha * dist_score + beta * freq_score + gamma * coh.  freq_score = log1p(f) / log1p(f_max) comes from the
// host (a dict lookup by token strings).  A sample equal to i or j is not kept (:149-151).
__global__ void __launch_bounds__(kScoreWarps * 32)
score_candidates_kernel(const float *__restrict__ E, int64_t ldE, const int32_t *__restrict__ ii,
                        const int32_t *__restrict__ jj, const int32_t *__restrict__ li, const int32_t *__restrict__ lj,
                        const int32_t *__restrict__ sample, int S, const float *__restrict__ dist,
                        const double *__restrict__ freq_score, double alpha, double beta, double gamma, double threshold,
                        double *__restrict__ score, double *__restrict__ coh_out, int64_t C, int D, float c,
                        int semantics, float sqrt_c, float sgn) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float *buf = smem + (size_t)w * (3 * D + 64);          // 2D scratch + D merged point + up to 64 kept distances
  float *merged = buf + 2 * D;
  float *vals = merged + D;
  for (int64_t r = (int64_t)blockIdx.x * kScoreWarps + w; r < C; r += (int64_t)gridDim.x * kScoreWarps) {
    const int32_t ci = ii[r], cj = jj[r];
    warp_midpoint(E + (int64_t)ci * ldE, E + (int64_t)cj * ldE, li[r], lj[r], D, c, semantics, false, buf, lane,
                  [&](int k, float v) { merged[k] = v; });
    __syncwarp();
    int kept = 0;
    for (int s = 0; s < S; ++s) {
      const int32_t row = sample[r * S + s];
      if (row == ci || row == cj) continue;                       // (uniform over the warp)
      const float m = warp_mdot(merged, E + (int64_t)row * ldE, D, lane);
      if (lane == 0) vals[kept] = dist_from_mdot(m, sgn, sqrt_c);
      ++kept;
    }
    __syncwarp();
    if (lane == 0) {
      double coh = 0.0;
      if (kept > 0) {
        double sum;
        if (kept < 8) {
          sum = 0.0;
          for (int q = 0; q < kept; ++q) sum = __dadd_rn(sum, (double)vals[q]);
        } else {
          double rr[8];
          for (int q = 0; q < 8; ++q) rr[q] = (double)vals[q];
          int q = 8;
          for (; q < kept - (kept % 8); q += 8)
            for (int t = 0; t < 8; ++t) rr[t] = __dadd_rn(rr[t], (double)vals[q + t]);
          sum = __dadd_rn(__dadd_rn(__dadd_rn(rr[0], rr[1]), __dadd_rn(rr[2], rr[3])),
                          __dadd_rn(__dadd_rn(rr[4], rr[5]), __dadd_rn(rr[6], rr[7])));
          for (; q < kept; ++q) sum = __dadd_rn(sum, (double)vals[q]);
        }
        const double avg = __ddiv_rn(sum, (double)kept);
        coh = __ddiv_rn(1.0, __dadd_rn(1.0, exp(__dsub_rn(avg, threshold))));
      }
      const double ds = __ddiv_rn(1.0, __dadd_rn(1.0, (double)dist[r]));
      score[r] = __dadd_rn(__dadd_rn(__dmul_rn(alpha, ds), __dmul_rn(beta, freq_score[r])), __dmul_rn(gamma, coh));
      if (coh_out) coh_out[r] = coh;
    }
    __syncwarp();
  }
}

}  // namespace hyp

using namespace hyp;

extern "C" int hyp_score_candidates(const float *E, int64_t ldE, const int32_t *idx_i, const int32_t *idx_j,
                                    const int32_t *len_i, const int32_t *len_j, const int32_t *sample, int S,
                                    const float *dist, const double *freq_score, double alpha, double beta, double gamma,
                                    double threshold, double *score, double *coherence, int64_t C, int D, float c,
                                    int semantics, void *stream) {
  if (C < 0 || S < 0 || S > 64 || D < 2 || D > HYP_MAX_D || !(c > 0.f)) {
    set_error("hyp_score_candidates: bad shape C=%lld S=%d D=%d (S <= 64)", (long long)C, S, D);
    return HYP_ERR_ARG;
  }
  if (C == 0) return HYP_OK;
  if (!E || !idx_i || !idx_j || !len_i || !len_j || (S > 0 && !sample) || !dist || !freq_score || !score) return HYP_ERR_ARG;
  const size_t smem = (size_t)kScoreWarps * (3 * D + 64) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(score_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("hyp_score_candidates: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return HYP_ERR_CUDA;
  }
  int64_t blocks = (C + kScoreWarps - 1) / kScoreWarps;
  if (blocks > 148 * 16) blocks = 148 * 16;
  score_candidates_kernel<<<(int)blocks, kScoreWarps * 32, smem, (cudaStream_t)stream>>>(
      E, ldE, idx_i, idx_j, len_i, len_j, sample, S, dist, freq_score, alpha, beta, gamma, threshold, score, coherence, C,
      D, c, semantics, sqrtf(c), semantics == HYP_SEM_REFERENCE ? -1.f : 1.f);
  return check_launch("hyp_score_candidates");
}

extern "C" int hyp_coherence_distances(const float *E, int64_t ldE, const int32_t *idx_i, const int32_t *idx_j,
                                       const int32_t *len_i, const int32_t *len_j, const int32_t *sample, int S,
                                       float *out, int64_t C, int D, float c, int semantics, void *stream) {
  if (C < 0 || S < 0 || D < 2 || D > HYP_MAX_D || !(c > 0.f)) {
    set_error("hyp_coherence_distances: bad shape C=%lld S=%d D=%d", (long long)C, S, D);
    return HYP_ERR_ARG;
  }
  if (C == 0 || S == 0) return HYP_OK;
  if (!E || !idx_i || !idx_j || !len_i || !len_j || !sample || !out) return HYP_ERR_ARG;
  const size_t smem = (size_t)kScoreWarps * 3 * D * sizeof(float);
  int64_t blocks = (C + kScoreWarps - 1) / kScoreWarps;
  if (blocks > 148 * 16) blocks = 148 * 16;
  coherence_kernel<<<(int)blocks, kScoreWarps * 32, smem, (cudaStream_t)stream>>>(
      E, ldE, idx_i, idx_j, len_i, len_j, sample, S, out, C, D, c, semantics, sqrtf(c),
      semantics == HYP_SEM_REFERENCE ? -1.f : 1.f);
  return check_launch("hyp_coherence_distances");
}
