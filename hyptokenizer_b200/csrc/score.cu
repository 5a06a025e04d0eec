// score.cu -- K7: the inner loop of frequency-aware candidate scoring (sm_100a).
//
// Replaces, for a whole candidate list at once, the per-candidate Python loop of
// FrequencyAwareHyperbolicTokenizer._compute_semantic_coherence
// (tokenizer/frequency_aware_hyperbolic_merge.py:114-166): the UN-projected weighted midpoint of the
// candidate's two rows (:139-141) and its Lorentz distance to up to 50 sampled rows (:149-153).
// The sample indices come from the host (torch.randperm on the reference's global CPU generator, so
// the RNG stream is consumed exactly as the reference does); the float64 mean / sigmoid / weighted
// score stay on the host in numpy, as in the reference (:160-199).
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

}  // namespace hyp

using namespace hyp;

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
