// pointwise.cu -- K1 fused Lorentz pointwise ops and K3 exact pair re-scoring (sm_100a).
// One warp per row pair; lane t reads elements t, t+32, ... (coalesced) and the reduction
// follows ATen's CPU order so the Minkowski product is bit-identical to the reference.
// Compiled with -fmad=false; all arithmetic that must round like torch uses *_rn intrinsics.
#include "common.cuh"

namespace hyp {

constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ int64_t warp_global_id() {
  return (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
mdot_kernel(const float *__restrict__ x, int64_t ldx, const float *__restrict__ y, int64_t ldy,
            float *__restrict__ out, int64_t n, int D) {
  const int lane = threadIdx.x & 31;
  for (int64_t r = warp_global_id(); r < n; r += (int64_t)gridDim.x * kWarpsPerBlock) {
    float m = warp_mdot(x + r * ldx, y + r * ldy, D, lane);
    if (lane == 0) out[r] = m;
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
distance_kernel(const float *__restrict__ x, int64_t ldx, const float *__restrict__ y, int64_t ldy,
                float *__restrict__ out, int64_t n, int D, float sqrt_c, float sgn) {
  const int lane = threadIdx.x & 31;
  for (int64_t r = warp_global_id(); r < n; r += (int64_t)gridDim.x * kWarpsPerBlock) {
    float m = warp_mdot(x + r * ldx, y + r * ldy, D, lane);
    if (lane == 0) out[r] = dist_from_mdot(m, sgn, sqrt_c);
  }
}

// backward of distance_kernel: gx = w * (y0, -ys), gy = w * (x0, -xs), w = g * d(distance)/d<x,y>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
distance_bwd_kernel(const float *__restrict__ x, int64_t ldx, const float *__restrict__ y, int64_t ldy,
                    const float *__restrict__ g, float *__restrict__ gx, float *__restrict__ gy, int64_t n, int D,
                    float sqrt_c, float sgn) {
  const int lane = threadIdx.x & 31;
  for (int64_t r = warp_global_id(); r < n; r += (int64_t)gridDim.x * kWarpsPerBlock) {
    const float *xr = x + r * ldx, *yr = y + r * ldy;
    const float m = warp_mdot(xr, yr, D, lane);
    const float w = dist_grad_from_mdot(m, sgn, sqrt_c, g[r]);
    for (int k = lane; k < D; k += 32) {
      const float wy = __fmul_rn(w, yr[k]), wx = __fmul_rn(w, xr[k]);
      if (gx) gx[r * D + k] = k == 0 ? wy : -wy;
      if (gy) gy[r * D + k] = k == 0 ? wx : -wx;
    }
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
rescore_kernel(const float *__restrict__ E, int64_t ldE, const int32_t *__restrict__ ii,
               const int32_t *__restrict__ jj, float *__restrict__ d_out, float *__restrict__ u_out,
               int64_t n, int D, float sqrt_c, float sgn) {
  const int lane = threadIdx.x & 31;
  for (int64_t r = warp_global_id(); r < n; r += (int64_t)gridDim.x * kWarpsPerBlock) {
    float m = warp_mdot(E + (int64_t)ii[r] * ldE, E + (int64_t)jj[r] * ldE, D, lane);
    if (lane == 0) {
      d_out[r] = dist_from_mdot(m, sgn, sqrt_c);
      if (u_out) u_out[r] = m;
    }
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
logmap_kernel(const float *__restrict__ x, int64_t ldx, const float *__restrict__ y, int64_t ldy,
              float *__restrict__ out, int64_t ldo, int64_t n, int D, int semantics) {
  const int lane = threadIdx.x & 31;
  for (int64_t r = warp_global_id(); r < n; r += (int64_t)gridDim.x * kWarpsPerBlock) {
    const float *xr = x + r * ldx, *yr = y + r * ldy;
    float m = warp_mdot(xr, yr, D, lane);
    float ms;
    float coef = logmap_coef(m, semantics, &ms);
    for (int k = lane; k < D; k += 32)
      out[r * ldo + k] = __fmul_rn(coef, __fadd_rn(yr[k], __fmul_rn(ms, xr[k])));
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
expmap_kernel(const float *__restrict__ x, int64_t ldx, const float *__restrict__ v, int64_t ldv,
              float *__restrict__ out, int64_t ldo, int64_t n, int D) {
  const int lane = threadIdx.x & 31;
  for (int64_t r = warp_global_id(); r < n; r += (int64_t)gridDim.x * kWarpsPerBlock) {
    float *o = out + r * ldo;
    warp_expmap(x + r * ldx, v + r * ldv, D, lane, [&](int k, float val) { o[k] = val; });
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
project_kernel(const float *__restrict__ x, int64_t ldx, float *__restrict__ out, int64_t ldo,
               int64_t n, int D, float c) {
  const int lane = threadIdx.x & 31;
  for (int64_t r = warp_global_id(); r < n; r += (int64_t)gridDim.x * kWarpsPerBlock) {
    const float *xr = x + r * ldx;
    float nr = warp_norm_aten([&](int e) { return xr[1 + e]; }, D - 1, lane);
    float x0 = __fsqrt_rn(__fadd_rn(1.0f, __fmul_rn(__fmul_rn(c, nr), nr)));
    // spatial part first (out may alias x), then the time component
    for (int k = lane; k < D; k += 32)
      if (k) out[r * ldo + k] = xr[k];
    __syncwarp();
    if (lane == 0) out[r * ldo] = x0;
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
midpoint_kernel(const float *__restrict__ E, int64_t ldE, const int32_t *__restrict__ ii,
                const int32_t *__restrict__ jj, const int32_t *__restrict__ li,
                const int32_t *__restrict__ lj, float *__restrict__ out, int64_t ldo, int64_t n,
                int D, float c, int semantics, int project) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float *buf = smem + (size_t)w * 2 * D;
  for (int64_t r = warp_global_id(); r < n; r += (int64_t)gridDim.x * kWarpsPerBlock) {
    float *o = out + r * ldo;
    warp_midpoint(E + (int64_t)ii[r] * ldE, E + (int64_t)jj[r] * ldE, li[r], lj[r], D, c, semantics,
                  project != 0, buf, lane, [&](int k, float val) { o[k] = val; });
  }
}

static inline int grid_for(int64_t n) {
  int64_t b = (n + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int64_t cap = 148 * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

static int bad_dims(int64_t n, int D) {
  if (n == 0 && D >= 2 && D <= HYP_MAX_D) return 0;
  if (n < 0 || D < 2 || D > HYP_MAX_D) {
    set_error("bad shape: n=%lld D=%d (need n>=0, 2<=D<=%d)", (long long)n, D, HYP_MAX_D);
    return 1;
  }
  return 0;
}

}  // namespace hyp

using namespace hyp;

extern "C" int hyp_minkowski_dot(const float *x, int64_t ldx, const float *y, int64_t ldy, float *out,
                                 int64_t n, int D, void *stream) {
  if (bad_dims(n, D)) return HYP_ERR_ARG;
  if (n == 0) return HYP_OK;
  if (!x || !y || !out) return HYP_ERR_ARG;
  mdot_kernel<<<grid_for(n), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(x, ldx, y, ldy, out, n, D);
  return check_launch("hyp_minkowski_dot");
}

extern "C" int hyp_distance(const float *x, int64_t ldx, const float *y, int64_t ldy, float *out,
                            int64_t n, int D, float c, int semantics, void *stream) {
  if (bad_dims(n, D)) return HYP_ERR_ARG;
  if (n == 0) return HYP_OK;
  if (!x || !y || !out) return HYP_ERR_ARG;
  distance_kernel<<<grid_for(n), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
      x, ldx, y, ldy, out, n, D, sqrtf(c), semantics == HYP_SEM_REFERENCE ? -1.f : 1.f);
  return check_launch("hyp_distance");
}

extern "C" int hyp_distance_backward(const float *x, int64_t ldx, const float *y, int64_t ldy, const float *grad_out,
                                     float *grad_x, float *grad_y, int64_t n, int D, float c, int semantics,
                                     void *stream) {
  if (bad_dims(n, D)) return HYP_ERR_ARG;
  if (n == 0) return HYP_OK;
  if (!x || !y || !grad_out || (!grad_x && !grad_y)) return HYP_ERR_ARG;
  distance_bwd_kernel<<<grid_for(n), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
      x, ldx, y, ldy, grad_out, grad_x, grad_y, n, D, sqrtf(c), semantics == HYP_SEM_REFERENCE ? -1.f : 1.f);
  return check_launch("hyp_distance_backward");
}

extern "C" int hyp_rescore_pairs(const float *E, int64_t ldE, const int32_t *idx_i, const int32_t *idx_j,
                                 float *d_out, float *u_out, int64_t n, int D, float c, int semantics,
                                 void *stream) {
  if (bad_dims(n, D)) return HYP_ERR_ARG;
  if (n == 0) return HYP_OK;
  if (!E || !idx_i || !idx_j || !d_out) return HYP_ERR_ARG;
  rescore_kernel<<<grid_for(n), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
      E, ldE, idx_i, idx_j, d_out, u_out, n, D, sqrtf(c), semantics == HYP_SEM_REFERENCE ? -1.f : 1.f);
  return check_launch("hyp_rescore_pairs");
}

extern "C" int hyp_log_map(const float *x, int64_t ldx, const float *y, int64_t ldy, float *out,
                           int64_t ldo, int64_t n, int D, int semantics, void *stream) {
  if (bad_dims(n, D)) return HYP_ERR_ARG;
  if (n == 0) return HYP_OK;
  if (!x || !y || !out) return HYP_ERR_ARG;
  logmap_kernel<<<grid_for(n), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(x, ldx, y, ldy, out, ldo, n,
                                                                              D, semantics);
  return check_launch("hyp_log_map");
}

extern "C" int hyp_exp_map(const float *x, int64_t ldx, const float *v, int64_t ldv, float *out,
                           int64_t ldo, int64_t n, int D, void *stream) {
  if (bad_dims(n, D)) return HYP_ERR_ARG;
  if (n == 0) return HYP_OK;
  if (!x || !v || !out) return HYP_ERR_ARG;
  expmap_kernel<<<grid_for(n), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(x, ldx, v, ldv, out, ldo, n, D);
  return check_launch("hyp_exp_map");
}

extern "C" int hyp_project(const float *x, int64_t ldx, float *out, int64_t ldo, int64_t n, int D, float c,
                           void *stream) {
  if (bad_dims(n, D)) return HYP_ERR_ARG;
  if (n == 0) return HYP_OK;
  if (!x || !out) return HYP_ERR_ARG;
  project_kernel<<<grid_for(n), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(x, ldx, out, ldo, n, D, c);
  return check_launch("hyp_project");
}

extern "C" int hyp_midpoint(const float *E, int64_t ldE, const int32_t *idx_i, const int32_t *idx_j,
                            const int32_t *len_i, const int32_t *len_j, float *out, int64_t ldo, int64_t n,
                            int D, float c, int semantics, int project, void *stream) {
  if (bad_dims(n, D)) return HYP_ERR_ARG;
  if (n == 0) return HYP_OK;
  if (!E || !idx_i || !idx_j || !len_i || !len_j || !out) return HYP_ERR_ARG;
  size_t smem = (size_t)kWarpsPerBlock * 2 * D * sizeof(float);
  // (per device, so set before every launch: a process-wide "already set" flag would skip it on a second GPU)
  if (cudaFuncSetAttribute(midpoint_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) != cudaSuccess) {
    set_error("hyp_midpoint: cudaFuncSetAttribute: %s", cudaGetErrorString(cudaGetLastError()));
    return HYP_ERR_CUDA;
  }
  midpoint_kernel<<<grid_for(n), kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
      E, ldE, idx_i, idx_j, len_i, len_j, out, ldo, n, D, c, semantics, project);
  return check_launch("hyp_midpoint");
}
