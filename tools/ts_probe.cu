// ts_probe.cu -- feasibility probe (not part of the library): does tcgen05.mma kind::tf32 take its A operand from
// TMEM ("TS" form) with the layout "row m on lane m, element k in column a_col + k", M = 128, N = 64, K = 8 per
// instruction?  One CTA, known inputs, result compared with the host.  Also runs the SS form (A from shared memory) on
// the same data as a control of the harness.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o ts_probe tools/ts_probe.cu && ./ts_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int M = 128, N = 64, K = 16;   // two k-steps of 8

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// element (r, k) of a K-major tile of 128-byte rows under SWIZZLE_128B (what TMA writes): 16-byte chunk index ^= r & 7
__device__ __forceinline__ uint32_t sw128_off(int r, int k) { return r * 128 + ((((k >> 2) ^ (r & 7)) & 7) << 4) + (k & 3) * 4; }

template <int NCOLS>
constexpr uint32_t idesc_tf32() {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NCOLS >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(128) probe(const float *A, const float *B, float *D_ts, float *D_ss) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *base = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *sA = base;                    // 128 rows x 128 B
  uint8_t *sB = base + 16384;            // 64 rows x 128 B
  uint64_t *bar = (uint64_t *)(base + 16384 + 8192);
  uint32_t *tmem_ptr = (uint32_t *)(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (16384 + 8192) / 4; i += 128) ((uint32_t *)base)[i] = 0;
  __syncthreads();
  for (int i = tid; i < M * K; i += 128) {
    const int r = i / K, k = i % K;
    *(float *)(sA + sw128_off(r, k)) = A[r * K + k];
  }
  for (int i = tid; i < N * K; i += 128) {
    const int r = i / K, k = i % K;
    *(float *)(sB + sw128_off(r, k)) = B[r * K + k];
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the MMA's async proxy
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_ptr;
  const uint32_t lane_addr = tmem + ((uint32_t)(32 * warp) << 16);
  constexpr uint32_t D_TS = 0, D_SS = 64, A_COL = 256;

  // A -> TMEM: thread = row (lane 32*warp + lane), 16 consecutive 32-bit columns
  {
    uint32_t r[16];
    const int row = 32 * warp + lane;
#pragma unroll
    for (int k = 0; k < 16; ++k) r[k] = __float_as_uint(A[row * K + k]);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(lane_addr + A_COL), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (tid == 0) {
    const uint32_t idesc = idesc_tf32<N>();
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint64_t bd = desc_sw128(smem_u32(sB) + ks * 32);
      const uint64_t ad = desc_sw128(smem_u32(sA) + ks * 32);
      const uint32_t acc = ks > 0;
      // TS form: A from TMEM
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                   ::"r"(tmem + D_TS), "r"(tmem + A_COL + 8 * ks), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      // SS form: A from shared memory
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem + D_SS), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  while (!mbar_try_wait(bar, 0)) {
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int which = 0; which < 2; ++which) {
    float *out = which == 0 ? D_ts : D_ss;
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t r[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(lane_addr + (which == 0 ? D_TS : D_SS) + c0)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int row = 32 * warp + lane;
      for (int c = 0; c < 32; ++c) out[row * N + c0 + c] = __uint_as_float(r[c]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  float *hA = (float *)malloc(M * K * 4), *hB = (float *)malloc(N * K * 4), *ref = (float *)malloc(M * N * 4);
  for (int i = 0; i < M * K; ++i) hA[i] = (float)((i * 7 + 3) % 17 - 8) * 0.125f;     // exact in TF32
  for (int i = 0; i < N * K; ++i) hB[i] = (float)((i * 5 + 1) % 13 - 6) * 0.25f;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0.f;
      for (int k = 0; k < K; ++k) s += hA[m * K + k] * hB[n * K + k];
      ref[m * N + n] = s;
    }
  float *dA, *dB, *dT, *dS;
  cudaMalloc(&dA, M * K * 4); cudaMalloc(&dB, N * K * 4); cudaMalloc(&dT, M * N * 4); cudaMalloc(&dS, M * N * 4);
  cudaMemcpy(dA, hA, M * K * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, N * K * 4, cudaMemcpyHostToDevice);
  cudaMemset(dT, 0xff, M * N * 4); cudaMemset(dS, 0xff, M * N * 4);
  const int smem = 1024 + 16384 + 8192 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(dA, dB, dT, dS);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  float *hT = (float *)malloc(M * N * 4), *hS = (float *)malloc(M * N * 4);
  cudaMemcpy(hT, dT, M * N * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hS, dS, M * N * 4, cudaMemcpyDeviceToHost);
  int bad_t = 0, bad_s = 0;
  for (int i = 0; i < M * N; ++i) {
    bad_t += hT[i] != ref[i];
    bad_s += hS[i] != ref[i];
  }
  printf("SS form (control): %d of %d elements differ\n", bad_s, M * N);
  printf("TS form (A in TMEM): %d of %d elements differ\n", bad_t, M * N);
  for (int m = 0; m < 3; ++m) printf("row %d: ref %g %g %g | ts %g %g %g | ss %g %g %g\n", m, ref[m * N], ref[m * N + 1], ref[m * N + 2],
                                     hT[m * N], hT[m * N + 1], hT[m * N + 2], hS[m * N], hS[m * N + 1], hS[m * N + 2]);
  return (bad_t || bad_s) ? 2 : 0;
}
