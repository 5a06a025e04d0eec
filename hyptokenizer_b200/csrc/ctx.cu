// ctx.cu -- hyp_ctx: multi-GPU context of the row-sharded all-pairs / top-k path (K8).
//
// The reference has no multi-GPU code at all; the north star shards the vocabulary rows over the GPUs of one
// NVSwitch box and merges the per-shard top-k lists by an all-gather.  Here the all-gather is not a separate
// collective: every rank owns a gather buffer (cudaMalloc, exported with CUDA IPC and mapped by every peer), and
// the kernel that PRODUCES a row of the result -- tc_finish_kernel / the exact redo kernel, through TopkSink --
// stores it into the buffers of all ranks over NVLink as it goes.  What remains of the collective is one
// barrier kernel: a release flag to every peer, an acquire spin on the local flags.
//
// Buffers are double-buffered by call parity: the barrier of call e+1 cannot complete before every rank has
// launched call e+1, i.e. (stream order) after its consumers of call e, so call e+2 may overwrite call e's buffer.
// One host thread per ctx; every call is stream-ordered on the caller's stream.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace hyp {
int gram_topk_check_args(int64_t n, int64_t row0, int64_t nrows, int D, float c, int k);
int gram_topk_run(const float *E, int64_t ldE, int64_t n, int64_t row0, int64_t nrows, int D, float c, int semantics,
                  int k, const TopkSink &sink, int32_t *row_flags, void *workspace, int64_t workspace_bytes,
                  cudaStream_t st);
}  // namespace hyp

struct hyp_ctx {
  int rank, world, device;
  int64_t slot_bytes;   // what one rank contributes to a gather
  size_t buf_bytes;     // one parity buffer: world slots
  size_t flags_off;     // unsigned int arrivals[HYP_MAX_PEERS], then int status
  size_t alloc_bytes;
  uint8_t *local;
  uint8_t *peer[HYP_MAX_PEERS];   // peer[rank] == local
  unsigned int epoch;   // gathers completed
  int connected;
};

namespace hyp {

struct PeerFlags {
  unsigned int *flags[HYP_MAX_PEERS];
};

// Thread p announces this rank's arrival at call `epoch` to rank p and waits for rank p's arrival here.  The stores
// of the kernels before it on the stream are complete (stream order); the system-scope fence + release store publish
// them to the peer that acquires the flag.  A peer that never arrives (crashed rank) ends the spin after ~20 s with
// status = 1 instead of hanging the GPU.
__global__ void ctx_barrier_kernel(PeerFlags pf, int rank, int world, unsigned int epoch, int *status) {
  const int p = threadIdx.x;
  if (p >= world) return;
  __threadfence_system();
  unsigned int *dst = pf.flags[p] + rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(epoch) : "memory");
  const unsigned int *src = pf.flags[rank] + p;
  unsigned int v = 0;
  unsigned int ns = 32;
  for (long long it = 0;; ++it) {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
    if ((int)(v - epoch) >= 0) break;
    if (it > (1ll << 24)) {            // ~20 s of 1 us naps
      *status = 1;
      break;
    }
    __nanosleep(ns);
    if (ns < 1024) ns <<= 1;
  }
}

// Generic push for results that were not written through a TopkSink: this rank's slot -> every rank's buffer.
__global__ void ctx_push_kernel(const int4 *__restrict__ src, int64_t n16, TopkSink sink) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
    const int4 v = src[i];
#pragma unroll
    for (int g = 0; g < HYP_MAX_PEERS; ++g)
      if (g < sink.n_pairs) reinterpret_cast<int4 *>(sink.pairs[g])[i] = v;
  }
}

static TopkSink ctx_sink(hyp_ctx *ctx) {
  TopkSink s{};
  const size_t off = (size_t)(ctx->epoch & 1) * ctx->buf_bytes + (size_t)ctx->rank * (size_t)ctx->slot_bytes;
  for (int g = 0; g < ctx->world; ++g) s.pairs[g] = reinterpret_cast<int2 *>(ctx->peer[g] + off);
  s.n_pairs = ctx->world;
  return s;
}

static int ctx_barrier(hyp_ctx *ctx, void **gathered, cudaStream_t st) {
  uint8_t *mine = ctx->local + (size_t)(ctx->epoch & 1) * ctx->buf_bytes;
  ctx->epoch += 1;
  if (ctx->world > 1) {
    PeerFlags pf{};
    for (int g = 0; g < ctx->world; ++g) pf.flags[g] = reinterpret_cast<unsigned int *>(ctx->peer[g] + ctx->flags_off);
    int *status = reinterpret_cast<int *>(ctx->local + ctx->flags_off + HYP_MAX_PEERS * sizeof(unsigned int));
    ctx_barrier_kernel<<<1, 32, 0, st>>>(pf, ctx->rank, ctx->world, ctx->epoch, status);
    int rc = check_launch("hyp_ctx barrier");
    if (rc) return rc;
  }
  if (gathered) *gathered = mine;
  return HYP_OK;
}

static int ctx_usable(hyp_ctx *ctx, const char *what) {
  if (!ctx) {
    set_error("%s: NULL context", what);
    return HYP_ERR_ARG;
  }
  if (!ctx->connected) {
    set_error("%s: hyp_ctx_connect has not been called", what);
    return HYP_ERR_ARG;
  }
  int dev = -1;
  cudaGetDevice(&dev);
  if (dev != ctx->device) {
    set_error("%s: context belongs to device %d, current device is %d", what, ctx->device, dev);
    return HYP_ERR_ARG;
  }
  return HYP_OK;
}

}  // namespace hyp

using namespace hyp;

extern "C" int hyp_ctx_create(hyp_ctx **out, int rank, int world, int64_t slot_bytes) {
  if (!out || world < 1 || world > HYP_MAX_PEERS || rank < 0 || rank >= world || slot_bytes < 256 || (slot_bytes & 255)) {
    set_error("hyp_ctx_create: bad arguments (rank=%d world=%d slot_bytes=%lld: world <= %d, slot_bytes a positive "
              "multiple of 256)", rank, world, (long long)slot_bytes, HYP_MAX_PEERS);
    return HYP_ERR_ARG;
  }
  hyp_ctx *ctx = (hyp_ctx *)calloc(1, sizeof(hyp_ctx));
  if (!ctx) return HYP_ERR_CUDA;
  ctx->rank = rank;
  ctx->world = world;
  cudaGetDevice(&ctx->device);
  ctx->slot_bytes = slot_bytes;
  ctx->buf_bytes = (size_t)world * (size_t)ctx->slot_bytes;
  ctx->flags_off = 2 * ctx->buf_bytes;
  ctx->alloc_bytes = ctx->flags_off + 256;
  cudaError_t e = cudaMalloc((void **)&ctx->local, ctx->alloc_bytes);
  if (e == cudaSuccess) e = cudaMemset(ctx->local + ctx->flags_off, 0, 256);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_error("hyp_ctx_create: %s", cudaGetErrorString(e));
    if (ctx->local) cudaFree(ctx->local);
    free(ctx);
    return HYP_ERR_CUDA;
  }
  ctx->peer[rank] = ctx->local;
  ctx->connected = world == 1;
  *out = ctx;
  return HYP_OK;
}

extern "C" int hyp_ctx_export(hyp_ctx *ctx, void *handle_out) {
  if (!ctx || !handle_out) return HYP_ERR_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == HYP_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ctx->local);
  if (e != cudaSuccess) {
    set_error("hyp_ctx_export: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    return HYP_ERR_CUDA;
  }
  memcpy(handle_out, &h, sizeof(h));
  return HYP_OK;
}

extern "C" int hyp_ctx_connect(hyp_ctx *ctx, const void *handles) {
  if (!ctx || (!handles && ctx->world > 1)) return HYP_ERR_ARG;
  for (int g = 0; g < ctx->world; ++g) {
    if (g == ctx->rank || ctx->peer[g]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const uint8_t *)handles + (size_t)g * HYP_IPC_HANDLE_BYTES, sizeof(h));
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_error("hyp_ctx_connect: cudaIpcOpenMemHandle(rank %d): %s", g, cudaGetErrorString(e));
      cudaGetLastError();
      return HYP_ERR_CUDA;
    }
    ctx->peer[g] = (uint8_t *)p;
  }
  ctx->connected = 1;
  return HYP_OK;
}

extern "C" int hyp_ctx_status(hyp_ctx *ctx, int *status_out) {
  if (!ctx || !status_out) return HYP_ERR_ARG;
  cudaError_t e = cudaMemcpy(status_out, ctx->local + ctx->flags_off + HYP_MAX_PEERS * sizeof(unsigned int), sizeof(int),
                             cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    set_error("hyp_ctx_status: %s", cudaGetErrorString(e));
    return HYP_ERR_CUDA;
  }
  return HYP_OK;
}

extern "C" int hyp_ctx_destroy(hyp_ctx *ctx) {
  if (!ctx) return HYP_OK;
  cudaDeviceSynchronize();
  for (int g = 0; g < ctx->world; ++g)
    if (g != ctx->rank && ctx->peer[g]) cudaIpcCloseMemHandle(ctx->peer[g]);
  if (ctx->local) cudaFree(ctx->local);
  free(ctx);
  return HYP_OK;
}

extern "C" int hyp_allgather_topk(hyp_ctx *ctx, const void *local, int64_t bytes, void **gathered, void *stream) {
  int rc = ctx_usable(ctx, "hyp_allgather_topk");
  if (rc) return rc;
  if (!local || bytes < 0 || bytes > ctx->slot_bytes || (bytes & 15) || ((uintptr_t)local & 15)) {
    set_error("hyp_allgather_topk: bytes=%lld must be a multiple of 16 and <= the context's slot (%lld), `local` "
              "16-byte aligned", (long long)bytes, (long long)ctx->slot_bytes);
    return HYP_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (bytes > 0) {
    const int64_t n16 = bytes / 16;
    int64_t blocks = (n16 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    ctx_push_kernel<<<(int)blocks, 256, 0, st>>>((const int4 *)local, n16, ctx_sink(ctx));
    rc = check_launch("hyp_allgather_topk(push)");
    if (rc) return rc;
  }
  return ctx_barrier(ctx, gathered, st);
}

extern "C" int hyp_gram_topk_allgather(hyp_ctx *ctx, const float *E, int64_t ldE, int64_t n, int D, float c,
                                       int semantics, int k, int32_t *row_flags, void *workspace,
                                       int64_t workspace_bytes, void **gathered, void *stream) {
  int rc = ctx_usable(ctx, "hyp_gram_topk_allgather");
  if (rc) return rc;
  const int64_t per = (n + ctx->world - 1) / ctx->world;
  const int64_t row0 = (int64_t)ctx->rank * per < n ? (int64_t)ctx->rank * per : n;
  const int64_t nrows = per < n - row0 ? per : n - row0;
  rc = gram_topk_check_args(n, row0, nrows, D, c, k);
  if (rc) return rc;
  if (per * k * 8 > ctx->slot_bytes) {
    set_error("hyp_gram_topk_allgather: a shard of %lld rows x k=%d needs %lld bytes, the context's slot has %lld",
              (long long)per, k, (long long)(per * k * 8), (long long)ctx->slot_bytes);
    return HYP_ERR_WORKSPACE;
  }
  if (!E || !gathered || (nrows > 0 && (!row_flags || !workspace))) return HYP_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (nrows > 0) {
    rc = gram_topk_run(E, ldE, n, row0, nrows, D, c, semantics, k, ctx_sink(ctx), row_flags, workspace, workspace_bytes,
                       st);
    if (rc) return rc;
  }
  return ctx_barrier(ctx, gathered, st);
}
