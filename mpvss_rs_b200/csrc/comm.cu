// Multi-GPU plumbing of the C ABI: one context per GPU (one process or thread each), joined by an NCCL
// communicator.  Participants are independent, so the only exchange of a phase is ONE all-gather of the
// fixed-size transcript rows (north_star: "results are gathered once per phase with an
// NCCL all-gather over NVLink"); it is issued from here, on the library's own stream, between the kernels
// and the copy to the host.
//
// libnccl is bound at run time (dlopen "libnccl.so.2"): a single-GPU user of libmpvss_b200.so needs no
// NCCL installation, and inside a PyTorch process the already loaded NCCL is reused.
#include <dlfcn.h>
#include <nccl.h>
#include "ctx.h"
#include "transcript.h"

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi* nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {getenv("MPVSS_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) {
      api.error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "not found");
      return;
    }
    auto sym = [&](const char* s) {
      void* p = dlsym(api.handle, s);
      if (!p && api.error.empty()) api.error = std::string("libnccl lacks ") + s;
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  });
  return &api;
}

int nccl_fail(mpvss_ctx* ctx, ncclResult_t r, const char* what) {
  const NcclApi* a = nccl();
  return mpvss_fail(ctx, MPVSS_ERR_COMM, std::string(what) + ": " + (a->GetErrorString ? a->GetErrorString(r) : "NCCL error"));
}

}  // namespace

int comm_allgather(mpvss_ctx* ctx, const void* src, void* dst, size_t bytes) {
  if (ctx->nranks <= 1 || !ctx->comm) {
    MPVSS_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return MPVSS_OK;
  }
  ncclResult_t r = nccl()->AllGather(src, dst, bytes, ncclUint8, static_cast<ncclComm_t>(ctx->comm), ctx->stream);
  if (r != ncclSuccess) return nccl_fail(ctx, r, "ncclAllGather");
  return MPVSS_OK;
}

namespace {
// one warp per row; rows are multiples of 4 bytes (1056 / 164 / 160)
__global__ void reorder_rows_kernel(const uint32_t* in, uint32_t* out, uint32_t n_total, uint32_t nranks, uint32_t rpr,
                                    uint32_t row_words) {
  const uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n_total) return;
  const uint32_t* src = in + ((size_t)(i % nranks) * rpr + i / nranks) * row_words;
  uint32_t* dst = out + (size_t)i * row_words;
  for (uint32_t k = threadIdx.x & 31u; k < row_words; k += 32) dst[k] = src[k];
}
}  // namespace

int comm_reorder_rows(mpvss_ctx* ctx, const void* gathered, void* ordered, size_t n_total, size_t row_bytes) {
  const uint32_t rpr = (uint32_t)transcript::rows_per_rank(n_total, ctx->nranks);
  reorder_rows_kernel<<<(unsigned)((n_total + 7) / 8), 256, 0, ctx->stream>>>(
      static_cast<const uint32_t*>(gathered), static_cast<uint32_t*>(ordered), (uint32_t)n_total, (uint32_t)ctx->nranks, rpr,
      (uint32_t)(row_bytes / 4));
  MPVSS_CUDA(ctx, cudaGetLastError());
  return MPVSS_OK;
}

void comm_release(mpvss_ctx* ctx) {
  if (ctx->comm) {
    nccl()->CommDestroy(static_cast<ncclComm_t>(ctx->comm));
    ctx->comm = nullptr;
  }
  ctx->nranks = 1;
  ctx->rank = 0;
}

extern "C" {

int mpvss_comm_unique_id(uint8_t* id_out, size_t id_len) {
  if (!id_out || id_len < MPVSS_COMM_ID_BYTES) return MPVSS_ERR_ARG;
  NcclApi* a = nccl();
  if (!a->error.empty()) return MPVSS_ERR_COMM;
  ncclUniqueId id;
  static_assert(sizeof(id) == MPVSS_COMM_ID_BYTES, "ncclUniqueId size");
  if (a->GetUniqueId(&id) != ncclSuccess) return MPVSS_ERR_COMM;
  memcpy(id_out, &id, sizeof id);
  return MPVSS_OK;
}

int mpvss_comm_init(mpvss_ctx* ctx, const uint8_t* id, size_t id_len, int nranks, int rank) {
  if (!ctx) return MPVSS_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  if (!id || id_len < MPVSS_COMM_ID_BYTES || nranks < 1 || rank < 0 || rank >= nranks)
    return mpvss_fail(ctx, MPVSS_ERR_ARG, "comm_init: bad arguments");
  NcclApi* a = nccl();
  if (!a->error.empty()) return mpvss_fail(ctx, MPVSS_ERR_COMM, a->error);
  MPVSS_CUDA(ctx, cudaSetDevice(ctx->device));
  comm_release(ctx);
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof uid);
  ncclComm_t c = nullptr;
  ncclResult_t r = a->CommInitRank(&c, nranks, uid, rank);
  if (r != ncclSuccess) return nccl_fail(ctx, r, "ncclCommInitRank");
  ctx->comm = c;
  ctx->nranks = nranks;
  ctx->rank = rank;
  ctx->v_n = 0;  // anything staged was sliced for the old communicator
  return MPVSS_OK;
}

int mpvss_comm_destroy(mpvss_ctx* ctx) {
  if (!ctx) return MPVSS_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  comm_release(ctx);
  ctx->v_n = 0;
  return MPVSS_OK;
}

int mpvss_comm_size(const mpvss_ctx* ctx) { return ctx ? ctx->nranks : 0; }
int mpvss_comm_rank(const mpvss_ctx* ctx) { return ctx ? ctx->rank : -1; }

}  // extern "C"
