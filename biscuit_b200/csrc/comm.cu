// Multi-GPU exchange step of the hot path behind the C ABI: one NCCL communicator per bq_ctx and
// fixed-size all-gathers of raw bytes on the ctx stream.  The path has exactly two exchange patterns -- per-slide
// aggregates (48 B per slide) after the local, bit-exact slide reduction, and per-tile (pred, unc, label) triples for
// the cohort-wide tile ROCs of `detect` -- and both are "every rank contributes a block, everyone needs all blocks".
// The reference has no collective at all (single process); this is what a non-Python host binds to drive 2/4/8 GPUs.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2, the copy PyTorch already loaded when there is one), like
// cuTensorMapEncodeTiled in model.cu: libbiscuit_b200.so keeps linking only the static CUDA runtime.
#include <dlfcn.h>

#include "common.cuh"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;                 // ncclSuccess == 0
constexpr int kNcclUint8 = 1;             // ncclDataType_t: ncclInt8 = 0, ncclUint8 = 1

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
  bool ok() const { return handle && GetUniqueId && CommInitRank && CommDestroy && AllGather; }
};

NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
    api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) {
    api.error = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
    return api;
  }
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
  api.AllGather = (decltype(api.AllGather))dlsym(api.handle, "ncclAllGather");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
  if (!api.ok()) api.error = "libnccl.so.2 lacks a required symbol";
  return api;
}

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  DevBuf send, recv;
};

std::vector<std::pair<bq_ctx*, Comm*>>& registry() {
  static std::vector<std::pair<bq_ctx*, Comm*>> r;
  return r;
}
Comm* find(bq_ctx* ctx) {
  for (auto& e : registry())
    if (e.first == ctx) return e.second;
  return nullptr;
}

int nccl_fail(bq_ctx* ctx, const char* what, ncclResult_t r) {
  NcclApi& a = nccl();
  return bq_fail(ctx, BQ_ERR_CUDA, "%s failed: %s", what, a.GetErrorString ? a.GetErrorString(r) : "NCCL error");
}

}  // namespace

extern "C" {

int bq_comm_unique_id(uint8_t id[BQ_COMM_ID_BYTES]) {
  NcclApi& a = nccl();
  if (!a.ok() || !id) return BQ_ERR_STATE;
  ncclUniqueId u;
  if (a.GetUniqueId(&u) != 0) return BQ_ERR_CUDA;
  memcpy(id, u.internal, BQ_COMM_ID_BYTES);
  return BQ_OK;
}

int bq_comm_init(bq_ctx* ctx, int32_t rank, int32_t world, const uint8_t id[BQ_COMM_ID_BYTES]) {
  if (!ctx) return BQ_ERR_ARG;
  if (!id || world < 1 || rank < 0 || rank >= world) return bq_fail(ctx, BQ_ERR_ARG, "bq_comm_init: bad argument");
  if (find(ctx)) return bq_fail(ctx, BQ_ERR_STATE, "bq_comm_init: this context already has a communicator");
  NcclApi& a = nccl();
  if (!a.ok()) return bq_fail(ctx, BQ_ERR_STATE, "NCCL unavailable: %s", a.error.c_str());
  BQ_CUDA(ctx, cudaSetDevice(ctx->device));
  Comm* c = new Comm();
  c->rank = rank;
  c->world = world;
  ncclUniqueId u;
  memcpy(u.internal, id, BQ_COMM_ID_BYTES);
  ncclResult_t r = a.CommInitRank(&c->comm, world, u, rank);
  if (r != 0) { delete c; return nccl_fail(ctx, "ncclCommInitRank", r); }
  registry().push_back({ctx, c});
  return BQ_OK;
}

int bq_comm_size(bq_ctx* ctx, int32_t* rank, int32_t* world) {
  if (!ctx) return BQ_ERR_ARG;
  Comm* c = find(ctx);
  if (rank) *rank = c ? c->rank : 0;
  if (world) *world = c ? c->world : 1;
  return BQ_OK;
}

void bq_comm_destroy(bq_ctx* ctx) {
  auto& reg = registry();
  for (size_t i = 0; i < reg.size(); ++i) {
    if (reg[i].first != ctx) continue;
    Comm* c = reg[i].second;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (c->comm) nccl().CommDestroy(c->comm);
    delete c;
    reg.erase(reg.begin() + i);
    return;
  }
}

int bq_allgather_bytes(bq_ctx* ctx, const void* send, int64_t nbytes, void* recv) {
  if (!ctx) return BQ_ERR_ARG;
  if (nbytes < 0 || (nbytes > 0 && (!send || !recv))) return bq_fail(ctx, BQ_ERR_ARG, "bq_allgather_bytes: bad argument");
  Comm* c = find(ctx);
  BQ_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!c || c->world == 1) {                         // no communicator: a world of one
    if (nbytes) BQ_CUDA(ctx, cudaMemcpyAsync(recv, send, (size_t)nbytes, cudaMemcpyDefault, ctx->stream));
    BQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BQ_OK;
  }
  if (nbytes == 0) return BQ_OK;
  const bool send_dev = bq_is_device_ptr(send), recv_dev = bq_is_device_ptr(recv);
  const void* s = send;
  void* r = recv;
  int rc;
  if (!send_dev) {
    if ((rc = bq_alloc(ctx, c->send, (size_t)nbytes))) return rc;
    BQ_CUDA(ctx, cudaMemcpyAsync(c->send.p, send, (size_t)nbytes, cudaMemcpyHostToDevice, ctx->stream));
    s = c->send.p;
  }
  if (!recv_dev) {
    if ((rc = bq_alloc(ctx, c->recv, (size_t)nbytes * c->world))) return rc;
    r = c->recv.p;
  }
  ncclResult_t nr = nccl().AllGather(s, r, (size_t)nbytes, kNcclUint8, c->comm, ctx->stream);
  if (nr != 0) return nccl_fail(ctx, "ncclAllGather", nr);
  if (!recv_dev) BQ_CUDA(ctx, cudaMemcpyAsync(recv, r, (size_t)nbytes * c->world, cudaMemcpyDeviceToHost, ctx->stream));
  BQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BQ_OK;
}

}  // extern "C"
