// K21-K23 — the collective layer: gradient all-reduce (SUM), loss reduce and weight broadcast of the data-parallel step over
// NCCL (NVLink 5 / NVSwitch), called directly — no framework in between.
//
// Replaces what MultiWorkerMirroredStrategy does inside optimizer.apply_gradients / strategy.reduce / strategy.scope
// (W:834, W:848, W:896; V:1246, V:1260, V:1266; CommunicationOptions(NCCL, timeout 120 s) V:1463-1475):
//   * one communicator per (process, GPU), created from a 128-byte unique id that the HOST exchanges (any out-of-band channel:
//     the Python host uses the torch.distributed store, a C host can use a file or MPI);
//   * every collective is enqueued on the CALLER's stream, so it orders with the kernels around it and can be captured into
//     the step's CUDA graph (no host round trip at the fwd/bwd -> reduce -> Adam boundaries);
//   * ts_comm_alloc hands out ncclMemAlloc memory registered with the communicator (ncclCommRegister): buffers the NVSwitch
//     can reduce in place (NVLS multicast / zero-copy); the gradient arenas live there;
//   * ts_comm_allreduce_bucket optionally pre-multiplies this rank's contribution by a DEVICE scalar inside the collective
//     (ncclRedOpCreatePreMulSum) — the reference's local clip_by_global_norm factor (V:1243) costs no extra pass;
//   * ts_comm_check polls ncclCommGetAsyncError (the reference's 120 s collective timeout becomes an error code, TS_ENCCL,
//     instead of a hang).
// libnccl is resolved at run time (dlopen: the copy already in the process — torch's — or the system one), so libtethys.so
// still loads on a box without NCCL or a GPU for the symbol-export test.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>
#include <vector>
#include "common.cuh"

namespace ts {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
  ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t*) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*MemAlloc)(void**, size_t) = nullptr;
  ncclResult_t (*MemFree)(void*) = nullptr;
  ncclResult_t (*CommRegister)(const ncclComm_t, void*, size_t, void**) = nullptr;
  ncclResult_t (*CommDeregister)(const ncclComm_t, void*) = nullptr;
  ncclResult_t (*RedOpCreatePreMulSum)(ncclRedOp_t*, void*, ncclDataType_t, ncclScalarResidence_t, ncclComm_t) = nullptr;
  ncclResult_t (*RedOpDestroy)(ncclRedOp_t, ncclComm_t) = nullptr;
};

static NcclApi g_nccl;
static std::string g_nccl_err;

static bool nccl_load() {
  if (g_nccl.lib) return true;
  const char* cand[4] = {getenv("TETHYS_NCCL_LIB"), "libnccl.so.2", "libnccl.so", nullptr};
  void* h = nullptr;
  for (int i = 0; i < 3 && !h; ++i)
    if (cand[i] && cand[i][0]) h = dlopen(cand[i], RTLD_NOW | RTLD_GLOBAL);
  if (!h) { g_nccl_err = std::string("libnccl not found: ") + (dlerror() ? dlerror() : "?"); return false; }
#define TS_SYM(field, name, required)                                              \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));        \
  if (!g_nccl.field && required) { g_nccl_err = std::string("libnccl lacks ") + name; return false; }
  TS_SYM(GetVersion, "ncclGetVersion", true)
  TS_SYM(GetUniqueId, "ncclGetUniqueId", true)
  TS_SYM(CommInitRank, "ncclCommInitRank", true)
  TS_SYM(CommDestroy, "ncclCommDestroy", true)
  TS_SYM(CommAbort, "ncclCommAbort", false)
  TS_SYM(CommGetAsyncError, "ncclCommGetAsyncError", true)
  TS_SYM(GetErrorString, "ncclGetErrorString", true)
  TS_SYM(AllReduce, "ncclAllReduce", true)
  TS_SYM(Broadcast, "ncclBroadcast", true)
  TS_SYM(ReduceScatter, "ncclReduceScatter", false)
  TS_SYM(AllGather, "ncclAllGather", false)
  TS_SYM(MemAlloc, "ncclMemAlloc", false)
  TS_SYM(MemFree, "ncclMemFree", false)
  TS_SYM(CommRegister, "ncclCommRegister", false)
  TS_SYM(CommDeregister, "ncclCommDeregister", false)
  TS_SYM(RedOpCreatePreMulSum, "ncclRedOpCreatePreMulSum", false)
  TS_SYM(RedOpDestroy, "ncclRedOpDestroy", false)
#undef TS_SYM
  g_nccl.lib = h;
  return true;
}

struct CommBuf { void* ptr; size_t bytes; void* reg; bool nccl_mem; };

struct Comm {
  Ctx* ctx = nullptr;
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0, version = 0;
  std::vector<CommBuf> bufs;
  long long collectives = 0;
};

static int nccl_fail(Ctx* ctx, ncclResult_t r, const char* what) {
  return set_err(ctx, TS_ENCCL, "%s failed: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl error");
}
#define TS_NCCL_OK(ctx, expr)                                \
  do {                                                       \
    ncclResult_t _r = (expr);                                \
    if (_r != ncclSuccess) return ts::nccl_fail((ctx), _r, #expr); \
  } while (0)

static bool to_nccl_dtype(int dt, ncclDataType_t* out) {
  switch (dt) {
    case TS_F32: *out = ncclFloat32; return true;
    case TS_BF16: *out = ncclBfloat16; return true;
    default: return false;
  }
}

}  // namespace ts

extern "C" {

int ts_comm_unique_id(ts_ctx* ctx_, void* out128) {
  ts::Ctx* ctx = reinterpret_cast<ts::Ctx*>(ctx_);
  if (!ctx || !out128) return TS_EINVAL;
  if (!ts::nccl_load()) return ts::set_err(ctx, TS_ENCCL, "%s", ts::g_nccl_err.c_str());
  static_assert(sizeof(ncclUniqueId) == 128, "the ABI carries the NCCL unique id as 128 opaque bytes");
  ncclUniqueId id;
  TS_NCCL_OK(ctx, ts::g_nccl.GetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
  return 0;
}

int ts_comm_init(ts_ctx* ctx_, const void* unique_id128, int nranks, int rank, ts_comm** out) {
  ts::Ctx* ctx = reinterpret_cast<ts::Ctx*>(ctx_);
  if (!ctx || !unique_id128 || !out || nranks < 1 || rank < 0 || rank >= nranks) return TS_EINVAL;
  if (!ts::nccl_load()) return ts::set_err(ctx, TS_ENCCL, "%s", ts::g_nccl_err.c_str());
  TS_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  ts::Comm* c = new ts::Comm();
  c->ctx = ctx; c->nranks = nranks; c->rank = rank;
  ts::g_nccl.GetVersion(&c->version);
  ncclUniqueId id;
  memcpy(&id, unique_id128, sizeof(id));
  ncclResult_t r = ts::g_nccl.CommInitRank(&c->comm, nranks, id, rank);
  if (r != ncclSuccess) { delete c; return ts::nccl_fail(ctx, r, "ncclCommInitRank"); }
  *out = reinterpret_cast<ts_comm*>(c);
  return 0;
}

int ts_comm_info(ts_comm* c_, int* nranks, int* rank, int* nccl_version, int* registered_buffers) {
  ts::Comm* c = reinterpret_cast<ts::Comm*>(c_);
  if (!c) return TS_EINVAL;
  if (nranks) *nranks = c->nranks;
  if (rank) *rank = c->rank;
  if (nccl_version) *nccl_version = c->version;
  if (registered_buffers) {
    int n = 0;
    for (auto& b : c->bufs) n += b.reg != nullptr;
    *registered_buffers = n;
  }
  return 0;
}

int ts_comm_alloc(ts_comm* c_, int64_t bytes, void** ptr) {
  ts::Comm* c = reinterpret_cast<ts::Comm*>(c_);
  if (!c || !ptr || bytes <= 0) return TS_EINVAL;
  ts::Ctx* ctx = c->ctx;
  ts::CommBuf b = {nullptr, (size_t)bytes, nullptr, false};
  // ncclMemAlloc gives memory the fabric can map for in-switch (NVLS) reductions; plain cudaMalloc is the fallback of an
  // NCCL build without it (registration is then skipped and NCCL stages through its own buffers)
  if (ts::g_nccl.MemAlloc && ts::g_nccl.MemAlloc(&b.ptr, b.bytes) == ncclSuccess) {
    b.nccl_mem = true;
    if (ts::g_nccl.CommRegister && ts::g_nccl.CommRegister(c->comm, b.ptr, b.bytes, &b.reg) != ncclSuccess) b.reg = nullptr;
  } else {
    TS_CUDA_OK(ctx, cudaMalloc(&b.ptr, b.bytes));
  }
  TS_CUDA_OK(ctx, cudaMemset(b.ptr, 0, b.bytes));
  c->bufs.push_back(b);
  *ptr = b.ptr;
  return 0;
}

int ts_comm_free(ts_comm* c_, void* ptr) {
  ts::Comm* c = reinterpret_cast<ts::Comm*>(c_);
  if (!c || !ptr) return TS_EINVAL;
  for (size_t i = 0; i < c->bufs.size(); ++i)
    if (c->bufs[i].ptr == ptr) {
      ts::CommBuf b = c->bufs[i];
      c->bufs.erase(c->bufs.begin() + i);
      cudaDeviceSynchronize();
      if (b.reg && ts::g_nccl.CommDeregister) ts::g_nccl.CommDeregister(c->comm, b.reg);
      if (b.nccl_mem) ts::g_nccl.MemFree(b.ptr);
      else cudaFree(b.ptr);
      return 0;
    }
  return ts::set_err(c->ctx, TS_EINVAL, "ts_comm_free: %p was not allocated by this communicator", ptr);
}

int ts_comm_broadcast(ts_comm* c_, void* buf, int64_t count, int dtype, int root, void* stream) {
  ts::Comm* c = reinterpret_cast<ts::Comm*>(c_);
  if (!c || !buf || count < 0 || root < 0 || root >= c->nranks) return TS_EINVAL;
  ncclDataType_t dt;
  if (!ts::to_nccl_dtype(dtype, &dt)) return ts::set_err(c->ctx, TS_EDTYPE, "ts_comm_broadcast: dtype %d", dtype);
  TS_NCCL_OK(c->ctx, ts::g_nccl.Broadcast(buf, buf, (size_t)count, dt, root, c->comm, (cudaStream_t)stream));
  c->collectives++;
  return 0;
}

int ts_comm_allreduce_bucket(ts_comm* c_, void* buf, int64_t count, int dtype, const void* premul_scale_dev, void* stream) {
  ts::Comm* c = reinterpret_cast<ts::Comm*>(c_);
  if (!c || !buf || count < 0) return TS_EINVAL;
  ncclDataType_t dt;
  if (!ts::to_nccl_dtype(dtype, &dt)) return ts::set_err(c->ctx, TS_EDTYPE, "ts_comm_allreduce_bucket: dtype %d", dtype);
  ncclRedOp_t op = ncclSum;
  bool custom = false;
  if (premul_scale_dev) {
    // sum_r scale_r * x_r: the scalar (same dtype as the data, in device memory) is read when the collective RUNS
    TS_REQUIRE(c->ctx, ts::g_nccl.RedOpCreatePreMulSum, TS_EUNSUPPORTED, "this NCCL build has no ncclRedOpCreatePreMulSum");
    TS_NCCL_OK(c->ctx, ts::g_nccl.RedOpCreatePreMulSum(&op, const_cast<void*>(premul_scale_dev), dt, ncclScalarDevice, c->comm));
    custom = true;
  }
  ncclResult_t r = ts::g_nccl.AllReduce(buf, buf, (size_t)count, dt, op, c->comm, (cudaStream_t)stream);
  if (custom) ts::g_nccl.RedOpDestroy(op, c->comm);   // NCCL keeps the op alive until the collectives enqueued with it finish
  if (r != ncclSuccess) return ts::nccl_fail(c->ctx, r, "ncclAllReduce");
  c->collectives++;
  return 0;
}

int ts_comm_check(ts_comm* c_) {
  ts::Comm* c = reinterpret_cast<ts::Comm*>(c_);
  if (!c) return TS_EINVAL;
  ncclResult_t async = ncclSuccess;
  TS_NCCL_OK(c->ctx, ts::g_nccl.CommGetAsyncError(c->comm, &async));
  if (async != ncclSuccess && async != ncclInProgress) return ts::nccl_fail(c->ctx, async, "asynchronous NCCL error (ncclCommGetAsyncError)");
  return 0;
}

int ts_comm_finalize(ts_comm* c_) {
  ts::Comm* c = reinterpret_cast<ts::Comm*>(c_);
  if (!c) return TS_EINVAL;
  cudaDeviceSynchronize();
  for (auto& b : c->bufs) {
    if (b.reg && ts::g_nccl.CommDeregister) ts::g_nccl.CommDeregister(c->comm, b.reg);
    if (b.nccl_mem) ts::g_nccl.MemFree(b.ptr);
    else cudaFree(b.ptr);
  }
  c->bufs.clear();
  ncclResult_t r = ts::g_nccl.CommDestroy(c->comm);
  ts::Ctx* ctx = c->ctx;
  delete c;
  if (r != ncclSuccess) return ts::nccl_fail(ctx, r, "ncclCommDestroy");
  return 0;
}

}  // extern "C"
