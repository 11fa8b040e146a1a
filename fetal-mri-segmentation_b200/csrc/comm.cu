// comm.cu — the data-parallel collectives of the hot path on raw NCCL (NVLink 5 / NVSwitch), behind the C ABI.
//
// The reference is single-process (fetal_net/training.py:115-117: workers=1, use_multiprocessing=False); these entry
// points are the extension SURVEY.md §8e asks for: one process per GPU, (1) all-reduce of the 8 soft-Dice sums in the
// forward pass (metrics.py:11-15 flattens the batch axis, so Dice is a whole-batch statistic), (2) bucketed SUM
// all-reduce of the flat fp32 gradient buffer on a side stream overlapped with the rest of backward, (3) one reduce of
// the float64 partial sums of patch-sharded sliding-window inference. No device pointer leaves the library.
//
// libnccl is resolved at run time (dlopen "libnccl.so.2": inside a process that has imported torch this is torch's
// bundled NCCL 2.28, already mapped under that soname; otherwise the system copy) so that libfetalb200.so keeps linking
// against libcudart only.
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "common.cuh"

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

NcclApi g_nccl;
std::once_flag g_nccl_once;
std::string g_nccl_err;

void load_nccl() {
  const char* env = getenv("FETAL_B200_NCCL_LIB");
  const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    if (!nm || !nm[0]) continue;
    g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  if (!g_nccl.handle) {
    g_nccl_err = std::string("libnccl.so.2 not found (dlopen: ") + (dlerror() ? dlerror() : "?") +
                 "); set FETAL_B200_NCCL_LIB";
    return;
  }
#define NCCL_SYM(field, sym)                                              \
  g_nccl.field = (decltype(g_nccl.field))dlsym(g_nccl.handle, sym);       \
  if (!g_nccl.field) {                                                    \
    g_nccl_err = std::string("libnccl: missing symbol ") + sym;          \
    return;                                                               \
  }
  NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  NCCL_SYM(CommInitRank, "ncclCommInitRank")
  NCCL_SYM(CommDestroy, "ncclCommDestroy")
  NCCL_SYM(AllReduce, "ncclAllReduce")
  NCCL_SYM(Reduce, "ncclReduce")
  NCCL_SYM(Broadcast, "ncclBroadcast")
  NCCL_SYM(GetErrorString, "ncclGetErrorString")
  NCCL_SYM(GetVersion, "ncclGetVersion")
#undef NCCL_SYM
}

int nccl_ready() {
  std::call_once(g_nccl_once, load_nccl);
  FM_CHECK(g_nccl_err.empty(), FM_ECOMM, "%s", g_nccl_err.c_str());
  return FM_OK;
}

#define FM_NCCL(expr)                                                                              \
  do {                                                                                             \
    ncclResult_t _r = (expr);                                                                      \
    if (_r != ncclSuccess) {                                                                       \
      fm_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r));       \
      return FM_ECOMM;                                                                             \
    }                                                                                              \
  } while (0)

static_assert(sizeof(ncclUniqueId) == FM_COMM_UID_BYTES, "ncclUniqueId is 128 bytes");

}  // namespace

extern "C" int fm_comm_unique_id(uint8_t out[FM_COMM_UID_BYTES]) {
  FM_CHECK(out, FM_EINVAL, "fm_comm_unique_id: NULL argument");
  FM_TRY(nccl_ready());
  ncclUniqueId id;
  FM_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(out, &id, sizeof(id));
  return FM_OK;
}

extern "C" int fm_comm_init(fm_ctx* ctx, int rank, int nranks, const uint8_t uid[FM_COMM_UID_BYTES]) {
  FM_CHECK(ctx && uid && nranks >= 1 && rank >= 0 && rank < nranks, FM_EINVAL, "fm_comm_init: rank %d of %d", rank,
           nranks);
  FM_CHECK(ctx->comm == nullptr, FM_ESTATE, "fm_comm_init: communicator already initialised");
  FM_TRY(nccl_ready());
  FM_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, uid, sizeof(id));
  ncclComm_t comm = nullptr;
  FM_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
  ctx->comm = comm;
  ctx->comm_rank = rank;
  ctx->comm_size = nranks;
  ctx->comm_enabled = true;
  if (!ctx->comm_stream) {
    // higher priority than the compute stream: a bucket that is ready should not queue behind persistent conv CTAs
    int lo = 0, hi = 0;
    FM_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    FM_CUDA(cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, hi));
    FM_CUDA(cudaEventCreateWithFlags(&ctx->comm_ev, cudaEventDisableTiming));
  }
  return FM_OK;
}

extern "C" int fm_comm_destroy(fm_ctx* ctx) {
  if (!ctx || !ctx->comm) return FM_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->comm_stream) cudaStreamSynchronize(ctx->comm_stream);
  g_nccl.CommDestroy((ncclComm_t)ctx->comm);
  ctx->comm = nullptr;
  ctx->comm_size = 1;
  ctx->comm_rank = 0;
  return FM_OK;
}

extern "C" int fm_comm_info(fm_ctx* ctx, int out[3]) {
  FM_CHECK(ctx && out, FM_EINVAL, "fm_comm_info: NULL argument");
  out[0] = ctx->comm_rank;
  out[1] = ctx->comm ? ctx->comm_size : 1;
  int v = 0;
  if (ctx->comm && g_nccl.GetVersion) g_nccl.GetVersion(&v);
  out[2] = v;
  return FM_OK;
}

extern "C" int fm_comm_enable(fm_ctx* ctx, int on) {
  FM_CHECK(ctx, FM_EINVAL, "fm_comm_enable: NULL ctx");
  ctx->comm_enabled = on != 0;
  return FM_OK;
}

extern "C" uint64_t fm_comm_stream(fm_ctx* ctx) { return ctx ? (uint64_t)(uintptr_t)ctx->comm_stream : 0; }

// dtype: 0 = float32, 1 = float64. In place, SUM. A context without a communicator (single GPU) is a no-op.
int comm_allreduce(fm_ctx* ctx, void* buf, size_t count, int dtype, cudaStream_t stream) {
  if (!ctx->comm || ctx->comm_size == 1 || !ctx->comm_enabled || count == 0) return FM_OK;
  FM_NCCL(g_nccl.AllReduce(buf, buf, count, dtype ? ncclDouble : ncclFloat, ncclSum, (ncclComm_t)ctx->comm, stream));
  return FM_OK;
}
int comm_reduce(fm_ctx* ctx, void* buf, size_t count, int dtype, int root, cudaStream_t stream) {
  if (!ctx->comm || ctx->comm_size == 1 || !ctx->comm_enabled || count == 0) return FM_OK;
  FM_NCCL(g_nccl.Reduce(buf, buf, count, dtype ? ncclDouble : ncclFloat, ncclSum, root, (ncclComm_t)ctx->comm, stream));
  return FM_OK;
}
int comm_broadcast(fm_ctx* ctx, void* buf, size_t count, int dtype, int root, cudaStream_t stream) {
  if (!ctx->comm || ctx->comm_size == 1 || count == 0) return FM_OK;
  FM_NCCL(g_nccl.Broadcast(buf, buf, count, dtype ? ncclDouble : ncclFloat, root, (ncclComm_t)ctx->comm, stream));
  return FM_OK;
}

// Times `iters` back-to-back in-place fp32 SUM all-reduces of `bytes` on the communication stream (CUDA events on
// that stream), after 3 warm-up rounds. bench.py derives bus GB/s = 2 (n-1)/n * bytes / t from it.
extern "C" int fm_comm_allreduce_bench(fm_ctx* ctx, int64_t bytes, int iters, float* ms_per_iter) {
  FM_CHECK(ctx && ctx->comm && bytes >= 4 && iters > 0 && ms_per_iter, FM_EINVAL, "fm_comm_allreduce_bench: bad argument");
  FM_CUDA(cudaSetDevice(ctx->device));
  float* buf = nullptr;
  FM_CUDA(cudaMalloc((void**)&buf, (size_t)bytes));
  FM_CUDA(cudaMemset(buf, 0, (size_t)bytes));
  cudaEvent_t e0, e1;
  FM_CUDA(cudaEventCreate(&e0));
  FM_CUDA(cudaEventCreate(&e1));
  int rc = FM_OK;
  for (int i = 0; i < 3 && rc == FM_OK; ++i) rc = comm_allreduce(ctx, buf, (size_t)bytes / 4, 0, ctx->comm_stream);
  if (rc == FM_OK) {
    cudaStreamSynchronize(ctx->comm_stream);
    cudaEventRecord(e0, ctx->comm_stream);
    for (int i = 0; i < iters && rc == FM_OK; ++i) rc = comm_allreduce(ctx, buf, (size_t)bytes / 4, 0, ctx->comm_stream);
    cudaEventRecord(e1, ctx->comm_stream);
    cudaStreamSynchronize(ctx->comm_stream);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    *ms_per_iter = ms / (float)iters;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  return rc;
}
