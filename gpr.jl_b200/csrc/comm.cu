// Multi-GPU plumbing of libgprb200: contexts for several devices of one process, NCCL
// communicators, and the ONE collective of the path - the final all-gather of the per-trial result rows that replaces
// the lock-guarded result callbacks of the reference (examples/parallel/core.jl:47-56).
//
// The path shards by trial (core.jl:28 iterates `jobid`; every trial owns its dataset, CPnoise.jl:13-17): nothing is
// exchanged while optimising or predicting.  The gather payload is a few hundred KB at most (theta*, mll, info,
// predictions per trial), i.e. latency bound; nothing follows it on the device, so there is nothing to fuse it with.
//
// NCCL is bound lazily (dlopen of libnccl.so.2 at the first communicator call): the library loads and runs single-GPU
// without NCCL installed, and inside a PyTorch process it reuses the libnccl that torch already loaded.
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <string.h>

#include <limits>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace gprb {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
#define GPRB_NCCL_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name))
    GPRB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    GPRB_NCCL_SYM(CommInitRank, "ncclCommInitRank");
    GPRB_NCCL_SYM(CommInitAll, "ncclCommInitAll");
    GPRB_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    GPRB_NCCL_SYM(AllGather, "ncclAllGather");
    GPRB_NCCL_SYM(GroupStart, "ncclGroupStart");
    GPRB_NCCL_SYM(GroupEnd, "ncclGroupEnd");
    GPRB_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef GPRB_NCCL_SYM
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommInitAll || !api.CommDestroy || !api.AllGather || !api.GroupStart ||
        !api.GroupEnd || !api.GetErrorString)
      api.handle = nullptr;
  });
  return api.handle ? &api : nullptr;
}

static int nccl_fail(ncclResult_t r, const char* what) {
  NcclApi* a = nccl_api();
  set_error(std::string("NCCL error: ") + (a ? a->GetErrorString(r) : "?") + " in " + what);
  return GPRB_ERR_NCCL;
}

#define GPRB_NCCL(call, what)                          \
  do {                                                 \
    ncclResult_t r__ = (call);                         \
    if (r__ != ncclSuccess) return nccl_fail(r__, what); \
  } while (0)

void comm_release(gprb_ctx* ctx) {
  if (!ctx) return;
  if (ctx->comm) {
    NcclApi* a = nccl_api();
    if (a) a->CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
  }
  cudaFree(ctx->gather_dev);
  if (ctx->gather_host) cudaFreeHost(ctx->gather_host);
  ctx->gather_dev = ctx->gather_host = nullptr;
  ctx->gather_cap = 0;
}

// staging of one rank: send block [per][width + 1] (column 0 = global row id, NaN = unused) and the receive block
static int gather_reserve(gprb_ctx* ctx, size_t send, size_t recv) {
  const size_t need = send + recv;
  if (ctx->gather_cap >= need) return 0;
  GPRB_CUDA(cudaSetDevice(ctx->device));
  cudaFree(ctx->gather_dev);
  if (ctx->gather_host) cudaFreeHost(ctx->gather_host);
  ctx->gather_dev = ctx->gather_host = nullptr;
  ctx->gather_cap = 0;
  GPRB_CUDA(cudaMalloc((void**)&ctx->gather_dev, sizeof(double) * need));
  GPRB_CUDA(cudaMallocHost((void**)&ctx->gather_host, sizeof(double) * need));
  ctx->gather_cap = need;
  return 0;
}

static void fill_send(double* send, int per, int width, int count, const int32_t* ids, const double* rows) {
  const double qnan = std::numeric_limits<double>::quiet_NaN();
  for (int k = 0; k < per; ++k) {
    double* r = send + (size_t)k * (width + 1);
    if (k < count) {
      r[0] = (double)ids[k];
      memcpy(r + 1, rows + (size_t)k * width, sizeof(double) * width);
    } else {
      for (int c = 0; c <= width; ++c) r[c] = qnan;
    }
  }
}

static void scatter_recv(const double* recv, int nblocks, int per, int width, int n_rows, double* out) {
  for (size_t k = 0; k < (size_t)n_rows * width; ++k) out[k] = std::numeric_limits<double>::quiet_NaN();
  for (int k = 0; k < nblocks * per; ++k) {
    const double* r = recv + (size_t)k * (width + 1);
    if (!(r[0] >= 0.0) || r[0] >= (double)n_rows) continue;  // NaN id: unused slot
    memcpy(out + (size_t)(int)r[0] * width, r + 1, sizeof(double) * width);
  }
}

static int check_rows(const char* fn, int n_rows, int width, int count, const int32_t* ids, const double* rows, int per) {
  if (n_rows < 1 || width < 1 || count < 0 || (count > 0 && (!ids || !rows))) {
    set_error(std::string(fn) + ": bad argument");
    return GPRB_ERR_ARG;
  }
  if (count > per) {
    set_error(std::string(fn) + ": count_local exceeds ceil(n_rows / nranks) - use a round-robin or block partition of the trials");
    return GPRB_ERR_ARG;
  }
  for (int k = 0; k < count; ++k)
    if (ids[k] < 0 || ids[k] >= n_rows) {
      set_error(std::string(fn) + ": row id out of range");
      return GPRB_ERR_ARG;
    }
  return 0;
}

}  // namespace gprb

using namespace gprb;

extern "C" {

int gprb_init_multi(gprb_ctx** ctxs, int32_t ngpus, const int* devs) {
  GPRB_REQUIRE(ctxs != nullptr && ngpus >= 1 && ngpus <= 64, "gprb_init_multi: bad argument");
  for (int g = 0; g < ngpus; ++g) ctxs[g] = nullptr;
  std::vector<int> dv(ngpus);
  for (int g = 0; g < ngpus; ++g) dv[g] = devs ? devs[g] : g;
  for (int g = 0; g < ngpus; ++g)
    for (int h = 0; h < g; ++h) GPRB_REQUIRE(dv[g] != dv[h], "gprb_init_multi: a device is listed twice");
  int rc = 0;
  for (int g = 0; g < ngpus && !rc; ++g) rc = gprb_init(&ctxs[g], dv[g]);
  if (!rc && ngpus > 1) {
    NcclApi* a = nccl_api();
    if (!a) {
      set_error("gprb_init_multi: libnccl.so.2 not found (needed for the final gather across GPUs)");
      rc = GPRB_ERR_NCCL;
    } else {
      std::vector<ncclComm_t> comms(ngpus);
      ncclResult_t r = a->CommInitAll(comms.data(), ngpus, dv.data());
      if (r != ncclSuccess) rc = nccl_fail(r, "ncclCommInitAll");
      else
        for (int g = 0; g < ngpus; ++g) { ctxs[g]->comm = comms[g]; ctxs[g]->rank = g; ctxs[g]->nranks = ngpus; }
    }
  }
  if (rc) {
    for (int g = 0; g < ngpus; ++g) { gprb_destroy(ctxs[g]); ctxs[g] = nullptr; }
  }
  return rc;
}

int gprb_comm_unique_id(void* id128) {
  GPRB_REQUIRE(id128 != nullptr, "gprb_comm_unique_id: NULL argument");
  static_assert(sizeof(ncclUniqueId) == 128, "the ABI hands NCCL unique ids around as 128 bytes");
  NcclApi* a = nccl_api();
  if (!a) { set_error("gprb_comm_unique_id: libnccl.so.2 not found"); return GPRB_ERR_NCCL; }
  ncclUniqueId id;
  GPRB_NCCL(a->GetUniqueId(&id), "ncclGetUniqueId");
  memcpy(id128, &id, 128);
  return GPRB_OK;
}

int gprb_comm_init_rank(gprb_ctx* ctx, int32_t nranks, int32_t rank, const void* id128) {
  GPRB_REQUIRE(ctx && id128 && nranks >= 1 && rank >= 0 && rank < nranks, "gprb_comm_init_rank: bad argument");
  GPRB_REQUIRE(ctx->comm == nullptr, "gprb_comm_init_rank: the context already has a communicator");
  NcclApi* a = nccl_api();
  if (!a) { set_error("gprb_comm_init_rank: libnccl.so.2 not found"); return GPRB_ERR_NCCL; }
  GPRB_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t comm;
  GPRB_NCCL(a->CommInitRank(&comm, nranks, id, rank), "ncclCommInitRank");
  ctx->comm = comm; ctx->rank = rank; ctx->nranks = nranks;
  return GPRB_OK;
}

int gprb_gather(gprb_ctx* ctx, int32_t n_rows, int32_t width, int32_t count_local, const int32_t* row_ids,
                const double* rows, double* out) {
  GPRB_REQUIRE(ctx && out, "gprb_gather: NULL argument");
  const int nr = ctx->comm ? ctx->nranks : 1;
  const int per = (n_rows + nr - 1) / nr;
  int rc = check_rows("gprb_gather", n_rows, width, count_local, row_ids, rows, per);
  if (rc) return rc;
  const size_t blk = (size_t)per * (width + 1);
  if ((rc = gather_reserve(ctx, blk, blk * nr))) return rc;
  double* hs = ctx->gather_host;
  double* hr = hs + blk;
  fill_send(hs, per, width, count_local, row_ids, rows);
  if (nr == 1) {  // no communicator: the local scatter
    scatter_recv(hs, 1, per, width, n_rows, out);
    return GPRB_OK;
  }
  NcclApi* a = nccl_api();
  GPRB_CUDA(cudaSetDevice(ctx->device));
  double* ds = ctx->gather_dev;
  double* dr = ds + blk;
  GPRB_CUDA(cudaMemcpyAsync(ds, hs, sizeof(double) * blk, cudaMemcpyHostToDevice, ctx->upload));
  GPRB_NCCL(a->AllGather(ds, dr, blk, ncclDouble, (ncclComm_t)ctx->comm, ctx->upload), "ncclAllGather");
  GPRB_CUDA(cudaMemcpyAsync(hr, dr, sizeof(double) * blk * nr, cudaMemcpyDeviceToHost, ctx->upload));
  GPRB_CUDA(cudaStreamSynchronize(ctx->upload));
  scatter_recv(hr, nr, per, width, n_rows, out);
  return GPRB_OK;
}

int gprb_gather_multi(gprb_ctx* const* ctxs, int32_t ngpus, int32_t n_rows, int32_t width, const int32_t* counts,
                      const int32_t* const* row_ids, const double* const* rows, double* out) {
  GPRB_REQUIRE(ctxs && counts && row_ids && rows && out && ngpus >= 1, "gprb_gather_multi: NULL argument");
  if (ngpus == 1) return gprb_gather(ctxs[0], n_rows, width, counts[0], row_ids[0], rows[0], out);
  const int per = (n_rows + ngpus - 1) / ngpus;
  const size_t blk = (size_t)per * (width + 1);
  int rc;
  for (int g = 0; g < ngpus; ++g) {
    GPRB_REQUIRE(ctxs[g] && ctxs[g]->comm && ctxs[g]->nranks == ngpus && ctxs[g]->rank == g,
                 "gprb_gather_multi: contexts must come from one gprb_init_multi call, in order");
    if ((rc = check_rows("gprb_gather_multi", n_rows, width, counts[g], row_ids[g], rows[g], per))) return rc;
    if ((rc = gather_reserve(ctxs[g], blk, blk * ngpus))) return rc;
    fill_send(ctxs[g]->gather_host, per, width, counts[g], row_ids[g], rows[g]);
  }
  NcclApi* a = nccl_api();
  for (int g = 0; g < ngpus; ++g) {
    GPRB_CUDA(cudaSetDevice(ctxs[g]->device));
    GPRB_CUDA(cudaMemcpyAsync(ctxs[g]->gather_dev, ctxs[g]->gather_host, sizeof(double) * blk, cudaMemcpyHostToDevice, ctxs[g]->upload));
  }
  GPRB_NCCL(a->GroupStart(), "ncclGroupStart");
  for (int g = 0; g < ngpus; ++g) {
    ncclResult_t r = a->AllGather(ctxs[g]->gather_dev, ctxs[g]->gather_dev + blk, blk, ncclDouble, (ncclComm_t)ctxs[g]->comm, ctxs[g]->upload);
    if (r != ncclSuccess) { a->GroupEnd(); return nccl_fail(r, "ncclAllGather"); }
  }
  GPRB_NCCL(a->GroupEnd(), "ncclGroupEnd");
  // every rank holds the full table; rank 0's copy is handed out
  GPRB_CUDA(cudaSetDevice(ctxs[0]->device));
  GPRB_CUDA(cudaMemcpyAsync(ctxs[0]->gather_host + blk, ctxs[0]->gather_dev + blk, sizeof(double) * blk * ngpus, cudaMemcpyDeviceToHost, ctxs[0]->upload));
  for (int g = 0; g < ngpus; ++g) {
    GPRB_CUDA(cudaSetDevice(ctxs[g]->device));
    GPRB_CUDA(cudaStreamSynchronize(ctxs[g]->upload));
  }
  scatter_recv(ctxs[0]->gather_host + blk, ngpus, per, width, n_rows, out);
  return GPRB_OK;
}

}  // extern "C"
