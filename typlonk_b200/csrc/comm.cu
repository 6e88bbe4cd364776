// Multi-GPU plumbing of libtyplonk_b200: ranks, the two exchange steps of the prover
// (an all-gather of small device buffers after every MSM batch, a device-buffer broadcast per quotient
// coset / witness slice), and the single-process device group behind tp_ctx_create_multi.
//
// Transports, chosen once per context:
//   * NCCL, owned by the library (dlopen of libnccl.so.2 -- inside a torch process that is the copy torch already
//     loaded): ncclCommInitRank from a unique id the host language distributes once (one process per GPU), or
//     ncclCommInitAll (one process, one worker thread per GPU).  Stream-ordered on the ctx stream; no host round trip.
//   * "local": ranks that live in one process WITHOUT NCCL -- several ranks on one device (how the single-GPU tests
//     exercise the whole sharded path), or a host without libnccl.  Peer copies plus a host barrier.
// The reference (plonk/src/proof.rs:26-57) is one call in one process: a group context keeps that shape -- the
// caller's thread hands each entry point to one worker per device and gets one result back.
#include <dlfcn.h>
#include <nccl.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "common.cuh"

namespace tp {

// ---- NCCL through dlopen ----------------------------------------------------------------------------
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  bool ok = false;
};
static NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {getenv("TP_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
    auto sym = [&](const char* s) { return dlsym(api.handle, s); };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.ok = api.GroupStart && api.GroupEnd && api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.AllGather && api.Broadcast &&
             api.GetErrorString;
  });
  return api;
}
static int nccl_fail(tp_ctx* ctx, const char* what, ncclResult_t r) {
  ctx->err = std::string(what) + ": " + nccl_api().GetErrorString(r);
  return TP_ERR_COLLECTIVE;
}

// ---- local transport ---------------------------------------------------------------------------------
struct LocalGroup {
  int world = 0;
  std::mutex m;
  std::condition_variable cv;
  int arrived = 0;
  uint64_t generation = 0;
  std::vector<void*> ptr;
  void barrier() {
    std::unique_lock<std::mutex> lk(m);
    const uint64_t gen = generation;
    if (++arrived == world) {
      arrived = 0;
      generation++;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return generation != gen; });
    }
  }
};

// ---- collectives the prover calls ----------------------------------------------------------------------
// All ranks call with the same sizes.  NCCL: enqueued on ctx->stream.  Local: returns with the data in place.
// In place (send_dev == recv_dev + rank * bytes) is allowed on every transport.
int comm_allgather(tp_ctx* ctx, const void* send_dev, void* recv_dev, size_t bytes) {
  if (ctx->world <= 1) {
    if (recv_dev != send_dev) TP_CUDA_OK(ctx, cudaMemcpyAsync(recv_dev, send_dev, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return TP_OK;
  }
  if (ctx->nccl) {
    ncclResult_t r = nccl_api().AllGather(send_dev, recv_dev, bytes, ncclChar, (ncclComm_t)ctx->nccl, ctx->stream);
    return r == ncclSuccess ? TP_OK : nccl_fail(ctx, "ncclAllGather", r);
  }
  if (ctx->local) {
    LocalGroup* g = ctx->local;
    TP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));   // my send buffer is final, my receive buffer idle
    g->ptr[ctx->rank] = recv_dev;
    g->barrier();
    for (int p = 0; p < ctx->world; p++) {
      char* dst = (char*)g->ptr[p] + (size_t)ctx->rank * bytes;
      if (dst != (const char*)send_dev)   // in place: my own block is already where it belongs
        TP_CUDA_OK(ctx, cudaMemcpyAsync(dst, send_dev, bytes, cudaMemcpyDefault, ctx->stream));
    }
    TP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    g->barrier();
    return TP_OK;
  }
  return fail(ctx, TP_ERR_COLLECTIVE, "sharded context without a communicator (tp_ctx_comm_init_rank)");
}
int comm_bcast(tp_ctx* ctx, void* dev_ptr, size_t bytes, int root) {
  if (ctx->world <= 1) return TP_OK;
  if (ctx->nccl) {
    ncclResult_t r = nccl_api().Broadcast(dev_ptr, dev_ptr, bytes, ncclChar, root, (ncclComm_t)ctx->nccl, ctx->stream);
    return r == ncclSuccess ? TP_OK : nccl_fail(ctx, "ncclBroadcast", r);
  }
  if (ctx->local) {
    LocalGroup* g = ctx->local;
    TP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    g->ptr[ctx->rank] = dev_ptr;
    g->barrier();
    if (ctx->rank != root) {
      TP_CUDA_OK(ctx, cudaMemcpyAsync(dev_ptr, g->ptr[root], bytes, cudaMemcpyDefault, ctx->stream));
      TP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    g->barrier();
    return TP_OK;
  }
  return fail(ctx, TP_ERR_COLLECTIVE, "sharded context without a communicator (tp_ctx_comm_init_rank)");
}

// Several broadcasts with different roots issued as ONE NCCL group run concurrently (every root's links are busy at
// the same time) instead of one after the other.  No-ops on the other transports.
int comm_group_begin(tp_ctx* ctx) {
  if (ctx->world > 1 && ctx->nccl) {
    ncclResult_t r = nccl_api().GroupStart();
    if (r != ncclSuccess) return nccl_fail(ctx, "ncclGroupStart", r);
  }
  return TP_OK;
}
int comm_group_end(tp_ctx* ctx) {
  if (ctx->world > 1 && ctx->nccl) {
    ncclResult_t r = nccl_api().GroupEnd();
    if (r != ncclSuccess) return nccl_fail(ctx, "ncclGroupEnd", r);
  }
  return TP_OK;
}

// ---- worker threads of a device group -------------------------------------------------------------------
struct Workers {
  struct Slot {
    std::thread th;
    std::function<int()> job;
    int rc = 0;
    bool has_job = false, done = false;
  };
  std::mutex m;
  std::condition_variable cv_job, cv_done;
  std::vector<Slot> slots;
  bool stop = false;
  void start(const std::vector<tp_ctx*>& children) {
    slots.resize(children.size());
    for (size_t i = 0; i < children.size(); i++) {
      const int device = children[i]->device;
      slots[i].th = std::thread([this, i, device] {
        cudaSetDevice(device);
        for (;;) {
          std::function<int()> job;
          {
            std::unique_lock<std::mutex> lk(m);
            cv_job.wait(lk, [&] { return stop || slots[i].has_job; });
            if (stop) return;
            job = std::move(slots[i].job);
            slots[i].has_job = false;
          }
          int rc = job();
          {
            std::lock_guard<std::mutex> lk(m);
            slots[i].rc = rc;
            slots[i].done = true;
          }
          cv_done.notify_all();
        }
      });
    }
  }
  // fn(r) on worker r for every r at once; returns the first non-zero status (by rank)
  int run(const std::function<int(int)>& fn, std::vector<int>* all = nullptr) {
    {
      std::lock_guard<std::mutex> lk(m);
      for (size_t i = 0; i < slots.size(); i++) {
        const int r = (int)i;
        slots[i].job = [fn, r] { return fn(r); };
        slots[i].has_job = true;
        slots[i].done = false;
      }
    }
    cv_job.notify_all();
    std::unique_lock<std::mutex> lk(m);
    cv_done.wait(lk, [&] {
      for (auto& s : slots)
        if (!s.done) return false;
      return true;
    });
    int rc = TP_OK;
    for (auto& s : slots) {
      if (all) all->push_back(s.rc);
      if (rc == TP_OK && s.rc != TP_OK) rc = s.rc;
    }
    return rc;
  }
  void shutdown() {
    {
      std::lock_guard<std::mutex> lk(m);
      stop = true;
    }
    cv_job.notify_all();
    for (auto& s : slots)
      if (s.th.joinable()) s.th.join();
  }
};

int group_run(tp_ctx* g, const std::function<int(tp_ctx*, int)>& fn) {
  Workers* w = (Workers*)g->workers;
  std::vector<int> all;
  int rc = w->run([&](int r) { return fn(g->children[r], r); }, &all);
  if (rc != TP_OK) {
    for (size_t r = 0; r < all.size(); r++)
      if (all[r] != TP_OK) {
        g->err = "rank " + std::to_string(r) + ": " + g->children[r]->err;
        break;
      }
  }
  return rc;
}

void comm_release(tp_ctx* ctx) {
  if (ctx->nccl) {
    nccl_api().CommDestroy((ncclComm_t)ctx->nccl);
    ctx->nccl = nullptr;
  }
  ctx->local = nullptr;  // owned by the group front
}

}  // namespace tp

using namespace tp;

extern "C" {

int tp_comm_unique_id(uint8_t out[TP_COMM_ID_BYTES]) {
  if (!out) return TP_ERR_INVALID_ARG;
  static_assert(TP_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "the id crosses the ABI as raw bytes");
  NcclApi& api = nccl_api();
  if (!api.ok) return TP_ERR_COLLECTIVE;
  ncclUniqueId id;
  if (api.GetUniqueId(&id) != ncclSuccess) return TP_ERR_COLLECTIVE;
  memcpy(out, id.internal, TP_COMM_ID_BYTES);
  return TP_OK;
}

int tp_ctx_comm_init_rank(tp_ctx* ctx, int rank, int world, const uint8_t id_bytes[TP_COMM_ID_BYTES]) {
  if (!ctx) return TP_ERR_INVALID_ARG;
  if (!ctx->children.empty()) return fail(ctx, TP_ERR_INVALID_ARG, "comm_init_rank: a device group already owns its communicator");
  if (world < 1 || rank < 0 || rank >= world) return fail(ctx, TP_ERR_INVALID_ARG, "comm_init_rank: bad rank / world");
  comm_release(ctx);
  ctx->rank = rank;
  ctx->world = world;
  if (world == 1) return TP_OK;
  if (!id_bytes) return fail(ctx, TP_ERR_INVALID_ARG, "comm_init_rank: world > 1 needs the unique id of rank 0");
  NcclApi& api = nccl_api();
  if (!api.ok) return fail(ctx, TP_ERR_COLLECTIVE, "comm_init_rank: libnccl.so.2 not found (set TP_NCCL_LIB)");
  ncclUniqueId id;
  memcpy(id.internal, id_bytes, TP_COMM_ID_BYTES);
  TP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  ncclComm_t comm = nullptr;
  ncclResult_t r = api.CommInitRank(&comm, world, id, rank);
  if (r != ncclSuccess) {
    ctx->rank = 0;
    ctx->world = 1;
    return nccl_fail(ctx, "ncclCommInitRank", r);
  }
  ctx->nccl = comm;
  return TP_OK;
}

int tp_ctx_create_multi(const int* devices, int ndev, tp_ctx** out) {
  if (!out) return TP_ERR_INVALID_ARG;
  *out = nullptr;
  if (!devices || ndev < 1 || ndev > TP_MAX_GROUP) return TP_ERR_INVALID_ARG;
  tp_ctx* g = new tp_ctx();
  g->device = devices[0];
  auto destroy_children = [&] {
    for (tp_ctx* c : g->children) tp_ctx_destroy(c);
    g->children.clear();
    delete g;
  };
  bool distinct = true;
  for (int r = 0; r < ndev; r++) {
    tp_ctx* c = nullptr;
    int rc = tp_ctx_create(devices[r], nullptr, &c);
    if (rc != TP_OK) {
      destroy_children();
      return rc;
    }
    c->rank = r;
    c->world = ndev;
    g->children.push_back(c);
    for (int q = 0; q < r; q++) distinct = distinct && devices[q] != devices[r];
  }
  cudaSetDevice(devices[0]);
  const char* force_local = getenv("TP_COMM_LOCAL");
  bool use_nccl = ndev > 1 && distinct && nccl_api().ok && !(force_local && *force_local == '1');
  if (use_nccl) {
    std::vector<ncclComm_t> comms(ndev);
    if (nccl_api().CommInitAll(comms.data(), ndev, devices) == ncclSuccess) {
      for (int r = 0; r < ndev; r++) g->children[r]->nccl = comms[r];
    } else {
      use_nccl = false;
    }
  }
  if (ndev > 1 && !use_nccl) {
    LocalGroup* lg = new LocalGroup();
    lg->world = ndev;
    lg->ptr.assign(ndev, nullptr);
    g->local = lg;
    for (tp_ctx* c : g->children) c->local = lg;
  }
  Workers* w = new Workers();
  w->start(g->children);
  g->workers = w;
  g->world = ndev;
  *out = g;
  return TP_OK;
}

int tp_ctx_group_size(const tp_ctx* ctx, int* ndev, int* uses_nccl) {
  if (!ctx) return TP_ERR_INVALID_ARG;
  if (ndev) *ndev = ctx->children.empty() ? ctx->world : (int)ctx->children.size();
  if (uses_nccl) *uses_nccl = ctx->children.empty() ? (ctx->nccl != nullptr) : (ctx->children[0]->nccl != nullptr);
  return TP_OK;
}

}  // extern "C"

namespace tp {
// called by tp_ctx_destroy on a group front
void group_destroy(tp_ctx* g) {
  if (g->workers) {
    ((Workers*)g->workers)->shutdown();
    delete (Workers*)g->workers;
    g->workers = nullptr;
  }
  LocalGroup* lg = g->local;
  for (tp_ctx* c : g->children) tp_ctx_destroy(c);
  g->children.clear();
  delete lg;
  g->local = nullptr;
}
}  // namespace tp
