// comm.cu — C1: exchange of the ranks' exact local top-k lists.
//
// The corpus is row-sharded, one process per GPU. After K4 each rank holds
// [B][k] rag_rec (48 B each, exact fp64 scores, global chunk ids) and every rank needs all of them
// for K5. Payload is B*k*48 B per rank (480 B at batch 1, 480 KB at B=1024, k=10): pure latency.
//
// Default path — peer-to-peer mailboxes, fused into K5 (k5_body.cuh::p2p_exchange): every rank owns a mailbox
// (cudaMalloc, exported with cudaIpcGetMemHandle, mapped by the peers); K5's warp for query b stores its
// records into all mailboxes (over NVLink between GPUs), fences, raises flag[src][b] in each, spins on its own
// flags and merges. Two parity halves make the mailbox safe to refill while a slower peer still reads the
// previous exchange (a rank cannot be two exchanges ahead: exchange s+1 completes only after every peer
// has STARTED s+1, i.e. finished reading s). No collective call and no extra launch on the hot path.
//
// Bootstrap: HOST-DRIVEN (rag_comm_p2p_export / rag_comm_p2p_import — the host carries the 64-byte IPC handles,
// no NCCL anywhere), or through NCCL (rag_comm_init: the handles ride an ncclAllGather, and the mailboxes may
// grow collectively). Fallback exchange — one ncclAllGather on the compute stream (RAGERA_COMM=nccl, or peers
// without P2P access). NCCL is bound at run time with dlopen so that libragera.so has no link-time NCCL
// dependency; single-GPU users and host-bootstrapped ranks never load it.
//
// A peer that never arrives does not hang or trap the device: the waiting warp gives up after
// RAGERA_P2P_TIMEOUT_MS, raises a host-mapped status word, the call returns RAG_ERR_TIMEOUT and the communicator
// is marked broken (the ranks' exchange counters can no longer be trusted) until it is bootstrapped again.
#include "common.cuh"

#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0 };

struct nccl_api {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

nccl_api g_nccl;

int load_nccl() {
  if (g_nccl.handle) return RAG_OK;
  // RTLD_NOLOAD first: if the host process (e.g. torch) already mapped libnccl, share it
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return rag_set_error(RAG_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                            \
  *(void**)(&g_nccl.field) = dlsym(h, name);                                        \
  if (!g_nccl.field) return rag_set_error(RAG_ERR_NCCL, "libnccl lacks %s", name);
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllGather, "ncclAllGather")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  g_nccl.handle = h;
  return RAG_OK;
}
}  // namespace

struct rag_comm {
  ncclComm_t comm = nullptr;    // null for a host-bootstrapped communicator
  // peer-to-peer mailboxes
  bool want_p2p = true;         // RAGERA_COMM=nccl turns it off
  bool p2p_failed = false;      // IPC mapping was refused once: stay on NCCL
  bool host_driven = false;     // rag_comm_p2p_export / _import: fixed capacity, no collective growth
  bool imported = false;
  bool broken = false;          // an exchange timed out: counters of the ranks may disagree
  unsigned char* mbox = nullptr;
  size_t half_bytes = 0, flags_off = 0;
  uint32_t flag_stride = 0;
  unsigned char* peer[8] = {nullptr};
  uint32_t step = 0;
  long long timeout_cycles = 0;
  uint32_t* h_status = nullptr;   // pinned + mapped: the kernels raise it, the host reads it after a sync
  uint32_t* d_status = nullptr;   // its device alias
  unsigned char* d_hs = nullptr;  // staging for the handle all-gather (NCCL bootstrap)
  unsigned char* d_hr = nullptr;
};

namespace {

void p2p_unmap(rag_index* idx) {
  rag_comm* c = idx->comm;
  for (int g = 0; g < idx->nranks && g < 8; g++) {
    if (g != idx->rank && c->peer[g]) cudaIpcCloseMemHandle(c->peer[g]);
    c->peer[g] = nullptr;
  }
}

// diagnostics: RAGERA_P2P_STEP0 starts the exchange counter there (every rank alike) so that a test can walk it across
// the 20-bit wrap in a few searches; even values only, so that the first exchange's parity is what a fresh counter gives
uint32_t initial_step() {
  const char* e = getenv("RAGERA_P2P_STEP0");
  if (!e) return 0;
  const uint32_t v = (uint32_t)strtoul(e, nullptr, 0);
  return v <= 0xFFFFEu ? (v & ~1u) : 0;
}

int comm_common_init(rag_comm* c) {
  double ms = 20000.0;
  if (const char* e = getenv("RAGERA_P2P_TIMEOUT_MS")) ms = atof(e) > 0 ? atof(e) : ms;
  c->timeout_cycles = (long long)(ms * 2.0e6);  // clock64 runs at <= ~2 GHz: the wait is at least `ms`
  RAG_CUDA(cudaHostAlloc((void**)&c->h_status, 64, cudaHostAllocMapped));
  *c->h_status = 0;
  RAG_CUDA(cudaHostGetDevicePointer((void**)&c->d_status, c->h_status, 0));
  return RAG_OK;
}

// carve one mailbox for (nranks, B, k): [records | flags] x 2 parity halves, zeroed (flag 0 is never awaited)
int mbox_alloc(rag_index* idx, size_t rec_cap, uint32_t stride) {
  rag_comm* c = idx->comm;
  if (c->mbox) RAG_CUDA(cudaFree(c->mbox));
  c->mbox = nullptr;
  c->flag_stride = stride;
  c->flags_off = rec_cap;
  c->half_bytes = rec_cap + (size_t)idx->nranks * stride * sizeof(uint32_t);
  RAG_CUDA(cudaMalloc((void**)&c->mbox, 2 * c->half_bytes));
  RAG_CUDA(cudaMemsetAsync(c->mbox, 0, 2 * c->half_bytes, idx->stream));
  RAG_CUDA(cudaStreamSynchronize(idx->stream));
  return RAG_OK;
}

// all ranks: allgather `bytes` from d_hs into d_hr on the compute stream and wait (doubles as a barrier)
int p2p_allgather_sync(rag_index* idx, size_t bytes) {
  rag_comm* c = idx->comm;
  ncclResult_t r = g_nccl.AllGather(c->d_hs, c->d_hr, bytes, ncclInt8, c->comm, idx->stream);
  if (r != 0) return rag_set_error(RAG_ERR_NCCL, "ncclAllGather (mailbox bootstrap): %s", g_nccl.GetErrorString(r));
  RAG_CUDA(cudaStreamSynchronize(idx->stream));
  return RAG_OK;
}

void comm_free(rag_index* idx) {
  rag_comm* c = idx->comm;
  if (!c) return;
  p2p_unmap(idx);
  if (c->mbox) cudaFree(c->mbox);
  if (c->d_hs) cudaFree(c->d_hs);
  if (c->d_hr) cudaFree(c->d_hr);
  if (c->h_status) cudaFreeHost(c->h_status);
  if (c->comm) g_nccl.CommDestroy(c->comm);
  delete c;
  idx->comm = nullptr;
  cudaGetLastError();
}
}  // namespace

bool comm_uses_p2p(const rag_index* idx) {
  const rag_comm* c = idx->comm;
  return idx->nranks > 1 && c && c->want_p2p && !c->p2p_failed && (!c->host_driven || c->imported);
}

// Every rank calls it with the same (B, k). NCCL-bootstrapped: collective — grows the mailboxes when one parity half
// cannot hold nranks*B*k records, then re-exchanges the IPC handles. Host-bootstrapped: the capacity is what
// rag_comm_p2p_export was given; a larger shape is an error.
int comm_p2p_ensure(rag_index* idx, uint32_t B, uint32_t k) {
  rag_comm* c = idx->comm;
  if (idx->nranks > 1 && c && c->broken)
    return rag_set_error(RAG_ERR_STATE, "sharded search: an earlier exchange timed out; bootstrap the communicator again on all ranks");
  if (idx->nranks > 1 && c && c->host_driven && !c->imported)
    return rag_set_error(RAG_ERR_STATE, "sharded search: rag_comm_p2p_import has not been called");
  if (!comm_uses_p2p(idx)) return RAG_OK;
  const size_t need_rec = (((size_t)idx->nranks * B * k * sizeof(rag_rec)) + 255) & ~(size_t)255;
  if (c->mbox && need_rec <= c->flags_off && B <= c->flag_stride) return RAG_OK;
  if (c->host_driven)
    return rag_set_error(RAG_ERR_STATE, "sharded search: batch %u x k %u exceeds the mailbox capacity given to rag_comm_p2p_export", B, k);
  const int G = idx->nranks;
  if (!c->d_hs) {
    RAG_CUDA(cudaMalloc((void**)&c->d_hs, sizeof(cudaIpcMemHandle_t)));
    RAG_CUDA(cudaMalloc((void**)&c->d_hr, sizeof(cudaIpcMemHandle_t) * 8));
  }
  // 1. nobody may still be inside an exchange that uses the old mailboxes; unmap, then barrier
  RAG_CUDA(cudaStreamSynchronize(idx->stream));
  p2p_unmap(idx);
  RAG_CHECK(p2p_allgather_sync(idx, 16));
  // 2. new mailbox (zeroed), handle out
  size_t rec_cap = need_rec + need_rec / 2;
  if (rec_cap < ((size_t)1 << 20)) rec_cap = (size_t)1 << 20;
  if (rec_cap < c->flags_off) rec_cap = c->flags_off;  // never shrink
  uint32_t stride = c->flag_stride > 4096u ? c->flag_stride : 4096u;
  while (stride < B) stride *= 2;
  RAG_CHECK(mbox_alloc(idx, rec_cap, stride));
  cudaIpcMemHandle_t mine;
  RAG_CUDA(cudaIpcGetMemHandle(&mine, c->mbox));
  RAG_CUDA(cudaMemcpyAsync(c->d_hs, &mine, sizeof(mine), cudaMemcpyHostToDevice, idx->stream));
  // 3. all-gather the handles (every rank has zeroed its mailbox before it contributes), map the peers
  RAG_CHECK(p2p_allgather_sync(idx, sizeof(mine)));
  cudaIpcMemHandle_t all[8];
  RAG_CUDA(cudaMemcpy(all, c->d_hr, sizeof(mine) * G, cudaMemcpyDeviceToHost));
  int failed = 0;
  for (int g = 0; g < G; g++) {
    if (g == idx->rank) { c->peer[g] = c->mbox; continue; }
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, all[g], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); failed = 1; break; }
    c->peer[g] = (unsigned char*)p;
  }
  // 4. agree on the outcome: one rank without peer access sends everybody to the NCCL path
  unsigned char flag[16] = {(unsigned char)failed};
  RAG_CUDA(cudaMemcpyAsync(c->d_hs, flag, 16, cudaMemcpyHostToDevice, idx->stream));
  RAG_CHECK(p2p_allgather_sync(idx, 16));
  unsigned char flags[16 * 8];
  RAG_CUDA(cudaMemcpy(flags, c->d_hr, (size_t)16 * G, cudaMemcpyDeviceToHost));
  for (int g = 0; g < G; g++) failed |= flags[16 * g];
  if (failed) {
    p2p_unmap(idx);
    c->p2p_failed = true;
  }
  c->step = initial_step();  // fresh, zeroed mailboxes on every rank
  return RAG_OK;
}

int comm_p2p_next(rag_index* idx, uint32_t B, uint32_t k, rag_p2p_view* v) {
  *v = rag_p2p_view();
  v->nranks = 1;
  if (!comm_uses_p2p(idx)) return RAG_OK;
  rag_comm* c = idx->comm;
  const size_t need_rec = (size_t)idx->nranks * B * k * sizeof(rag_rec);
  if (!c->mbox || need_rec > c->flags_off || B > c->flag_stride) return rag_set_error(RAG_ERR_STATE, "sharded search: mailboxes not sized (comm_p2p_ensure)");
  for (int g = 0; g < idx->nranks; g++) v->base[g] = c->peer[g];
  v->nranks = (uint32_t)idx->nranks;
  v->rank = (uint32_t)idx->rank;
  c->step = rag_p2p_next_step(c->step);  // never 0, parity alternates across the wrap (common.cuh)
  v->step = c->step;
  v->flag = (c->step << 12) | ((B * 131u + k) & 0xFFFu);
  v->half_bytes = c->half_bytes;
  v->flags_off = c->flags_off;
  v->flag_stride = c->flag_stride;
  v->timeout_cycles = c->timeout_cycles;
  v->status = c->d_status;
  return RAG_OK;
}

// after the stream has been synchronised: did any query of an exchange give up waiting for a peer?
int comm_check_status(rag_index* idx) {
  rag_comm* c = idx->comm;
  if (!c || !c->h_status || *(volatile uint32_t*)c->h_status == 0) return RAG_OK;
  *c->h_status = 0;
  c->broken = true;
  return rag_set_error(RAG_ERR_TIMEOUT, "sharded search: a peer rank did not arrive at the exchange within the timeout "
                       "(ranks must issue the same sequence of calls); the communicator is disabled until it is bootstrapped again");
}

extern "C" int rag_comm_p2p_export(rag_index* idx, int nranks, int rank, uint32_t max_batch, uint32_t max_k,
                                   uint8_t handle[RAG_COMM_HANDLE_BYTES]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == RAG_COMM_HANDLE_BYTES, "IPC handle size");
  if (!idx || !handle || nranks < 1 || rank < 0 || rank >= nranks || nranks > 8)
    return rag_set_error(RAG_ERR_INVALID, "rag_comm_p2p_export: bad nranks/rank (1..8 ranks supported)");
  RAG_LOCK(idx);
  if (max_batch == 0 || max_batch > 4096 || max_k == 0 || max_k > RAG_MAX_TOPK)
    return rag_set_error(RAG_ERR_INVALID, "rag_comm_p2p_export: max_batch must be 1..4096, max_k 1..%d", RAG_MAX_TOPK);
  RAG_CUDA(cudaSetDevice(idx->device));
  rag_comm_destroy(idx);
  memset(handle, 0, RAG_COMM_HANDLE_BYTES);
  if (nranks == 1) return RAG_OK;
  rag_comm* c = new rag_comm();
  c->host_driven = true;
  idx->comm = c;
  idx->nranks = nranks;
  idx->rank = rank;
  int rc = comm_common_init(c);
  if (rc == RAG_OK) {
    const size_t rec = (((size_t)nranks * max_batch * max_k * sizeof(rag_rec)) + 255) & ~(size_t)255;
    rc = mbox_alloc(idx, rec, (max_batch + 31u) & ~31u);
  }
  if (rc == RAG_OK) {
    cudaIpcMemHandle_t mine;
    cudaError_t e = cudaIpcGetMemHandle(&mine, c->mbox);
    if (e != cudaSuccess) rc = rag_set_error(RAG_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    else memcpy(handle, &mine, sizeof(mine));
  }
  if (rc != RAG_OK) rag_comm_destroy(idx);
  return rc;
}

extern "C" int rag_comm_p2p_import(rag_index* idx, const uint8_t* handles) {
  if (!idx || !handles) return rag_set_error(RAG_ERR_INVALID, "rag_comm_p2p_import: null argument");
  RAG_LOCK(idx);
  if (idx->nranks == 1 && !idx->comm) return RAG_OK;
  rag_comm* c = idx->comm;
  if (!c || !c->host_driven || !c->mbox) return rag_set_error(RAG_ERR_STATE, "rag_comm_p2p_import without rag_comm_p2p_export");
  RAG_CUDA(cudaSetDevice(idx->device));
  p2p_unmap(idx);
  for (int g = 0; g < idx->nranks; g++) {
    if (g == idx->rank) { c->peer[g] = c->mbox; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)g * RAG_COMM_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      p2p_unmap(idx);
      return rag_set_error(RAG_ERR_UNSUPPORTED, "cannot map the mailbox of rank %d (%s): no peer access — use the NCCL exchange on all ranks",
                           g, cudaGetErrorString(e));
    }
    c->peer[g] = (unsigned char*)p;
  }
  c->imported = true;
  c->step = initial_step();
  return RAG_OK;
}

// Two-phase teardown: CUDA leaves freeing an exported allocation while an importer still maps it undefined, so
// every rank first closes its mappings of the peers' mailboxes (this call), the host runs a barrier, and only then
// may any rank free its own mailbox (rag_comm_destroy, rag_index_destroy, a new rag_comm_p2p_export).
extern "C" int rag_comm_detach(rag_index* idx) {
  if (!idx) return rag_set_error(RAG_ERR_INVALID, "null index handle");
  RAG_LOCK(idx);
  if (!idx->comm) return RAG_OK;
  RAG_CUDA(cudaSetDevice(idx->device));
  RAG_CUDA(cudaStreamSynchronize(idx->stream));
  p2p_unmap(idx);
  idx->comm->imported = false;
  if (!idx->comm->host_driven) idx->comm->p2p_failed = true;  // NCCL-bootstrapped: stay on the all-gather from here on
  return RAG_OK;
}

extern "C" int rag_comm_unique_id(uint8_t id[RAG_COMM_ID_BYTES]) {
  RAG_CHECK(load_nccl());
  ncclUniqueId u;
  ncclResult_t r = g_nccl.GetUniqueId(&u);
  if (r != 0) return rag_set_error(RAG_ERR_NCCL, "ncclGetUniqueId: %s", g_nccl.GetErrorString(r));
  memcpy(id, u.internal, RAG_COMM_ID_BYTES);
  return RAG_OK;
}

extern "C" int rag_comm_init(rag_index* idx, int nranks, int rank, const uint8_t id[RAG_COMM_ID_BYTES]) {
  if (!idx || nranks < 1 || rank < 0 || rank >= nranks || nranks > 8)
    return rag_set_error(RAG_ERR_INVALID, "rag_comm_init: bad nranks/rank (1..8 ranks supported)");
  RAG_LOCK(idx);
  rag_comm_destroy(idx);
  if (nranks == 1) return RAG_OK;
  RAG_CHECK(load_nccl());
  RAG_CUDA(cudaSetDevice(idx->device));
  ncclUniqueId u;
  memcpy(u.internal, id, RAG_COMM_ID_BYTES);
  rag_comm* c = new rag_comm();
  ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, u, rank);
  if (r != 0) {
    delete c;
    return rag_set_error(RAG_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
  }
  if (const char* e = getenv("RAGERA_COMM")) c->want_p2p = strcmp(e, "nccl") != 0;
  idx->comm = c;
  idx->nranks = nranks;
  idx->rank = rank;
  const int rc = comm_common_init(c);
  if (rc != RAG_OK) rag_comm_destroy(idx);
  return rc;
}

extern "C" uint32_t rag_debug_p2p_next_step(uint32_t step) { return rag_p2p_next_step(step); }

extern "C" int rag_comm_destroy(rag_index* idx) {
  if (!idx) return RAG_OK;
  RAG_LOCK(idx);
  if (idx && idx->comm) {
    cudaSetDevice(idx->device);
    cudaStreamSynchronize(idx->stream);
    comm_free(idx);
  }
  if (idx) { idx->nranks = 1; idx->rank = 0; }
  return RAG_OK;
}

int comm_allgather_local(rag_index* idx, uint32_t B, uint32_t k) {
  if (idx->nranks <= 1 || comm_uses_p2p(idx)) return RAG_OK;  // peer-to-peer: K5 does the exchange itself
  rag_prof_scope ps(idx, RAG_PROF_COMM);
  if (!idx->comm || !idx->comm->comm)
    return rag_set_error(RAG_ERR_STATE, "sharded search without a communicator (rag_comm_p2p_export/import or rag_comm_init)");
  const size_t bytes = (size_t)B * k * sizeof(rag_rec);
  ncclResult_t r = g_nccl.AllGather(idx->cur->d_local, idx->cur->d_gather, bytes, ncclInt8, idx->comm->comm, idx->stream);
  if (r != 0) return rag_set_error(RAG_ERR_NCCL, "ncclAllGather: %s", g_nccl.GetErrorString(r));
  idx->launches++;
  return RAG_OK;
}
