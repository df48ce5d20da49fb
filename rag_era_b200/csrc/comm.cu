// comm.cu — C1: exchange of the ranks' exact local top-k lists.
//
// The corpus is row-sharded, one process per GPU. After K4 each rank holds
// [B][k] rag_rec (48 B each, exact fp64 scores, global chunk ids) and every rank needs all of them
// for K5. Payload is B*k*48 B per rank (480 B at batch 1, 480 KB at B=1024, k=10): pure latency.
//
// Default path — peer-to-peer mailboxes, fused into K5 (k5_fuse.cu): every rank owns a mailbox
// (cudaMalloc, exported with cudaIpcGetMemHandle, mapped by the peers); K5's warp for query b stores its
// records into all mailboxes over NVLink, fences, raises flag[src][b] = step in each, spins on its own
// flags and merges. Two parity halves make the mailbox safe to refill while a slower peer still reads the
// previous exchange (a rank cannot be two exchanges ahead: exchange s+1 completes only after every peer
// has STARTED s+1, i.e. finished reading s). No collective call and no extra launch on the hot path.
// Fallback — one ncclAllGather on the compute stream (RAGERA_COMM=nccl, or peers without P2P access).
//
// NCCL is bound at run time with dlopen so that libragera.so has no link-time NCCL
// dependency; single-GPU users never load it. It bootstraps the mailbox handles as well.
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
  ncclComm_t comm = nullptr;
  // peer-to-peer mailboxes
  bool want_p2p = true;         // RAGERA_COMM=nccl turns it off
  bool p2p_failed = false;      // IPC mapping was refused once: stay on NCCL
  unsigned char* mbox = nullptr;
  size_t half_bytes = 0, flags_off = 0;
  uint32_t flag_stride = 0;
  unsigned char* peer[8] = {nullptr};
  uint32_t step = 0;
  unsigned char* d_hs = nullptr;  // staging for the handle all-gather
  unsigned char* d_hr = nullptr;
};

namespace {

void p2p_unmap(rag_index* idx) {
  rag_comm* c = idx->comm;
  for (int g = 0; g < idx->nranks; g++) {
    if (g != idx->rank && c->peer[g]) cudaIpcCloseMemHandle(c->peer[g]);
    c->peer[g] = nullptr;
  }
}

// all ranks: allgather `bytes` from d_hs into d_hr on the compute stream and wait (doubles as a barrier)
int p2p_allgather_sync(rag_index* idx, size_t bytes) {
  rag_comm* c = idx->comm;
  ncclResult_t r = g_nccl.AllGather(c->d_hs, c->d_hr, bytes, ncclInt8, c->comm, idx->stream);
  if (r != 0) return rag_set_error(RAG_ERR_NCCL, "ncclAllGather (mailbox bootstrap): %s", g_nccl.GetErrorString(r));
  RAG_CUDA(cudaStreamSynchronize(idx->stream));
  return RAG_OK;
}
}  // namespace

bool comm_uses_p2p(const rag_index* idx) {
  return idx->nranks > 1 && idx->comm && idx->comm->want_p2p && !idx->comm->p2p_failed;
}

// Collective: every rank calls it with the same (B, k). Grows the mailboxes when one parity half cannot hold
// nranks*B*k records, then re-exchanges the IPC handles.
int comm_p2p_ensure(rag_index* idx, uint32_t B, uint32_t k) {
  if (!comm_uses_p2p(idx)) return RAG_OK;
  rag_comm* c = idx->comm;
  const size_t need_rec = (((size_t)idx->nranks * B * k * sizeof(rag_rec)) + 255) & ~(size_t)255;
  if (c->mbox && need_rec <= c->flags_off && B <= c->flag_stride) return RAG_OK;
  const int G = idx->nranks;
  if (!c->d_hs) {
    RAG_CUDA(cudaMalloc((void**)&c->d_hs, sizeof(cudaIpcMemHandle_t)));
    RAG_CUDA(cudaMalloc((void**)&c->d_hr, sizeof(cudaIpcMemHandle_t) * 8));
  }
  // 1. nobody may still be inside an exchange that uses the old mailboxes; unmap, then barrier
  RAG_CUDA(cudaStreamSynchronize(idx->stream));
  p2p_unmap(idx);
  RAG_CHECK(p2p_allgather_sync(idx, 16));
  // 2. new mailbox (zeroed: flag 0 never equals a step), handle out
  if (c->mbox) RAG_CUDA(cudaFree(c->mbox));
  c->mbox = nullptr;
  size_t rec_cap = need_rec + need_rec / 2;
  if (rec_cap < ((size_t)1 << 20)) rec_cap = (size_t)1 << 20;
  if (rec_cap < c->flags_off) rec_cap = c->flags_off;  // never shrink
  uint32_t stride = c->flag_stride > 4096u ? c->flag_stride : 4096u;
  while (stride < B) stride *= 2;
  c->flag_stride = stride;
  c->flags_off = rec_cap;
  c->half_bytes = rec_cap + (size_t)G * stride * sizeof(uint32_t);
  RAG_CUDA(cudaMalloc((void**)&c->mbox, 2 * c->half_bytes));
  RAG_CUDA(cudaMemsetAsync(c->mbox, 0, 2 * c->half_bytes, idx->stream));
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
  v->step = ++c->step;
  if (v->step == 0) v->step = ++c->step;  // 0 is the "never written" flag value
  v->half_bytes = c->half_bytes;
  v->flags_off = c->flags_off;
  v->flag_stride = c->flag_stride;
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
  if (nranks == 1) { idx->nranks = 1; idx->rank = 0; return RAG_OK; }
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
  return RAG_OK;
}

extern "C" int rag_comm_destroy(rag_index* idx) {
  if (idx && idx->comm) {
    cudaSetDevice(idx->device);
    cudaStreamSynchronize(idx->stream);
    p2p_unmap(idx);
    if (idx->comm->mbox) cudaFree(idx->comm->mbox);
    if (idx->comm->d_hs) cudaFree(idx->comm->d_hs);
    if (idx->comm->d_hr) cudaFree(idx->comm->d_hr);
    if (idx->comm->comm) g_nccl.CommDestroy(idx->comm->comm);
    delete idx->comm;
    idx->comm = nullptr;
  }
  if (idx) { idx->nranks = 1; idx->rank = 0; }
  return RAG_OK;
}

int comm_allgather_local(rag_index* idx, uint32_t B, uint32_t k) {
  if (idx->nranks <= 1 || comm_uses_p2p(idx)) return RAG_OK;  // peer-to-peer: K5 does the exchange itself
  rag_prof_scope ps(idx, RAG_PROF_COMM);
  if (!idx->comm) return rag_set_error(RAG_ERR_STATE, "sharded search without rag_comm_init");
  const size_t bytes = (size_t)B * k * sizeof(rag_rec);
  ncclResult_t r = g_nccl.AllGather(idx->cur->d_local, idx->cur->d_gather, bytes, ncclInt8, idx->comm->comm, idx->stream);
  if (r != 0) return rag_set_error(RAG_ERR_NCCL, "ncclAllGather: %s", g_nccl.GetErrorString(r));
  idx->launches++;
  return RAG_OK;
}
