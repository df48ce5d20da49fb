// comm.cu — C1: exchange of the ranks' exact local top-k lists.
//
// The corpus is row-sharded, one process per GPU. After K4 each rank holds
// [B][k] rag_rec (48 B each, exact fp64 scores, global chunk ids). One
// ncclAllGather over NVLink/NVSwitch gives every rank all lists; K5 then merges them.
// Payload is B*k*48 B per rank (480 KB at B=1024, k=10): latency-bound, so it is
// issued on the compute stream with no staging copy (K4 writes the send buffer, K5 reads
// the receive buffer).
//
// NCCL is bound at run time with dlopen so that libragera.so has no link-time NCCL
// dependency; single-GPU users never load it.
#include "common.cuh"

#include <dlfcn.h>
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
};

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
  idx->comm = c;
  idx->nranks = nranks;
  idx->rank = rank;
  return RAG_OK;
}

extern "C" int rag_comm_destroy(rag_index* idx) {
  if (idx && idx->comm) {
    if (idx->comm->comm) g_nccl.CommDestroy(idx->comm->comm);
    delete idx->comm;
    idx->comm = nullptr;
  }
  if (idx) { idx->nranks = 1; idx->rank = 0; }
  return RAG_OK;
}

int comm_allgather_local(rag_index* idx, uint32_t B, uint32_t k) {
  if (idx->nranks <= 1) return RAG_OK;
  rag_prof_scope ps(idx, RAG_PROF_COMM);
  if (!idx->comm) return rag_set_error(RAG_ERR_STATE, "sharded search without rag_comm_init");
  const size_t bytes = (size_t)B * k * sizeof(rag_rec);
  ncclResult_t r = g_nccl.AllGather(idx->cur->d_local, idx->cur->d_gather, bytes, ncclInt8, idx->comm->comm, idx->stream);
  if (r != 0) return rag_set_error(RAG_ERR_NCCL, "ncclAllGather: %s", g_nccl.GetErrorString(r));
  idx->launches++;
  return RAG_OK;
}
