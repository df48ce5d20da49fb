// k2_tensor.cu — K2: batched scoring on the tcgen05 tensor cores (placeholder until the
// kernel lands; the API reports the path as unavailable rather than falling back).
#include "common.cuh"

int k2_available(const rag_index* idx) { (void)idx; return 0; }
int k2_plan(rag_index* idx, uint32_t B, uint32_t kp, uint32_t* parts) {
  (void)idx; (void)B; (void)kp; (void)parts;
  return rag_set_error(RAG_ERR_UNSUPPORTED, "tensor path (K2) is not built into this library");
}
int k2_launch(rag_index* idx, uint32_t B, uint32_t kp, uint32_t parts) {
  (void)idx; (void)B; (void)kp; (void)parts;
  return rag_set_error(RAG_ERR_UNSUPPORTED, "tensor path (K2) is not built into this library");
}
void k2_destroy(rag_index* idx) { (void)idx; }
