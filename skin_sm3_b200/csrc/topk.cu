// N1: similarity top-k retrieval (placeholder until the kernel lands; see include/sm3_b200.h).
#include "common.cuh"
extern "C" int sm3_sim_topk(const void* query, const void* bank, int64_t n_query, int64_t n_bank, int D, int dtype,
                            int k, int64_t exclude_self_offset, float* vals, int64_t* idx, void* stream) {
  (void)query; (void)bank; (void)n_query; (void)n_bank; (void)D; (void)dtype; (void)k; (void)exclude_self_offset;
  (void)vals; (void)idx; (void)stream;
  sm3::set_error("sim_topk: not implemented yet");
  return SM3_ERR_DTYPE;
}
