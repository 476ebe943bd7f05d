// tcgen05 tensor-core GEMM (placeholder until the TMA/TMEM kernel lands in this file).
#include "common.cuh"

void svla_tmap_cache_free(void* cache) { (void)cache; }
bool svla_gemm_tc_supported(const svla_gemm_desc* d) { (void)d; return false; }
int svla_gemm_tc(svla_ctx* ctx, const svla_gemm_desc* d, cudaStream_t st) {
  (void)ctx; (void)d; (void)st;
  svla_set_error("tcgen05 GEMM not built");
  return SVLA_ERR_INTERNAL;
}
