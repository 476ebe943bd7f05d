// Context, error reporting.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";
unsigned long long g_svla_launches = 0;

void svla_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void svla_tmap_cache_free(void* cache);  // gemm_tc.cu

extern "C" {

const char* svla_last_error(void) { return g_err; }
int svla_version(void) { return 100; }

int svla_ctx_create(int device, svla_ctx** out) {
  SVLA_CHECK_ARG(out != nullptr, "out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    svla_set_error("no CUDA device visible (%s): this library has no CPU fallback", cudaGetErrorString(e));
    return SVLA_ERR_NO_DEVICE;
  }
  SVLA_CHECK_ARG(device >= 0 && device < n, "bad device ordinal");
  cudaDeviceProp prop;
  SVLA_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    svla_set_error("device %d is sm_%d%d; libsafevla_b200 is built for sm_100a only", device, prop.major,
                   prop.minor);
    return SVLA_ERR_UNSUPPORTED_ARCH;
  }
  int prev = 0;
  SVLA_CUDA(cudaGetDevice(&prev));
  SVLA_CUDA(cudaSetDevice(device));
  svla_ctx* c = new svla_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->tmap_cache = nullptr;
  SVLA_CUDA(cudaMalloc(&c->partials, sizeof(float) * kMaxPartialBlocks * 16));
  SVLA_CUDA(cudaMalloc(&c->tickets, sizeof(unsigned int) * 16));
  SVLA_CUDA(cudaMemset(c->tickets, 0, sizeof(unsigned int) * 16));
  c->ws_bytes = (size_t)128 << 20;
  SVLA_CUDA(cudaMalloc(&c->ws, c->ws_bytes));
  SVLA_CUDA(cudaSetDevice(prev));
  *out = c;
  return SVLA_OK;
}

int svla_ctx_destroy(svla_ctx* ctx) {
  if (!ctx) return SVLA_OK;
  cudaFree(ctx->partials);
  cudaFree(ctx->tickets);
  cudaFree(ctx->ws);
  svla_tmap_cache_free(ctx->tmap_cache);
  delete ctx;
  return SVLA_OK;
}

int svla_sm_count(svla_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
unsigned long long svla_launch_count(void) { return g_svla_launches; }
int svla_launch_count_add(unsigned long long n) {  // launches replayed from a CUDA graph captured over this library's calls
  g_svla_launches += n;
  return SVLA_OK;
}

}  // extern "C"
