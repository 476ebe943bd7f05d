// Shared helpers for libsafevla_b200 (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/safevla_b200.h"

struct svla_ctx {
  int device;
  int sm_count;
  float* partials;       // [kMaxPartialBlocks * 16] deterministic two-stage reductions
  unsigned int* tickets; // [16] last-block tickets (self-resetting)
  void* tmap_cache;      // tensor-map cache (gemm_tc.cu)
  void* ws;              // scratch: split-K slices, per-block partials of column reductions
  size_t ws_bytes;
};

constexpr int kMaxPartialBlocks = 2048;

void svla_set_error(const char* fmt, ...);

#define SVLA_CHECK_ARG(cond, msg)                                   \
  do {                                                              \
    if (!(cond)) {                                                  \
      svla_set_error("%s:%d: %s", __FILE__, __LINE__, msg);         \
      return SVLA_ERR_BAD_ARG;                                      \
    }                                                               \
  } while (0)

#define SVLA_CUDA(call)                                                              \
  do {                                                                               \
    cudaError_t _e = (call);                                                         \
    if (_e != cudaSuccess) {                                                         \
      svla_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return (int)_e;                                                                \
    }                                                                                \
  } while (0)

// one call per kernel launch: checks the launch and counts it (svla_launch_count, bench.py's gpu_launches)
extern unsigned long long g_svla_launches;
#define SVLA_LAUNCH_CHECK()            \
  do {                                 \
    ++g_svla_launches;                 \
    SVLA_CUDA(cudaGetLastError());     \
  } while (0)

static inline cudaStream_t as_stream(svla_stream s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- dropout spec -> device arguments (csrc/philox.cuh) ---------------------------------------
#include "philox.cuh"
static inline bool svla_dropout_ok(const svla_dropout* d) { return !d || (d->p >= 0.f && d->p < 1.f); }
static inline DropArgs make_drop_args(const svla_dropout* d) {
  DropArgs a{};
  if (d && d->p > 0.f) {
    a.thr = (uint32_t)(d->p * 65536.0 + 0.5);
    a.scale = 1.f / (1.f - d->p);
    a.key0 = (uint32_t)(d->seed & 0xFFFFFFFFull);
    a.key1 = (uint32_t)(d->seed >> 32);
    a.site = d->site;
    a.step = d->step;
  } else {
    a.scale = 1.f;
  }
  a.row0 = d ? d->row0 : 0u;
  a.row_stride = (d && d->row_stride) ? d->row_stride : 1u;
  return a;
}

// ---- dtype helpers -----------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 4 consecutive elements <-> float4 (16 B for fp32, 8 B for bf16); pointers must be so aligned
template <typename T> __device__ __forceinline__ float4 load4(const T* p);
template <> __device__ __forceinline__ float4 load4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <> __device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <typename T> __device__ __forceinline__ void store4(T* p, float4 v);
template <> __device__ __forceinline__ void store4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// exact (erf) GELU, nn.GELU() default -- the DINOv2 MLP activation
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum (all threads get the result); `red` is >= 32 floats of shared memory
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

__device__ __forceinline__ long long map_row(svla_rowmap m, long long r) {
  if (m.group <= 0) return r;
  return (r / m.group) * (long long)m.group_stride + m.group_offset + (r % m.group);
}

// dispatch a templated launcher on svla_dtype
#define SVLA_DISPATCH_DTYPE(dt, T, ...)                                  \
  do {                                                                   \
    if ((dt) == SVLA_F32) { using T = float; __VA_ARGS__; }              \
    else if ((dt) == SVLA_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
    else { svla_set_error("bad dtype %d", (int)(dt)); return SVLA_ERR_BAD_ARG; } \
  } while (0)
