// Single-step decoder attention against the KV cache (rollout-side T = 1 inference; reference
// training/online/third_party_models/llama/model.py:224-247,279-317 with the episode-start mask of
// allenact_dino_transformer.py:386-397): the step's query attends to cache positions
// [max(pos - time_step[n], 0), pos] of its sampler, pos = the model's time_step_counter.
//
// One warp per (sampler, head).  A cached K / V row of the head (64 elements) is read by eight lanes (16 bytes each
// for bf16, 32 for fp32), so a warp instruction covers four cache positions; every row group keeps an online
// softmax (running max / sum / weighted V slice) and the four groups are merged at the end.  Memory bound: the
// cache slice is read exactly once.
#include "common.cuh"

namespace {

constexpr int DH = 64;

template <typename T> __device__ __forceinline__ void load8(const T* p, float* f);
template <> __device__ __forceinline__ void load8<float>(const float* p, float* f) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float* f) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 t = __bfloat1622float2(h[e]);
    f[2 * e] = t.x; f[2 * e + 1] = t.y;
  }
}
template <typename T> __device__ __forceinline__ void store8(T* p, const float* f);
template <> __device__ __forceinline__ void store8<float>(float* p, const float* f) {
  reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}
template <> __device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

template <typename T>
__global__ void __launch_bounds__(128) attn_decode_kernel(const T* __restrict__ q, long long ldq,
                                                          const T* __restrict__ cache_k, const T* __restrict__ cache_v,
                                                          long long cache_rows, long long ldc,
                                                          const int64_t* __restrict__ time_step, int pos,
                                                          T* __restrict__ o, long long ldo, int N, int H, float scale,
                                                          int q_per_cache) {
  const int lane = threadIdx.x & 31, r = lane >> 3, c = lane & 7;
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (wid >= (long long)N * H) return;
  const int n = (int)(wid / H), h = (int)(wid % H);
  float qv[8];
  load8<T>(q + (long long)n * ldq + h * DH + c * 8, qv);
#pragma unroll
  for (int d = 0; d < 8; ++d) qv[d] *= scale;
  long long start = time_step ? (long long)pos - time_step[n] : 0;  // no time_step: every row [0, pos]
  if (start < 0) start = 0;
  const long long cn = n / q_per_cache;  // queries that share one cached sequence (plain attention: S per sequence)
  const T* kb = cache_k + cn * cache_rows * ldc + h * DH + c * 8;
  const T* vb = cache_v + cn * cache_rows * ldc + h * DH + c * 8;
  float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) acc[d] = 0.f;
  for (long long s0 = start; s0 <= pos; s0 += 8) {  // two key quads per iteration
    float kf[2][8], vf[2][8];
    bool ok[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long s = s0 + u * 4 + r;
      ok[u] = s <= pos;
      if (ok[u]) {
        load8<T>(kb + s * ldc, kf[u]);
        load8<T>(vb + s * ldc, vf[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float sc = 0.f;
      if (ok[u]) {
#pragma unroll
        for (int d = 0; d < 8; ++d) sc = fmaf(qv[d], kf[u][d], sc);
      }
      sc += __shfl_xor_sync(0xffffffffu, sc, 1);
      sc += __shfl_xor_sync(0xffffffffu, sc, 2);
      sc += __shfl_xor_sync(0xffffffffu, sc, 4);
      if (ok[u]) {
        const float mn = fmaxf(m, sc);
        const float corr = __expf(m - mn), p = __expf(sc - mn);
        l = l * corr + p;
#pragma unroll
        for (int d = 0; d < 8; ++d) acc[d] = fmaf(p, vf[u][d], acc[d] * corr);
        m = mn;
      }
    }
  }
  // merge the four row groups
  float mg = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
  mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, 16));
  const float w = (m == -INFINITY) ? 0.f : __expf(m - mg);
  l *= w;
  l += __shfl_xor_sync(0xffffffffu, l, 8);
  l += __shfl_xor_sync(0xffffffffu, l, 16);
  const float inv = 1.f / l;
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    float a = acc[d] * w;
    a += __shfl_xor_sync(0xffffffffu, a, 8);
    a += __shfl_xor_sync(0xffffffffu, a, 16);
    acc[d] = a * inv;
  }
  if (r == 0) store8<T>(o + (long long)n * ldo + h * DH + c * 8, acc);
}

}  // namespace

extern "C" int svla_attn_decode(svla_ctx* ctx, const void* q, long long ldq, const void* cache_k, const void* cache_v,
                                long long cache_rows, long long ldc, const int64_t* time_step, int pos, void* o,
                                long long ldo, int dtype, int N, int H, int dh, float scale, int q_per_cache,
                                svla_stream stream) {
  SVLA_CHECK_ARG(ctx && q && cache_k && cache_v && o, "NULL argument");
  SVLA_CHECK_ARG(dh == DH, "head dim must be 64");
  SVLA_CHECK_ARG(pos >= 0 && pos < cache_rows, "position outside the cache");
  SVLA_CHECK_ARG(ldq % 8 == 0 && ldc % 8 == 0 && ldo % 8 == 0, "leading dims must be multiples of 8");
  SVLA_CHECK_ARG(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(cache_k) |
                   reinterpret_cast<uintptr_t>(cache_v) | reinterpret_cast<uintptr_t>(o)) & 31) == 0,
                 "buffers must be 32-byte aligned");
  SVLA_CHECK_ARG(q_per_cache >= 1, "q_per_cache must be >= 1");
  if (N <= 0) return SVLA_OK;
  const long long threads = (long long)N * H * 32;
  SVLA_DISPATCH_DTYPE(dtype, T, (attn_decode_kernel<T><<<(unsigned)((threads + 127) / 128), 128, 0, as_stream(stream)>>>(
                                    (const T*)q, ldq, (const T*)cache_k, (const T*)cache_v, cache_rows, ldc, time_step, pos,
                                    (T*)o, ldo, N, H, scale, q_per_cache)));
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}
