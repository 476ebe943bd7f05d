// Flat-arena optimizer path: sum-of-squares, fused global-norm clip + Adam (+ bf16 shadow refresh,
// + grad zeroing), Lagrange-multiplier update.  HBM bound: clip_adam reads p,g,m,v and writes
// p,m,v = 24 B per parameter (+4 B zeroing g, +2 B bf16 shadow when enabled).
#include <algorithm>

#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) sq_norm_kernel(const float* __restrict__ x, long long n, float* partials,
                                                      unsigned int* ticket, float* out) {
  __shared__ float red[32];
  __shared__ unsigned int s_ticket;
  float s = 0.f;
  const long long n4 = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x4 + i);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = x[(n4 << 2) + threadIdx.x];
    s += v * v;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
  __threadfence();
  if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1u);
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  __threadfence();
  if (threadIdx.x < 32) {
    double t = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += 32) t += (double)partials[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) {
      out[0] = (float)t;
      *ticket = 0u;
    }
  }
}

struct AdamArgs {
  float* p; float* g; float* m; float* v; __nv_bfloat16* pb;
  long long n; const float* sq_norm;
  svla_adam_hparams hp;
  float bc1, bc2_sqrt_inv;
};

__global__ void __launch_bounds__(256) clip_adam_kernel(AdamArgs a) {
  float coef = a.hp.grad_prescale;
  if (a.sq_norm != nullptr && a.hp.max_grad_norm > 0.f) {
    const float norm = sqrtf(*a.sq_norm);  // norm of the pre-scaled gradient
    coef *= fminf(a.hp.max_grad_norm / (norm + 1e-6f), 1.f);
  }
  const float b1 = a.hp.beta1, b2 = a.hp.beta2, step_size = a.hp.lr / a.bc1;
  const long long n4 = a.n >> 2;
  float4* p4 = reinterpret_cast<float4*>(a.p);
  float4* g4 = reinterpret_cast<float4*>(a.g);
  float4* m4 = reinterpret_cast<float4*>(a.m);
  float4* v4 = reinterpret_cast<float4*>(a.v);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 p = p4[i], g = g4[i], m = m4[i], v = v4[i];
    float* pp = &p.x; float* gg = &g.x; float* mm = &m.x; float* vv = &v.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = gg[j] * coef;
      mm[j] = b1 * mm[j] + (1.f - b1) * gr;
      vv[j] = b2 * vv[j] + (1.f - b2) * gr * gr;
      const float denom = sqrtf(vv[j]) * a.bc2_sqrt_inv + a.hp.eps;
      pp[j] = pp[j] - step_size * (mm[j] / denom);
    }
    p4[i] = p; m4[i] = m; v4[i] = v;
    if (a.hp.zero_grad) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.pb) store4<__nv_bfloat16>(a.pb + (i << 2), p);
  }
}

__global__ void lagrange_kernel(float* lam, float* st, const float* cost_sum_cnt, float limit, float lr, float ub) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float cnt = cost_sum_cnt[1];
  // a rollout in which no episode finished carries no episode-cost information: lambda and its Adam state stay
  // untouched (treating it as Jc = 0 would push lambda towards 0 and weaken the constraint); st[3] keeps the last Jc
  if (!(cnt >= 1.f)) return;
  const float jc = cost_sum_cnt[0] / cnt;
  const float g = -(jc - limit);  // d/dlambda of -lambda * (Jc - d)
  float m = st[0], v = st[1];
  const float t = st[2] + 1.f;
  m = 0.9f * m + 0.1f * g;
  v = 0.999f * v + 0.001f * g * g;
  const float bc1 = 1.f - powf(0.9f, t), bc2 = 1.f - powf(0.999f, t);
  const float denom = sqrtf(v) / sqrtf(bc2) + 1e-8f;
  float l = *lam - (lr / bc1) * (m / denom);
  l = fmaxf(l, 0.f);
  if (ub >= 0.f) l = fminf(l, ub);
  *lam = l;
  st[0] = m; st[1] = v; st[2] = t;
  st[3] = jc;
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                        long long n) {
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
    store4<__nv_bfloat16>(y + (i << 2), __ldg(reinterpret_cast<const float4*>(x) + i));
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) y[(n4 << 2) + threadIdx.x] = __float2bfloat16_rn(x[(n4 << 2) + threadIdx.x]);
}

inline int grid_for(svla_ctx* ctx, long long n4, int per_sm) {
  long long b = (n4 + 255) / 256;
  b = std::min<long long>(b, (long long)ctx->sm_count * per_sm);
  return (int)std::max<long long>(b, 1);
}

}  // namespace

extern "C" int svla_sq_norm(svla_ctx* ctx, const float* x, long long n, float* out_dev, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && x && out_dev, "NULL argument");
  SVLA_CHECK_ARG(n >= 0, "negative n");
  SVLA_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0, "x must be 16-byte aligned");
  int grid = std::min(grid_for(ctx, n >> 2, 8), kMaxPartialBlocks);
  sq_norm_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, n, ctx->partials + kMaxPartialBlocks * 8, ctx->tickets + 1,
                                                       out_dev);
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

extern "C" int svla_clip_adam(svla_ctx* ctx, float* p, float* g, float* m, float* v, void* p_bf16, long long n,
                              const float* sq_norm_dev, const svla_adam_hparams* hp, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && p && g && m && v && hp, "NULL argument");
  SVLA_CHECK_ARG(n >= 0 && (n & 3) == 0, "arena length must be a multiple of 4");
  SVLA_CHECK_ARG(hp->step >= 1, "Adam step is 1-based");
  AdamArgs a;
  a.p = p; a.g = g; a.m = m; a.v = v; a.pb = reinterpret_cast<__nv_bfloat16*>(p_bf16);
  a.n = n; a.sq_norm = sq_norm_dev; a.hp = *hp;
  a.bc1 = (float)(1.0 - pow((double)hp->beta1, (double)hp->step));
  a.bc2_sqrt_inv = (float)(1.0 / sqrt(1.0 - pow((double)hp->beta2, (double)hp->step)));
  clip_adam_kernel<<<grid_for(ctx, n >> 2, 8), 256, 0, as_stream(stream)>>>(a);
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

extern "C" int svla_lagrange_update(svla_ctx* ctx, float* lambda_dev, float* state_dev, const float* cost_sum_cnt_dev,
                                    float cost_limit, float lr, float upper_bound, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && lambda_dev && state_dev && cost_sum_cnt_dev, "NULL argument");
  lagrange_kernel<<<1, 32, 0, as_stream(stream)>>>(lambda_dev, state_dev, cost_sum_cnt_dev, cost_limit, lr,
                                                    upper_bound);
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

extern "C" int svla_cast_bf16(svla_ctx* ctx, const float* x, void* y, long long n, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && x && y, "NULL argument");
  cast_bf16_kernel<<<grid_for(ctx, n >> 2, 8), 256, 0, as_stream(stream)>>>(x, reinterpret_cast<__nv_bfloat16*>(y), n);
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}
