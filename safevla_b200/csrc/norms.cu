// LayerNorm / RMSNorm forward + backward, one warp per row, 128-bit accesses, fp32 statistics.
// HBM bound: fwd reads x (+res) and writes y; bwd reads dy and x, writes dx; the per-feature
// gradients (dgamma/dbeta/dtoken/dw) are reduced deterministically: register accumulation per
// warp -> fixed block partials in the context workspace -> fixed-order fold.
// Kernels are specialised on NV = D / 128 (float4 per lane) so the row lives in exactly NV registers x 4.
#include <algorithm>

#include "common.cuh"
#include "fold.cuh"

namespace {

constexpr int kWarps = 8;

template <typename TI, typename TO, int NV>
__global__ void __launch_bounds__(kWarps * 32)
layernorm_fwd_kernel(const TI* __restrict__ x, const TI* __restrict__ res, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const float* __restrict__ token, int relu, float eps,
                     TO* __restrict__ y, svla_rowmap ymap, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                     long long rows) {
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarps;
  for (long long r = warp0; r < rows; r += nwarps) {
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const long long off = r * D + (i * 32 + lane) * 4;
      v[i] = load4<TI>(x + off);
      if (res) {
        const float4 q = load4<TI>(res + off);
        v[i].x += q.x; v[i].y += q.y; v[i].z += q.z; v[i].w += q.w;
      }
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
    if (lane == 0) {
      if (mean_out) mean_out[r] = mean;
      if (rstd_out) rstd_out[r] = rstd;
    }
    const long long orow = map_row(ymap, r);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x;
      o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z;
      o.w = (v[i].w - mean) * rstd * g.w + b.w;
      if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      if (token) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(token + c));
        o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
      }
      store4<TO>(y + orow * D + c, o);
    }
  }
}

// partial layout in workspace: [gridDim.x][3][D]  (dgamma, dbeta, dtoken)
// 4 warps per block and no gamma/beta register copies (they sit in L1): ~100 registers, 5 blocks per SM, so
// enough row loads are in flight to cover HBM latency.
constexpr int kBwdWarps = 4;
template <typename TDY, typename TI, typename TDX, int NV, bool RELU, bool TOKEN>
__global__ void __launch_bounds__(kBwdWarps * 32)
layernorm_bwd_kernel(const TDY* __restrict__ dy, svla_rowmap dymap, const TI* __restrict__ x,
                     const TI* __restrict__ res, const float* __restrict__ gamma, const float* __restrict__ beta,
                     const float* __restrict__ mean_in, const float* __restrict__ rstd_in, TDX* __restrict__ dx,
                     float* __restrict__ partial, long long rows) {
  constexpr int D = NV * 128;
  extern __shared__ float sm[];  // [kBwdWarps][3][D]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long warp0 = (long long)blockIdx.x * kBwdWarps + w;
  const long long nwarps = (long long)gridDim.x * kBwdWarps;
  float4 ag[NV], ab[NV], at[TOKEN ? NV : 1];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (TOKEN) at[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long r = warp0; r < rows; r += nwarps) {
    const long long drow = map_row(dymap, r);
    float4 xv[NV], d[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {  // all loads of the row first
      const int c = (i * 32 + lane) * 4;
      xv[i] = load4<TI>(x + r * D + c);
      d[i] = load4<TDY>(dy + drow * D + c);
    }
    const float mean = mean_in[r], rstd = rstd_in[r];
    if (res) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 q = load4<TI>(res + r * D + (i * 32 + lane) * 4);
        xv[i].x += q.x; xv[i].y += q.y; xv[i].z += q.z; xv[i].w += q.w;
      }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
      float4 h;
      h.x = (xv[i].x - mean) * rstd; h.y = (xv[i].y - mean) * rstd;
      h.z = (xv[i].z - mean) * rstd; h.w = (xv[i].w - mean) * rstd;
      if (TOKEN) { at[i].x += d[i].x; at[i].y += d[i].y; at[i].z += d[i].z; at[i].w += d[i].w; }
      if (RELU) {
        const float4 bt = __ldg(reinterpret_cast<const float4*>(beta + c));
        if (h.x * gm.x + bt.x <= 0.f) d[i].x = 0.f;
        if (h.y * gm.y + bt.y <= 0.f) d[i].y = 0.f;
        if (h.z * gm.z + bt.z <= 0.f) d[i].z = 0.f;
        if (h.w * gm.w + bt.w <= 0.f) d[i].w = 0.f;
      }
      ag[i].x += d[i].x * h.x; ag[i].y += d[i].y * h.y; ag[i].z += d[i].z * h.z; ag[i].w += d[i].w * h.w;
      ab[i].x += d[i].x; ab[i].y += d[i].y; ab[i].z += d[i].z; ab[i].w += d[i].w;
      d[i].x *= gm.x; d[i].y *= gm.y; d[i].z *= gm.z; d[i].w *= gm.w;  // d xhat
      s1 += d[i].x + d[i].y + d[i].z + d[i].w;
      s2 += d[i].x * h.x + d[i].y * h.y + d[i].z * h.z + d[i].w * h.w;
      xv[i] = h;
    }
    s1 = warp_sum(s1) / (float)D;
    s2 = warp_sum(s2) / (float)D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      float4 o;
      o.x = rstd * (d[i].x - s1 - xv[i].x * s2);
      o.y = rstd * (d[i].y - s1 - xv[i].y * s2);
      o.z = rstd * (d[i].z - s1 - xv[i].z * s2);
      o.w = rstd * (d[i].w - s1 - xv[i].w * s2);
      store4<TDX>(dx + r * D + c, o);
    }
  }
  // block partials
  float* mine = sm + (size_t)w * 3 * D;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    *reinterpret_cast<float4*>(mine + c) = ag[i];
    *reinterpret_cast<float4*>(mine + D + c) = ab[i];
    *reinterpret_cast<float4*>(mine + 2 * D + c) = TOKEN ? at[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 3 * D; e += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int ww = 0; ww < kBwdWarps; ++ww) s += sm[(size_t)ww * 3 * D + e];
    partial[(size_t)blockIdx.x * 3 * D + e] = s;
  }
}

// out[j][d] += sum_b partial[b][j][d]   (fixed order)
__global__ void fold_partials_kernel(const float* __restrict__ partial, int nb, int nvec, int D, float* o0, float* o1,
                                     float* o2) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nvec * D) return;
  const int j = e / D, d = e % D;
  float* o = j == 0 ? o0 : (j == 1 ? o1 : o2);
  if (!o) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int b = 0;
  for (; b + 3 < nb; b += 4) {
    s0 += partial[(size_t)b * nvec * D + e];
    s1 += partial[(size_t)(b + 1) * nvec * D + e];
    s2 += partial[(size_t)(b + 2) * nvec * D + e];
    s3 += partial[(size_t)(b + 3) * nvec * D + e];
  }
  for (; b < nb; ++b) s0 += partial[(size_t)b * nvec * D + e];
  o[d] += (s0 + s1) + (s2 + s3);
}

template <typename TI, typename TO, int NV>
__global__ void __launch_bounds__(kWarps * 32)
rmsnorm_fwd_kernel(const TI* __restrict__ x, const float* __restrict__ w, float eps, TO* __restrict__ y,
                   float* __restrict__ rstd_out, long long rows) {
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarps;
  for (long long r = warp0; r < rows; r += nwarps) {
    float4 v[NV];
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i] = load4<TI>(x + r * D + (i * 32 + lane) * 4);
      q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
    if (lane == 0 && rstd_out) rstd_out[r] = rstd;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 g = __ldg(reinterpret_cast<const float4*>(w + c));
      store4<TO>(y + r * D + c,
                 make_float4(v[i].x * rstd * g.x, v[i].y * rstd * g.y, v[i].z * rstd * g.z, v[i].w * rstd * g.w));
    }
  }
}

template <typename TDY, typename TI, typename TDX, int NV>
__global__ void __launch_bounds__(kWarps * 32)
rmsnorm_bwd_kernel(const TDY* __restrict__ dy, const TI* __restrict__ x, const float* __restrict__ w,
                   const float* __restrict__ rstd_in, TDX* __restrict__ dx, int accumulate_dx,
                   float* __restrict__ partial, long long rows) {
  constexpr int D = NV * 128;
  extern __shared__ float sm[];  // [kWarps][D]
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const long long warp0 = (long long)blockIdx.x * kWarps + wi;
  const long long nwarps = (long long)gridDim.x * kWarps;
  float4 aw[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) aw[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long r = warp0; r < rows; r += nwarps) {
    const float rstd = rstd_in[r];
    float4 xh[NV], dh[NV];
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 xv = load4<TI>(x + r * D + c);
      float4 d = load4<TDY>(dy + r * D + c);
      const float4 g = __ldg(reinterpret_cast<const float4*>(w + c));
      float4 h = make_float4(xv.x * rstd, xv.y * rstd, xv.z * rstd, xv.w * rstd);
      aw[i].x += d.x * h.x; aw[i].y += d.y * h.y; aw[i].z += d.z * h.z; aw[i].w += d.w * h.w;
      d.x *= g.x; d.y *= g.y; d.z *= g.z; d.w *= g.w;
      s2 += d.x * h.x + d.y * h.y + d.z * h.z + d.w * h.w;
      xh[i] = h; dh[i] = d;
    }
    s2 = warp_sum(s2) / (float)D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      float4 o;
      o.x = rstd * (dh[i].x - xh[i].x * s2);
      o.y = rstd * (dh[i].y - xh[i].y * s2);
      o.z = rstd * (dh[i].z - xh[i].z * s2);
      o.w = rstd * (dh[i].w - xh[i].w * s2);
      if (accumulate_dx) {
        const float4 p = load4<TDX>(dx + r * D + c);
        o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
      }
      store4<TDX>(dx + r * D + c, o);
    }
  }
  float* mine = sm + (size_t)wi * D;
#pragma unroll
  for (int i = 0; i < NV; ++i) *reinterpret_cast<float4*>(mine + (i * 32 + lane) * 4) = aw[i];
  __syncthreads();
  for (int e = threadIdx.x; e < D; e += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int ww = 0; ww < kWarps; ++ww) s += sm[(size_t)ww * D + e];
    partial[(size_t)blockIdx.x * D + e] = s;
  }
}

inline int norm_grid(svla_ctx* ctx, long long rows, int per_sm) {
  const long long b = (rows + kWarps - 1) / kWarps;
  return (int)std::max<long long>(1, std::min<long long>(b, (long long)ctx->sm_count * per_sm));
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

#define DISPATCH2(dtA, TA_, dtB, TB_, ...)                                              \
  do {                                                                                  \
    if (dtA == SVLA_F32 && dtB == SVLA_F32) { using TA_ = float; using TB_ = float; __VA_ARGS__; } \
    else if (dtA == SVLA_F32 && dtB == SVLA_BF16) { using TA_ = float; using TB_ = __nv_bfloat16; __VA_ARGS__; } \
    else if (dtA == SVLA_BF16 && dtB == SVLA_F32) { using TA_ = __nv_bfloat16; using TB_ = float; __VA_ARGS__; } \
    else if (dtA == SVLA_BF16 && dtB == SVLA_BF16) { using TA_ = __nv_bfloat16; using TB_ = __nv_bfloat16; __VA_ARGS__; } \
    else { svla_set_error("bad dtype"); return SVLA_ERR_BAD_ARG; }                      \
  } while (0)

// the towers use D = 512 (and 384-wide DINO features upstream); other multiples of 128 up to 1024 are accepted
#define DISPATCH_NV(D_, NV_, ...)                                                                 \
  do {                                                                                            \
    switch ((D_) / 128) {                                                                         \
      case 1: { constexpr int NV_ = 1; __VA_ARGS__; } break;                                      \
      case 2: { constexpr int NV_ = 2; __VA_ARGS__; } break;                                      \
      case 3: { constexpr int NV_ = 3; __VA_ARGS__; } break;                                      \
      case 4: { constexpr int NV_ = 4; __VA_ARGS__; } break;                                      \
      case 6: { constexpr int NV_ = 6; __VA_ARGS__; } break;                                      \
      case 8: { constexpr int NV_ = 8; __VA_ARGS__; } break;                                      \
      default: svla_set_error("unsupported feature width %d", (int)(D_)); return SVLA_ERR_BAD_SHAPE; \
    }                                                                                             \
  } while (0)

extern "C" int svla_layernorm_fwd(svla_ctx* ctx, const void* x, const void* res, int dtype_in, const float* gamma,
                                  const float* beta, const float* token, int relu, float eps, void* y, int dtype_out,
                                  svla_rowmap ymap, float* mean, float* rstd, long long rows, int D,
                                  svla_stream stream) {
  SVLA_CHECK_ARG(ctx && x && gamma && beta && y, "NULL argument");
  SVLA_CHECK_ARG(D % 128 == 0 && D <= 1024, "D must be a multiple of 128, <= 1024");
  SVLA_CHECK_ARG(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta), "unaligned buffer");
  if (rows <= 0) return SVLA_OK;
  DISPATCH2(dtype_in, TI, dtype_out, TO, DISPATCH_NV(D, NV, (layernorm_fwd_kernel<TI, TO, NV>
            <<<norm_grid(ctx, rows, 8), kWarps * 32, 0, as_stream(stream)>>>(
                (const TI*)x, (const TI*)res, gamma, beta, token, relu, eps, (TO*)y, ymap, mean, rstd, rows))));
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

extern "C" int svla_layernorm_bwd(svla_ctx* ctx, const void* dy, int dtype_dy, svla_rowmap dymap, const void* x,
                                  const void* res, int dtype_in, const float* gamma, const float* beta, int relu,
                                  const float* mean, const float* rstd, void* dx, int dtype_dx, float* dgamma,
                                  float* dbeta, float* dtoken, long long rows, int D, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && dy && x && gamma && beta && mean && rstd && dx, "NULL argument");
  SVLA_CHECK_ARG(D % 128 == 0 && D <= 1024, "D must be a multiple of 128, <= 1024");
  SVLA_CHECK_ARG(dtype_dx == dtype_dy, "dx and dy must share a dtype");
  if (rows <= 0) return SVLA_OK;
  const int grid = (int)std::max<long long>(
      1, std::min<long long>((rows + kBwdWarps - 1) / kBwdWarps, (long long)ctx->sm_count * 5));
  float* partial = reinterpret_cast<float*>(ctx->ws);
  SVLA_CHECK_ARG((size_t)grid * 3 * D * sizeof(float) <= ctx->ws_bytes, "workspace too small");
  const size_t smem = sizeof(float) * kBwdWarps * 3 * D;
#define SVLA_LN_BWD(RELU_, TOKEN_)                                                                                  \
  DISPATCH2(dtype_dy, TDY, dtype_in, TI, DISPATCH_NV(D, NV, {                                                        \
    auto kern = layernorm_bwd_kernel<TDY, TI, TDY, NV, RELU_, TOKEN_>;                                               \
    SVLA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                   \
    kern<<<grid, kBwdWarps * 32, smem, as_stream(stream)>>>((const TDY*)dy, dymap, (const TI*)x, (const TI*)res, gamma, \
                                                            beta, mean, rstd, (TDY*)dx, partial, rows);             \
  }))
  if (relu && dtoken) SVLA_LN_BWD(true, true);
  else if (relu) SVLA_LN_BWD(true, false);
  else if (dtoken) SVLA_LN_BWD(false, true);
  else SVLA_LN_BWD(false, false);
#undef SVLA_LN_BWD
  SVLA_LAUNCH_CHECK();
  svla_launch_fold(partial, grid, 3 * D, D, dgamma, dbeta, dtoken, 1, as_stream(stream));
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

extern "C" int svla_rmsnorm_fwd(svla_ctx* ctx, const void* x, int dtype_in, const float* w, float eps, void* y,
                                int dtype_out, float* rstd, long long rows, int D, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && x && w && y, "NULL argument");
  SVLA_CHECK_ARG(D % 128 == 0 && D <= 1024, "D must be a multiple of 128, <= 1024");
  if (rows <= 0) return SVLA_OK;
  DISPATCH2(dtype_in, TI, dtype_out, TO, DISPATCH_NV(D, NV, (rmsnorm_fwd_kernel<TI, TO, NV>
            <<<norm_grid(ctx, rows, 8), kWarps * 32, 0, as_stream(stream)>>>((const TI*)x, w, eps, (TO*)y, rstd, rows))));
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

extern "C" int svla_rmsnorm_bwd(svla_ctx* ctx, const void* dy, int dtype_dy, const void* x, int dtype_in,
                                const float* w, const float* rstd, void* dx, int dtype_dx, int accumulate_dx,
                                float* dw, long long rows, int D, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && dy && x && w && rstd && dx, "NULL argument");
  SVLA_CHECK_ARG(D % 128 == 0 && D <= 1024, "D must be a multiple of 128, <= 1024");
  SVLA_CHECK_ARG(dtype_dx == dtype_dy || dtype_dx == SVLA_F32, "dx must be fp32 or share dy's dtype");
  if (rows <= 0) return SVLA_OK;
  const int grid = norm_grid(ctx, rows, 2);
  float* partial = reinterpret_cast<float*>(ctx->ws);
  const size_t smem = sizeof(float) * kWarps * D;
  if (dtype_dx == dtype_dy) {
    DISPATCH2(dtype_dy, TDY, dtype_in, TI, DISPATCH_NV(D, NV, (rmsnorm_bwd_kernel<TDY, TI, TDY, NV>
              <<<grid, kWarps * 32, smem, as_stream(stream)>>>((const TDY*)dy, (const TI*)x, w, rstd, (TDY*)dx,
                                                               accumulate_dx, partial, rows))));
  } else {
    DISPATCH2(dtype_dy, TDY, dtype_in, TI, DISPATCH_NV(D, NV, (rmsnorm_bwd_kernel<TDY, TI, float, NV>
              <<<grid, kWarps * 32, smem, as_stream(stream)>>>((const TDY*)dy, (const TI*)x, w, rstd, (float*)dx,
                                                               accumulate_dx, partial, rows))));
  }
  SVLA_LAUNCH_CHECK();
  if (dw) {
    svla_launch_fold(partial, grid, D, D, dw, nullptr, nullptr, 1, as_stream(stream));
    SVLA_LAUNCH_CHECK();
  }
  return SVLA_OK;
}
