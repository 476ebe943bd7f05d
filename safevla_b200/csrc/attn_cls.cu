// Single-query attention for the last fusion layer: only output token 0 of the fusion transformer is
// consumed (reference allenact_dino_transformer.py:708), so in layer 3 the query, the out-projection
// and the FFN are needed for the CLS row alone while K/V cover all S tokens.  One warp per
// (sequence, head): lanes own keys for the score/softmax part and head-dim columns for the
// weighted sums; no shared memory.  Memory bound (reads K and V once: 2*S*64 elements per warp).
#include "common.cuh"

namespace {

constexpr int DH = 64;
constexpr int kMaxChunks = 8;  // S <= 256

template <typename T> __device__ __forceinline__ void load_row64(const T* p, float* out) {
#pragma unroll
  for (int i = 0; i < DH; i += 4) {
    const float4 v = load4<T>(p + i);
    out[i] = v.x; out[i + 1] = v.y; out[i + 2] = v.z; out[i + 3] = v.w;
  }
}

template <typename T>
__global__ void __launch_bounds__(128) attn_cls_fwd_kernel(const T* __restrict__ q, long long ldq, const T* __restrict__ k,
                                                           const T* __restrict__ v, long long ldkv, T* __restrict__ o,
                                                           long long ldo, float* __restrict__ lse, int B, int S, int H,
                                                           float scale) {
  const int lane = threadIdx.x & 31;
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (wid >= (long long)B * H) return;
  const int b = (int)(wid / H), h = (int)(wid % H);
  float qv[DH];
  load_row64<T>(q + (long long)b * ldq + h * DH, qv);
  const T* kb = k + (long long)b * S * ldkv + h * DH;
  const T* vb = v + (long long)b * S * ldkv + h * DH;
  float sc[kMaxChunks];
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    sc[c] = -INFINITY;
    const int j = c * 32 + lane;
    if (j < S) {
      float kr[DH];
      load_row64<T>(kb + (long long)j * ldkv, kr);
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) acc = fmaf(qv[d], kr[d], acc);
      sc[c] = acc * scale;
      mx = fmaxf(mx, sc[c]);
    }
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    sc[c] = (sc[c] == -INFINITY) ? 0.f : __expf(sc[c] - mx);
    sum += sc[c];
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  float o0 = 0.f, o1 = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    if (c * 32 < S) {
      for (int l = 0; l < 32; ++l) {
        const int j = c * 32 + l;
        const float p = __shfl_sync(0xffffffffu, sc[c], l);
        if (j < S) {
          o0 = fmaf(p, to_f<T>(vb[(long long)j * ldkv + lane]), o0);
          o1 = fmaf(p, to_f<T>(vb[(long long)j * ldkv + lane + 32]), o1);
        }
      }
    }
  }
  T* ob = o + (long long)b * ldo + h * DH;
  ob[lane] = from_f<T>(o0 * inv);
  ob[lane + 32] = from_f<T>(o1 * inv);
  if (lane == 0) lse[wid] = mx + __logf(sum);
}

template <typename T>
__global__ void __launch_bounds__(128) attn_cls_bwd_kernel(const T* __restrict__ q, long long ldq, const T* __restrict__ k,
                                                           const T* __restrict__ v, long long ldkv,
                                                           const T* __restrict__ o, const T* __restrict__ d_o,
                                                           long long ldo, T* __restrict__ dq, long long lddq,
                                                           T* __restrict__ dk, T* __restrict__ dv, long long lddkv,
                                                           const float* __restrict__ lse, int B, int S, int H,
                                                           float scale) {
  const int lane = threadIdx.x & 31;
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (wid >= (long long)B * H) return;
  const int b = (int)(wid / H), h = (int)(wid % H);
  float qv[DH], dov[DH];
  load_row64<T>(q + (long long)b * ldq + h * DH, qv);
  load_row64<T>(d_o + (long long)b * ldo + h * DH, dov);
  float delta = 0.f;
  {
    const T* ob = o + (long long)b * ldo + h * DH;
    const float part = to_f<T>(ob[lane]) * to_f<T>(d_o[(long long)b * ldo + h * DH + lane]) +
                 to_f<T>(ob[lane + 32]) * to_f<T>(d_o[(long long)b * ldo + h * DH + lane + 32]);
    delta = warp_sum(part);
  }
  const float l_ = lse[wid];
  const T* kb = k + (long long)b * S * ldkv + h * DH;
  const T* vb = v + (long long)b * S * ldkv + h * DH;
  T* dkb = dk + (long long)b * S * lddkv + h * DH;
  T* dvb = dv + (long long)b * S * lddkv + h * DH;
  float ds[kMaxChunks];
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    ds[c] = 0.f;
    const int j = c * 32 + lane;
    if (j < S) {
      float kr[DH];
      load_row64<T>(kb + (long long)j * ldkv, kr);
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) s = fmaf(qv[d], kr[d], s);
      const float p = __expf(s * scale - l_);
      float vr[DH];
      load_row64<T>(vb + (long long)j * ldkv, vr);
      float dp = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) dp = fmaf(dov[d], vr[d], dp);
      const float dsj = p * (dp - delta);
      ds[c] = dsj;
      T* dkr = dkb + (long long)j * lddkv;
      T* dvr = dvb + (long long)j * lddkv;
#pragma unroll
      for (int d = 0; d < DH; d += 4) {
        store4<T>(dkr + d, make_float4(scale * dsj * qv[d], scale * dsj * qv[d + 1], scale * dsj * qv[d + 2],
                                       scale * dsj * qv[d + 3]));
        store4<T>(dvr + d, make_float4(p * dov[d], p * dov[d + 1], p * dov[d + 2], p * dov[d + 3]));
      }
    }
  }
  float q0 = 0.f, q1 = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    if (c * 32 < S) {
      for (int l = 0; l < 32; ++l) {
        const int j = c * 32 + l;
        const float d_ = __shfl_sync(0xffffffffu, ds[c], l);
        if (j < S) {
          q0 = fmaf(d_, to_f<T>(kb[(long long)j * ldkv + lane]), q0);
          q1 = fmaf(d_, to_f<T>(kb[(long long)j * ldkv + lane + 32]), q1);
        }
      }
    }
  }
  T* dqb = dq + (long long)b * lddq + h * DH;
  dqb[lane] = from_f<T>(q0 * scale);
  dqb[lane + 32] = from_f<T>(q1 * scale);
}

}  // namespace

extern "C" int svla_attn_cls_fwd(svla_ctx* ctx, const void* q, long long ldq, const void* k, const void* v,
                                 long long ldkv, void* o, long long ldo, int dtype, float* lse, int B, int S, int H,
                                 int dh, float scale, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && q && k && v && o && lse, "NULL argument");
  SVLA_CHECK_ARG(dh == DH && S >= 1 && S <= 32 * kMaxChunks, "head dim must be 64 and S <= 256");
  SVLA_CHECK_ARG(ldq % 4 == 0 && ldkv % 4 == 0 && ldo % 4 == 0, "leading dims must be multiples of 4");
  if (B <= 0) return SVLA_OK;
  const long long threads = (long long)B * H * 32;
  SVLA_DISPATCH_DTYPE(dtype, T, (attn_cls_fwd_kernel<T><<<(unsigned)((threads + 127) / 128), 128, 0, as_stream(stream)>>>(
                                    (const T*)q, ldq, (const T*)k, (const T*)v, ldkv, (T*)o, ldo, lse, B, S, H, scale)));
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

extern "C" int svla_attn_cls_bwd(svla_ctx* ctx, const void* q, long long ldq, const void* k, const void* v,
                                 long long ldkv, const void* o, const void* d_o, long long ldo, void* dq,
                                 long long lddq, void* dk, void* dv, long long lddkv, int dtype, const float* lse,
                                 int B, int S, int H, int dh, float scale, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && q && k && v && o && d_o && dq && dk && dv && lse, "NULL argument");
  SVLA_CHECK_ARG(dh == DH && S >= 1 && S <= 32 * kMaxChunks, "head dim must be 64 and S <= 256");
  if (B <= 0) return SVLA_OK;
  const long long threads = (long long)B * H * 32;
  SVLA_DISPATCH_DTYPE(dtype, T, (attn_cls_bwd_kernel<T><<<(unsigned)((threads + 127) / 128), 128, 0, as_stream(stream)>>>(
                                    (const T*)q, ldq, (const T*)k, (const T*)v, ldkv, (const T*)o, (const T*)d_o, ldo,
                                    (T*)dq, lddq, (T*)dk, (T*)dv, lddkv, lse, B, S, H, scale)));
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}
