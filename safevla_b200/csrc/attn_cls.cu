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


// ---- bf16 fast path ------------------------------------------------------------------------------------------
// A warp still owns one (sequence, head), but a K / V row (64 bf16 = 128 B) is read by EIGHT lanes with one 16-byte
// load each, so a warp instruction covers four whole rows (four full 128-byte lines) instead of 32 partial ones.
// lane = (r, c): r = lane / 8 picks the row of the quad, c = lane % 8 the 8-column slice.  Dot products reduce over
// the 8 lanes of a row group (3 shuffles); the weighted sums reduce over the 4 row groups at the very end.
constexpr int kQuads = 64;  // S <= 256

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 t = __bfloat1622float2(h[e]);
    f[2 * e] = t.x; f[2 * e + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
  return u;
}
__device__ __forceinline__ float group8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}
__device__ __forceinline__ float rows4_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}

// KB: quads (4 keys x 128 B) with loads in flight together per warp.  16 (8 KB per warp, three CTAs per SM) halves the
// number of dependent DRAM round trips of an item against 8 (4 KB, four CTAs per SM).
template <int NQ, bool DROP, int KB>  // key quads: S <= 4 * NQ; DROP: dropout on the probabilities
__global__ void __launch_bounds__(128, KB == 16 ? 3 : 4) attn_cls_fwd_bf16_kernel(const __nv_bfloat16* __restrict__ q, long long ldq,
                                                                const __nv_bfloat16* __restrict__ k,
                                                                const __nv_bfloat16* __restrict__ v, long long ldkv,
                                                                __nv_bfloat16* __restrict__ o, long long ldo,
                                                                float* __restrict__ lse, int B, int S, int H, float scale,
                                                                DropArgs drop) {
  const int lane = threadIdx.x & 31, r = lane >> 3, c = lane & 7;
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (wid >= (long long)B * H) return;
  const int b = (int)(wid / H), h = (int)(wid % H);
  float qv[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(q + (long long)b * ldq + h * DH + c * 8)), qv);
  const __nv_bfloat16* kb = k + (long long)b * S * ldkv + h * DH + c * 8;
  const __nv_bfloat16* vb = v + (long long)b * S * ldkv + h * DH + c * 8;
  constexpr int kBatch = KB;
  float sc[NQ];
  float mx = -INFINITY;
#pragma unroll
  for (int i0 = 0; i0 < NQ; i0 += kBatch) {
    uint4 kr[kBatch];
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {
      const int j = 4 * (i0 + i) + r;
      if (j < S) kr[i] = __ldg(reinterpret_cast<const uint4*>(kb + (long long)j * ldkv));
    }
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {
      const int j = 4 * (i0 + i) + r;
      float kf[8], acc = 0.f;
      if (j < S) {
        unpack8(kr[i], kf);
#pragma unroll
        for (int d = 0; d < 8; ++d) acc = fmaf(qv[d], kf[d], acc);
      }
      acc = group8_sum(acc) * scale;
      sc[i0 + i] = (j < S) ? acc : -INFINITY;
      mx = fmaxf(mx, sc[i0 + i]);
    }
  }
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
  float sum = 0.f, acc[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) acc[d] = 0.f;
#pragma unroll
  for (int i0 = 0; i0 < NQ; i0 += kBatch) {
    uint4 vr[kBatch];
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {
      const int j = 4 * (i0 + i) + r;
      if (j < S) vr[i] = __ldg(reinterpret_cast<const uint4*>(vb + (long long)j * ldkv));
    }
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {
      const int j = 4 * (i0 + i) + r;
      if (j < S) {
        float p = __expf(sc[i0 + i] - mx);
        sum += p;  // the normaliser is the undropped row sum
        if (DROP)  // mask row = the CLS row of this (sequence, head) in the full-sequence mask
          p = ((dropout_keep8(drop, drop.row0 + (uint32_t)wid * (S > 128 ? 256u : 128u), (uint32_t)(j >> 3)) >> (j & 7)) & 1u) ? p * drop.scale : 0.f;
        float vf[8];
        unpack8(vr[i], vf);
#pragma unroll
        for (int d = 0; d < 8; ++d) acc[d] = fmaf(p, vf[d], acc[d]);
      }
    }
  }
  sum = rows4_sum(sum);  // every lane of a row group holds the same p, so this is the full row sum
  const float inv = 1.f / sum;
#pragma unroll
  for (int d = 0; d < 8; ++d) acc[d] = rows4_sum(acc[d]) * inv;
  if (r == 0) *reinterpret_cast<uint4*>(o + (long long)b * ldo + h * DH + c * 8) = pack8(acc);
  if (lane == 0) lse[wid] = mx + __logf(sum);
}

template <int NQ>
__global__ void __launch_bounds__(128, 4) attn_cls_bwd_bf16_kernel(
    const __nv_bfloat16* __restrict__ q, long long ldq, const __nv_bfloat16* __restrict__ k,
    const __nv_bfloat16* __restrict__ v, long long ldkv, const __nv_bfloat16* __restrict__ o,
    const __nv_bfloat16* __restrict__ d_o, long long ldo, __nv_bfloat16* __restrict__ dq, long long lddq,
    __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv, long long lddkv, const float* __restrict__ lse, int B,
    int S, int H, float scale, DropArgs drop) {
  const int lane = threadIdx.x & 31, r = lane >> 3, c = lane & 7;
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (wid >= (long long)B * H) return;
  const int b = (int)(wid / H), h = (int)(wid % H);
  float qv[8], dov[8], ov[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(q + (long long)b * ldq + h * DH + c * 8)), qv);
  unpack8(__ldg(reinterpret_cast<const uint4*>(d_o + (long long)b * ldo + h * DH + c * 8)), dov);
  unpack8(__ldg(reinterpret_cast<const uint4*>(o + (long long)b * ldo + h * DH + c * 8)), ov);
  float delta = 0.f;
#pragma unroll
  for (int d = 0; d < 8; ++d) delta = fmaf(ov[d], dov[d], delta);
  delta = group8_sum(delta);
  const float l_ = lse[wid];
  const __nv_bfloat16* kb = k + (long long)b * S * ldkv + h * DH + c * 8;
  const __nv_bfloat16* vb = v + (long long)b * S * ldkv + h * DH + c * 8;
  __nv_bfloat16* dkb = dk + (long long)b * S * lddkv + h * DH + c * 8;
  __nv_bfloat16* dvb = dv + (long long)b * S * lddkv + h * DH + c * 8;
  float dqa[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) dqa[d] = 0.f;
  constexpr int kBatch = 8;  // quads with loads in flight together
#pragma unroll 1
  for (int i0 = 0; i0 < NQ; i0 += kBatch) {
    if (4 * i0 >= S) break;
    uint4 kr[kBatch], vr[kBatch];
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {
      const int j = 4 * (i0 + i) + r;
      if (j < S) {
        kr[i] = __ldg(reinterpret_cast<const uint4*>(kb + (long long)j * ldkv));
        vr[i] = __ldg(reinterpret_cast<const uint4*>(vb + (long long)j * ldkv));
      }
    }
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {
      const int j = 4 * (i0 + i) + r;
      const bool ok = j < S;
      float kf[8], vf[8], s = 0.f, dp = 0.f;
      if (ok) {
        unpack8(kr[i], kf);
        unpack8(vr[i], vf);
#pragma unroll
        for (int d = 0; d < 8; ++d) {
          s = fmaf(qv[d], kf[d], s);
          dp = fmaf(dov[d], vf[d], dp);
        }
      }
      s = group8_sum(s);
      dp = group8_sum(dp);
      if (ok) {
        const float p = __expf(s * scale - l_);
        float pm = p;  // P~ = P * keep / (1 - p_drop): delta = dO . O = rowsum(P~ dP) holds with O computed from P~
        if (drop.thr != 0u)
          pm = ((dropout_keep8(drop, drop.row0 + (uint32_t)wid * (S > 128 ? 256u : 128u), (uint32_t)(j >> 3)) >> (j & 7)) & 1u) ? p * drop.scale : 0.f;
        const float dsj = fmaf(pm, dp, -p * delta);
        float a[8], bb[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) {
          a[d] = scale * dsj * qv[d];
          bb[d] = pm * dov[d];
          dqa[d] = fmaf(dsj, kf[d], dqa[d]);
        }
        *reinterpret_cast<uint4*>(dkb + (long long)j * lddkv) = pack8(a);
        *reinterpret_cast<uint4*>(dvb + (long long)j * lddkv) = pack8(bb);
      }
    }
  }
#pragma unroll
  for (int d = 0; d < 8; ++d) dqa[d] = rows4_sum(dqa[d]) * scale;
  if (r == 0) *reinterpret_cast<uint4*>(dq + (long long)b * lddq + h * DH + c * 8) = pack8(dqa);
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" int svla_attn_cls_fwd(svla_ctx* ctx, const void* q, long long ldq, const void* k, const void* v,
                                 long long ldkv, void* o, long long ldo, int dtype, float* lse, int B, int S, int H,
                                 int dh, float scale, const svla_dropout* drop, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && q && k && v && o && lse, "NULL argument");
  SVLA_CHECK_ARG(svla_dropout_ok(drop), "dropout p must be in [0, 1)");
  const DropArgs da = make_drop_args(drop);
  SVLA_CHECK_ARG(dh == DH && S >= 1 && S <= 32 * kMaxChunks, "head dim must be 64 and S <= 256");
  SVLA_CHECK_ARG(ldq % 4 == 0 && ldkv % 4 == 0 && ldo % 4 == 0, "leading dims must be multiples of 4");
  if (B <= 0) return SVLA_OK;
  const long long threads = (long long)B * H * 32;
  if (dtype == SVLA_BF16 && ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 8 == 0 && al16(q) && al16(k) && al16(v) && al16(o)) {
    const unsigned grid = (unsigned)((threads + 127) / 128);
    const auto st = as_stream(stream);
#define SVLA_CLS_FWD(NQ_, DR_)                                                                                      \
  attn_cls_fwd_bf16_kernel<NQ_, DR_, (NQ_ >= 32 && !(DR_)) ? 16 : 8><<<grid, 128, 0, st>>>((const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k,    \
                                                           (const __nv_bfloat16*)v, ldkv, (__nv_bfloat16*)o, ldo, lse, \
                                                           B, S, H, scale, da)
    SVLA_CHECK_ARG(da.thr == 0u || S <= 256, "CLS-row attention with dropout: S <= 256");
    if (da.thr != 0u) {  // mask row of sequence-head w: w * 128 (S <= 128) or w * 256, as in the full-sequence kernels
      if (S <= 64) SVLA_CLS_FWD(16, true);
      else if (S <= 128) SVLA_CLS_FWD(32, true);
      else SVLA_CLS_FWD(64, true);
    } else if (S <= 64) SVLA_CLS_FWD(16, false);
    else if (S <= 128) SVLA_CLS_FWD(32, false);
    else SVLA_CLS_FWD(64, false);
#undef SVLA_CLS_FWD
    SVLA_LAUNCH_CHECK();
    return SVLA_OK;
  }
  SVLA_CHECK_ARG(da.thr == 0u, "CLS-row attention with dropout: bf16, 16-byte aligned operands");
  SVLA_DISPATCH_DTYPE(dtype, T, (attn_cls_fwd_kernel<T><<<(unsigned)((threads + 127) / 128), 128, 0, as_stream(stream)>>>(
                                    (const T*)q, ldq, (const T*)k, (const T*)v, ldkv, (T*)o, ldo, lse, B, S, H, scale)));
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

extern "C" int svla_attn_cls_bwd(svla_ctx* ctx, const void* q, long long ldq, const void* k, const void* v,
                                 long long ldkv, const void* o, const void* d_o, long long ldo, void* dq,
                                 long long lddq, void* dk, void* dv, long long lddkv, int dtype, const float* lse,
                                 int B, int S, int H, int dh, float scale, const svla_dropout* drop, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && q && k && v && o && d_o && dq && dk && dv && lse, "NULL argument");
  SVLA_CHECK_ARG(svla_dropout_ok(drop), "dropout p must be in [0, 1)");
  const DropArgs da = make_drop_args(drop);
  SVLA_CHECK_ARG(dh == DH && S >= 1 && S <= 32 * kMaxChunks, "head dim must be 64 and S <= 256");
  if (B <= 0) return SVLA_OK;
  const long long threads = (long long)B * H * 32;
  if (dtype == SVLA_BF16 && ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 8 == 0 && lddq % 8 == 0 && lddkv % 8 == 0 && al16(q) &&
      al16(k) && al16(v) && al16(o) && al16(d_o) && al16(dq) && al16(dk) && al16(dv)) {
    SVLA_CHECK_ARG(da.thr == 0u || S <= 256, "CLS-row attention with dropout: S <= 256");
    attn_cls_bwd_bf16_kernel<64><<<(unsigned)((threads + 127) / 128), 128, 0, as_stream(stream)>>>(
        (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, ldkv, (const __nv_bfloat16*)o,
        (const __nv_bfloat16*)d_o, ldo, (__nv_bfloat16*)dq, lddq, (__nv_bfloat16*)dk, (__nv_bfloat16*)dv, lddkv, lse, B, S,
        H, scale, da);
    SVLA_LAUNCH_CHECK();
    return SVLA_OK;
  }
  SVLA_CHECK_ARG(da.thr == 0u, "CLS-row attention with dropout: bf16, 16-byte aligned operands");
  SVLA_DISPATCH_DTYPE(dtype, T, (attn_cls_bwd_kernel<T><<<(unsigned)((threads + 127) / 128), 128, 0, as_stream(stream)>>>(
                                    (const T*)q, ldq, (const T*)k, (const T*)v, ldkv, (const T*)o, (const T*)d_o, ldo,
                                    (T*)dq, lddq, (T*)dk, (T*)dv, lddkv, lse, B, S, H, scale)));
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}
