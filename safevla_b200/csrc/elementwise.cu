// Memory-bound glue of the towers: SwiGLU gate, action/in-hand/time embedding, NCHW -> token-major
// transpose(+cast), row gather/scatter with dtype conversion, column sums (bias gradients),
// goal-byte row hashing.  All coalesced, 128-bit where the layout allows.
#include <algorithm>

#include "common.cuh"
#include "fold.cuh"

namespace {

inline int ew_grid(svla_ctx* ctx, long long work_items, int threads, int per_sm = 8) {
  const long long b = (work_items + threads - 1) / threads;
  return (int)std::max<long long>(1, std::min<long long>(b, (long long)ctx->sm_count * per_sm));
}

// ---- SwiGLU -------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) swiglu_fwd_kernel(const T* __restrict__ ab, T* __restrict__ g, long long rows,
                                                         int F) {
  const long long total = rows * (F / 4);
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / (F / 4);
    const int c = (int)(e % (F / 4)) * 4;
    const float4 a = load4<T>(ab + r * 2 * F + c), b = load4<T>(ab + r * 2 * F + F + c);
    float4 o;
    o.x = a.x / (1.f + __expf(-a.x)) * b.x;
    o.y = a.y / (1.f + __expf(-a.y)) * b.y;
    o.z = a.z / (1.f + __expf(-a.z)) * b.z;
    o.w = a.w / (1.f + __expf(-a.w)) * b.w;
    store4<T>(g + r * F + c, o);
  }
}
__device__ __forceinline__ void swiglu_grad(float a, float b, float dg, float& da, float& db) {
  const float s = 1.f / (1.f + __expf(-a));
  const float silu = a * s;
  da = dg * b * (s + silu * (1.f - s));
  db = dg * silu;
}
template <typename T>
__global__ void __launch_bounds__(256) swiglu_bwd_kernel(const T* __restrict__ ab, const T* __restrict__ dg,
                                                         T* __restrict__ dab, long long rows, int F) {
  const long long total = rows * (F / 4);
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / (F / 4);
    const int c = (int)(e % (F / 4)) * 4;
    const float4 a = load4<T>(ab + r * 2 * F + c), b = load4<T>(ab + r * 2 * F + F + c), d = load4<T>(dg + r * F + c);
    float4 da, db;
    swiglu_grad(a.x, b.x, d.x, da.x, db.x);
    swiglu_grad(a.y, b.y, d.y, da.y, db.y);
    swiglu_grad(a.z, b.z, d.z, da.z, db.z);
    swiglu_grad(a.w, b.w, d.w, da.w, db.w);
    store4<T>(dab + r * 2 * F + c, da);
    store4<T>(dab + r * 2 * F + F + c, db);
  }
}

// ---- embeddings + sinusoidal time encoding --------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) embed_time_fwd_kernel(const T* __restrict__ obs, const int64_t* __restrict__ prev,
                                                             const float* __restrict__ masks,
                                                             const int64_t* __restrict__ in_hand,
                                                             const int64_t* __restrict__ time_step,
                                                             const float* __restrict__ Ea, const float* __restrict__ Eh,
                                                             const float* __restrict__ div_term, float* __restrict__ x,
                                                             int T_, int N, int A, int D) {
  const int row = blockIdx.x;  // t*N + n
  const int t = row / N, n = row % N;
  const int ai = (masks[row] != 0.f) ? (int)prev[row] : A;
  const int hi = in_hand ? (int)in_hand[row] : -1;
  const float pos = (float)time_step[row];
  float* out = x + ((long long)n * T_ + t) * D;
  for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) {
    float4 v = load4<T>(obs + (long long)row * D + c);
    const float4 e = __ldg(reinterpret_cast<const float4*>(Ea + (long long)ai * D + c));
    v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
    if (hi >= 0) {
      const float4 h = __ldg(reinterpret_cast<const float4*>(Eh + (long long)hi * D + c));
      v.x += h.x; v.y += h.y; v.z += h.z; v.w += h.w;
    }
    // pe[2i] = sin(pos * div[i]), pe[2i+1] = cos(pos * div[i])   (precise sinf/cosf: parity gate)
    const float a0 = pos * __ldg(div_term + c / 2), a1 = pos * __ldg(div_term + c / 2 + 1);
    v.x += sinf(a0); v.y += cosf(a0); v.z += sinf(a1); v.w += cosf(a1);
    *reinterpret_cast<float4*>(out + c) = v;
  }
}

// dx [N,T,D] -> d_obs [T,N,D]; embedding-table gradients through a deterministic ordered reduction:
// one block per table row k gathers, in (t,n) order, every row whose index is k.
template <typename T>
__global__ void __launch_bounds__(128) embed_time_bwd_copy_kernel(const float* __restrict__ dx, T* __restrict__ dobs,
                                                                  int T_, int N, int D) {
  const int row = blockIdx.x;
  const int t = row / N, n = row % N;
  const float* src = dx + ((long long)n * T_ + t) * D;
  for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4)
    store4<T>(dobs + (long long)row * D + c, *reinterpret_cast<const float4*>(src + c));
}
// table gradients: grid (table rows, row slices); every block walks its slice of (t, n) rows in order and writes a
// partial; a second kernel folds the slices in order (deterministic, no atomics)
constexpr int kEmbedSlices = 32;
__global__ void __launch_bounds__(128) embed_table_grad_kernel(const float* __restrict__ dx,
                                                               const int64_t* __restrict__ prev,
                                                               const float* __restrict__ masks,
                                                               const int64_t* __restrict__ in_hand,
                                                               float* __restrict__ partial, int T_, int N, int A, int D) {
  // blockIdx.x in [0, A+2): action table rows; [A+2, A+5): in-hand table rows
  const int k = blockIdx.x, slice = blockIdx.y;
  const bool is_hand = k >= A + 2;
  const int kk = is_hand ? k - (A + 2) : k;
  const int R = T_ * N, chunk = (R + kEmbedSlices - 1) / kEmbedSlices;
  const int r0 = slice * chunk, r1 = min(R, r0 + chunk);
  float* dst = partial + ((long long)slice * (A + 5) + k) * D;
  for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!(is_hand && in_hand == nullptr)) {
      for (int row = r0; row < r1; ++row) {
        const int idx = is_hand ? (int)in_hand[row] : ((masks[row] != 0.f) ? (int)prev[row] : A);
        if (idx != kk) continue;
        const int t = row / N, n = row % N;
        const float4 v = *reinterpret_cast<const float4*>(dx + ((long long)n * T_ + t) * D + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    *reinterpret_cast<float4*>(dst + c) = acc;
  }
}
__global__ void embed_table_fold_kernel(const float* __restrict__ partial, float* __restrict__ dEa, float* __restrict__ dEh,
                                        int A, int D) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (A + 5) * D) return;
  const int k = e / D, d = e % D;
  float s = 0.f;
  for (int sl = 0; sl < kEmbedSlices; ++sl) s += partial[(long long)sl * (A + 5) * D + e];
  if (k < A + 2) dEa[(long long)k * D + d] += s;
  else if (dEh) dEh[(long long)(k - (A + 2)) * D + d] += s;
}

// ---- NCHW -> token-major -----------------------------------------------------------------
// x [R, C, P] fp32 -> y [R, P, C]; 32x32 shared-memory transpose tiles per r.
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_tokens_kernel(const float* __restrict__ x, T* __restrict__ y, long long R,
                                                             int C, int P) {
  __shared__ float tile[32][33];
  const int tiles_c = (C + 31) / 32, tiles_p = (P + 31) / 32;
  const long long ntiles = R * tiles_c * tiles_p;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (long long ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
    const long long r = ti / (tiles_c * tiles_p);
    const int rem = (int)(ti % (tiles_c * tiles_p));
    const int c0 = (rem / tiles_p) * 32, p0 = (rem % tiles_p) * 32;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + ty + i * 8, p = p0 + tx;
      tile[ty + i * 8][tx] = (c < C && p < P) ? __ldg(x + (r * C + c) * P + p) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = p0 + ty + i * 8, c = c0 + tx;
      if (p < P && c < C) y[(r * P + p) * C + c] = from_f<T>(tile[tx][ty + i * 8]);
    }
  }
}

// ---- row gather / scatter ----------------------------------------------------------------
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) copy_rows_kernel(const TS* __restrict__ src, long long lds, svla_rowmap smap,
                                                        const int64_t* __restrict__ idx, TD* __restrict__ dst,
                                                        long long ldd, svla_rowmap dmap, long long rows, int D,
                                                        int accumulate) {
  const int vec = D / 4;
  const long long total = rows * vec;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / vec;
    const int c = (int)(e % vec) * 4;
    const long long sr = idx ? idx[r] : map_row(smap, r);
    float4 v = load4<TS>(src + sr * lds + c);
    TD* d = dst + map_row(dmap, r) * ldd + c;
    if (accumulate) {
      const float4 o = load4<TD>(d);
      v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
    }
    store4<TD>(d, v);
  }
}
template <typename TD>
__global__ void __launch_bounds__(256) fill_rows_kernel(const float* __restrict__ vecp, TD* __restrict__ dst,
                                                        long long ldd, svla_rowmap dmap, long long rows, int D) {
  const int vec = D / 4;
  const long long total = rows * vec;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / vec;
    const int c = (int)(e % vec) * 4;
    store4<TD>(dst + map_row(dmap, r) * ldd + c, __ldg(reinterpret_cast<const float4*>(vecp + c)));
  }
}

// ---- column sums -------------------------------------------------------------------------
// stage 1: block b sums rows [b*chunk, (b+1)*chunk) for all N columns (thread per 4 columns,
// rows walked sequentially -> coalesced 128-bit row segments); stage 2 folds the partials.
template <typename T>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const T* __restrict__ x, long long M, int N, long long ldx,
                                                             long long chunk, float* __restrict__ partial) {
  // 64 column groups (4 columns each) x 4 row lanes; every thread keeps 4 independent row loads in flight
  __shared__ float4 red[4][64];
  const long long r0 = blockIdx.x * chunk, r1 = min(M, r0 + chunk);
  const int cg = threadIdx.x & 63, rl = threadIdx.x >> 6;
  for (int c0 = 0; c0 < N; c0 += 256) {
    const int c = c0 + cg * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < N) {
      long long r = r0 + rl;
      for (; r + 12 < r1; r += 16) {
        const float4 v0 = load4<T>(x + r * ldx + c), v1 = load4<T>(x + (r + 4) * ldx + c);
        const float4 v2 = load4<T>(x + (r + 8) * ldx + c), v3 = load4<T>(x + (r + 12) * ldx + c);
        s.x += (v0.x + v1.x) + (v2.x + v3.x); s.y += (v0.y + v1.y) + (v2.y + v3.y);
        s.z += (v0.z + v1.z) + (v2.z + v3.z); s.w += (v0.w + v1.w) + (v2.w + v3.w);
      }
      for (; r < r1; r += 4) {
        const float4 v = load4<T>(x + r * ldx + c);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
    }
    __syncthreads();
    red[rl][cg] = s;
    __syncthreads();
    if (rl == 0 && c < N) {
      const float4 a = red[0][cg], b = red[1][cg], d = red[2][cg], e = red[3][cg];
      *reinterpret_cast<float4*>(partial + (long long)blockIdx.x * N + c) =
          make_float4((a.x + b.x) + (d.x + e.x), (a.y + b.y) + (d.y + e.y), (a.z + b.z) + (d.z + e.z),
                      (a.w + b.w) + (d.w + e.w));
    }
  }
}
template <typename T>
__global__ void __launch_bounds__(256) colsum_partial_scalar_kernel(const T* __restrict__ x, long long M, int N,
                                                                    long long ldx, long long chunk,
                                                                    float* __restrict__ partial) {
  const long long r0 = blockIdx.x * chunk, r1 = min(M, r0 + chunk);
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    float s = 0.f;
    for (long long r = r0; r < r1; ++r) s += to_f<T>(x[r * ldx + c]);
    partial[(long long)blockIdx.x * N + c] = s;
  }
}
__global__ void colsum_fold_kernel(const float* __restrict__ partial, int nb, int N, float* __restrict__ out,
                                   int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  float s = 0.f;
  for (int b = 0; b < nb; ++b) s += partial[(long long)b * N + c];
  out[c] = accumulate ? out[c] + s : s;
}

// ---- goal-byte row hash (FNV-1a over 8-byte words, then avalanche) ---------------------------
__global__ void __launch_bounds__(256) hash_rows_kernel(const uint8_t* __restrict__ rows, long long R, int L,
                                                        uint64_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (r >= R) return;
  const uint8_t* p = rows + r * L;
  // each lane folds words lane, lane+32, ...; lanes are then combined in lane order
  uint64_t h = 1469598103934665603ull ^ (uint64_t)lane;
  const int nwords = L / 8;
  const bool al = ((reinterpret_cast<uintptr_t>(p) & 7) == 0);
  for (int wv = lane; wv < nwords; wv += 32) {
    uint64_t v;
    if (al) v = __ldg(reinterpret_cast<const unsigned long long*>(p) + wv);
    else {
      v = 0;
      for (int b = 0; b < 8; ++b) v |= (uint64_t)p[wv * 8 + b] << (8 * b);
    }
    h = (h ^ v) * 1099511628211ull;
    h ^= h >> 29;
  }
  for (int b = nwords * 8 + lane; b < L; b += 32) h = (h ^ p[b]) * 1099511628211ull;
  // ordered combine
  uint64_t acc = 0;
  for (int l = 0; l < 32; ++l) {
    const uint64_t hl = __shfl_sync(0xffffffffu, h, l);
    acc = (acc ^ hl) * 0x9E3779B97F4A7C15ull;
    acc ^= acc >> 32;
  }
  if (lane == 0) out[r] = acc;
}

__global__ void __launch_bounds__(256) scale_by_kernel(float* __restrict__ x, long long n, const float* __restrict__ s) {
  const float f = *s;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] *= f;
}


// ---- stand-alone dropout ------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) dropout_rows_kernel(const TI* __restrict__ x, long long ldx, TO* __restrict__ out,
                                                           long long ldo, long long rows, int cols, DropArgs d) {
  const int g8 = cols >> 3;  // column groups of eight: one Philox call each
  const long long total = rows * g8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / g8;
    const int cg = (int)(i - r * g8);
    const uint32_t keep = dropout_keep8(d, d.row0 + (uint32_t)r * d.row_stride, (uint32_t)cg);
    const float4 a = load4<TI>(x + r * ldx + cg * 8), b = load4<TI>(x + r * ldx + cg * 8 + 4);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = ((keep >> e) & 1u) ? v[e] * d.scale : 0.f;
    store4<TO>(out + r * ldo + cg * 8, make_float4(o[0], o[1], o[2], o[3]));
    store4<TO>(out + r * ldo + cg * 8 + 4, make_float4(o[4], o[5], o[6], o[7]));
  }
}

}  // namespace

extern "C" int svla_scale_by(svla_ctx* ctx, float* x, long long n, const float* scale_dev, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && x && scale_dev, "NULL argument");
  if (n <= 0) return SVLA_OK;
  scale_by_kernel<<<ew_grid(ctx, n, 256), 256, 0, as_stream(stream)>>>(x, n, scale_dev);
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

#define DISPATCH2E(dtA, TA_, dtB, TB_, ...)                                             \
  do {                                                                                  \
    if (dtA == SVLA_F32 && dtB == SVLA_F32) { using TA_ = float; using TB_ = float; __VA_ARGS__; } \
    else if (dtA == SVLA_F32 && dtB == SVLA_BF16) { using TA_ = float; using TB_ = __nv_bfloat16; __VA_ARGS__; } \
    else if (dtA == SVLA_BF16 && dtB == SVLA_F32) { using TA_ = __nv_bfloat16; using TB_ = float; __VA_ARGS__; } \
    else if (dtA == SVLA_BF16 && dtB == SVLA_BF16) { using TA_ = __nv_bfloat16; using TB_ = __nv_bfloat16; __VA_ARGS__; } \
    else { svla_set_error("bad dtype"); return SVLA_ERR_BAD_ARG; }                      \
  } while (0)

extern "C" int svla_dropout_rows(svla_ctx* ctx, const void* x, int dtype_x, long long ldx, void* out, int dtype_out,
                                 long long ldo, long long rows, int cols, const svla_dropout* drop, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && x && out && drop, "NULL argument");
  SVLA_CHECK_ARG(svla_dropout_ok(drop), "dropout p must be in [0, 1)");
  SVLA_CHECK_ARG(rows >= 0 && cols >= 0 && cols % 8 == 0 && ldx % 4 == 0 && ldo % 4 == 0, "cols must be a multiple of 8");
  if (rows == 0 || cols == 0) return SVLA_OK;
  const DropArgs d = make_drop_args(drop);
  DISPATCH2E(dtype_x, TI, dtype_out, TO,
             (dropout_rows_kernel<TI, TO><<<ew_grid(ctx, rows * (cols / 8), 256), 256, 0, as_stream(stream)>>>(
                 reinterpret_cast<const TI*>(x), ldx, reinterpret_cast<TO*>(out), ldo, rows, cols, d)));
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

extern "C" int svla_swiglu_fwd(svla_ctx* ctx, const void* ab, void* g, int dtype, long long rows, int F,
                               svla_stream stream) {
  SVLA_CHECK_ARG(ctx && ab && g, "NULL argument");
  SVLA_CHECK_ARG(F % 4 == 0, "F must be a multiple of 4");
  if (rows <= 0) return SVLA_OK;
  SVLA_DISPATCH_DTYPE(dtype, T, (swiglu_fwd_kernel<T><<<ew_grid(ctx, rows * (F / 4), 256), 256, 0, as_stream(stream)>>>(
                                    (const T*)ab, (T*)g, rows, F)));
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}
extern "C" int svla_swiglu_bwd(svla_ctx* ctx, const void* ab, const void* dg, void* dab, int dtype, long long rows,
                               int F, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && ab && dg && dab, "NULL argument");
  SVLA_CHECK_ARG(F % 4 == 0, "F must be a multiple of 4");
  if (rows <= 0) return SVLA_OK;
  SVLA_DISPATCH_DTYPE(dtype, T, (swiglu_bwd_kernel<T><<<ew_grid(ctx, rows * (F / 4), 256), 256, 0, as_stream(stream)>>>(
                                    (const T*)ab, (const T*)dg, (T*)dab, rows, F)));
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

extern "C" int svla_embed_time_fwd(svla_ctx* ctx, const void* obs_embed, int dtype_in, const int64_t* prev_actions,
                                   const float* masks, const int64_t* in_hand, const int64_t* time_step,
                                   const float* E_a, const float* E_h, const float* div_term, float* x_out, int T, int N,
                                   int A, int D, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && obs_embed && prev_actions && masks && time_step && E_a && div_term && x_out, "NULL argument");
  SVLA_CHECK_ARG(!in_hand || E_h, "in_hand without its table");
  SVLA_CHECK_ARG(D % 4 == 0, "D must be a multiple of 4");
  if (T * N <= 0) return SVLA_OK;
  SVLA_DISPATCH_DTYPE(dtype_in, TI, (embed_time_fwd_kernel<TI><<<T * N, 128, 0, as_stream(stream)>>>(
                                        (const TI*)obs_embed, prev_actions, masks, in_hand, time_step, E_a, E_h,
                                        div_term, x_out, T, N, A, D)));
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

extern "C" int svla_embed_time_bwd(svla_ctx* ctx, const float* dx, const int64_t* prev_actions, const float* masks,
                                   const int64_t* in_hand, void* d_obs_embed, int dtype_out, float* dE_a, float* dE_h,
                                   int T, int N, int A, int D, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && dx && prev_actions && masks && d_obs_embed && dE_a, "NULL argument");
  SVLA_CHECK_ARG(D % 4 == 0, "D must be a multiple of 4");
  if (T * N <= 0) return SVLA_OK;
  SVLA_DISPATCH_DTYPE(dtype_out, TO, (embed_time_bwd_copy_kernel<TO><<<T * N, 128, 0, as_stream(stream)>>>(
                                         dx, (TO*)d_obs_embed, T, N, D)));
  SVLA_LAUNCH_CHECK();
  float* partial = reinterpret_cast<float*>(ctx->ws);
  SVLA_CHECK_ARG((size_t)kEmbedSlices * (A + 5) * D * sizeof(float) <= ctx->ws_bytes, "workspace too small");
  embed_table_grad_kernel<<<dim3(A + 5, kEmbedSlices), 128, 0, as_stream(stream)>>>(dx, prev_actions, masks, in_hand,
                                                                                     partial, T, N, A, D);
  SVLA_LAUNCH_CHECK();
  embed_table_fold_kernel<<<((A + 5) * D + 255) / 256, 256, 0, as_stream(stream)>>>(partial, dE_a, dE_h, A, D);
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

extern "C" int svla_nchw_to_tokens(svla_ctx* ctx, const float* x, void* y, int dtype_out, long long R, int C, int P,
                                   svla_stream stream) {
  SVLA_CHECK_ARG(ctx && x && y, "NULL argument");
  if (R <= 0) return SVLA_OK;
  const long long ntiles = R * ((C + 31) / 32) * ((P + 31) / 32);
  const int grid = (int)std::min<long long>(ntiles, (long long)ctx->sm_count * 16);
  SVLA_DISPATCH_DTYPE(dtype_out, TO,
                      (nchw_to_tokens_kernel<TO><<<grid, 256, 0, as_stream(stream)>>>(x, (TO*)y, R, C, P)));
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

extern "C" int svla_copy_rows(svla_ctx* ctx, const void* src, int dtype_src, long long lds, svla_rowmap smap,
                              const int64_t* idx, void* dst, int dtype_dst, long long ldd, svla_rowmap dmap,
                              long long rows, int D, int accumulate, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && src && dst, "NULL argument");
  SVLA_CHECK_ARG(D % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0, "D and leading dims must be multiples of 4");
  if (rows <= 0) return SVLA_OK;
  DISPATCH2E(dtype_src, TS, dtype_dst, TD,
             (copy_rows_kernel<TS, TD><<<ew_grid(ctx, rows * (D / 4), 256), 256, 0, as_stream(stream)>>>(
                 (const TS*)src, lds, smap, idx, (TD*)dst, ldd, dmap, rows, D, accumulate)));
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

extern "C" int svla_fill_rows(svla_ctx* ctx, const float* vec, void* dst, int dtype_dst, long long ldd,
                              svla_rowmap dmap, long long rows, int D, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && vec && dst, "NULL argument");
  SVLA_CHECK_ARG(D % 4 == 0 && ldd % 4 == 0, "D and ldd must be multiples of 4");
  if (rows <= 0) return SVLA_OK;
  SVLA_DISPATCH_DTYPE(dtype_dst, TD, (fill_rows_kernel<TD><<<ew_grid(ctx, rows * (D / 4), 256), 256, 0, as_stream(stream)>>>(
                                         vec, (TD*)dst, ldd, dmap, rows, D)));
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

extern "C" int svla_colsum(svla_ctx* ctx, const void* x, int dtype, long long M, int N, long long ldx, float* out,
                           int accumulate, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && x && out, "NULL argument");
  if (M <= 0) return SVLA_OK;
  const bool vec = (N % 4 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  int nb = (int)std::min<long long>((M + 63) / 64, (long long)ctx->sm_count * 4);
  nb = (int)std::min<size_t>((size_t)nb, ctx->ws_bytes / (sizeof(float) * (size_t)N));
  const long long chunk = (M + nb - 1) / nb;
  nb = (int)((M + chunk - 1) / chunk);
  float* partial = reinterpret_cast<float*>(ctx->ws);
  if (vec) {
    SVLA_DISPATCH_DTYPE(dtype, T, (colsum_partial_kernel<T><<<nb, 256, 0, as_stream(stream)>>>((const T*)x, M, N, ldx,
                                                                                               chunk, partial)));
  } else {
    SVLA_DISPATCH_DTYPE(dtype, T, (colsum_partial_scalar_kernel<T><<<nb, 256, 0, as_stream(stream)>>>(
                                      (const T*)x, M, N, ldx, chunk, partial)));
  }
  SVLA_LAUNCH_CHECK();
  svla_launch_fold(partial, nb, N, N, out, nullptr, nullptr, accumulate, as_stream(stream));
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

extern "C" int svla_hash_rows(svla_ctx* ctx, const uint8_t* rows, long long R, int L, uint64_t* out,
                              svla_stream stream) {
  SVLA_CHECK_ARG(ctx && rows && out, "NULL argument");
  if (R <= 0) return SVLA_OK;
  const long long threads = R * 32;
  hash_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, as_stream(stream)>>>(rows, R, L, out);
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}


// ---- episode-cost bookkeeping of the storage (Jc of the Lagrange update) -------------------------------------------
// One block, fixed summation order (deterministic): episode_cost[n] += cost[n]; samplers whose episode ended at this
// step (mask_next[n] == 0) add their episode total to sum_cnt[0], bump sum_cnt[1] and restart at zero.
__global__ void __launch_bounds__(256) episode_cost_step_kernel(const float* __restrict__ costs,
                                                                const float* __restrict__ mask_next,
                                                                float* __restrict__ episode_cost,
                                                                float* __restrict__ sum_cnt, int N) {
  __shared__ float red[32];
  float s = 0.f, c = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float e = episode_cost[n] + costs[n];
    if (mask_next[n] == 0.f) {
      s += e;
      c += 1.f;
      e = 0.f;
    }
    episode_cost[n] = e;
  }
  s = block_sum(s, red);
  c = block_sum(c, red);
  if (threadIdx.x == 0) {
    sum_cnt[0] += s;
    sum_cnt[1] += c;
  }
}

extern "C" int svla_episode_cost_step(svla_ctx* ctx, const float* costs, const float* mask_next, float* episode_cost,
                                      float* sum_cnt, int N, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && costs && mask_next && episode_cost && sum_cnt, "NULL argument");
  if (N <= 0) return SVLA_OK;
  episode_cost_step_kernel<<<1, 256, 0, as_stream(stream)>>>(costs, mask_next, episode_cost, sum_cnt, N);
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}


// ---- K cost channels: fold the per-channel cost advantages and multipliers into the single pair the fused loss takes
//   (A - sum_k lambda_k A_c,k) / (1 + sum_k lambda_k)  ==  (A - L * A_eff) / (1 + L),  L = sum_k lambda_k,
//   A_eff = sum_k lambda_k A_c,k / L   (A_eff = 0 when L = 0).   c_adv is channel-major [K, R]; 16-byte accesses.
__global__ void __launch_bounds__(256) combine_cost_adv_kernel(const float* __restrict__ c_adv,
                                                               const float* __restrict__ lambdas, int K, long long R,
                                                               float* __restrict__ out, float* __restrict__ lambda_eff) {
  float lam[8];
  float L = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    lam[k] = k < K ? lambdas[k] : 0.f;
    L += lam[k];
  }
  const float inv = L > 0.f ? 1.f / L : 0.f;
  if (blockIdx.x == 0 && threadIdx.x == 0) lambda_eff[0] = L;
  const long long R4 = R >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < R4; i += (long long)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < K; ++k) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(c_adv + (long long)k * R) + i);
      acc.x = fmaf(lam[k], v.x, acc.x); acc.y = fmaf(lam[k], v.y, acc.y);
      acc.z = fmaf(lam[k], v.z, acc.z); acc.w = fmaf(lam[k], v.w, acc.w);
    }
    reinterpret_cast<float4*>(out)[i] = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
  }
  if (blockIdx.x == 0) {  // tail (R % 4 rows)
    for (long long i = (R4 << 2) + threadIdx.x; i < R; i += blockDim.x) {
      float acc = 0.f;
      for (int k = 0; k < K; ++k) acc = fmaf(lam[k], c_adv[(long long)k * R + i], acc);
      out[i] = acc * inv;
    }
  }
}

extern "C" int svla_combine_cost_advantages(svla_ctx* ctx, const float* c_adv, const float* lambdas_dev, int K,
                                            long long R, float* c_adv_eff, float* lambda_eff_dev, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && c_adv && lambdas_dev && c_adv_eff && lambda_eff_dev, "NULL argument");
  SVLA_CHECK_ARG(K >= 1 && K <= 8, "1 <= K <= 8 cost channels");
  if (R <= 0) return SVLA_OK;
  // channel k starts at c_adv + k * R: 16-byte aligned vector loads need R % 4 == 0 for k > 0
  SVLA_CHECK_ARG(K == 1 || R % 4 == 0, "R must be a multiple of 4 for K > 1 (16-byte channel alignment)");
  const int grid = (int)std::min<long long>((R / 4 + 255) / 256 + 1, (long long)ctx->sm_count * 8);
  combine_cost_adv_kernel<<<grid, 256, 0, as_stream(stream)>>>(c_adv, lambdas_dev, K, R, c_adv_eff, lambda_eff_dev);
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}
