// Small-sequence multi-head attention on CUDA cores (fp32 math): the parity-mode kernel for the
// three attention flavours of the towers (S <= 256, head dim 64):
//   FULL         fusion block   softmax(q k^T / 8) v                     (nn.MultiheadAttention)
//   TRAJ_CAUSAL  decoder        mask[i,j] = traj[i]==traj[j] && j<=i, computed from traj_index in
//                               shared memory -- the [N,1,T,T] mask of allenact_dino_transformer.py:399-402
//                               is never materialised
//   T5_BIAS      T5 encoder     unscaled scores + relative bias [H,S,S] + key padding mask
// One CTA per (sequence, head); Q/K/V (and dO) of that head live in shared memory, rows padded to
// kill bank conflicts.  Forward saves the log-sum-exp per row; backward recomputes P from it.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int DH = 64;
constexpr int kThreads = 128;  // 4 warps
constexpr int kMaxS = 256;
constexpr int kMaxChunks = kMaxS / 32;

template <typename T> struct Pad { static constexpr int v = 1; };
template <> struct Pad<__nv_bfloat16> { static constexpr int v = 2; };

struct AttnArgs {
  int mode;
  const void* q; const void* k; const void* v; long long ld;
  void* o; const void* d_o; long long ldo;
  void* dq; void* dk; void* dv; long long ldd;
  float* lse;
  const int64_t* traj; const float* bias; const int64_t* keymask;
  int B, S, H;
  float scale;
};

template <typename T>
__device__ __forceinline__ void load_head(T* dst, const T* src, long long ld, int S, int lds) {
  // src points at row 0 of this (sequence, head); copy [S, 64] -> dst[S][lds]
  for (int e = threadIdx.x; e < S * (DH / 4); e += kThreads) {
    const int r = e / (DH / 4), c = (e % (DH / 4)) * 4;
    const float4 v = load4<T>(src + (long long)r * ld + c);
    T* d = dst + r * lds + c;
    d[0] = from_f<T>(v.x); d[1] = from_f<T>(v.y); d[2] = from_f<T>(v.z); d[3] = from_f<T>(v.w);
  }
}

__device__ __forceinline__ bool allowed(const AttnArgs& a, const int* s_traj, const int* s_kmask, int i, int j) {
  if (a.mode == SVLA_ATTN_TRAJ_CAUSAL) return j <= i && s_traj[i] == s_traj[j];
  if (a.mode == SVLA_ATTN_T5_BIAS) return s_kmask == nullptr || s_kmask[j] != 0;
  return true;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) attn_fwd_kernel(AttnArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int LDS = DH + Pad<T>::v;
  const int S = a.S, b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  T* sK = reinterpret_cast<T*>(smem_raw);
  T* sV = sK + S * LDS;
  float* sQ = reinterpret_cast<float*>(sV + S * LDS);  // [4][64] (2*S*LDS elements: 4-byte aligned)
  float* sP = sQ + 4 * DH;                                             // [4][S]
  int* sTraj = reinterpret_cast<int*>(sP + 4 * S);                     // [S] (traj or keymask)
  const long long row0 = (long long)b * S;
  const T* gq = reinterpret_cast<const T*>(a.q) + row0 * a.ld + h * DH;
  load_head<T>(sK, reinterpret_cast<const T*>(a.k) + row0 * a.ld + h * DH, a.ld, S, LDS);
  load_head<T>(sV, reinterpret_cast<const T*>(a.v) + row0 * a.ld + h * DH, a.ld, S, LDS);
  if (a.mode == SVLA_ATTN_TRAJ_CAUSAL)
    for (int i = threadIdx.x; i < S; i += kThreads) sTraj[i] = (int)a.traj[row0 + i];
  if (a.mode == SVLA_ATTN_T5_BIAS && a.keymask)
    for (int i = threadIdx.x; i < S; i += kThreads) sTraj[i] = (int)a.keymask[row0 + i];
  __syncthreads();
  const int* s_kmask = (a.mode == SVLA_ATTN_T5_BIAS && a.keymask) ? sTraj : nullptr;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nch = (S + 31) / 32;
  float* myQ = sQ + w * DH;
  float* myP = sP + w * S;
  for (int i = w; i < S; i += 4) {
    __syncwarp();
    myQ[lane] = to_f<T>(gq[(long long)i * a.ld + lane]);
    myQ[lane + 32] = to_f<T>(gq[(long long)i * a.ld + lane + 32]);
    __syncwarp();
    float sc[kMaxChunks];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      sc[c] = -INFINITY;
      const int j = c * 32 + lane;
      if (c < nch && j < S && allowed(a, sTraj, s_kmask, i, j)) {
        const T* kr = sK + j * LDS;
        float acc = 0.f;
#pragma unroll 16
        for (int d = 0; d < DH; ++d) acc = fmaf(myQ[d], to_f<T>(kr[d]), acc);
        acc *= a.scale;
        if (a.mode == SVLA_ATTN_T5_BIAS && a.bias) acc += __ldg(a.bias + ((long long)h * S + i) * S + j);
        sc[c] = acc;
        mx = fmaxf(mx, acc);
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      const int j = c * 32 + lane;
      if (c < nch && j < S) {
        const float p = (sc[c] == -INFINITY) ? 0.f : __expf(sc[c] - mx);
        myP[j] = p;
        sum += p;
      }
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.f / sum;
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < S; ++j) {
      const float p = myP[j];
      o0 = fmaf(p, to_f<T>(sV[j * LDS + lane]), o0);
      o1 = fmaf(p, to_f<T>(sV[j * LDS + lane + 32]), o1);
    }
    T* go = reinterpret_cast<T*>(a.o) + (row0 + i) * a.ldo + h * DH;
    go[lane] = from_f<T>(o0 * inv);
    go[lane + 32] = from_f<T>(o1 * inv);
    if (lane == 0 && a.lse) a.lse[((long long)b * a.H + h) * S + i] = mx + __logf(sum);
  }
}

// Backward.  Phase A (key ownership): dK_j, dV_j.  Phase B (query ownership): dQ_i.
template <typename T>
__global__ void __launch_bounds__(kThreads) attn_bwd_kernel(AttnArgs a) {
  // Two [S, 64] tiles are resident at a time: phase A (one key per warp) needs Q and dO of every query plus ONE row of
  // K and V per warp; phase B (one query per warp) needs K and V of every key plus ONE row of Q and dO per warp.  With
  // all four tiles resident fp32 stopped at S = 208; this way S = 256 (the 256-step decoder window of BASELINE config
  // 4) fits in fp32 as well.
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int LDS = DH + Pad<T>::v;
  const int S = a.S, b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  T* sT0 = reinterpret_cast<T*>(smem_raw);      // phase A: Q, phase B: K
  T* sT1 = sT0 + S * LDS;                       // phase A: dO, phase B: V
  float* sLse = reinterpret_cast<float*>(sT1 + S * LDS);  // [S]
  float* sDelta = sLse + S;                                                    // [S]
  float* sBufA = sDelta + S;                                                   // [4][S]  p column / dS row
  float* sBufB = sBufA + 4 * S;                                                // [4][S]  dS column
  float* sRow = sBufB + 4 * S;                                                 // [4][2][64] per-warp row pair
  int* sTraj = reinterpret_cast<int*>(sRow + 4 * 2 * DH);                      // [S]
  const long long row0 = (long long)b * S;
  const T* gq = reinterpret_cast<const T*>(a.q) + row0 * a.ld + h * DH;
  const T* gk = reinterpret_cast<const T*>(a.k) + row0 * a.ld + h * DH;
  const T* gv = reinterpret_cast<const T*>(a.v) + row0 * a.ld + h * DH;
  const T* gdo = reinterpret_cast<const T*>(a.d_o) + row0 * a.ldo + h * DH;
  T* sQ = sT0;
  T* sdO = sT1;
  load_head<T>(sQ, gq, a.ld, S, LDS);
  load_head<T>(sdO, gdo, a.ldo, S, LDS);
  if (a.mode == SVLA_ATTN_TRAJ_CAUSAL)
    for (int i = threadIdx.x; i < S; i += kThreads) sTraj[i] = (int)a.traj[row0 + i];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // delta_i = dO_i . O_i ; lse_i
  const T* go = reinterpret_cast<const T*>(a.o) + row0 * a.ldo + h * DH;
  __syncthreads();
  for (int i = w; i < S; i += 4) {
    float d = to_f<T>(sdO[i * LDS + lane]) * to_f<T>(go[(long long)i * a.ldo + lane]) +
              to_f<T>(sdO[i * LDS + lane + 32]) * to_f<T>(go[(long long)i * a.ldo + lane + 32]);
    d = warp_sum(d);
    if (lane == 0) {
      sDelta[i] = d;
      sLse[i] = a.lse[((long long)b * a.H + h) * S + i];
    }
  }
  __syncthreads();
  const int nch = (S + 31) / 32;
  float* colP = sBufA + w * S;
  float* colDS = sBufB + w * S;
  float* rowA = sRow + w * 2 * DH;  // K_j (phase A) / Q_i (phase B)
  float* rowB = rowA + DH;          // V_j (phase A) / dO_i (phase B)
  // ---- phase A: each warp owns keys j = w, w+4, ...
  for (int j = w; j < S; j += 4) {
    __syncwarp();
    rowA[lane] = to_f<T>(gk[(long long)j * a.ld + lane]);
    rowA[lane + 32] = to_f<T>(gk[(long long)j * a.ld + lane + 32]);
    rowB[lane] = to_f<T>(gv[(long long)j * a.ld + lane]);
    rowB[lane + 32] = to_f<T>(gv[(long long)j * a.ld + lane + 32]);
    __syncwarp();
    for (int c = 0; c < nch; ++c) {
      const int i = c * 32 + lane;
      if (i < S) {
        float p = 0.f, ds = 0.f;
        if (allowed(a, sTraj, nullptr, i, j)) {
          float s = 0.f, dp = 0.f;
          const T* qr = sQ + i * LDS;
          const T* dor = sdO + i * LDS;
#pragma unroll 16
          for (int d = 0; d < DH; ++d) {
            s = fmaf(to_f<T>(qr[d]), rowA[d], s);
            dp = fmaf(to_f<T>(dor[d]), rowB[d], dp);
          }
          p = __expf(s * a.scale - sLse[i]);
          ds = p * (dp - sDelta[i]);
        }
        colP[i] = p;
        colDS[i] = ds;
      }
    }
    __syncwarp();
    float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
    for (int i = 0; i < S; ++i) {
      const float p = colP[i], ds = colDS[i];
      k0 = fmaf(ds, to_f<T>(sQ[i * LDS + lane]), k0);
      k1 = fmaf(ds, to_f<T>(sQ[i * LDS + lane + 32]), k1);
      v0 = fmaf(p, to_f<T>(sdO[i * LDS + lane]), v0);
      v1 = fmaf(p, to_f<T>(sdO[i * LDS + lane + 32]), v1);
    }
    T* gdk = reinterpret_cast<T*>(a.dk) + (row0 + j) * a.ldd + h * DH;
    T* gdv = reinterpret_cast<T*>(a.dv) + (row0 + j) * a.ldd + h * DH;
    gdk[lane] = from_f<T>(k0 * a.scale);
    gdk[lane + 32] = from_f<T>(k1 * a.scale);
    gdv[lane] = from_f<T>(v0);
    gdv[lane + 32] = from_f<T>(v1);
  }
  // ---- phase B: the tiles now hold K and V; each warp owns queries i = w, w+4, ...
  __syncthreads();
  T* sK = sT0;
  T* sV = sT1;
  load_head<T>(sK, gk, a.ld, S, LDS);
  load_head<T>(sV, gv, a.ld, S, LDS);
  __syncthreads();
  float* rowDS = sBufA + w * S;
  for (int i = w; i < S; i += 4) {
    __syncwarp();
    rowA[lane] = to_f<T>(gq[(long long)i * a.ld + lane]);
    rowA[lane + 32] = to_f<T>(gq[(long long)i * a.ld + lane + 32]);
    rowB[lane] = to_f<T>(gdo[(long long)i * a.ldo + lane]);
    rowB[lane + 32] = to_f<T>(gdo[(long long)i * a.ldo + lane + 32]);
    __syncwarp();
    for (int c = 0; c < nch; ++c) {
      const int j = c * 32 + lane;
      if (j < S) {
        float ds = 0.f;
        if (allowed(a, sTraj, nullptr, i, j)) {
          float s = 0.f, dp = 0.f;
          const T* kr = sK + j * LDS;
          const T* vr = sV + j * LDS;
#pragma unroll 16
          for (int d = 0; d < DH; ++d) {
            s = fmaf(rowA[d], to_f<T>(kr[d]), s);
            dp = fmaf(rowB[d], to_f<T>(vr[d]), dp);
          }
          ds = __expf(s * a.scale - sLse[i]) * (dp - sDelta[i]);
        }
        rowDS[j] = ds;
      }
    }
    __syncwarp();
    float q0 = 0.f, q1 = 0.f;
    for (int j = 0; j < S; ++j) {
      const float ds = rowDS[j];
      q0 = fmaf(ds, to_f<T>(sK[j * LDS + lane]), q0);
      q1 = fmaf(ds, to_f<T>(sK[j * LDS + lane + 32]), q1);
    }
    T* gdq = reinterpret_cast<T*>(a.dq) + (row0 + i) * a.ldd + h * DH;
    gdq[lane] = from_f<T>(q0 * a.scale);
    gdq[lane + 32] = from_f<T>(q1 * a.scale);
  }
}

template <typename T> size_t fwd_smem(int S) {
  constexpr int LDS = DH + Pad<T>::v;
  size_t e = (size_t)2 * S * LDS;
  return e * sizeof(T) + sizeof(float) * (4 * DH + 4 * S) + sizeof(int) * S + 16;
}
template <typename T> size_t bwd_smem(int S) {
  constexpr int LDS = DH + Pad<T>::v;
  size_t e = (size_t)2 * S * LDS;
  return e * sizeof(T) + sizeof(float) * (2 * S + 8 * S + 8 * DH) + sizeof(int) * S + 16;
}

}  // namespace

bool svla_attn_tc_supported(int mode, int dtype, int S, int dh, long long ld, long long ldo, const void* q,
                            const void* k, const void* v, const void* o);
int svla_attn_tc_fwd(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld, void* o,
                     long long ldo, float* lse, const int64_t* traj, int B, int S, int H, float scale, cudaStream_t st);
int svla_attn_tc_bwd(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld, const void* o,
                     const void* d_o, long long ldo, void* dq, void* dk, void* dv, long long ldd, const float* lse,
                     const int64_t* traj, int B, int S, int H, float scale, cudaStream_t st);
// attn_tc2.cu: 128 < S <= 256
bool svla_attn_tc2_supported(int mode, int dtype, int S, int dh, long long ld, long long ldo, const void* q,
                             const void* k, const void* v, const void* o);
int svla_attn_tc2_fwd(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld, void* o,
                      long long ldo, float* lse, const int64_t* traj, int B, int S, int H, float scale, cudaStream_t st);
int svla_attn_tc2_bwd(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld, const void* o,
                      const void* d_o, long long ldo, void* dq, void* dk, void* dv, long long ldd, const float* lse,
                      const int64_t* traj, int B, int S, int H, float scale, cudaStream_t st);
// attn_flash.cu: S > 256, forward only
bool svla_attn_flash_supported(int mode, int dtype, int S, int dh, long long ld, long long ldo, const void* q,
                               const void* k, const void* v, const void* o);
int svla_attn_flash_fwd(svla_ctx* ctx, const void* q, const void* k, const void* v, long long ld, void* o, long long ldo,
                        float* lse, int B, int S, int H, float scale, cudaStream_t st);
extern "C" int svla_attn_decode(svla_ctx* ctx, const void* q, long long ldq, const void* cache_k, const void* cache_v,
                                long long cache_rows, long long ldc, const int64_t* time_step, int pos, void* o,
                                long long ldo, int dtype, int N, int H, int dh, float scale, int q_per_cache,
                                svla_stream stream);
static inline bool lse_needed_beyond_flash(const float* lse, int dtype) { return lse != nullptr && dtype != SVLA_BF16; }
// attn_ws.cu: warp-specialised persistent tcgen05 kernels (S <= 128), the default for bf16
bool svla_attn_ws_supported(int mode, int dtype, int S, int dh, long long ld, long long ldo, const void* q, const void* k,
                            const void* v, const void* o);
int svla_attn_ws_fwd(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld, void* o,
                     long long ldo, float* lse, const int64_t* traj, int B, int S, int H, float scale, cudaStream_t st);
int svla_attn_ws_bwd(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld, const void* d_o,
                     long long ldo, void* dq, void* dk, void* dv, long long ldd, const float* lse, const int64_t* traj,
                     int B, int S, int H, float scale, cudaStream_t st);
// 0 auto, 1 CUDA-core kernel, 2 tcgen05 kernel (error when unsupported), 3 the round-1 one-CTA-per-item tcgen05 kernels
// of attn_tc.cu instead of the warp-specialised ones (A/B comparisons)
static int g_attn_impl = 0;
extern "C" int svla_set_attn_impl(int impl) {
  g_attn_impl = impl;
  return SVLA_OK;
}

extern "C" int svla_attn_fwd(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld,
                             void* o, long long ldo, int dtype, float* lse, const int64_t* traj, const float* bias,
                             const int64_t* keymask, int B, int S, int H, int dh, float scale, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && q && k && v && o, "NULL argument");
  SVLA_CHECK_ARG(dh == DH, "head dim must be 64");
  SVLA_CHECK_ARG(S >= 1, "S must be >= 1");
  SVLA_CHECK_ARG(mode != SVLA_ATTN_TRAJ_CAUSAL || traj, "TRAJ_CAUSAL needs traj");
  SVLA_CHECK_ARG(ld % 4 == 0 && ldo % 4 == 0, "leading dims must be multiples of 4");
  if (B <= 0) return SVLA_OK;
  if (S > kMaxS) {
    // longer than the single-/two-tile kernels take (the vision preprocessor's 257- / 433-token ViT blocks):
    // unmasked forward only -- tcgen05 flash kernel for bf16, the warp-per-query cache kernel otherwise
    SVLA_CHECK_ARG(mode == SVLA_ATTN_FULL && !lse_needed_beyond_flash(lse, dtype), "S > 256: unmasked forward without lse (bf16: lse available)");
    if (g_attn_impl != 1 && svla_attn_flash_supported(mode, dtype, S, dh, ld, ldo, q, k, v, o))
      return svla_attn_flash_fwd(ctx, q, k, v, ld, o, ldo, lse, B, S, H, scale, as_stream(stream));
    SVLA_CHECK_ARG(lse == nullptr, "S > 256 on the CUDA-core path does not produce lse");
    return svla_attn_decode(ctx, q, ld, k, v, S, ld, nullptr, S - 1, o, ldo, dtype, B * S, H, dh, scale, S, stream);
  }
  {
    const bool tc2_ok = svla_attn_tc2_supported(mode, dtype, S, dh, ld, ldo, q, k, v, o);
    if (tc2_ok && g_attn_impl != 1)
      return svla_attn_tc2_fwd(ctx, mode, q, k, v, ld, o, ldo, lse, traj, B, S, H, scale, as_stream(stream));
    if (g_attn_impl != 1 && g_attn_impl != 3 && svla_attn_ws_supported(mode, dtype, S, dh, ld, ldo, q, k, v, o))
      return svla_attn_ws_fwd(ctx, mode, q, k, v, ld, o, ldo, lse, traj, B, S, H, scale, as_stream(stream));
    const bool tc_ok = svla_attn_tc_supported(mode, dtype, S, dh, ld, ldo, q, k, v, o);
    if (g_attn_impl == 2 && !tc_ok) {
      svla_set_error("svla_attn_fwd: tcgen05 attention requested for an unsupported case (S=%d dtype=%d mode=%d)", S,
                     dtype, mode);
      return SVLA_ERR_BAD_SHAPE;
    }
    if (tc_ok && g_attn_impl != 1)
      return svla_attn_tc_fwd(ctx, mode, q, k, v, ld, o, ldo, lse, traj, B, S, H, scale, as_stream(stream));
  }
  AttnArgs a{};
  a.mode = mode; a.q = q; a.k = k; a.v = v; a.ld = ld; a.o = o; a.ldo = ldo; a.lse = lse;
  a.traj = traj; a.bias = bias; a.keymask = keymask; a.B = B; a.S = S; a.H = H; a.scale = scale;
  SVLA_DISPATCH_DTYPE(dtype, T, {
    const size_t smem = fwd_smem<T>(S);
    SVLA_CHECK_ARG(smem <= 227 * 1024, "sequence too long for the shared-memory attention kernel at this dtype");
    SVLA_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_fwd_kernel<T><<<B * H, kThreads, smem, as_stream(stream)>>>(a);
  });
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

extern "C" int svla_attn_bwd(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld,
                             const void* o, const void* d_o, long long ldo, void* dq, void* dk, void* dv,
                             long long ldd, int dtype, const float* lse, const int64_t* traj, int B, int S, int H,
                             int dh, float scale, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && q && k && v && o && d_o && dq && dk && dv && lse, "NULL argument");
  SVLA_CHECK_ARG(dh == DH, "head dim must be 64");
  SVLA_CHECK_ARG(S >= 1 && S <= kMaxS, "S must be in [1, 256]");
  SVLA_CHECK_ARG(mode == SVLA_ATTN_FULL || mode == SVLA_ATTN_TRAJ_CAUSAL, "backward: FULL or TRAJ_CAUSAL only");
  SVLA_CHECK_ARG(mode != SVLA_ATTN_TRAJ_CAUSAL || traj, "TRAJ_CAUSAL needs traj");
  if (B <= 0) return SVLA_OK;
  {
    const bool grads_ok = ldd % 8 == 0 && ((reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) |
                                            reinterpret_cast<uintptr_t>(dv) | reinterpret_cast<uintptr_t>(d_o)) & 15) == 0;
    if (grads_ok && g_attn_impl != 1 && svla_attn_tc2_supported(mode, dtype, S, dh, ld, ldo, q, k, v, o))
      return svla_attn_tc2_bwd(ctx, mode, q, k, v, ld, o, d_o, ldo, dq, dk, dv, ldd, lse, traj, B, S, H, scale,
                               as_stream(stream));
    if (grads_ok && g_attn_impl != 1 && g_attn_impl != 3 && svla_attn_ws_supported(mode, dtype, S, dh, ld, ldo, q, k, v, o))
      return svla_attn_ws_bwd(ctx, mode, q, k, v, ld, d_o, ldo, dq, dk, dv, ldd, lse, traj, B, S, H, scale,
                              as_stream(stream));
    const bool tc_ok = svla_attn_tc_supported(mode, dtype, S, dh, ld, ldo, q, k, v, o) && grads_ok;
    if (g_attn_impl == 2 && !tc_ok) {
      svla_set_error("svla_attn_bwd: tcgen05 attention requested for an unsupported case (S=%d dtype=%d mode=%d)", S,
                     dtype, mode);
      return SVLA_ERR_BAD_SHAPE;
    }
    if (tc_ok && g_attn_impl != 1)
      return svla_attn_tc_bwd(ctx, mode, q, k, v, ld, o, d_o, ldo, dq, dk, dv, ldd, lse, traj, B, S, H, scale,
                              as_stream(stream));
  }
  AttnArgs a{};
  a.mode = mode; a.q = q; a.k = k; a.v = v; a.ld = ld; a.o = const_cast<void*>(o); a.d_o = d_o; a.ldo = ldo;
  a.dq = dq; a.dk = dk; a.dv = dv; a.ldd = ldd; a.lse = const_cast<float*>(lse); a.traj = traj;
  a.B = B; a.S = S; a.H = H; a.scale = scale;
  SVLA_DISPATCH_DTYPE(dtype, T, {
    const size_t smem = bwd_smem<T>(S);
    SVLA_CHECK_ARG(smem <= 227 * 1024, "sequence too long for the shared-memory attention backward at this dtype");
    SVLA_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_bwd_kernel<T><<<B * H, kThreads, smem, as_stream(stream)>>>(a);
  });
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}
