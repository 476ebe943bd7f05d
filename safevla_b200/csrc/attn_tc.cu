// tcgen05 attention for the towers' short sequences (S <= 128 tokens, head dim 64, bf16):
//   fusion block (S = 117, no mask) and decoder (T <= 128, trajectory-causal mask from traj_index).
// One (sequence, head) per work item, persistent CTAs of 128 threads; thread i owns query row i.
//
// forward:   TMA(Q,K,V) -> S = Q K^T (tcgen05, fp32 in TMEM) -> row softmax from TMEM (exp2, fp32) -> P (bf16,
//            128B-swizzled in shared memory) -> O = P V (tcgen05) -> O / rowsum -> global; saves log-sum-exp.
// backward:  TMA(Q,K,V,dO) -> S = Q K^T, dP = dO V^T -> P = exp(S - lse), dS = P (dP - delta) ->
//            dV = P^T dO, dK = dS^T Q, dQ = dS K (three tcgen05 GEMMs out of the same shared-memory tiles: P/dS are
//            written once in a layout that is K-major for dQ and MN-major for dV/dK; Q, K, V, dO are consumed
//            exactly as TMA lands them).
// Rows/columns >= S of the 128-wide tile come from the neighbouring sequence (or TMA zero fill): masked to exact
// zeros before they can reach an accumulator.
#include <cuda.h>

#include <algorithm>

#include "common.cuh"

int svla_make_tmap_bf16(svla_ctx* ctx, const void* ptr, long long inner, long long outer, long long ld, int bi, int bo,
                        CUtensorMap* out);  // gemm_tc.cu

#include "attn_tc_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------ forward
// Shared memory per CTA: Q | K | V (3 x 16 KB); once S = Q K^T has retired, Q|K are dead and P (32 KB) is written
// over them.  TMEM: 128 columns -- S, then O reuses columns [0, 64).  48.6 KB + 128 columns => 4 CTAs per SM, whose
// load / MMA / softmax phases overlap each other.
template <int MODE>
__global__ void __launch_bounds__(128)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mk,
                   const __grid_constant__ CUtensorMap mv, AttnTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;             // 16 KB each
  uint8_t* sK = smem + 16384;
  uint8_t* sV = smem + 32768;
  uint8_t* sP = smem;             // 32 KB, aliases Q|K
  int* sTraj = reinterpret_cast<int*>(smem + 49152);                       // [128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 49152 + 512);        // load, mma
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr uint32_t kCols = 128;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const int S = a.S;
  const float sl2 = a.scale * kLog2e;
  uint32_t ph_load = 0, ph_mma = 0;
  for (int w = blockIdx.x; w < a.B * a.H; w += gridDim.x) {
    const int b = w / a.H, h = w % a.H;
    const int row0 = b * S;
    if (tid == 0) {
      mbar_expect_tx(&bars[0], 3 * 16384);
      tma_load_2d(sQ, &mq, &bars[0], h * DH, row0);
      tma_load_2d(sK, &mk, &bars[0], h * DH, row0);
      tma_load_2d(sV, &mv, &bars[0], h * DH, row0);
    }
    if (MODE == SVLA_ATTN_TRAJ_CAUSAL) sTraj[tid] = (tid < S) ? (int)a.traj[row0 + tid] : -1 - tid;
    mbar_wait(&bars[0], ph_load);
    ph_load ^= 1;
    if (tid == 0) {
      tc_fence_after();
      const uint32_t q = smem_u32(sQ), k = smem_u32(sK);
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk)
        umma_bf16(tmem, desc_kmajor(q, kk), desc_kmajor(k, kk), idesc(128, 128, false, false), kk > 0);
      umma_commit(&bars[1]);
    }
    __syncthreads();  // sTraj visible
    mbar_wait(&bars[1], ph_mma);  // S complete: Q and K are dead from here on
    ph_mma ^= 1;
    tc_fence_after();
    // ---- softmax of row i = tid
    const int i = tid;
    const int my_traj = (MODE == SVLA_ATTN_TRAJ_CAUSAL) ? sTraj[i] : 0;
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      if (c * 32 >= S) break;
      uint32_t r[32];
      tmem_ld32(tmem + lane_base + c * 32, r);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = c * 32 + j;
        bool ok = col < S;
        if (MODE == SVLA_ATTN_TRAJ_CAUSAL) ok = ok && col <= i && sTraj[col] == my_traj;
        if (ok) mx = fmaxf(mx, __uint_as_float(r[j]));
      }
    }
    const float mxs = (mx == -INFINITY) ? 0.f : mx * sl2;
    float sum = 0.f;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      if (c * 32 < S) {
        tmem_ld32(tmem + lane_base + c * 32, r);
        tmem_wait_ld();
      }
#pragma unroll
      for (int j8 = 0; j8 < 4; ++j8) {
        float p[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int col = c * 32 + j8 * 8 + e;
          bool ok = col < S && i < S;
          if (MODE == SVLA_ATTN_TRAJ_CAUSAL) ok = ok && col <= i && sTraj[col] == my_traj;
          p[e] = ok ? exp2f(__uint_as_float(r[j8 * 8 + e]) * sl2 - mxs) : 0.f;
          sum += p[e];
        }
        store_p8(sP, i, c * 4 + j8, p);
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();  // P complete, every thread is done reading S
    if (tid == 0) {
      tc_fence_after();
      const uint32_t p = smem_u32(sP), v = smem_u32(sV);
#pragma unroll
      for (int kk = 0; kk < TS / 16; ++kk)  // O overwrites TMEM columns [0, 64)
        umma_bf16(tmem, desc_p_kmajor(p, kk), desc_mnmajor64(v, kk), idesc(128, 64, false, true), kk > 0);
      umma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], ph_mma);
    ph_mma ^= 1;
    tc_fence_after();
    {
      uint32_t r0[32], r1[32];
      tmem_ld32(tmem + lane_base, r0);
      tmem_ld32(tmem + lane_base + 32, r1);
      tmem_wait_ld();
      if (i < S) {
        store_row64(a.o + (long long)(row0 + i) * a.ldo + h * DH, r0, r1, 1.f / sum);
        if (a.lse) a.lse[((long long)b * a.H + h) * S + i] = mx * a.scale + __logf(sum);
      }
    }
    tc_fence_before();
    __syncthreads();  // all TMEM reads / smem reads of this item are done before the next item reuses them
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ backward
// Shared memory per CTA: Q | K | V | dO (4 x 16 KB) + ONE 32 KB tile that first holds P (for dV = P^T dO) and then,
// once that GEMM has retired, dS (for dK = dS^T Q and dQ = dS K); dS waits in registers (packed bf16) meanwhile.
// TMEM: S [0,128) and dP [128,256) are consumed into P/dS before dV [0,64), dK [64,128), dQ [128,192) reuse
// their columns.  96.6 KB + 256 columns => 2 CTAs per SM.
template <int MODE>
__global__ void __launch_bounds__(128)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mk,
                   const __grid_constant__ CUtensorMap mv, const __grid_constant__ CUtensorMap mdo, AttnTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 16384;
  uint8_t* sV = smem + 32768;
  uint8_t* sdO = smem + 49152;
  uint8_t* sP = smem + 65536;    // 32 KB: P, then dS
  int* sTraj = reinterpret_cast<int*>(smem + 98304);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 98304 + 512);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr uint32_t kCols = 256;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const int S = a.S;
  const float sl2 = a.scale * kLog2e;
  uint32_t ph_load = 0, ph_mma = 0;
  for (int w = blockIdx.x; w < a.B * a.H; w += gridDim.x) {
    const int b = w / a.H, h = w % a.H;
    const int row0 = b * S;
    const int i = tid;
    if (tid == 0) {
      mbar_expect_tx(&bars[0], 4 * 16384);
      tma_load_2d(sQ, &mq, &bars[0], h * DH, row0);
      tma_load_2d(sK, &mk, &bars[0], h * DH, row0);
      tma_load_2d(sV, &mv, &bars[0], h * DH, row0);
      tma_load_2d(sdO, &mdo, &bars[0], h * DH, row0);
    }
    if (MODE == SVLA_ATTN_TRAJ_CAUSAL) sTraj[tid] = (tid < S) ? (int)a.traj[row0 + tid] : -1 - tid;
    // delta_i = dO_i . O_i and lse_i straight from global memory (overlaps the TMA)
    float delta = 0.f, lse2 = 0.f;
    if (i < S) {
      const uint4* po = reinterpret_cast<const uint4*>(a.o_in + (long long)(row0 + i) * a.ldo + h * DH);
      const uint4* pd = reinterpret_cast<const uint4*>(a.d_o + (long long)(row0 + i) * a.ldo + h * DH);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 uo = __ldg(po + j), ud = __ldg(pd + j);
        const __nv_bfloat162* ho = reinterpret_cast<const __nv_bfloat162*>(&uo);
        const __nv_bfloat162* hd = reinterpret_cast<const __nv_bfloat162*>(&ud);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 fo = __bfloat1622float2(ho[e]), fd = __bfloat1622float2(hd[e]);
          delta = fmaf(fo.x, fd.x, delta);
          delta = fmaf(fo.y, fd.y, delta);
        }
      }
      lse2 = a.lse[((long long)b * a.H + h) * S + i] * kLog2e;
    }
    mbar_wait(&bars[0], ph_load);
    ph_load ^= 1;
    if (tid == 0) {
      tc_fence_after();
      const uint32_t q = smem_u32(sQ), k = smem_u32(sK), v = smem_u32(sV), d = smem_u32(sdO);
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk)
        umma_bf16(tmem, desc_kmajor(q, kk), desc_kmajor(k, kk), idesc(128, 128, false, false), kk > 0);
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk)
        umma_bf16(tmem + 128, desc_kmajor(d, kk), desc_kmajor(v, kk), idesc(128, 128, false, false), kk > 0);
      umma_commit(&bars[1]);
    }
    __syncthreads();
    mbar_wait(&bars[1], ph_mma);
    ph_mma ^= 1;
    tc_fence_after();
    const int my_traj = (MODE == SVLA_ATTN_TRAJ_CAUSAL) ? sTraj[i] : 0;
    uint32_t dsp[64];  // this row's dS, packed bf16x2
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t rs[32], rp[32];
      if (c * 32 < S) {
        tmem_ld32(tmem + lane_base + c * 32, rs);
        tmem_ld32(tmem + lane_base + 128 + c * 32, rp);
        tmem_wait_ld();
      }
#pragma unroll
      for (int j8 = 0; j8 < 4; ++j8) {
        float p[8], ds[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int col = c * 32 + j8 * 8 + e;
          bool ok = col < S && i < S;
          if (MODE == SVLA_ATTN_TRAJ_CAUSAL) ok = ok && col <= i && sTraj[col] == my_traj;
          p[e] = ok ? exp2f(__uint_as_float(rs[j8 * 8 + e]) * sl2 - lse2) : 0.f;
          ds[e] = ok ? p[e] * (__uint_as_float(rp[j8 * 8 + e]) - delta) : 0.f;
        }
        store_p8(sP, i, c * 4 + j8, p);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __nv_bfloat162 hb = __floats2bfloat162_rn(ds[2 * e], ds[2 * e + 1]);
          dsp[(c * 4 + j8) * 4 + e] = *reinterpret_cast<const uint32_t*>(&hb);
        }
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();  // P complete; S and dP fully consumed
    if (tid == 0) {
      tc_fence_after();
      const uint32_t d = smem_u32(sdO), p = smem_u32(sP);
#pragma unroll
      for (int kk = 0; kk < TS / 16; ++kk)  // dV[keys, dh] = P^T dO  -> TMEM [0, 64)
        umma_bf16(tmem, desc_p_mnmajor(p, kk), desc_mnmajor64(d, kk), idesc(128, 64, true, true), kk > 0);
      umma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], ph_mma);  // dV retired: the tile may now take dS
    ph_mma ^= 1;
    tc_fence_after();
#pragma unroll
    for (int c8 = 0; c8 < 16; ++c8) {
      const int chunk = c8 >> 3, cc = c8 & 7;
      *reinterpret_cast<uint4*>(sP + chunk * 16384 + i * 128 + ((cc ^ (i & 7)) << 4)) =
          make_uint4(dsp[c8 * 4], dsp[c8 * 4 + 1], dsp[c8 * 4 + 2], dsp[c8 * 4 + 3]);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t q = smem_u32(sQ), k = smem_u32(sK), s_ = smem_u32(sP);
#pragma unroll
      for (int kk = 0; kk < TS / 16; ++kk)  // dK[keys, dh] = dS^T Q -> TMEM [64, 128)
        umma_bf16(tmem + 64, desc_p_mnmajor(s_, kk), desc_mnmajor64(q, kk), idesc(128, 64, true, true), kk > 0);
#pragma unroll
      for (int kk = 0; kk < TS / 16; ++kk)  // dQ[queries, dh] = dS K -> TMEM [128, 192)
        umma_bf16(tmem + 128, desc_p_kmajor(s_, kk), desc_mnmajor64(k, kk), idesc(128, 64, false, true), kk > 0);
      umma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], ph_mma);
    ph_mma ^= 1;
    tc_fence_after();
    {
      uint32_t r0[32], r1[32];
      const long long orow = (long long)(row0 + i) * a.ldd + h * DH;
      tmem_ld32(tmem + lane_base, r0);
      tmem_ld32(tmem + lane_base + 32, r1);
      tmem_wait_ld();
      if (i < S) store_row64(a.dv + orow, r0, r1, 1.f);
      tmem_ld32(tmem + lane_base + 64, r0);
      tmem_ld32(tmem + lane_base + 96, r1);
      tmem_wait_ld();
      if (i < S) store_row64(a.dk + orow, r0, r1, a.scale);
      tmem_ld32(tmem + lane_base + 128, r0);
      tmem_ld32(tmem + lane_base + 160, r1);
      tmem_wait_ld();
      if (i < S) store_row64(a.dq + orow, r0, r1, a.scale);
    }
    tc_fence_before();
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kCols) : "memory");
  }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

bool svla_attn_tc_supported(int mode, int dtype, int S, int dh, long long ld, long long ldo, const void* q,
                            const void* k, const void* v, const void* o) {
  return dtype == SVLA_BF16 && dh == DH && S >= 1 && S <= TS && (mode == SVLA_ATTN_FULL || mode == SVLA_ATTN_TRAJ_CAUSAL) &&
         ld % 8 == 0 && ldo % 8 == 0 && al16(q) && al16(k) && al16(v) && al16(o);
}

int svla_attn_tc_fwd(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld, void* o,
                     long long ldo, float* lse, const int64_t* traj, int B, int S, int H, float scale, cudaStream_t st) {
  CUtensorMap mq, mk, mv;
  const long long rows = (long long)B * S;
  int rc;
  if ((rc = svla_make_tmap_bf16(ctx, q, (long long)H * DH, rows, ld, DH, TS, &mq))) return rc;
  if ((rc = svla_make_tmap_bf16(ctx, k, (long long)H * DH, rows, ld, DH, TS, &mk))) return rc;
  if ((rc = svla_make_tmap_bf16(ctx, v, (long long)H * DH, rows, ld, DH, TS, &mv))) return rc;
  AttnTcArgs a{};
  a.mode = mode; a.B = B; a.S = S; a.H = H; a.scale = scale; a.traj = traj; a.lse = lse;
  a.o = reinterpret_cast<__nv_bfloat16*>(o); a.ldo = ldo;
  constexpr size_t smem = 49152 + 512 + 64 + 1024;
  static bool attr = false;
  if (!attr) {
    SVLA_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel<SVLA_ATTN_FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    SVLA_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel<SVLA_ATTN_TRAJ_CAUSAL>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const int grid = std::min(B * H, 4 * ctx->sm_count);
  if (mode == SVLA_ATTN_FULL) attn_tc_fwd_kernel<SVLA_ATTN_FULL><<<grid, 128, smem, st>>>(mq, mk, mv, a);
  else attn_tc_fwd_kernel<SVLA_ATTN_TRAJ_CAUSAL><<<grid, 128, smem, st>>>(mq, mk, mv, a);
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

int svla_attn_tc_bwd(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld, const void* o,
                     const void* d_o, long long ldo, void* dq, void* dk, void* dv, long long ldd, const float* lse,
                     const int64_t* traj, int B, int S, int H, float scale, cudaStream_t st) {
  CUtensorMap mq, mk, mv, mdo;
  const long long rows = (long long)B * S;
  int rc;
  if ((rc = svla_make_tmap_bf16(ctx, q, (long long)H * DH, rows, ld, DH, TS, &mq))) return rc;
  if ((rc = svla_make_tmap_bf16(ctx, k, (long long)H * DH, rows, ld, DH, TS, &mk))) return rc;
  if ((rc = svla_make_tmap_bf16(ctx, v, (long long)H * DH, rows, ld, DH, TS, &mv))) return rc;
  if ((rc = svla_make_tmap_bf16(ctx, d_o, (long long)H * DH, rows, ldo, DH, TS, &mdo))) return rc;
  AttnTcArgs a{};
  a.mode = mode; a.B = B; a.S = S; a.H = H; a.scale = scale; a.traj = traj; a.lse = const_cast<float*>(lse);
  a.o_in = reinterpret_cast<const __nv_bfloat16*>(o); a.d_o = reinterpret_cast<const __nv_bfloat16*>(d_o); a.ldo = ldo;
  a.dq = reinterpret_cast<__nv_bfloat16*>(dq); a.dk = reinterpret_cast<__nv_bfloat16*>(dk);
  a.dv = reinterpret_cast<__nv_bfloat16*>(dv); a.ldd = ldd;
  constexpr size_t smem = 98304 + 512 + 64 + 1024;
  static bool attr = false;
  if (!attr) {
    SVLA_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel<SVLA_ATTN_FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    SVLA_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel<SVLA_ATTN_TRAJ_CAUSAL>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const int grid = std::min(B * H, 2 * ctx->sm_count);
  if (mode == SVLA_ATTN_FULL) attn_tc_bwd_kernel<SVLA_ATTN_FULL><<<grid, 128, smem, st>>>(mq, mk, mv, mdo, a);
  else attn_tc_bwd_kernel<SVLA_ATTN_TRAJ_CAUSAL><<<grid, 128, smem, st>>>(mq, mk, mv, mdo, a);
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}
