// tcgen05 attention for 128 < S <= 256 tokens (head dim 64, bf16): the two-camera fusion block (S = 201,
// BASELINE configs 4-5) and decoders over 256-step rollouts.  Same arithmetic as attn_tc.cu, two 128-row tiles
// per sequence.
//
// forward  (work item = sequence x head x query tile, 128 threads, thread = query row):
//   TMA(Q_i [128 x 64], K [256 x 64], V [256 x 64]) -> S = Q_i K^T as ONE N = 256 MMA chain (256 TMEM columns) ->
//   exact row softmax over all 256 columns -> P [128 x 256] bf16 (four swizzled 64-column chunks, over the dead
//   Q|K) -> O = P V (K = 256) -> O / rowsum.  96 KB + 256 columns => 2 CTAs per SM.
//
// backward (work item = sequence x head, 256 threads = 4 lane quarters x 2 column halves, 1 CTA per SM):
//   Q, K, V, dO [256 x 64] resident in shared memory; key tile j outer, query tile i inner:
//     S_ij = Q_i K_j^T, dP_ij = dO_i V_j^T                      TMEM [0,128) [128,256)
//     P = exp(S - lse), dS = P (dP - delta) -> two bf16 tiles in shared memory
//     dV_j += P^T dO_i, dK_j += dS^T Q_i                        TMEM [256,320) [320,384)   (accumulate over i)
//     dQ_i += dS K_j                                            TMEM [384,448) [448,512)   (accumulate over j)
//   All 512 TMEM columns are in use; the next block's S / dP MMAs are issued in the same batch as this block's three
//   gradient MMAs so the tensor pipe runs ahead of the element-wise stage.
// Rows / columns >= S come from the neighbouring sequence (or TMA zero fill) and are masked to exact zeros before
// they reach an accumulator.
#include <cuda.h>

#include <algorithm>

#include "common.cuh"

int svla_make_tmap_bf16(svla_ctx* ctx, const void* ptr, long long inner, long long outer, long long ld, int bi, int bo,
                        CUtensorMap* out);  // gemm_tc.cu

#include "attn_tc_common.cuh"

namespace {

constexpr int S2 = 256;  // maximum sequence length of these kernels

// ------------------------------------------------------------------------------------------ forward
template <int MODE, bool DROP>  // DROP: dropout on P (compiled out of the plain kernels: its presence alone cost 20 %)
__global__ void __launch_bounds__(128)
attn_tc2_fwd_kernel(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mk,
                    const __grid_constant__ CUtensorMap mv, AttnTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;               // 16 KB
  uint8_t* sK = smem + 16384;       // 32 KB
  uint8_t* sV = smem + 65536;       // 32 KB  (16 KB of padding before it completes the P overlay)
  uint8_t* sP = smem;               // 64 KB over Q | K | pad
  int* sTraj = reinterpret_cast<int*>(smem + 98304);                     // [256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 98304 + 1024);     // load, mma
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
  const int nqt = (S + TS - 1) / TS;       // query tiles per sequence
  const int nch = (S + 31) / 32;           // 32-column chunks that hold real keys
  const float sl2 = a.scale * kLog2e;
  uint32_t ph_load = 0, ph_mma = 0;
  const long long total = (long long)a.B * a.H * nqt;
  for (long long w = blockIdx.x; w < total; w += gridDim.x) {
    const int qt = (int)(w % nqt);
    const int h = (int)((w / nqt) % a.H), b = (int)(w / ((long long)nqt * a.H));
    const int row0 = b * S;
    if (tid == 0) {
      mbar_expect_tx(&bars[0], 5 * 16384);
      tma_load_2d(sQ, &mq, &bars[0], h * DH, row0 + qt * TS);
      tma_load_2d(sK, &mk, &bars[0], h * DH, row0);
      tma_load_2d(sK + 16384, &mk, &bars[0], h * DH, row0 + TS);
      tma_load_2d(sV, &mv, &bars[0], h * DH, row0);
      tma_load_2d(sV + 16384, &mv, &bars[0], h * DH, row0 + TS);
    }
    if (MODE == SVLA_ATTN_TRAJ_CAUSAL) {
      sTraj[tid] = (tid < S) ? (int)a.traj[row0 + tid] : -1 - tid;
      sTraj[tid + 128] = (tid + 128 < S) ? (int)a.traj[row0 + tid + 128] : -1 - (tid + 128);
    }
    mbar_wait(&bars[0], ph_load);
    ph_load ^= 1;
    if (tid == 0) {
      tc_fence_after();
      const uint32_t q = smem_u32(sQ), k = smem_u32(sK);
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk)
        umma_bf16(tmem, desc_kmajor(q, kk), desc_kmajor(k, kk), idesc(128, 256, false, false), kk > 0);
      umma_commit(&bars[1]);
    }
    __syncthreads();  // sTraj visible
    mbar_wait(&bars[1], ph_mma);  // S complete: Q and K are dead from here on
    ph_mma ^= 1;
    tc_fence_after();
    const int i = qt * TS + tid;  // query row inside the sequence
    const int my_traj = (MODE == SVLA_ATTN_TRAJ_CAUSAL) ? sTraj[min(i, S2 - 1)] : 0;
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < nch; ++c) {
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
    for (int c = 0; c < 8; ++c) {
      uint32_t r[32];
      if (c < nch) {
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
        if (DROP) {  // the normaliser stays the undropped row sum
          const uint32_t keep = dropout_keep8(a.drop, a.drop.row0 + (uint32_t)((b * a.H + h) * S2 + i), (uint32_t)(c * 4 + j8));
#pragma unroll
          for (int e = 0; e < 8; ++e) p[e] = ((keep >> e) & 1u) ? p[e] * a.drop.scale : 0.f;
        }
        store_p8(sP, tid, c * 4 + j8, p);
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();  // P complete, every thread is done reading S
    if (tid == 0) {
      tc_fence_after();
      const uint32_t p = smem_u32(sP), v = smem_u32(sV);
#pragma unroll
      for (int kk = 0; kk < S2 / 16; ++kk)  // O overwrites TMEM columns [0, 64)
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
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ backward
constexpr uint32_t kColS = 0, kColDP = 128, kColDV = 256, kColDK = 320, kColDQ = 384;

template <int MODE>
__global__ void __launch_bounds__(256, 1)
attn_tc2_bwd_kernel(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mk,
                    const __grid_constant__ CUtensorMap mv, const __grid_constant__ CUtensorMap mdo, AttnTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                // 32 KB each: rows 0..255 of this (sequence, head)
  uint8_t* sK = smem + 32768;
  uint8_t* sV = smem + 65536;
  uint8_t* sdO = smem + 98304;
  uint8_t* sP = smem + 131072;       // 32 KB  [128 queries x 128 keys]
  uint8_t* sdS = smem + 163840;      // 32 KB
  int* sTraj = reinterpret_cast<int*>(smem + 196608);                    // [256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 196608 + 1024);    // load, mma
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int qrow = tid & 127;        // row of the 128-row tile this thread owns (TMEM lane)
  const int chalf = tid >> 7;        // which 64 columns of a 128-column block this thread handles
  constexpr uint32_t kCols = 512;
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
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const int S = a.S;
  const int nt = (S + TS - 1) / TS;  // tiles per sequence (query and key): 2 here, 1 also works
  const float sl2 = a.scale * kLog2e;
  uint32_t ph_load = 0, ph_mma = 0;

  for (int w = blockIdx.x; w < a.B * a.H; w += gridDim.x) {
    const int b = w / a.H, h = w % a.H;
    const int row0 = b * S;
    if (tid == 0) {
      mbar_expect_tx(&bars[0], 8 * 16384);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        tma_load_2d(sQ + t * 16384, &mq, &bars[0], h * DH, row0 + t * TS);
        tma_load_2d(sK + t * 16384, &mk, &bars[0], h * DH, row0 + t * TS);
        tma_load_2d(sV + t * 16384, &mv, &bars[0], h * DH, row0 + t * TS);
        tma_load_2d(sdO + t * 16384, &mdo, &bars[0], h * DH, row0 + t * TS);
      }
    }
    if (MODE == SVLA_ATTN_TRAJ_CAUSAL) sTraj[tid] = (tid < S) ? (int)a.traj[row0 + tid] : -1 - tid;
    // delta_i = dO_i . O_i and lse_i for this thread's row of both query tiles (overlaps the TMA)
    float delta[2] = {0.f, 0.f}, lse2[2] = {0.f, 0.f};
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int i = t * TS + qrow;
      if (i < S) {
        const uint4* po = reinterpret_cast<const uint4*>(a.o_in + (long long)(row0 + i) * a.ldo + h * DH);
        const uint4* pd = reinterpret_cast<const uint4*>(a.d_o + (long long)(row0 + i) * a.ldo + h * DH);
        float d = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 uo = __ldg(po + j), ud = __ldg(pd + j);
          const __nv_bfloat162* ho = reinterpret_cast<const __nv_bfloat162*>(&uo);
          const __nv_bfloat162* hd = reinterpret_cast<const __nv_bfloat162*>(&ud);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 fo = __bfloat1622float2(ho[e]), fd = __bfloat1622float2(hd[e]);
            d = fmaf(fo.x, fd.x, d);
            d = fmaf(fo.y, fd.y, d);
          }
        }
        delta[t] = d;
        lse2[t] = a.lse[((long long)b * a.H + h) * S + i] * kLog2e;
      }
    }
    mbar_wait(&bars[0], ph_load);
    ph_load ^= 1;
    const uint32_t uq = smem_u32(sQ), uk = smem_u32(sK), uv = smem_u32(sV), ud_ = smem_u32(sdO);
    const uint32_t up = smem_u32(sP), us = smem_u32(sdS);
    if (tid == 0) {  // S_00, dP_00
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk)
        umma_bf16(tmem + kColS, desc_kmajor(uq, kk), desc_kmajor(uk, kk), idesc(128, 128, false, false), kk > 0);
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk)
        umma_bf16(tmem + kColDP, desc_kmajor(ud_, kk), desc_kmajor(uv, kk), idesc(128, 128, false, false), kk > 0);
      umma_commit(&bars[1]);
    }
    __syncthreads();  // sTraj visible
    for (int j = 0; j < nt; ++j) {
      for (int i = 0; i < nt; ++i) {
        // S_ij / dP_ij of this block are ready, and every MMA that read the P / dS tiles has retired
        mbar_wait(&bars[1], ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
        const int gi = i * TS + qrow;  // query row in the sequence
        const int my_traj = (MODE == SVLA_ATTN_TRAJ_CAUSAL) ? sTraj[gi] : 0;
        const float dl = i ? delta[1] : delta[0], l2 = i ? lse2[1] : lse2[0];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int cb = chalf * 64 + c * 32;  // first column (key inside the tile) of this 32-column chunk
          uint32_t rs[32], rp[32];
          if (j * TS + cb < S) {
            tmem_ld32(tmem + lane_base + kColS + cb, rs);
            tmem_ld32(tmem + lane_base + kColDP + cb, rp);
            tmem_wait_ld();
          }
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            float p[8], ds[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int col = j * TS + cb + j8 * 8 + e;  // key index in the sequence
              bool ok = col < S && gi < S;
              if (MODE == SVLA_ATTN_TRAJ_CAUSAL) ok = ok && col <= gi && sTraj[col] == my_traj;
              p[e] = ok ? exp2f(__uint_as_float(rs[j8 * 8 + e]) * sl2 - l2) : 0.f;
              ds[e] = ok ? p[e] * (__uint_as_float(rp[j8 * 8 + e]) - dl) : 0.f;
            }
            if (MODE == SVLA_ATTN_FULL && a.drop.thr != 0u) {
              // P~ = P keep / (1 - p) feeds dV; dS = P~ dP - P delta, delta = dO . O = rowsum(P~ dP) (O is the dropped output)
              const uint32_t keep = dropout_keep8(a.drop, a.drop.row0 + (uint32_t)(w * S2 + gi),
                                                  (uint32_t)(((j * TS + cb) >> 3) + j8));
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float pm = ((keep >> e) & 1u) ? p[e] * a.drop.scale : 0.f;
                ds[e] = fmaf(pm, __uint_as_float(rp[j8 * 8 + e]), -p[e] * dl);  // p = 0 where masked: ds = 0 there
                p[e] = pm;
              }
            }
            store_p8_sts(sP, qrow, (cb >> 3) + j8, p);
            store_p8_sts(sdS, qrow, (cb >> 3) + j8, ds);
          }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();  // P / dS complete; S and dP fully consumed
        if (tid == 0) {
          tc_fence_after();
          const uint32_t qi = uq + i * 16384, kj = uk + j * 16384, doi = ud_ + i * 16384;
#pragma unroll
          for (int kk = 0; kk < TS / 16; ++kk)  // dV_j[keys, dh] += P^T dO_i
            umma_bf16(tmem + kColDV, desc_p_mnmajor(up, kk), desc_mnmajor64(doi, kk), idesc(128, 64, true, true),
                      (i > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < TS / 16; ++kk)  // dK_j[keys, dh] += dS^T Q_i
            umma_bf16(tmem + kColDK, desc_p_mnmajor(us, kk), desc_mnmajor64(qi, kk), idesc(128, 64, true, true),
                      (i > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < TS / 16; ++kk)  // dQ_i[queries, dh] += dS K_j
            umma_bf16(tmem + kColDQ + i * 64, desc_p_kmajor(us, kk), desc_mnmajor64(kj, kk), idesc(128, 64, false, true),
                      (j > 0 || kk > 0) ? 1u : 0u);
          // run ahead: S / dP of the next block (their TMEM columns are free again)
          int ni = i + 1, nj = j;
          if (ni == nt) { ni = 0; ++nj; }
          if (nj < nt) {
            const uint32_t qn = uq + ni * 16384, kn = uk + nj * 16384, don = ud_ + ni * 16384, vn = uv + nj * 16384;
#pragma unroll
            for (int kk = 0; kk < DH / 16; ++kk)
              umma_bf16(tmem + kColS, desc_kmajor(qn, kk), desc_kmajor(kn, kk), idesc(128, 128, false, false), kk > 0);
#pragma unroll
            for (int kk = 0; kk < DH / 16; ++kk)
              umma_bf16(tmem + kColDP, desc_kmajor(don, kk), desc_kmajor(vn, kk), idesc(128, 128, false, false), kk > 0);
          }
          umma_commit(&bars[1]);
        }
      }
      // dV_j, dK_j complete once the last batch has retired; that same commit also gates the next block, so peek at
      // it here without consuming the phase twice: wait, drain, and let the next block's wait fall through
      mbar_wait(&bars[1], ph_mma);
      tc_fence_after();
      {
        const int gk = j * TS + qrow;  // key row
        const long long orow = (long long)(row0 + gk) * a.ldd + h * DH + chalf * 32;
        uint32_t r[32];
        tmem_ld32(tmem + lane_base + kColDV + chalf * 32, r);
        tmem_wait_ld();
        if (gk < S) {
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            uint4 u;
            __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              hh[e] = __floats2bfloat162_rn(__uint_as_float(r[c8 * 8 + 2 * e]), __uint_as_float(r[c8 * 8 + 2 * e + 1]));
            *reinterpret_cast<uint4*>(a.dv + orow + c8 * 8) = u;
          }
        }
        tmem_ld32(tmem + lane_base + kColDK + chalf * 32, r);
        tmem_wait_ld();
        if (gk < S) {
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            uint4 u;
            __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              hh[e] = __floats2bfloat162_rn(__uint_as_float(r[c8 * 8 + 2 * e]) * a.scale,
                                            __uint_as_float(r[c8 * 8 + 2 * e + 1]) * a.scale);
            *reinterpret_cast<uint4*>(a.dk + orow + c8 * 8) = u;
          }
        }
      }
      tc_fence_before();
      __syncthreads();  // dV_j / dK_j drained before the next key tile's first accumulate-off MMA overwrites them
      if (j + 1 == nt) {
        ph_mma ^= 1;  // nothing follows: consume the phase that was only peeked at
      }
    }
    // dQ_0, dQ_1
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      if (t < nt) {
        const int gi = t * TS + qrow;
        uint32_t r[32];
        tmem_ld32(tmem + lane_base + kColDQ + t * 64 + chalf * 32, r);
        tmem_wait_ld();
        if (gi < S) {
          const long long orow = (long long)(row0 + gi) * a.ldd + h * DH + chalf * 32;
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            uint4 u;
            __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              hh[e] = __floats2bfloat162_rn(__uint_as_float(r[c8 * 8 + 2 * e]) * a.scale,
                                            __uint_as_float(r[c8 * 8 + 2 * e + 1]) * a.scale);
            *reinterpret_cast<uint4*>(a.dq + orow + c8 * 8) = u;
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // all TMEM / smem reads of this item are done before the next item's TMA and MMAs
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

int svla_attn_tc2_fwd_drop(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld, void* o,
                           long long ldo, float* lse, const int64_t* traj, int B, int S, int H, float scale,
                           const svla_dropout* drop, cudaStream_t st);
int svla_attn_tc2_bwd_drop(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld,
                           const void* o, const void* d_o, long long ldo, void* dq, void* dk, void* dv, long long ldd,
                           const float* lse, const int64_t* traj, int B, int S, int H, float scale,
                           const svla_dropout* drop, cudaStream_t st);

bool svla_attn_tc2_supported(int mode, int dtype, int S, int dh, long long ld, long long ldo, const void* q,
                             const void* k, const void* v, const void* o) {
  return dtype == SVLA_BF16 && dh == DH && S > TS && S <= S2 &&
         (mode == SVLA_ATTN_FULL || mode == SVLA_ATTN_TRAJ_CAUSAL) && ld % 8 == 0 && ldo % 8 == 0 && al16(q) &&
         al16(k) && al16(v) && al16(o);
}

int svla_attn_tc2_fwd(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld, void* o,
                      long long ldo, float* lse, const int64_t* traj, int B, int S, int H, float scale, cudaStream_t st) {
  return svla_attn_tc2_fwd_drop(ctx, mode, q, k, v, ld, o, ldo, lse, traj, B, S, H, scale, nullptr, st);
}

int svla_attn_tc2_fwd_drop(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld, void* o,
                           long long ldo, float* lse, const int64_t* traj, int B, int S, int H, float scale,
                           const svla_dropout* drop, cudaStream_t st) {
  CUtensorMap mq, mk, mv;
  const long long rows = (long long)B * S;
  int rc;
  if ((rc = svla_make_tmap_bf16(ctx, q, (long long)H * DH, rows, ld, DH, TS, &mq))) return rc;
  if ((rc = svla_make_tmap_bf16(ctx, k, (long long)H * DH, rows, ld, DH, TS, &mk))) return rc;
  if ((rc = svla_make_tmap_bf16(ctx, v, (long long)H * DH, rows, ld, DH, TS, &mv))) return rc;
  AttnTcArgs a{};
  a.mode = mode; a.B = B; a.S = S; a.H = H; a.scale = scale; a.traj = traj; a.lse = lse;
  a.o = reinterpret_cast<__nv_bfloat16*>(o); a.ldo = ldo;
  a.drop = make_drop_args(mode == SVLA_ATTN_FULL ? drop : nullptr);
  constexpr size_t smem = 98304 + 1024 + 64 + 1024;
  static bool attr = false;
  if (!attr) {
    SVLA_CUDA(cudaFuncSetAttribute(attn_tc2_fwd_kernel<SVLA_ATTN_FULL, false>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SVLA_CUDA(cudaFuncSetAttribute(attn_tc2_fwd_kernel<SVLA_ATTN_FULL, true>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SVLA_CUDA(cudaFuncSetAttribute(attn_tc2_fwd_kernel<SVLA_ATTN_TRAJ_CAUSAL, false>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const long long items = (long long)B * H * ((S + TS - 1) / TS);
  const int grid = (int)std::min<long long>(items, 2LL * ctx->sm_count);
  if (mode == SVLA_ATTN_FULL && a.drop.thr != 0u) attn_tc2_fwd_kernel<SVLA_ATTN_FULL, true><<<grid, 128, smem, st>>>(mq, mk, mv, a);
  else if (mode == SVLA_ATTN_FULL) attn_tc2_fwd_kernel<SVLA_ATTN_FULL, false><<<grid, 128, smem, st>>>(mq, mk, mv, a);
  else attn_tc2_fwd_kernel<SVLA_ATTN_TRAJ_CAUSAL, false><<<grid, 128, smem, st>>>(mq, mk, mv, a);
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

int svla_attn_tc2_bwd(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld, const void* o,
                      const void* d_o, long long ldo, void* dq, void* dk, void* dv, long long ldd, const float* lse,
                      const int64_t* traj, int B, int S, int H, float scale, cudaStream_t st) {
  return svla_attn_tc2_bwd_drop(ctx, mode, q, k, v, ld, o, d_o, ldo, dq, dk, dv, ldd, lse, traj, B, S, H, scale, nullptr, st);
}

int svla_attn_tc2_bwd_drop(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld,
                           const void* o, const void* d_o, long long ldo, void* dq, void* dk, void* dv, long long ldd,
                           const float* lse, const int64_t* traj, int B, int S, int H, float scale,
                           const svla_dropout* drop, cudaStream_t st) {
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
  a.drop = make_drop_args(mode == SVLA_ATTN_FULL ? drop : nullptr);
  constexpr size_t smem = 196608 + 1024 + 64 + 1024;
  static bool attr = false;
  if (!attr) {
    SVLA_CUDA(cudaFuncSetAttribute(attn_tc2_bwd_kernel<SVLA_ATTN_FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    SVLA_CUDA(cudaFuncSetAttribute(attn_tc2_bwd_kernel<SVLA_ATTN_TRAJ_CAUSAL>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const int grid = std::min(B * H, ctx->sm_count);
  if (mode == SVLA_ATTN_FULL) attn_tc2_bwd_kernel<SVLA_ATTN_FULL><<<grid, 256, smem, st>>>(mq, mk, mv, mdo, a);
  else attn_tc2_bwd_kernel<SVLA_ATTN_TRAJ_CAUSAL><<<grid, 256, smem, st>>>(mq, mk, mv, mdo, a);
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}
