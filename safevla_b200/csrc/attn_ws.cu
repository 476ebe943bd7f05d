// Warp-specialised, persistent tcgen05 attention for the towers' short sequences (S <= 128, head dim 64):
// fusion block (S = 117, no mask), decoder (T <= 128, trajectory-causal mask from traj_index).  Replaces
// nn.MultiheadAttention's core (allenact_dino_transformer.py:545-552 via nn.TransformerEncoderLayer) and the decoder's
// scaled_dot_product_attention (training/online/third_party_models/llama/model.py:317) forward + backward.
//
// One CTA per SM walks (sequence, head) items; roles per warp:
//   warp 0       TMA producer: the item's Q | K | V (| dO) tiles through 3-D tensor maps (rows beyond the sequence are
//                zero-filled, never the neighbouring sequence) into an mbarrier ring, up to NS items ahead;
//   warp 1       tcgen05.mma issuer (one elected thread); accumulators in TMEM;
//   warp 2       TMEM allocation;
//   warps 4..    compute groups (TMEM lane = query row); registers are re-balanced towards them (setmaxnreg).
// forward (384 threads):  items alternate between two softmax groups of four warps (thread = query row: the whole
//           128-column score row is pulled from TMEM into registers once; max / exp2 / sum / bf16 pack run on it with
//           independent accumulators); S = Q K^T of item i+1 is issued before O = P V of item i, so the tensor pipe
//           computes the next score tile while one group exponentiates and the other drains the previous O (TMEM:
//           S_A, S_B, O_A, O_B).  P is written (bf16, 128B-swizzled) over the item's dead Q | K tiles; the stage
//           returns to the producer when P V has retired.
// backward (512 threads): group A = eight warps, two per TMEM lane quadrant, each thread holding 64 columns of S and
//           of dP = dO V^T in registers: P = exp2(S - lse), delta = rowsum(P * dP) (halves exchanged through shared
//           memory, so O is never re-read), dS = P (dP - delta), both packed to bf16 tiles; group B (four warps)
//           drains dV / dK / dQ of the previous item to HBM meanwhile.  TMEM holds two items (S | dP, overwritten by
//           dV | dK | dQ once P / dS are written); P / dS live in one 64 KB shared-memory pair.
//
// NP = 1: bf16 operands (the fast path).  NP = 3: every operand arrives as a (hi, lo) bf16 pair of an fp32 tensor
// (svla_split_concat) and every product is evaluated as hi*hi + lo*hi + hi*lo in the same accumulator; P / dS are
// split in registers -- fp32 attention to ~2^-16 on the tensor cores (parity-grade mode), fp32 outputs.
#include <cuda.h>

#include <algorithm>

#include "common.cuh"

int svla_make_tmap3_bf16(svla_ctx* ctx, const void* ptr, long long d0, long long d1, long long d2, long long ld1, int b0,
                         int b1, CUtensorMap* out);  // gemm_tc.cu

#include "attn_tc_common.cuh"

namespace {

constexpr int kFwdThreads = 384, kBwdThreads = 512;
constexpr int kTile = 16384;  // [128 rows x 64] bf16, 128B swizzle

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// release arrive: the shared-memory writes (P / dS tiles, made visible to the async proxy by the fence before it) and
// the TMEM reads of the arriving thread are ordered before the waiter's acquire
__device__ __forceinline__ void mbar_arrive_release(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// register re-balancing between the warpgroups of a CTA: reg_inc / reg_dec (tc_common.cuh)
__device__ __forceinline__ float ex2_approx(float x) {  // 2^x (ex2.approx.ftz: 2^-inf = +0)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void sts_f32(uint32_t saddr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

struct WsMaps {
  CUtensorMap q[2], k[2], v[2], d[2];  // [part]: hi (or the bf16 tensor itself), lo
};

struct WsArgs {
  int mode, B, S, H;
  float scale;
  const int64_t* traj;
  float* lse;
  void* o; long long ldo;                  // forward output (bf16: NP = 1, fp32: NP = 3)
  void* dq; void* dk; void* dv; long long ldd;
  DropArgs drop;        // dropout on the attention probabilities (thr == 0: off); mask row = drop.row0 + item * 128 + query
};

// 8 consecutive P values of row i (columns c8*8 ..) into the hi (and lo) tile
template <int NP>
__device__ __forceinline__ void store_p8_parts(uint8_t* base, int i, int c8, const float* v) {
  store_p8(base, i, c8, v);
  if (NP == 3) {
    float lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) lo[e] = v[e] - __bfloat162float(__float2bfloat16_rn(v[e]));
    store_p8(base + 2 * kTile, i, c8, lo);
  }
}

template <int NP>
__device__ __forceinline__ void store_out_row64(void* dst_base, long long elem_off, const uint32_t* r0, const uint32_t* r1,
                                                float mul) {
  if (NP == 1) {
    store_row64(reinterpret_cast<__nv_bfloat16*>(dst_base) + elem_off, r0, r1, mul);
  } else {
    float4* d = reinterpret_cast<float4*>(reinterpret_cast<float*>(dst_base) + elem_off);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      d[j] = make_float4(__uint_as_float(r0[4 * j]) * mul, __uint_as_float(r0[4 * j + 1]) * mul,
                         __uint_as_float(r0[4 * j + 2]) * mul, __uint_as_float(r0[4 * j + 3]) * mul);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      d[8 + j] = make_float4(__uint_as_float(r1[4 * j]) * mul, __uint_as_float(r1[4 * j + 1]) * mul,
                             __uint_as_float(r1[4 * j + 2]) * mul, __uint_as_float(r1[4 * j + 3]) * mul);
  }
}

// ------------------------------------------------------------------------------------------ forward
template <int MODE, int NP>
__global__ void __launch_bounds__(kFwdThreads, 1) attn_ws_fwd_kernel(const __grid_constant__ WsMaps m, WsArgs a) {
  constexpr int NPARTS = NP == 1 ? 1 : 2;
  constexpr int NS = NP == 1 ? 4 : 2;
  constexpr uint32_t kStage = 3 * NPARTS * kTile;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  int* sTraj = reinterpret_cast<int*>(smem + NS * kStage);                // [group 2][buffer 2][128]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NS * kStage + 2048);
  uint64_t* empty = full + NS;
  uint64_t* s_full = empty + NS;   // [2] S_g ready
  uint64_t* p_full = s_full + 2;   // [2] P_g written, S_g consumed
  uint64_t* o_full = p_full + 2;   // [2] O_g ready
  uint64_t* o_free = o_full + 2;   // [2] O_g consumed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_free + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_full[g], 128);
      mbar_init(&o_full[g], 1);
      mbar_init(&o_free[g], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const int S = a.S, H = a.H;
  const int total = a.B * H;
  const int n_items = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // blockIdx.x < total

  if (warp < 4) {
    reg_dec<64>();  // 64 + 216 + 216 = 496 of the 512 registers per thread-quad slot
    if (warp == 0) {
      // ================================ TMA producer ================================
      if (elect_one()) {
        for (int it = 0; it < n_items; ++it) {
          const int w = blockIdx.x + it * gridDim.x, b = w / H, h = w % H, s = it % NS;
          mbar_wait(&empty[s], ((it / NS) & 1) ^ 1);
          uint8_t* st = smem + s * kStage;
          mbar_expect_tx(&full[s], kStage);
#pragma unroll
          for (int p = 0; p < NPARTS; ++p) {
            tma_load_3d(st + p * kTile, &m.q[p], &full[s], h * DH, 0, b);
            tma_load_3d(st + (NPARTS + p) * kTile, &m.k[p], &full[s], h * DH, 0, b);
            tma_load_3d(st + (2 * NPARTS + p) * kTile, &m.v[p], &full[s], h * DH, 0, b);
          }
        }
      }
    } else if (warp == 1) {
      // ================================ MMA issuer ================================
      for (int it = 0; it <= n_items; ++it) {
        if (it < n_items) {
          const int s = it % NS, g = it & 1;
          mbar_wait(&full[s], (it / NS) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t q = smem_u32(smem + s * kStage), k = q + NPARTS * kTile;
            const uint32_t d = tmem + g * 128;
#pragma unroll
            for (int pr = 0; pr < NP; ++pr) {  // (Q_hi, K_hi), (Q_lo, K_hi), (Q_hi, K_lo)
              const uint32_t qa = q + (pr == 1 ? kTile : 0), kb = k + (pr == 2 ? kTile : 0);
#pragma unroll
              for (int kk = 0; kk < DH / 16; ++kk)
                umma_bf16(d, desc_kmajor(qa, kk), desc_kmajor(kb, kk), idesc(128, 128, false, false), (pr > 0 || kk > 0) ? 1u : 0u);
            }
            umma_commit(&s_full[g]);
          }
          __syncwarp();
        }
        if (it >= 1) {
          const int j = it - 1, s = j % NS, g = j & 1;
          mbar_wait(&p_full[g], (j >> 1) & 1);
          mbar_wait(&o_free[g], ((j >> 1) & 1) ^ 1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t p = smem_u32(smem + s * kStage), v = p + 2 * NPARTS * kTile;
            const uint32_t d = tmem + 256 + g * 64;
#pragma unroll
            for (int pr = 0; pr < NP; ++pr) {  // (P_hi, V_hi), (P_lo, V_hi), (P_hi, V_lo)
              const uint32_t pa = p + (pr == 1 ? 2 * kTile : 0), vb = v + (pr == 2 ? kTile : 0);
#pragma unroll
              for (int kk = 0; kk < TS / 16; ++kk)
                umma_bf16(d, desc_p_kmajor(pa, kk), desc_mnmajor64(vb, kk), idesc(128, 64, false, true), (pr > 0 || kk > 0) ? 1u : 0u);
            }
            umma_commit(&o_full[g]);
            umma_commit(&empty[s]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ================================ softmax groups ================================
    reg_inc<216>();
    const int g = (warp - 4) >> 2;
    const int i = ((warp & 3) << 5) + lane;  // query row == TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const float sl2 = a.scale * kLog2e;
    constexpr uint32_t kNegInf = 0xff800000u;
    for (int it = g; it < n_items; it += 2) {
      const int w = blockIdx.x + it * gridDim.x, b = w / H, h = w % H, s = it % NS;
      const int u = it >> 1;  // use index of this group's barriers
      const long long row0 = (long long)b * S;
      const int* traj = sTraj + (g * 2 + (u & 1)) * 128;
      if (MODE == SVLA_ATTN_TRAJ_CAUSAL) {
        sTraj[(g * 2 + (u & 1)) * 128 + i] = (i < S) ? (int)a.traj[row0 + i] : -1 - i;
        named_bar_sync(1 + g, 128);
      }
      uint8_t* sP = smem + s * kStage;
      mbar_wait(&s_full[g], u & 1);
      tc_fence_after();
      const uint32_t ts = tmem + lane_base + g * 128;
      uint32_t r[128];  // the whole score row of query i
      tmem_ld32(ts, r);
      tmem_ld32(ts + 32, r + 32);
      tmem_ld32(ts + 64, r + 64);
      tmem_ld32(ts + 96, r + 96);
      tmem_wait_ld();
      if (MODE == SVLA_ATTN_TRAJ_CAUSAL) {
        const int my_traj = traj[i];
#pragma unroll
        for (int j = 0; j < 128; ++j)
          if (!(j <= i && traj[j] == my_traj)) r[j] = kNegInf;  // columns >= S carry -1 - j: never equal
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c * 32 + 32 > S) {  // warp-uniform: only the chunk that straddles S (and the empty ones behind it)
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c * 32 + j >= S) r[c * 32 + j] = kNegInf;
          }
      }
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int j = 0; j < 128; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(r[j]));
      const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      const float mxs = (mx == -INFINITY) ? 0.f : mx * sl2;
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c8 = 0; c8 < 16; ++c8) {
        float p[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          p[e] = ex2_approx(fmaf(__uint_as_float(r[c8 * 8 + e]), sl2, -mxs));
          s4[e & 3] += p[e];
        }
        if (MODE == SVLA_ATTN_FULL && a.drop.thr != 0u) {  // normaliser = undropped row sum; P V sees the dropped row
          const uint32_t keep = dropout_keep8(a.drop, a.drop.row0 + (uint32_t)(w * 128 + i), (uint32_t)c8);
#pragma unroll
          for (int e = 0; e < 8; ++e) p[e] = ((keep >> e) & 1u) ? p[e] * a.drop.scale : 0.f;
        }
        store_p8_parts<NP>(sP, i, c8, p);
      }
      const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      fence_async_smem();
      tc_fence_before();
      mbar_arrive_release(&p_full[g]);
      // ---- O of this item
      mbar_wait(&o_full[g], u & 1);
      tc_fence_after();
      tmem_ld32(tmem + lane_base + 256 + g * 64, r);
      tmem_ld32(tmem + lane_base + 256 + g * 64 + 32, r + 32);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive_release(&o_free[g]);
      if (i < S) {
        store_out_row64<NP>(a.o, (row0 + i) * a.ldo + h * DH, r, r + 32, 1.f / sum);
        if (a.lse) a.lse[((long long)b * H + h) * S + i] = mx * a.scale + __logf(sum);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ backward
// Shared memory: 2 stages x (Q | K | V | dO) = 128 KB, P | dS = 64 KB.  TMEM: item buffer t = it & 1 at columns
// [256 t, 256 t + 256): S | dP, overwritten by dV [0, 64) | dK [64, 128) | dQ [128, 192) once P / dS are written.
template <int MODE>
__global__ void __launch_bounds__(kBwdThreads, 1) attn_ws_bwd_kernel(const __grid_constant__ WsMaps m, WsArgs a) {
  constexpr int NS = 2;
  constexpr uint32_t kStage = 4 * kTile;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sP = smem + NS * kStage;   // 32 KB
  uint8_t* sdS = sP + 2 * kTile;      // 32 KB
  int* sTraj = reinterpret_cast<int*>(sdS + 2 * kTile);          // [buffer 2][128]
  float* sDelta = reinterpret_cast<float*>(sTraj + 256);          // [buffer 2][half 2][128]
  uint64_t* full = reinterpret_cast<uint64_t*>(sDelta + 512);
  uint64_t* empty = full + NS;
  uint64_t* sdp_full = empty + NS;    // [2] S | dP of buffer t ready
  uint64_t* out_full = sdp_full + 2;  // [2] dV | dK | dQ of buffer t ready
  uint64_t* out_free = out_full + 2;  // [2] ... drained
  uint64_t* pds_full = out_free + 2;  // P and dS written
  uint64_t* pds_free = pds_full + 1;  // the gradient MMAs have read them
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pds_free + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&sdp_full[t], 1);
      mbar_init(&out_full[t], 1);
      mbar_init(&out_free[t], 128);
    }
    mbar_init(pds_full, 256);
    mbar_init(pds_free, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const int S = a.S, H = a.H;
  const int total = a.B * H;
  const int n_items = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp < 4) {
    reg_dec<56>();  // 56 + 176 + 176 + 104 = 512 of the 512 registers per thread-quad slot
    if (warp == 0) {
      // ================================ TMA producer ================================
      if (elect_one()) {
        for (int it = 0; it < n_items; ++it) {
          const int w = blockIdx.x + it * gridDim.x, b = w / H, h = w % H, s = it % NS;
          mbar_wait(&empty[s], ((it / NS) & 1) ^ 1);
          uint8_t* st = smem + s * kStage;
          mbar_expect_tx(&full[s], kStage);
          tma_load_3d(st, &m.q[0], &full[s], h * DH, 0, b);
          tma_load_3d(st + kTile, &m.k[0], &full[s], h * DH, 0, b);
          tma_load_3d(st + 2 * kTile, &m.v[0], &full[s], h * DH, 0, b);
          tma_load_3d(st + 3 * kTile, &m.d[0], &full[s], h * DH, 0, b);
        }
      }
    } else if (warp == 1) {
      // ================================ MMA issuer ================================
      for (int it = 0; it <= n_items; ++it) {
        if (it < n_items) {
          const int s = it % NS, t = it & 1;
          mbar_wait(&full[s], (it / NS) & 1);
          mbar_wait(&out_free[t], ((it >> 1) & 1) ^ 1);  // gradients of item it - 2 (same TMEM columns) are drained
          tc_fence_after();
          if (elect_one()) {
            const uint32_t q = smem_u32(smem + s * kStage), k = q + kTile, v = q + 2 * kTile, d = q + 3 * kTile;
            const uint32_t tb = tmem + t * 256;
#pragma unroll
            for (int kk = 0; kk < DH / 16; ++kk)
              umma_bf16(tb, desc_kmajor(q, kk), desc_kmajor(k, kk), idesc(128, 128, false, false), kk > 0);
#pragma unroll
            for (int kk = 0; kk < DH / 16; ++kk)
              umma_bf16(tb + 128, desc_kmajor(d, kk), desc_kmajor(v, kk), idesc(128, 128, false, false), kk > 0);
            umma_commit(&sdp_full[t]);
          }
          __syncwarp();
        }
        if (it >= 1) {
          const int j = it - 1, s = j % NS, t = j & 1;
          mbar_wait(pds_full, j & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t q = smem_u32(smem + s * kStage), k = q + kTile, d = q + 3 * kTile;
            const uint32_t p = smem_u32(sP), ds = smem_u32(sdS);
            const uint32_t tb = tmem + t * 256;
#pragma unroll
            for (int kk = 0; kk < TS / 16; ++kk)  // dV[keys, dh] = P^T dO
              umma_bf16(tb, desc_p_mnmajor(p, kk), desc_mnmajor64(d, kk), idesc(128, 64, true, true), kk > 0);
#pragma unroll
            for (int kk = 0; kk < TS / 16; ++kk)  // dK[keys, dh] = dS^T Q
              umma_bf16(tb + 64, desc_p_mnmajor(ds, kk), desc_mnmajor64(q, kk), idesc(128, 64, true, true), kk > 0);
#pragma unroll
            for (int kk = 0; kk < TS / 16; ++kk)  // dQ[queries, dh] = dS K
              umma_bf16(tb + 128, desc_p_kmajor(ds, kk), desc_mnmajor64(k, kk), idesc(128, 64, false, true), kk > 0);
            umma_commit(&out_full[t]);
            umma_commit(pds_free);
            umma_commit(&empty[s]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < 12) {
    // ================================ group A: P and dS (8 warps, 64 columns per thread) ================================
    reg_inc<176>();
    const int i = ((warp & 3) << 5) + lane;
    const int hf = (warp - 4) >> 2;  // column half [64 hf, 64 hf + 64)
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const float sl2 = a.scale * kLog2e;
    // the row's log-sum-exp is requested one item ahead: loaded next to its use it put an L2 / DRAM round trip in
    // front of every item (the accumulators are usually ready when this group gets to them)
    float lse_next = (n_items > 0 && i < S) ? a.lse[(long long)blockIdx.x * S + i] : 0.f;  // item w: rows w * S ..
    for (int it = 0; it < n_items; ++it) {
      const int w = blockIdx.x + it * gridDim.x, b = w / H, h = w % H, t = it & 1;
      const long long row0 = (long long)b * S;
      const int* traj = sTraj + t * 128;
      if (MODE == SVLA_ATTN_TRAJ_CAUSAL && hf == 0) sTraj[t * 128 + i] = (i < S) ? (int)a.traj[row0 + i] : -1 - i;
      const float lse2 = lse_next * kLog2e;
      mbar_wait(&sdp_full[t], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tb = tmem + lane_base + t * 256 + hf * 64;
      uint32_t rs[64], rp[64];
      tmem_ld32(tb, rs);
      tmem_ld32(tb + 32, rs + 32);
      tmem_ld32(tb + 128, rp);
      tmem_ld32(tb + 160, rp + 32);
      if (it + 1 < n_items && i < S) lse_next = a.lse[(long long)(w + gridDim.x) * S + i];
      if (MODE == SVLA_ATTN_TRAJ_CAUSAL) named_bar_sync(1, 256);  // sTraj visible (overlaps the TMEM loads)
      tmem_wait_ld();
      // P in place of S (fp32); key columns >= S (zero-filled K / V rows) and masked pairs give exactly 0
      float d4[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t kb[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};  // keep bits of this lane's 64 columns (dropout on P)
      if (MODE == SVLA_ATTN_FULL && a.drop.thr != 0u) {  // the decoder (TRAJ_CAUSAL) has no dropout
        kb[0] = kb[1] = 0u;
#pragma unroll
        for (int l8 = 0; l8 < 8; ++l8)
          kb[l8 >> 2] |= dropout_keep8(a.drop, a.drop.row0 + (uint32_t)(w * 128 + i), (uint32_t)(hf * 8 + l8)) << ((l8 & 3) * 8);
      }
      const float dsc = (MODE == SVLA_ATTN_FULL) ? a.drop.scale : 1.f;  // 1 when dropout is off
      if (MODE == SVLA_ATTN_TRAJ_CAUSAL) {
        const int my_traj = traj[i];
#pragma unroll
        for (int e = 0; e < 64; ++e) {
          const int col = hf * 64 + e;
          const bool ok = col <= i && traj[col] == my_traj && i < S;
          const float p = ok ? ex2_approx(fmaf(__uint_as_float(rs[e]), sl2, -lse2)) : 0.f;
          rs[e] = __float_as_uint(p);
          const float pm = ((kb[e >> 5] >> (e & 31)) & 1u) ? p * dsc : 0.f;
          d4[e & 3] = fmaf(pm, __uint_as_float(rp[e]), d4[e & 3]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col0 = hf * 64 + c * 32;
          if (col0 + 32 <= S) {  // warp-uniform
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const float p = ex2_approx(fmaf(__uint_as_float(rs[c * 32 + e]), sl2, -lse2));
              rs[c * 32 + e] = __float_as_uint(p);
              const float pm = ((kb[c] >> e) & 1u) ? p * dsc : 0.f;
              d4[e & 3] = fmaf(pm, __uint_as_float(rp[c * 32 + e]), d4[e & 3]);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const float p = (col0 + e < S) ? ex2_approx(fmaf(__uint_as_float(rs[c * 32 + e]), sl2, -lse2)) : 0.f;
              rs[c * 32 + e] = __float_as_uint(p);
              const float pm = ((kb[c] >> e) & 1u) ? p * dsc : 0.f;
              d4[e & 3] = fmaf(pm, __uint_as_float(rp[c * 32 + e]), d4[e & 3]);
            }
          }
        }
      }
      // delta_i = sum over both column halves, in a fixed order
      const uint32_t sd = smem_u32(sDelta + t * 256);
      sts_f32(sd + (hf * 128 + i) * 4, (d4[0] + d4[1]) + (d4[2] + d4[3]));
      named_bar_sync(2 + (warp & 3), 64);  // only the two warps of this TMEM lane quadrant exchange their halves
      const float delta = lds_f32(sd + i * 4) + lds_f32(sd + (128 + i) * 4);
      mbar_wait(pds_free, (it & 1) ^ 1);  // the previous item's gradient MMAs no longer read the P / dS tiles
#pragma unroll
      for (int l8 = 0; l8 < 8; ++l8) {
        const int c8 = hf * 8 + l8, chunk = c8 >> 3, cc = c8 & 7;
        const uint32_t off = chunk * kTile + i * 128 + ((cc ^ (i & 7)) << 4);
        uint4 up, ud;
        uint32_t* pw = reinterpret_cast<uint32_t*>(&up);
        uint32_t* dw = reinterpret_cast<uint32_t*>(&ud);
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          const int e = l8 * 8 + 2 * e2;
          const float p0 = __uint_as_float(rs[e]), p1 = __uint_as_float(rs[e + 1]);
          // P~ = P * keep / (1 - p) feeds dV = P~^T dO; dS = P~ dP - P delta (delta = rowsum(P~ dP))
          const float m0 = ((kb[e >> 5] >> (e & 31)) & 1u) ? p0 * dsc : 0.f;
          const float m1 = ((kb[(e + 1) >> 5] >> ((e + 1) & 31)) & 1u) ? p1 * dsc : 0.f;
          const __nv_bfloat162 hp = __floats2bfloat162_rn(m0, m1);
          const __nv_bfloat162 hd = __floats2bfloat162_rn(fmaf(m0, __uint_as_float(rp[e]), -p0 * delta),
                                                          fmaf(m1, __uint_as_float(rp[e + 1]), -p1 * delta));
          pw[e2] = *reinterpret_cast<const uint32_t*>(&hp);
          dw[e2] = *reinterpret_cast<const uint32_t*>(&hd);
        }
        sts128(smem_u32(sP) + off, up);
        sts128(smem_u32(sdS) + off, ud);
      }
      fence_async_smem();
      tc_fence_before();
      mbar_arrive_release(pds_full);
    }
  } else {
    // ================================ group B: dV | dK | dQ -> HBM ================================
    reg_dec<104>();
    const int i = ((warp & 3) << 5) + lane;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    for (int it = 0; it < n_items; ++it) {
      const int w = blockIdx.x + it * gridDim.x, b = w / H, h = w % H, t = it & 1;
      const long long orow = ((long long)b * S + i) * a.ldd + h * DH;
      mbar_wait(&out_full[t], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tb = tmem + lane_base + t * 256;
      uint32_t r0[32], r1[32];
      tmem_ld32(tb, r0);
      tmem_ld32(tb + 32, r1);
      tmem_wait_ld();
      if (i < S) store_row64(reinterpret_cast<__nv_bfloat16*>(a.dv) + orow, r0, r1, 1.f);
      tmem_ld32(tb + 64, r0);
      tmem_ld32(tb + 96, r1);
      tmem_wait_ld();
      if (i < S) store_row64(reinterpret_cast<__nv_bfloat16*>(a.dk) + orow, r0, r1, a.scale);
      tmem_ld32(tb + 128, r0);
      tmem_ld32(tb + 160, r1);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive_release(&out_free[t]);
      if (i < S) store_row64(reinterpret_cast<__nv_bfloat16*>(a.dq) + orow, r0, r1, a.scale);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

constexpr size_t kBwdSmem = 2 * 4 * kTile + 4 * kTile + 1024 + 2048 + 256 + 1024;

template <int MODE>
int launch_bwd(const WsMaps& m, const WsArgs& a, int grid, cudaStream_t st) {
  auto kern = attn_ws_bwd_kernel<MODE>;
  static bool attr = false;
  if (!attr) {
    SVLA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem));
    attr = true;
  }
  kern<<<grid, kBwdThreads, kBwdSmem, st>>>(m, a);
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}


// ------------------------------------------------------------------------------------------ backward, split operands
// Parity-grade variant (NP = 3): Q, K, V, dO arrive as (hi, lo) bf16 pairs (8 tiles = 128 KB, one stage), every product
// is hi*hi + lo*hi + hi*lo.  P (hi, lo) and dS (hi, lo) take turns in ONE 64 KB buffer: P -> dV = P^T dO retires -> dS
// (kept in registers meanwhile) -> dK = dS^T Q, dQ = dS K.  fp32 gradients out.
template <int MODE>
__global__ void __launch_bounds__(kBwdThreads, 1) attn_ws_bwd_x3_kernel(const __grid_constant__ WsMaps m, WsArgs a) {
  constexpr uint32_t kStage = 8 * kTile;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sPB = smem + kStage;       // 64 KB: P hi | P lo, then dS hi | dS lo
  int* sTraj = reinterpret_cast<int*>(sPB + 4 * kTile);          // [buffer 2][128]
  float* sDelta = reinterpret_cast<float*>(sTraj + 256);          // [buffer 2][half 2][128]
  uint64_t* full = reinterpret_cast<uint64_t*>(sDelta + 512);
  uint64_t* empty = full + 1;
  uint64_t* sdp_full = empty + 1;     // [2]
  uint64_t* out_full = sdp_full + 2;  // [2]
  uint64_t* out_free = out_full + 2;  // [2]
  uint64_t* p_full = out_free + 2;    // P written
  uint64_t* dv_done = p_full + 1;     // dV MMAs retired: the buffer may take dS
  uint64_t* ds_full = dv_done + 1;    // dS written
  uint64_t* pb_free = ds_full + 1;    // dK / dQ MMAs retired: the buffer may take the next item's P
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pb_free + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(full, 1);
    mbar_init(empty, 1);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&sdp_full[t], 1);
      mbar_init(&out_full[t], 1);
      mbar_init(&out_free[t], 128);
    }
    mbar_init(p_full, 256);
    mbar_init(dv_done, 1);
    mbar_init(ds_full, 256);
    mbar_init(pb_free, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const int S = a.S, H = a.H;
  const int total = a.B * H;
  const int n_items = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp < 4) {
    reg_dec<56>();
    if (warp == 0) {
      if (elect_one()) {
        for (int it = 0; it < n_items; ++it) {
          const int w = blockIdx.x + it * gridDim.x, b = w / H, h = w % H;
          mbar_wait(empty, (it & 1) ^ 1);
          mbar_expect_tx(full, kStage);
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            tma_load_3d(smem + p * kTile, &m.q[p], full, h * DH, 0, b);
            tma_load_3d(smem + (2 + p) * kTile, &m.k[p], full, h * DH, 0, b);
            tma_load_3d(smem + (4 + p) * kTile, &m.v[p], full, h * DH, 0, b);
            tma_load_3d(smem + (6 + p) * kTile, &m.d[p], full, h * DH, 0, b);
          }
        }
      }
    } else if (warp == 1) {
      const uint32_t q = smem_u32(smem), k = q + 2 * kTile, v = q + 4 * kTile, d = q + 6 * kTile, pb = smem_u32(sPB);
      for (int it = 0; it < n_items; ++it) {
        const int t = it & 1;
        const uint32_t tb = tmem + t * 256;
        mbar_wait(full, it & 1);
        mbar_wait(&out_free[t], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int pr = 0; pr < 3; ++pr) {  // S = Q K^T
            const uint32_t xa = q + (pr == 1 ? kTile : 0), xb = k + (pr == 2 ? kTile : 0);
#pragma unroll
            for (int kk = 0; kk < DH / 16; ++kk)
              umma_bf16(tb, desc_kmajor(xa, kk), desc_kmajor(xb, kk), idesc(128, 128, false, false), (pr > 0 || kk > 0) ? 1u : 0u);
          }
#pragma unroll
          for (int pr = 0; pr < 3; ++pr) {  // dP = dO V^T
            const uint32_t xa = d + (pr == 1 ? kTile : 0), xb = v + (pr == 2 ? kTile : 0);
#pragma unroll
            for (int kk = 0; kk < DH / 16; ++kk)
              umma_bf16(tb + 128, desc_kmajor(xa, kk), desc_kmajor(xb, kk), idesc(128, 128, false, false), (pr > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(&sdp_full[t]);
        }
        __syncwarp();
        mbar_wait(p_full, it & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int pr = 0; pr < 3; ++pr) {  // dV[keys, dh] = P^T dO
            const uint32_t xa = pb + (pr == 1 ? 2 * kTile : 0), xb = d + (pr == 2 ? kTile : 0);
#pragma unroll
            for (int kk = 0; kk < TS / 16; ++kk)
              umma_bf16(tb, desc_p_mnmajor(xa, kk), desc_mnmajor64(xb, kk), idesc(128, 64, true, true), (pr > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(dv_done);
        }
        __syncwarp();
        mbar_wait(ds_full, it & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int pr = 0; pr < 3; ++pr) {  // dK[keys, dh] = dS^T Q
            const uint32_t xa = pb + (pr == 1 ? 2 * kTile : 0), xb = q + (pr == 2 ? kTile : 0);
#pragma unroll
            for (int kk = 0; kk < TS / 16; ++kk)
              umma_bf16(tb + 64, desc_p_mnmajor(xa, kk), desc_mnmajor64(xb, kk), idesc(128, 64, true, true), (pr > 0 || kk > 0) ? 1u : 0u);
          }
#pragma unroll
          for (int pr = 0; pr < 3; ++pr) {  // dQ[queries, dh] = dS K
            const uint32_t xa = pb + (pr == 1 ? 2 * kTile : 0), xb = k + (pr == 2 ? kTile : 0);
#pragma unroll
            for (int kk = 0; kk < TS / 16; ++kk)
              umma_bf16(tb + 128, desc_p_kmajor(xa, kk), desc_mnmajor64(xb, kk), idesc(128, 64, false, true), (pr > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(&out_full[t]);
          umma_commit(pb_free);
          umma_commit(empty);
        }
        __syncwarp();
      }
    }
  } else if (warp < 12) {
    reg_inc<176>();
    const int i = ((warp & 3) << 5) + lane;
    const int hf = (warp - 4) >> 2;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const float sl2 = a.scale * kLog2e;
    // the row's log-sum-exp is requested one item ahead: loaded next to its use it put an L2 / DRAM round trip in
    // front of every item (the accumulators are usually ready when this group gets to them)
    float lse_next = (n_items > 0 && i < S) ? a.lse[(long long)blockIdx.x * S + i] : 0.f;  // item w: rows w * S ..
    for (int it = 0; it < n_items; ++it) {
      const int w = blockIdx.x + it * gridDim.x, b = w / H, h = w % H, t = it & 1;
      const long long row0 = (long long)b * S;
      const int* traj = sTraj + t * 128;
      if (MODE == SVLA_ATTN_TRAJ_CAUSAL && hf == 0) sTraj[t * 128 + i] = (i < S) ? (int)a.traj[row0 + i] : -1 - i;
      const float lse2 = lse_next * kLog2e;
      mbar_wait(&sdp_full[t], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tb = tmem + lane_base + t * 256 + hf * 64;
      uint32_t rs[64], rp[64];
      tmem_ld32(tb, rs);
      tmem_ld32(tb + 32, rs + 32);
      tmem_ld32(tb + 128, rp);
      tmem_ld32(tb + 160, rp + 32);
      if (it + 1 < n_items && i < S) lse_next = a.lse[(long long)(w + gridDim.x) * S + i];
      if (MODE == SVLA_ATTN_TRAJ_CAUSAL) named_bar_sync(1, 256);
      tmem_wait_ld();
      float d4[4] = {0.f, 0.f, 0.f, 0.f};
      const int my_traj = (MODE == SVLA_ATTN_TRAJ_CAUSAL) ? traj[i] : 0;
#pragma unroll
      for (int e = 0; e < 64; ++e) {
        const int col = hf * 64 + e;
        bool ok = col < S;
        if (MODE == SVLA_ATTN_TRAJ_CAUSAL) ok = col <= i && traj[col] == my_traj && i < S;
        const float p = ok ? ex2_approx(fmaf(__uint_as_float(rs[e]), sl2, -lse2)) : 0.f;
        rs[e] = __float_as_uint(p);
        d4[e & 3] = fmaf(p, __uint_as_float(rp[e]), d4[e & 3]);
      }
      const uint32_t sd = smem_u32(sDelta + t * 256);
      sts_f32(sd + (hf * 128 + i) * 4, (d4[0] + d4[1]) + (d4[2] + d4[3]));
      named_bar_sync(2 + (warp & 3), 64);  // only the two warps of this TMEM lane quadrant exchange their halves
      const float delta = lds_f32(sd + i * 4) + lds_f32(sd + (128 + i) * 4);
      mbar_wait(pb_free, (it & 1) ^ 1);
#pragma unroll
      for (int l8 = 0; l8 < 8; ++l8) {
        float p[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) p[e] = __uint_as_float(rs[l8 * 8 + e]);
        store_p8_parts<3>(sPB, i, hf * 8 + l8, p);
      }
      fence_async_smem();
      tc_fence_before();
      mbar_arrive_release(p_full);
      // dS = P (dP - delta), in place of dP, while dV = P^T dO runs
#pragma unroll
      for (int e = 0; e < 64; ++e) rp[e] = __float_as_uint(__uint_as_float(rs[e]) * (__uint_as_float(rp[e]) - delta));
      mbar_wait(dv_done, it & 1);
#pragma unroll
      for (int l8 = 0; l8 < 8; ++l8) {
        float ds[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) ds[e] = __uint_as_float(rp[l8 * 8 + e]);
        store_p8_parts<3>(sPB, i, hf * 8 + l8, ds);
      }
      fence_async_smem();
      tc_fence_before();
      mbar_arrive_release(ds_full);
    }
  } else {
    reg_dec<104>();
    const int i = ((warp & 3) << 5) + lane;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    for (int it = 0; it < n_items; ++it) {
      const int w = blockIdx.x + it * gridDim.x, b = w / H, h = w % H, t = it & 1;
      const long long orow = ((long long)b * S + i) * a.ldd + h * DH;
      mbar_wait(&out_full[t], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tb = tmem + lane_base + t * 256;
      uint32_t r0[32], r1[32];
      tmem_ld32(tb, r0);
      tmem_ld32(tb + 32, r1);
      tmem_wait_ld();
      if (i < S) store_out_row64<3>(a.dv, orow, r0, r1, 1.f);
      tmem_ld32(tb + 64, r0);
      tmem_ld32(tb + 96, r1);
      tmem_wait_ld();
      if (i < S) store_out_row64<3>(a.dk, orow, r0, r1, a.scale);
      tmem_ld32(tb + 128, r0);
      tmem_ld32(tb + 160, r1);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive_release(&out_free[t]);
      if (i < S) store_out_row64<3>(a.dq, orow, r0, r1, a.scale);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

constexpr size_t kBwdX3Smem = 8 * kTile + 4 * kTile + 1024 + 2048 + 256 + 1024;

template <int MODE>
int launch_bwd_x3(const WsMaps& m, const WsArgs& a, int grid, cudaStream_t st) {
  auto kern = attn_ws_bwd_x3_kernel<MODE>;
  static bool attr = false;
  if (!attr) {
    SVLA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdX3Smem));
    attr = true;
  }
  kern<<<grid, kBwdThreads, kBwdX3Smem, st>>>(m, a);
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

template <int NP> constexpr size_t fwd_smem_bytes() {
  return (size_t)(NP == 1 ? 4 : 2) * 3 * (NP == 1 ? 1 : 2) * kTile + 2048 + 256 + 1024;
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int MODE, int NP>
int launch_fwd(const WsMaps& m, const WsArgs& a, int grid, cudaStream_t st) {
  constexpr size_t smem = fwd_smem_bytes<NP>();
  auto kern = attn_ws_fwd_kernel<MODE, NP>;
  static bool attr = false;
  if (!attr) {
    SVLA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  kern<<<grid, kFwdThreads, smem, st>>>(m, a);
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

}  // namespace

bool svla_attn_ws_supported(int mode, int dtype, int S, int dh, long long ld, long long ldo, const void* q, const void* k,
                            const void* v, const void* o) {
  return dtype == SVLA_BF16 && dh == DH && S >= 1 && S <= TS && (mode == SVLA_ATTN_FULL || mode == SVLA_ATTN_TRAJ_CAUSAL) &&
         ld % 8 == 0 && ldo % 8 == 0 && al16(q) && al16(k) && al16(v) && al16(o);
}

static int attn_ws_fwd_impl(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld, void* o,
                            long long ldo, float* lse, const int64_t* traj, int B, int S, int H, float scale,
                            const svla_dropout* drop, cudaStream_t st) {
  WsMaps m{};
  int rc;
  if ((rc = svla_make_tmap3_bf16(ctx, q, (long long)H * DH, S, B, ld, DH, TS, &m.q[0]))) return rc;
  if ((rc = svla_make_tmap3_bf16(ctx, k, (long long)H * DH, S, B, ld, DH, TS, &m.k[0]))) return rc;
  if ((rc = svla_make_tmap3_bf16(ctx, v, (long long)H * DH, S, B, ld, DH, TS, &m.v[0]))) return rc;
  WsArgs a{};
  a.mode = mode; a.B = B; a.S = S; a.H = H; a.scale = scale; a.traj = traj; a.lse = lse; a.o = o; a.ldo = ldo;
  a.drop = make_drop_args(drop);
  const int grid = std::min(B * H, ctx->sm_count);
  if (mode == SVLA_ATTN_FULL) return launch_fwd<SVLA_ATTN_FULL, 1>(m, a, grid, st);
  return launch_fwd<SVLA_ATTN_TRAJ_CAUSAL, 1>(m, a, grid, st);
}

int svla_attn_ws_fwd(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld, void* o,
                     long long ldo, float* lse, const int64_t* traj, int B, int S, int H, float scale, cudaStream_t st) {
  return attn_ws_fwd_impl(ctx, mode, q, k, v, ld, o, ldo, lse, traj, B, S, H, scale, nullptr, st);
}

static int attn_ws_bwd_impl(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld,
                            const void* d_o, long long ldo, void* dq, void* dk, void* dv, long long ldd, const float* lse,
                            const int64_t* traj, int B, int S, int H, float scale, const svla_dropout* drop,
                            cudaStream_t st) {
  WsMaps m{};
  int rc;
  if ((rc = svla_make_tmap3_bf16(ctx, q, (long long)H * DH, S, B, ld, DH, TS, &m.q[0]))) return rc;
  if ((rc = svla_make_tmap3_bf16(ctx, k, (long long)H * DH, S, B, ld, DH, TS, &m.k[0]))) return rc;
  if ((rc = svla_make_tmap3_bf16(ctx, v, (long long)H * DH, S, B, ld, DH, TS, &m.v[0]))) return rc;
  if ((rc = svla_make_tmap3_bf16(ctx, d_o, (long long)H * DH, S, B, ldo, DH, TS, &m.d[0]))) return rc;
  WsArgs a{};
  a.mode = mode; a.B = B; a.S = S; a.H = H; a.scale = scale; a.traj = traj; a.lse = const_cast<float*>(lse);
  a.dq = dq; a.dk = dk; a.dv = dv; a.ldd = ldd;
  a.drop = make_drop_args(drop);
  const int grid = std::min(B * H, ctx->sm_count);
  if (mode == SVLA_ATTN_FULL) return launch_bwd<SVLA_ATTN_FULL>(m, a, grid, st);
  return launch_bwd<SVLA_ATTN_TRAJ_CAUSAL>(m, a, grid, st);
}

int svla_attn_ws_bwd(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld, const void* d_o,
                     long long ldo, void* dq, void* dk, void* dv, long long ldd, const float* lse, const int64_t* traj,
                     int B, int S, int H, float scale, cudaStream_t st) {
  return attn_ws_bwd_impl(ctx, mode, q, k, v, ld, d_o, ldo, dq, dk, dv, ldd, lse, traj, B, S, H, scale, nullptr, st);
}

// attn_tc2.cu: 128 < S <= 256 (the two-camera fusion block)
bool svla_attn_tc2_supported(int mode, int dtype, int S, int dh, long long ld, long long ldo, const void* q,
                             const void* k, const void* v, const void* o);
int svla_attn_tc2_fwd_drop(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld, void* o,
                           long long ldo, float* lse, const int64_t* traj, int B, int S, int H, float scale,
                           const svla_dropout* drop, cudaStream_t st);
int svla_attn_tc2_bwd_drop(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld,
                           const void* o, const void* d_o, long long ldo, void* dq, void* dk, void* dv, long long ldd,
                           const float* lse, const int64_t* traj, int B, int S, int H, float scale,
                           const svla_dropout* drop, cudaStream_t st);

extern "C" int svla_attn_drop_fwd(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld,
                                  void* o, long long ldo, float* lse, const int64_t* traj, int B, int S, int H, int dh,
                                  float scale, const svla_dropout* drop, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && q && k && v && o, "NULL argument");
  SVLA_CHECK_ARG(svla_dropout_ok(drop), "dropout p must be in [0, 1)");
  const bool two_tile = S > TS && svla_attn_tc2_supported(mode, SVLA_BF16, S, dh, ld, ldo, q, k, v, o);
  SVLA_CHECK_ARG(two_tile || svla_attn_ws_supported(mode, SVLA_BF16, S, dh, ld, ldo, q, k, v, o),
                 "attention with dropout: bf16, S <= 256, head dim 64, FULL / TRAJ_CAUSAL, 16-byte aligned operands");
  SVLA_CHECK_ARG(mode != SVLA_ATTN_TRAJ_CAUSAL || traj, "TRAJ_CAUSAL needs traj");
  SVLA_CHECK_ARG(mode == SVLA_ATTN_FULL || !drop || drop->p == 0.f, "dropout exists for mode FULL (the fusion block) only");
  if (B <= 0) return SVLA_OK;
  if (two_tile)
    return svla_attn_tc2_fwd_drop(ctx, mode, q, k, v, ld, o, ldo, lse, traj, B, S, H, scale, drop, as_stream(stream));
  return attn_ws_fwd_impl(ctx, mode, q, k, v, ld, o, ldo, lse, traj, B, S, H, scale, drop, as_stream(stream));
}

extern "C" int svla_attn_drop_bwd(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld,
                                  const void* o, const void* d_o, long long ldo, void* dq, void* dk, void* dv,
                                  long long ldd, const float* lse, const int64_t* traj, int B, int S, int H, int dh,
                                  float scale, const svla_dropout* drop, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && q && k && v && d_o && dq && dk && dv && lse, "NULL argument");
  SVLA_CHECK_ARG(svla_dropout_ok(drop), "dropout p must be in [0, 1)");
  const bool two_tile = S > TS && svla_attn_tc2_supported(mode, SVLA_BF16, S, dh, ld, ldo, q, k, v, d_o) &&
                        ldd % 8 == 0 && al16(dq) && al16(dk) && al16(dv);
  SVLA_CHECK_ARG(two_tile || (svla_attn_ws_supported(mode, SVLA_BF16, S, dh, ld, ldo, q, k, v, d_o) && ldd % 8 == 0 &&
                              al16(dq) && al16(dk) && al16(dv)),
                 "attention with dropout: bf16, S <= 256, head dim 64, FULL / TRAJ_CAUSAL, 16-byte aligned operands");
  SVLA_CHECK_ARG(mode != SVLA_ATTN_TRAJ_CAUSAL || traj, "TRAJ_CAUSAL needs traj");
  SVLA_CHECK_ARG(mode == SVLA_ATTN_FULL || !drop || drop->p == 0.f, "dropout exists for mode FULL (the fusion block) only");
  if (B <= 0) return SVLA_OK;
  if (two_tile) {  // the two-tile kernel takes delta from dO . O: it needs the forward output
    SVLA_CHECK_ARG(o && al16(o), "attention backward for 128 < S <= 256 needs the forward output o");
    return svla_attn_tc2_bwd_drop(ctx, mode, q, k, v, ld, o, d_o, ldo, dq, dk, dv, ldd, lse, traj, B, S, H, scale, drop,
                                  as_stream(stream));
  }
  return attn_ws_bwd_impl(ctx, mode, q, k, v, ld, d_o, ldo, dq, dk, dv, ldd, lse, traj, B, S, H, scale, drop,
                          as_stream(stream));
}

// ---- split-operand (parity-grade) entry points: x_lo = x_hi + lo_off elements (svla_split_concat, axis 1, {0, 1})
static bool split_args_ok(int mode, int S, int dh, long long ld, long long lo_off, const void* q, const void* k,
                          const void* v) {
  return dh == DH && S >= 1 && S <= TS && (mode == SVLA_ATTN_FULL || mode == SVLA_ATTN_TRAJ_CAUSAL) && ld % 8 == 0 &&
         lo_off % 8 == 0 && al16(q) && al16(k) && al16(v);
}

extern "C" int svla_attn_split_fwd(svla_ctx* ctx, int mode, const void* q_hi, const void* k_hi, const void* v_hi,
                                   long long lo_off, long long ld, float* o, long long ldo, float* lse,
                                   const int64_t* traj, int B, int S, int H, int dh, float scale, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && q_hi && k_hi && v_hi && o, "NULL argument");
  SVLA_CHECK_ARG(split_args_ok(mode, S, dh, ld, lo_off, q_hi, k_hi, v_hi) && ldo % 4 == 0 && al16(o),
                 "split attention: S <= 128, head dim 64, FULL / TRAJ_CAUSAL, 16-byte aligned operands");
  SVLA_CHECK_ARG(mode != SVLA_ATTN_TRAJ_CAUSAL || traj, "TRAJ_CAUSAL needs traj");
  if (B <= 0) return SVLA_OK;
  WsMaps m{};
  int rc;
  const __nv_bfloat16* qs[2] = {reinterpret_cast<const __nv_bfloat16*>(q_hi), reinterpret_cast<const __nv_bfloat16*>(q_hi) + lo_off};
  const __nv_bfloat16* ks[2] = {reinterpret_cast<const __nv_bfloat16*>(k_hi), reinterpret_cast<const __nv_bfloat16*>(k_hi) + lo_off};
  const __nv_bfloat16* vs[2] = {reinterpret_cast<const __nv_bfloat16*>(v_hi), reinterpret_cast<const __nv_bfloat16*>(v_hi) + lo_off};
  for (int p = 0; p < 2; ++p) {
    if ((rc = svla_make_tmap3_bf16(ctx, qs[p], (long long)H * DH, S, B, ld, DH, TS, &m.q[p]))) return rc;
    if ((rc = svla_make_tmap3_bf16(ctx, ks[p], (long long)H * DH, S, B, ld, DH, TS, &m.k[p]))) return rc;
    if ((rc = svla_make_tmap3_bf16(ctx, vs[p], (long long)H * DH, S, B, ld, DH, TS, &m.v[p]))) return rc;
  }
  WsArgs a{};
  a.mode = mode; a.B = B; a.S = S; a.H = H; a.scale = scale; a.traj = traj; a.lse = lse; a.o = o; a.ldo = ldo;
  const int grid = std::min(B * H, ctx->sm_count);
  if (mode == SVLA_ATTN_FULL) return launch_fwd<SVLA_ATTN_FULL, 3>(m, a, grid, as_stream(stream));
  return launch_fwd<SVLA_ATTN_TRAJ_CAUSAL, 3>(m, a, grid, as_stream(stream));
}

extern "C" int svla_attn_split_bwd(svla_ctx* ctx, int mode, const void* q_hi, const void* k_hi, const void* v_hi,
                                   long long lo_off, long long ld, const void* do_hi, long long do_lo_off, long long lddo,
                                   float* dq, float* dk, float* dv, long long ldd, const float* lse, const int64_t* traj,
                                   int B, int S, int H, int dh, float scale, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && q_hi && k_hi && v_hi && do_hi && dq && dk && dv && lse, "NULL argument");
  SVLA_CHECK_ARG(split_args_ok(mode, S, dh, ld, lo_off, q_hi, k_hi, v_hi) && lddo % 8 == 0 && do_lo_off % 8 == 0 &&
                     al16(do_hi) && ldd % 4 == 0 && al16(dq) && al16(dk) && al16(dv),
                 "split attention: S <= 128, head dim 64, FULL / TRAJ_CAUSAL, 16-byte aligned operands");
  SVLA_CHECK_ARG(mode != SVLA_ATTN_TRAJ_CAUSAL || traj, "TRAJ_CAUSAL needs traj");
  if (B <= 0) return SVLA_OK;
  WsMaps m{};
  int rc;
  const __nv_bfloat16* qs[2] = {reinterpret_cast<const __nv_bfloat16*>(q_hi), reinterpret_cast<const __nv_bfloat16*>(q_hi) + lo_off};
  const __nv_bfloat16* ks[2] = {reinterpret_cast<const __nv_bfloat16*>(k_hi), reinterpret_cast<const __nv_bfloat16*>(k_hi) + lo_off};
  const __nv_bfloat16* vs[2] = {reinterpret_cast<const __nv_bfloat16*>(v_hi), reinterpret_cast<const __nv_bfloat16*>(v_hi) + lo_off};
  const __nv_bfloat16* ds[2] = {reinterpret_cast<const __nv_bfloat16*>(do_hi), reinterpret_cast<const __nv_bfloat16*>(do_hi) + do_lo_off};
  for (int p = 0; p < 2; ++p) {
    if ((rc = svla_make_tmap3_bf16(ctx, qs[p], (long long)H * DH, S, B, ld, DH, TS, &m.q[p]))) return rc;
    if ((rc = svla_make_tmap3_bf16(ctx, ks[p], (long long)H * DH, S, B, ld, DH, TS, &m.k[p]))) return rc;
    if ((rc = svla_make_tmap3_bf16(ctx, vs[p], (long long)H * DH, S, B, ld, DH, TS, &m.v[p]))) return rc;
    if ((rc = svla_make_tmap3_bf16(ctx, ds[p], (long long)H * DH, S, B, lddo, DH, TS, &m.d[p]))) return rc;
  }
  WsArgs a{};
  a.mode = mode; a.B = B; a.S = S; a.H = H; a.scale = scale; a.traj = traj; a.lse = const_cast<float*>(lse);
  a.dq = dq; a.dk = dk; a.dv = dv; a.ldd = ldd;
  const int grid = std::min(B * H, ctx->sm_count);
  if (mode == SVLA_ATTN_FULL) return launch_bwd_x3<SVLA_ATTN_FULL>(m, a, grid, as_stream(stream));
  return launch_bwd_x3<SVLA_ATTN_TRAJ_CAUSAL>(m, a, grid, as_stream(stream));
}
