// tcgen05 GEMM with CTA pairs (cta_group::2): two SMs of a TPC cooperate on a 256 x 256 output tile.
// Each CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256 N rows), so per FLOP the pair moves
// a third less data through L2 and shared memory than two independent 128 x 256 tiles -- which is what bounds the
// single-CTA kernel on the K = 512 GEMMs of the towers.  Structure per CTA as in gemm_tc.cu (TMA producer warp,
// MMA issuer, 8 epilogue warps, double-buffered TMEM accumulators), plus:
//   * TMA loads of BOTH CTAs signal the leader's full barrier (cp.async.bulk.tensor ... cta_group::2);
//   * only the leader issues tcgen05.mma.cta_group::2 (M = 256: rows 0-127 accumulate in the leader's TMEM, rows
//     128-255 in the peer's); tcgen05.commit ... multicast::cluster releases the smem stage / publishes the
//     accumulator in both CTAs;
//   * the peer's epilogue warps arrive remotely on the leader's "accumulator drained" barrier.
#include <algorithm>
#include <cstdlib>

#include "tc_common.cuh"

int svla_make_tmap(svla_ctx* ctx, const void* ptr, long long inner, long long outer, long long ld, int bi, int bo,
                   CUtensorMap* out, int kind);

namespace {

constexpr int BN2 = 256;               // tile N (each CTA stages BN2 / 2 rows of B)
constexpr int kStages2 = 6;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> leader CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {  // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  // relaxed: the accumulator hand-off is ordered by tcgen05.fence::before_thread_sync, not by a memory fence (a
  // .release.cluster arrive costs a MEMBAR + ERRBAR per tile and warp)
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerMask) : "memory");
}

// ASUM (weight-gradient launches, A = dY stored [rows, N_out]): the kernel also produces the bias gradient
// asum[m] += sum_k A^T[m, k] on the tensor cores -- for the tn == 0 tile of every row block one extra N = 16 MMA per
// k-step multiplies the same A tile with a shared-memory tile of ones into 16 spare TMEM columns.  It replaces a
// separate pass over dY (svla_colsum).  These launches run 5 pipeline stages (the sixth slot holds the ones) and a
// single accumulator stage (a weight gradient has at most a couple of tiles per CTA pair: nothing to overlap).
template <bool AMN, bool BMN, bool ASUM>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
svla_gemm_tc2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                     const __grid_constant__ CUtensorMap mapC, const __grid_constant__ CUtensorMap mapC2, TcArgs g) {
  constexpr uint32_t kABytes = BM * BK * 2, kBBytes = (BN2 / 2) * BK * 2;
  constexpr uint32_t kStageBytes = kABytes + kBBytes;  // 32 KB per CTA
  constexpr uint32_t kTmemCols = 2 * BN2;
  // instruction descriptor: D=f32, A=B=bf16, majors, N=256, M=256 (pair)
  constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((AMN ? 1u : 0u) << 15) | ((BMN ? 1u : 0u) << 16) |
                              ((uint32_t)(BN2 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  constexpr uint32_t kIdescOnes = (1u << 4) | (1u << 7) | (1u << 10) | ((AMN ? 1u : 0u) << 15) |
                                  ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  constexpr int kNS = ASUM ? kStages2 - 1 : kStages2;  // pipeline stages in use
  constexpr int kAcc = ASUM ? 1 : 2;                   // accumulator stages in use
  constexpr uint32_t kSumCol = BN2;                    // TMEM column of the ones-product (ASUM)

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem + kStages2 * kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_base + 8 * kStgBytes);
  uint64_t* empty_bar = full_bar + kStages2;
  uint64_t* tfull_bar = empty_bar + kStages2;
  uint64_t* tempty_bar = tfull_bar + 2;  // the leader's copy is the one that is waited on
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (warp == 0 && elect_one()) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapC) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapC2) : "memory");
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < kStages2; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 16);  // 8 epilogue warps of each CTA
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (ASUM && warp == 3) {  // [8 rows x 64] bf16 ones: this CTA's half of the N = 16 operand (any layout: all ones)
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem + kNS * kStageBytes);
    for (int i = lane; i < 256; i += 32) ones[i] = 0x3F803F80u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();  // CTA-local ordering of the TMEM base / barrier init ahead of the cluster barrier (racecheck still
                    // reports the pair-wide tcgen05.alloc.cta_group::2 itself: both CTAs' allocators write the base
                    // into both CTAs' shared memory, a protocol the tool does not model; profiles/sanitizer_r02.md)
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int tiles_m2 = (g.M + 255) / 256;
  const int total_work = tiles_m2 * g.tiles_n * g.splits;
  const int kb_total = (g.K + BK - 1) / BK;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp < kEpiWarp0) {
    reg_dec<40>();  // 4 x 40 + 8 x 232 registers per thread-quad slot: the epilogue warps hold a tile of side operand
  }
  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = cluster_id; w < total_work; w += num_clusters) {
        const int tn = w % g.tiles_n, tm2 = (w / g.tiles_n) % tiles_m2, sp = w / (g.tiles_n * tiles_m2);
        const int kb0 = sp * g.kb_per_split, kb1 = min(kb_total, kb0 + g.kb_per_split);
        const int mrow = tm2 * 256 + (int)rank * BM, nrow = tn * BN2 + (int)rank * (BN2 / 2);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * kStageBytes);  // bytes of both CTAs land on this barrier
          if (!AMN) {
            tma_load_2d_2sm(sa, &mapA, &full_bar[stage], kb * BK, mrow);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d_2sm(sa + j * (BK * 128), &mapA, &full_bar[stage], mrow + j * 64, kb * BK);
          }
          if (!BMN) {
            tma_load_2d_2sm(sb, &mapB, &full_bar[stage], kb * BK, nrow);
          } else {
#pragma unroll
            for (int j = 0; j < (BN2 / 2) / 64; ++j)
              tma_load_2d_2sm(sb + j * (BK * 128), &mapB, &full_bar[stage], nrow + j * 64, kb * BK);
          }
          if (++stage == kNS) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ================================ MMA issuer (leader CTA only) ================================
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = cluster_id; w < total_work; w += num_clusters) {
      const int sp = w / (g.tiles_n * tiles_m2);
      const bool do_sum = ASUM && g.asum != nullptr && (w % g.tiles_n) == 0;
      const int kb0 = sp * g.kb_per_split, kb1 = min(kb_total, kb0 + g.kb_per_split);
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);  // both CTAs have drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN2;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = AMN ? umma_desc(sa + k * 2048, BK * 128, 1024) : umma_desc(sa + k * 32, 16, 1024);
            const uint64_t db = BMN ? umma_desc(sb + k * 2048, BK * 128, 1024) : umma_desc(sb + k * 32, 16, 1024);
            umma_bf16_2sm(tmem_d, da, db, kIdesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if (do_sum) {
            const uint64_t dones = umma_desc(smem_u32(smem + kNS * kStageBytes), 16, 1024);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t da = AMN ? umma_desc(sa + k * 2048, BK * 128, 1024) : umma_desc(sa + k * 32, 16, 1024);
              umma_bf16_2sm(tmem_base + kSumCol, da, dones, kIdescOnes, (kb > kb0 || k > 0) ? 1u : 0u);
            }
          }
        }
        __syncwarp();
        if (elect_one()) umma_commit_2sm(&empty_bar[stage]);
        __syncwarp();
        if (++stage == kNS) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit_2sm(&tfull_bar[acc]);
      __syncwarp();
      if (++acc == kAcc) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= kEpiWarp0) {
    // ================================ epilogue (both CTAs, own 128 rows) ================================
    reg_inc<232>();
    const int q = warp & 3;
    const int half = (warp - kEpiWarp0) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool part = g.splits > 1;
    const int dtC = part ? (int)SVLA_F32 : g.dtypeC;
    const bool staged = (!g.residual || g.dtypeR == dtC) && (!g.aux || g.dtypeAux == dtC);
    const int cb = half * (BN2 / 2), ce = cb + BN2 / 2;
    SidePre pre;
    pre.valid = 0;
    pre.mb = make_uint4(0u, 0u, 0u, 0u);
    for (int w = cluster_id; w < total_work; w += num_clusters) {
      const int tn = w % g.tiles_n, tm2 = (w / g.tiles_n) % tiles_m2, sp = w / (g.tiles_n * tiles_m2);
      const int m0 = tm2 * 256 + (int)rank * BM + q * 32;
      const int m = m0 + lane;
      const bool row_ok = m < g.M;
      const int wn = w + num_clusters;  // the tile this warp drains next: its side operand is requested early
      const int next_m0 = wn < total_work ? ((wn / g.tiles_n) % tiles_m2) * 256 + (int)rank * BM + q * 32 : -1;
      const int next_nt0 = (wn % g.tiles_n) * BN2;
      bias_prefetch(pre, g, tn * BN2 + cb, lane, 4);
      bits_prefetch(pre, g, m0, tn * BN2 + cb, lane, ce - cb);
      if (staged) side_prefetch_first(pre, g, m0, tn * BN2 + cb, lane, 4);  // in flight while the MMAs finish
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (g.dbg == 1) {
      } else if (g.dbg == 2) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN2);
        uint32_t keep = 0;
        for (int c0 = cb; c0 < ce; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(taddr + c0, r);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e) keep ^= r[e];
        }
        if (keep == 0x12345678u) reinterpret_cast<uint32_t*>(g.C)[0] = keep;
      } else {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN2);
        uint8_t* stg = stage_base + (warp - kEpiWarp0) * kStgBytes;
        if (!staged) epilogue_direct(g, taddr, m, row_ok, tn * BN2, cb, ce, sp);
        else if (dtC == SVLA_F32)
          epilogue_staged_t<true>(g, &mapC, stg, taddr, m0, tn * BN2, cb, ce, sp, lane, pre, next_m0, next_nt0, &mapC2);
        else
          epilogue_staged_t<false>(g, &mapC, stg, taddr, m0, tn * BN2, cb, ce, sp, lane, pre, next_m0, next_nt0, &mapC2);
        if (ASUM && g.asum != nullptr && tn == 0 && half == 0) {  // every column of the ones-product is the row sum
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + kSumCol, r);
          tmem_wait_ld();
          if (row_ok) {
            if (part) g.asum_ws[(size_t)sp * g.M + m] = __uint_as_float(r[0]);
            else g.asum[m] += __uint_as_float(r[0]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tempty_bar[acc]);
      if (++acc == kAcc) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

__global__ void __launch_bounds__(256) tc2_splitk_reduce_kernel(TcArgs g) {
  const long long total = (long long)g.M * g.N;
  if (g.asum) {
    for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < g.M; m += (long long)gridDim.x * blockDim.x) {
      float v = 0.f;
      for (int s = 0; s < g.splits; ++s) v += g.asum_ws[(size_t)s * g.M + m];
      g.asum[m] += v;
    }
  }
  // weight gradients (every split launch of the towers): fp32 C, no bias / epilogue / residual -- four elements per
  // thread and 16-byte accesses; the same split order per element as the scalar loop below, so identical bits
  if (g.dtypeC == SVLA_F32 && !g.bias && g.epilogue == SVLA_EPI_NONE && !g.residual && (g.ldc & 3) == 0 &&
      (reinterpret_cast<uintptr_t>(g.C) & 15) == 0) {
    float* C = reinterpret_cast<float*>(g.C);
    const long long total4 = total >> 2;  // N is a multiple of 64
    for (long long i4 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i4 < total4; i4 += (long long)gridDim.x * blockDim.x) {
      const long long i = i4 << 2;
      const int m = (int)(i / g.N), n = (int)(i % g.N);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s = 0; s < g.splits; ++s) {
        const float4 p = __ldcs(reinterpret_cast<const float4*>(g.ws + (size_t)s * total + i));
        v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
      }
      v.x *= g.alpha; v.y *= g.alpha; v.z *= g.alpha; v.w *= g.alpha;
      float4* c = reinterpret_cast<float4*>(C + (long long)m * g.ldc + n);
      if (g.accumulate) {
        const float4 o = *c;
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
      }
      *c = v;
    }
    return;
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / g.N), n = (int)(i % g.N);
    float v = 0.f;
    for (int s = 0; s < g.splits; ++s) v += g.ws[(size_t)s * total + i];
    v *= g.alpha;
    if (g.bias) v += __ldg(g.bias + n);
    if (g.epilogue == SVLA_EPI_RELU) v = fmaxf(v, 0.f);
    else if (g.epilogue == SVLA_EPI_GELU) v = gelu_erf(v);
    else if (g.epilogue == SVLA_EPI_RELU_MASK) v = ld_elem(g.aux, g.dtypeAux, (long long)m * g.ldaux + n) > 0.f ? v : 0.f;
    if (g.residual) v += ld_elem(g.residual, g.dtypeR, (long long)m * g.ldr + n);
    const long long ci = (long long)m * g.ldc + n;
    if (g.accumulate) v += ld_elem(g.C, g.dtypeC, ci);
    if (g.dtypeC == SVLA_F32) reinterpret_cast<float*>(g.C)[ci] = v;
    else reinterpret_cast<__nv_bfloat16*>(g.C)[ci] = __float2bfloat16_rn(v);
  }
}

template <bool AMN, bool BMN, bool ASUM>
int launch_tc2(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mc2, const TcArgs& g,
               int grid, cudaStream_t st) {
  constexpr size_t smem = (size_t)kStages2 * (BM * BK * 2 + (BN2 / 2) * BK * 2) + 1024 + 8 * kStgBytes + 512;
  auto kern = svla_gemm_tc2_kernel<AMN, BMN, ASUM>;
  static bool attr_set = false;
  if (!attr_set) {
    SVLA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  kern<<<grid, kThreads, smem, st>>>(ma, mb, mc, mc2, g);  // cluster shape comes from __cluster_dims__
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

}  // namespace

bool svla_gemm_tc2_supported(const svla_gemm_desc* d) {
  // the single-CTA kernel's conditions are checked by the caller; the pair additionally wants a full 256-wide tile
  return d->M >= 256 && d->N >= 256;
}

int svla_gemm_tc2(svla_ctx* ctx, const svla_gemm_desc* d, cudaStream_t st) {
  const bool amn = d->transA != 0, bmn = d->transB == 0;
  TcArgs g;
  g.M = d->M; g.N = d->N; g.K = d->K;
  const int tiles_m2 = (d->M + 255) / 256;
  g.tiles_m = tiles_m2;
  g.tiles_n = (d->N + BN2 - 1) / BN2;
  const int kb_total = (d->K + BK - 1) / BK;
  const int clusters = ctx->sm_count / 2;
  int splits = 1;
  const int tiles = tiles_m2 * g.tiles_n;
  if (d->transA && tiles * 2 <= clusters && kb_total >= 32) {  // weight gradients only (see gemm_tc.cu)
    // split K so that tiles * splits fills whole waves of CTA pairs: the smallest split count within 3 % of the best
    // wave occupancy (16 tiles on 74 pairs: 4 splits leave 14 % of the chip idle, 9 splits 3 %)
    const size_t per = (size_t)d->M * ((size_t)d->N + 1) * sizeof(float);
    const int smax = (int)std::min<size_t>((size_t)std::min(kb_total / 8, 32), ctx->ws_bytes / std::max<size_t>(per, 1));
    double best = 0.0;
    for (int s = 1; s <= smax; ++s) {
      const int work = tiles * s, waves = (work + clusters - 1) / clusters;
      best = std::max(best, (double)work / ((double)waves * clusters));
    }
    for (int s = 1; s <= smax; ++s) {
      const int work = tiles * s, waves = (work + clusters - 1) / clusters;
      if ((double)work / ((double)waves * clusters) >= best - 0.03) { splits = s; break; }
    }
  }
  g.kb_per_split = (kb_total + splits - 1) / splits;
  g.splits = (kb_total + g.kb_per_split - 1) / g.kb_per_split;
  g.C = d->C; g.ldc = d->ldc; g.dtypeC = d->dtypeC;
  g.bias = d->bias;
  g.residual = d->residual; g.ldr = d->ldr; g.dtypeR = d->dtypeR;
  g.aux = d->aux; g.ldaux = d->ldaux; g.dtypeAux = d->dtypeAux;
  g.epilogue = d->epilogue; g.accumulate = d->accumulate; g.alpha = d->alpha;
  g.drop = make_drop_args(d->epilogue == SVLA_EPI_RELU_BITS ? d->dropout : nullptr);
  g.ws = reinterpret_cast<float*>(ctx->ws);
  g.asum = (amn && bmn) ? d->colsum_a : nullptr;
  g.asum_ws = g.ws + (size_t)g.splits * d->M * d->N;
  static const int dbg_env = getenv("SVLA_TC_DBG") ? atoi(getenv("SVLA_TC_DBG")) : 0;
  g.dbg = dbg_env;

  CUtensorMap ma, mb, mc;
  int rc;
  if (!amn) rc = svla_make_tmap(ctx, d->A, d->K, d->M, d->lda, BK, BM, &ma, 0);   // [M rows][K]  box {64, 128}
  else rc = svla_make_tmap(ctx, d->A, d->M, d->K, d->lda, 64, BK, &ma, 0);        // [K rows][M]  box {64, 64}
  if (rc) return rc;
  if (!bmn) rc = svla_make_tmap(ctx, d->B, d->K, d->N, d->ldb, BK, BN2 / 2, &mb, 0);  // [N rows][K]  box {64, 128}
  else rc = svla_make_tmap(ctx, d->B, d->N, d->K, d->ldb, 64, BK, &mb, 0);            // [K rows][N]  box {64, 64}
  if (rc) return rc;
  mc = ma;
  CUtensorMap mc2 = ma;
  g.tma_store = 0;
  if (g.splits == 1) {
    rc = svla_make_tmap(ctx, d->C, d->N, d->M, d->ldc, 32, 32, &mc, d->dtypeC == SVLA_F32 ? 2 : 1);
    if (rc) return rc;
    mc2 = mc;
    if (d->dtypeC == SVLA_BF16) {  // 64-column blocks: [32 rows x 128 B] boxes, 128B swizzle
      rc = svla_make_tmap(ctx, d->C, d->N, d->M, d->ldc, 64, 32, &mc2, 0);
      if (rc) return rc;
    }
    g.tma_store = 1;
  }
  const int grid = 2 * std::min(tiles * g.splits, clusters);
  if (!amn && !bmn) rc = launch_tc2<false, false, false>(ma, mb, mc, mc2, g, grid, st);
  else if (!amn && bmn) rc = launch_tc2<false, true, false>(ma, mb, mc, mc2, g, grid, st);
  else if (g.asum) rc = launch_tc2<true, true, true>(ma, mb, mc, mc2, g, grid, st);
  else rc = launch_tc2<true, true, false>(ma, mb, mc, mc2, g, grid, st);
  if (rc) return rc;
  if (g.splits > 1) {
    const long long total = (long long)d->M * d->N;
    const int rb = (int)std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 8);
    tc2_splitk_reduce_kernel<<<rb, 256, 0, st>>>(g);
    SVLA_LAUNCH_CHECK();
  }
  return SVLA_OK;
}
