// tcgen05 tensor-core GEMM for sm_100a: bf16 operands, fp32 accumulation in TMEM.
//
//   C[M,N] = epi(alpha * op(A) op(B) + bias) [+ residual]          (see svla_gemm_desc)
//
// One persistent kernel, warp-specialised:
//   warp 0     TMA producer   cp.async.bulk.tensor (128B swizzle) global -> shared, 4-stage mbarrier ring
//   warp 1     MMA issuer     one elected thread issues tcgen05.mma (cta_group::1, M=128, N=BN, K=16 per
//                             instruction), accumulators double-buffered in TMEM (2 x BN columns)
//   warp 2     TMEM allocator
//   warps 4-7  epilogue       tcgen05.ld -> registers -> bias / ReLU / ReLU-mask / residual / accumulate ->
//                             128-bit global stores; overlaps the next tile's main loop
// Operand layouts: both K-major (forward: activations x nn.Linear weight), K-major x MN-major (dgrad:
// dY x W), MN-major x MN-major (wgrad: dY^T x X) -- the UMMA shared-memory descriptors and the TMA boxes
// differ, the pipeline does not, so no operand is ever transposed in HBM.
// Split-K (wgrad: K = rows of the batch) writes fp32 slices to the context workspace; a fixed-order
// reduction kernel applies the epilogue (deterministic).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

#include "tc_common.cuh"

namespace {

// AMN / BMN: operand is MN-major (its M / N dimension is the contiguous one in global memory)
template <int BN, bool AMN, bool BMN>
__global__ void __launch_bounds__(kThreads, 1)
svla_gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    const __grid_constant__ CUtensorMap mapC, const __grid_constant__ CUtensorMap mapC2, TcArgs g) {
  constexpr int kStages = (BN == 256) ? 4 : 6;
  constexpr uint32_t kABytes = BM * BK * 2, kBBytes = BN * BK * 2;
  constexpr uint32_t kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kTmemCols = 2 * BN;  // double-buffered accumulator
  // instruction descriptor: D=f32, A=B=bf16, majors, N>>3, M>>4
  constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((AMN ? 1u : 0u) << 15) | ((BMN ? 1u : 0u) << 16) |
                              ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem + kStages * kStageBytes;  // 8 epilogue warps x 4 KB staging (1024-byte aligned)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_base + 8 * kStgBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;   // [2] accumulator ready for the epilogue
  uint64_t* tempty_bar = tfull_bar + 2;        // [2] accumulator drained
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && elect_one()) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapC) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapC2) : "memory");
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int total_work = g.tiles_m * g.tiles_n * g.splits;
  const int kb_total = (g.K + BK - 1) / BK;

  if (warp < kEpiWarp0) {
    reg_dec<40>();  // 4 x 40 + 8 x 232 registers per thread-quad slot: the epilogue warps hold a tile of side operand
  }
  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int tn = w % g.tiles_n, tm = (w / g.tiles_n) % g.tiles_m, sp = w / (g.tiles_n * g.tiles_m);
        const int kb0 = sp * g.kb_per_split, kb1 = min(kb_total, kb0 + g.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          mbar_expect_tx(&full_bar[stage], kStageBytes);
          if (!AMN) {
            tma_load_2d(sa, &mapA, &full_bar[stage], kb * BK, tm * BM);  // box {64 k, 128 rows}
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)  // boxes {64 m, 64 k}
              tma_load_2d(sa + j * (BK * 128), &mapA, &full_bar[stage], tm * BM + j * 64, kb * BK);
          }
          if (!BMN) {
            tma_load_2d(sb, &mapB, &full_bar[stage], kb * BK, tn * BN);  // box {64 k, BN rows}
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sb + j * (BK * 128), &mapB, &full_bar[stage], tn * BN + j * 64, kb * BK);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const int sp = w / (g.tiles_n * g.tiles_m);
      const int kb0 = sp * g.kb_per_split, kb1 = min(kb_total, kb0 + g.kb_per_split);
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: 8-row groups 1024 B apart, +32 B per 16-element k step inside the 128 B swizzle row.
            // MN-major: 64-element chunks BK*128 B apart (LBO), 8-k-row groups 1024 B apart (SBO),
            //           +16 rows * 128 B per k step.
            const uint64_t da = AMN ? umma_desc(sa + k * 2048, BK * 128, 1024) : umma_desc(sa + k * 32, 16, 1024);
            const uint64_t db = BMN ? umma_desc(sb + k * 2048, BK * 128, 1024) : umma_desc(sb + k * 32, 16, 1024);
            umma_bf16(tmem_d, da, db, kIdesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
        }
        __syncwarp();
        if (elect_one()) umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= kEpiWarp0) {
    reg_inc<232>();
    // ================================ epilogue ================================
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int half = (warp - kEpiWarp0) >> 2;  // which half of the tile's columns this warp drains
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool part = g.splits > 1;
    const int dtC = part ? (int)SVLA_F32 : g.dtypeC;
    const bool staged = (!g.residual || g.dtypeR == dtC) && (!g.aux || g.dtypeAux == dtC);
    const int cb = half * (BN / 2), ce = cb + BN / 2;
    SidePre pre;
    pre.valid = 0;
    pre.mb = make_uint4(0u, 0u, 0u, 0u);
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const int tn = w % g.tiles_n, tm = (w / g.tiles_n) % g.tiles_m, sp = w / (g.tiles_n * g.tiles_m);
      const int m = tm * BM + q * 32 + lane;
      const bool row_ok = m < g.M;
      const int wn = w + gridDim.x;  // the tile this warp drains next: its side operand is requested early
      const int next_m0 = wn < total_work ? ((wn / g.tiles_n) % g.tiles_m) * BM + q * 32 : -1;
      const int next_nt0 = (wn % g.tiles_n) * BN;
      bias_prefetch(pre, g, tn * BN + cb, lane, (ce - cb) / 32);
      bits_prefetch(pre, g, tm * BM + q * 32, tn * BN + cb, lane, ce - cb);
      if (staged && g.dbg == 0) side_prefetch_first(pre, g, tm * BM + q * 32, tn * BN + cb, lane, (ce - cb) / 32);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (g.dbg == 1) {
      } else if (g.dbg == 2) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
        uint32_t keep = 0;
        for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
          uint32_t r[32];
          tmem_ld32(taddr + c0, r);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e) keep ^= r[e];
        }
        if (keep == 0x12345678u) reinterpret_cast<uint32_t*>(g.C)[0] = keep;
      } else {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
        uint8_t* stg = stage_base + (warp - kEpiWarp0) * kStgBytes;
        if (!staged) epilogue_direct(g, taddr, m, row_ok, tn * BN, cb, ce, sp);
        else if (dtC == SVLA_F32)
          epilogue_staged_t<true>(g, &mapC, stg, taddr, tm * BM + q * 32, tn * BN, cb, ce, sp, lane, pre, next_m0, next_nt0, &mapC2);
        else
          epilogue_staged_t<false>(g, &mapC, stg, taddr, tm * BM + q * 32, tn * BN, cb, ce, sp, lane, pre, next_m0, next_nt0, &mapC2);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// fixed-order split-K fold + epilogue
__global__ void __launch_bounds__(256) tc_splitk_reduce_kernel(TcArgs g) {
  const long long total = (long long)g.M * g.N;
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

// ---------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct TmapCache {
  std::mutex mu;
  std::map<std::tuple<const void*, long long, long long, long long, int, int, int>, CUtensorMap> maps;
};

// 2-D tensor map: inner (contiguous) extent, outer extent, outer stride ld (elements), box {bi, bo}.
// kind 0: bf16, 128B swizzle (operands); kind 1: bf16, 64B swizzle (32-column output tiles); kind 2: fp32, 128B swizzle
int make_tmap(svla_ctx* ctx, const void* ptr, long long inner, long long outer, long long ld, int bi, int bo,
              CUtensorMap* out, int kind = 0) {
  if (!ctx->tmap_cache) ctx->tmap_cache = new TmapCache();
  TmapCache* tc = reinterpret_cast<TmapCache*>(ctx->tmap_cache);
  const auto key = std::make_tuple(ptr, inner, outer, ld, bi, bo, kind);
  {
    std::lock_guard<std::mutex> lk(tc->mu);
    auto it = tc->maps.find(key);
    if (it != tc->maps.end()) {
      *out = it->second;
      return SVLA_OK;
    }
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    svla_set_error("cuTensorMapEncodeTiled is not available from the driver");
    return SVLA_ERR_INTERNAL;
  }
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * (kind == 2 ? 4 : 2)};
  cuuint32_t box[2] = {(cuuint32_t)bi, (cuuint32_t)bo};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, kind == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   kind == 1 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    svla_set_error("cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%lld outer=%lld ld=%lld box=%dx%d", (int)r, ptr,
                   inner, outer, ld, bi, bo);
    return SVLA_ERR_INTERNAL;
  }
  std::lock_guard<std::mutex> lk(tc->mu);
  if (tc->maps.size() > 8192) tc->maps.clear();
  tc->maps[key] = *out;
  return SVLA_OK;
}

// 3-D bf16 tensor map over [d2][d1][d0] (d0 contiguous; row stride ld1 elements, d2 stride d1 * ld1), box {b0, b1, 1}, 128B
// swizzle.  The attention kernels address a (sequence, head) tile as (head * 64, 0, sequence): rows beyond the
// sequence (d1) are zero-filled on loads and clipped on stores instead of touching the neighbouring sequence.
int make_tmap3(svla_ctx* ctx, const void* ptr, long long d0, long long d1, long long d2, long long ld1, int b0, int b1,
               CUtensorMap* out) {
  if (!ctx->tmap_cache) ctx->tmap_cache = new TmapCache();
  TmapCache* tc = reinterpret_cast<TmapCache*>(ctx->tmap_cache);
  if (d1 >= 4096) {
    svla_set_error("make_tmap3: sequence extent %lld too large", d1);
    return SVLA_ERR_BAD_ARG;
  }
  const long long ld2 = d1 * ld1;  // sequences are stored back to back
  const auto key = std::make_tuple(ptr, d0 * 4096 + d1, d2, ld1, b0, b1, 3);
  {
    std::lock_guard<std::mutex> lk(tc->mu);
    auto it = tc->maps.find(key);
    if (it != tc->maps.end()) {
      *out = it->second;
      return SVLA_OK;
    }
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    svla_set_error("cuTensorMapEncodeTiled is not available from the driver");
    return SVLA_ERR_INTERNAL;
  }
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)ld1 * 2, (cuuint64_t)ld2 * 2};
  cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    svla_set_error("cuTensorMapEncodeTiled (3-D) failed (%d): ptr=%p dims=%lld x %lld x %lld ld=%lld,%lld box=%dx%d", (int)r,
                   ptr, d0, d1, d2, ld1, ld2, b0, b1);
    return SVLA_ERR_INTERNAL;
  }
  std::lock_guard<std::mutex> lk(tc->mu);
  if (tc->maps.size() > 8192) tc->maps.clear();
  tc->maps[key] = *out;
  return SVLA_OK;
}

template <int BN, bool AMN, bool BMN>
int launch_tc(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mc2, const TcArgs& g,
              int grid, cudaStream_t st) {
  constexpr int kStages = (BN == 256) ? 4 : 6;
  constexpr size_t smem = (size_t)kStages * (BM * BK * 2 + BN * BK * 2) + 1024 + 8 * kStgBytes + 512;
  auto kern = svla_gemm_tc_kernel<BN, AMN, BMN>;
  static bool attr_set = false;
  if (!attr_set) {
    SVLA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  kern<<<grid, kThreads, smem, st>>>(ma, mb, mc, mc2, g);
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

void svla_tmap_cache_free(void* cache) { delete reinterpret_cast<TmapCache*>(cache); }

int svla_make_tmap_bf16(svla_ctx* ctx, const void* ptr, long long inner, long long outer, long long ld, int bi, int bo,
                        CUtensorMap* out) {
  return make_tmap(ctx, ptr, inner, outer, ld, bi, bo, out);
}
int svla_make_tmap3_bf16(svla_ctx* ctx, const void* ptr, long long d0, long long d1, long long d2, long long ld1, int b0,
                         int b1, CUtensorMap* out) {
  return make_tmap3(ctx, ptr, d0, d1, d2, ld1, b0, b1, out);
}
int svla_make_tmap(svla_ctx* ctx, const void* ptr, long long inner, long long outer, long long ld, int bi, int bo,
                   CUtensorMap* out, int kind) {
  return make_tmap(ctx, ptr, inner, outer, ld, bi, bo, out, kind);
}
int svla_gemm_tc2(svla_ctx* ctx, const svla_gemm_desc* d, cudaStream_t st);  // gemm_tc2.cu (cta_group::2)
bool svla_gemm_tc2_supported(const svla_gemm_desc* d);

bool svla_gemm_tc_supported(const svla_gemm_desc* d) {
  if (d->dtypeA != SVLA_BF16 || d->dtypeB != SVLA_BF16) return false;
  if (d->transA && !d->transB) {
    // MN x MN (wgrad)
  } else if (!d->transA) {
    // K x K (transB) or K x MN (!transB)
  } else {
    return false;  // MN-major A with K-major B is not needed by the towers
  }
  if (d->M < 64 || d->N < 64 || d->N % 64 != 0 || d->K < 64) return false;
  if (d->lda % 8 || d->ldb % 8 || !al16(d->A) || !al16(d->B)) return false;
  const int esC = d->dtypeC == SVLA_F32 ? 4 : 2;
  if (!al16(d->C) || (d->ldc * esC) % 16) return false;
  if (d->bias && !al16(d->bias)) return false;
  if (d->residual && (!al16(d->residual) || (d->ldr * (d->dtypeR == SVLA_F32 ? 4 : 2)) % 16)) return false;
  if (d->aux && (!al16(d->aux) || (d->ldaux * (d->dtypeAux == SVLA_F32 ? 4 : 2)) % 16)) return false;
  return true;
}

static bool use_tc2() {
  static const int v = getenv("SVLA_TC2") ? atoi(getenv("SVLA_TC2")) : 1;
  return v != 0;
}

// true when svla_gemm_tc produces d->colsum_a itself (pair kernel, weight-gradient operand layout)
bool svla_gemm_tc_fuses_colsum(const svla_gemm_desc* d) {
  return d->colsum_a && d->transA && !d->transB && use_tc2() && svla_gemm_tc_supported(d) && svla_gemm_tc2_supported(d);
}

int svla_gemm_tc(svla_ctx* ctx, const svla_gemm_desc* d, cudaStream_t st) {
  if (use_tc2() && svla_gemm_tc2_supported(d)) return svla_gemm_tc2(ctx, d, st);
  const bool amn = d->transA != 0, bmn = d->transB == 0;
  const int BN = (d->N >= 256) ? 256 : 128;
  TcArgs g;
  g.M = d->M; g.N = d->N; g.K = d->K;
  g.tiles_m = (d->M + BM - 1) / BM;
  g.tiles_n = (d->N + BN - 1) / BN;
  const int kb_total = (d->K + BK - 1) / BK;
  int splits = 1;
  const int tiles = g.tiles_m * g.tiles_n;
  // split K only when K is the batch (row) dimension, i.e. weight-gradient launches: forward / dgrad results then do
  // not depend on how many rows the launch carries, so a sampler shard reproduces the full batch bit for bit
  if (d->transA && tiles * 2 <= ctx->sm_count && kb_total >= 32) {
    splits = std::min({ctx->sm_count / tiles, kb_total / 8, 32});
    const size_t per = (size_t)d->M * d->N * sizeof(float);
    splits = (int)std::min<size_t>((size_t)splits, ctx->ws_bytes / std::max<size_t>(per, 1));
    splits = std::max(splits, 1);
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
  g.asum = nullptr; g.asum_ws = nullptr;
  static const int dbg_env = getenv("SVLA_TC_DBG") ? atoi(getenv("SVLA_TC_DBG")) : 0;
  g.dbg = dbg_env;

  CUtensorMap ma, mb;
  int rc;
  if (!amn) rc = make_tmap(ctx, d->A, d->K, d->M, d->lda, BK, BM, &ma);       // [M rows][K]  box {64, 128}
  else rc = make_tmap(ctx, d->A, d->M, d->K, d->lda, 64, BK, &ma);            // [K rows][M]  box {64, 64}
  if (rc) return rc;
  if (!bmn) rc = make_tmap(ctx, d->B, d->K, d->N, d->ldb, BK, BN, &mb);       // [N rows][K]  box {64, BN}
  else rc = make_tmap(ctx, d->B, d->N, d->K, d->ldb, 64, BK, &mb);            // [K rows][N]  box {64, 64}
  if (rc) return rc;

  // output tiles leave through TMA stores (32 x 32 boxes out of the swizzled staging tile) unless this launch
  // writes split-K partials
  CUtensorMap mc = ma, mc2 = ma;
  g.tma_store = 0;
  if (g.splits == 1) {
    rc = make_tmap(ctx, d->C, d->N, d->M, d->ldc, 32, 32, &mc, d->dtypeC == SVLA_F32 ? 2 : 1);
    if (rc) return rc;
    mc2 = mc;
    if (d->dtypeC == SVLA_BF16) {  // 64-column blocks: [32 rows x 128 B] boxes, 128B swizzle
      rc = make_tmap(ctx, d->C, d->N, d->M, d->ldc, 64, 32, &mc2, 0);
      if (rc) return rc;
    }
    g.tma_store = 1;
  }
  const int grid = std::min(tiles * g.splits, ctx->sm_count);
  if (BN == 256) {
    if (!amn && !bmn) rc = launch_tc<256, false, false>(ma, mb, mc, mc2, g, grid, st);
    else if (!amn && bmn) rc = launch_tc<256, false, true>(ma, mb, mc, mc2, g, grid, st);
    else rc = launch_tc<256, true, true>(ma, mb, mc, mc2, g, grid, st);
  } else {
    if (!amn && !bmn) rc = launch_tc<128, false, false>(ma, mb, mc, mc2, g, grid, st);
    else if (!amn && bmn) rc = launch_tc<128, false, true>(ma, mb, mc, mc2, g, grid, st);
    else rc = launch_tc<128, true, true>(ma, mb, mc, mc2, g, grid, st);
  }
  if (rc) return rc;
  if (g.splits > 1) {
    const long long total = (long long)d->M * d->N;
    const int rb = (int)std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 8);
    tc_splitk_reduce_kernel<<<rb, 256, 0, st>>>(g);
    SVLA_LAUNCH_CHECK();
  }
  return SVLA_OK;
}
