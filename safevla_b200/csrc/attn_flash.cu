// tcgen05 flash-attention forward for sequences longer than 256 tokens (head dim 64, bf16, no mask): the
// DINOv2 ViT-S/14 blocks of the rollout-side vision preprocessor (433 tokens for a 224 x 378 crop, 257 for
// 224 x 224; reference architecture/allenact_preprocessors/dino_preprocessors.py:20-38).
//
// Work item = (sequence, head, 128-row query tile); 128 threads, thread = query row.  Key tiles of 128 rows stream
// through shared memory:  S = Q K_j^T (tcgen05, fp32 in TMEM) -> online softmax in registers (running max / sum,
// exp2) -> P (bf16, swizzled) -> PV = P V_j (tcgen05, 64 TMEM columns) -> O = O * corr + PV in registers.
// 80 KB + 256 TMEM columns => 2 CTAs per SM, whose load / MMA / softmax phases overlap.
#include <cuda.h>

#include <algorithm>

#include "common.cuh"

int svla_make_tmap_bf16(svla_ctx* ctx, const void* ptr, long long inner, long long outer, long long ld, int bi, int bo,
                        CUtensorMap* out);  // gemm_tc.cu

#include "attn_tc_common.cuh"

namespace {

__global__ void __launch_bounds__(128)
attn_flash_fwd_kernel(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mk,
                      const __grid_constant__ CUtensorMap mv, AttnTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;             // 16 KB each
  uint8_t* sK = smem + 16384;
  uint8_t* sV = smem + 32768;
  uint8_t* sP = smem + 49152;     // 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 81920);  // load, mma
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr uint32_t kCols = 256;  // S [0,128), PV [128,192)
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
  const int nt = (S + TS - 1) / TS;
  const float sl2 = a.scale * kLog2e;
  uint32_t ph_load = 0, ph_mma = 0;
  const long long total = (long long)a.B * a.H * nt;
  for (long long w = blockIdx.x; w < total; w += gridDim.x) {
    const int qt = (int)(w % nt);
    const int h = (int)((w / nt) % a.H), b = (int)(w / ((long long)nt * a.H));
    const int row0 = b * S;
    const int i = qt * TS + tid;  // query row inside the sequence
    float m = -INFINITY, l = 0.f, o[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] = 0.f;
    for (int j = 0; j < nt; ++j) {
      if (tid == 0) {
        mbar_expect_tx(&bars[0], (j == 0 ? 3 : 2) * 16384);
        if (j == 0) tma_load_2d(sQ, &mq, &bars[0], h * DH, row0 + qt * TS);
        tma_load_2d(sK, &mk, &bars[0], h * DH, row0 + j * TS);
        tma_load_2d(sV, &mv, &bars[0], h * DH, row0 + j * TS);
      }
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
      mbar_wait(&bars[1], ph_mma);
      ph_mma ^= 1;
      tc_fence_after();
      const int ncols = min(TS, S - j * TS);  // real keys in this tile
      float mx = m;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        if (c * 32 >= ncols) break;
        uint32_t r[32];
        tmem_ld32(tmem + lane_base + c * 32, r);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (c * 32 + e < ncols) mx = fmaxf(mx, __uint_as_float(r[e]));
      }
      const float mxs = mx * sl2;  // ncols >= 1, so mx is finite
      const float corr = (m == -INFINITY) ? 0.f : exp2f(m * sl2 - mxs);
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        if (c * 32 < ncols) {
          tmem_ld32(tmem + lane_base + c * 32, r);
          tmem_wait_ld();
        }
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          float p[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int col = c * 32 + j8 * 8 + e;
            p[e] = (col < ncols) ? exp2f(__uint_as_float(r[j8 * 8 + e]) * sl2 - mxs) : 0.f;
            sum += p[e];
          }
          store_p8(sP, tid, c * 4 + j8, p);
        }
      }
      m = mx;
      l = l * corr + sum;
      fence_async_smem();
      tc_fence_before();
      __syncthreads();  // P complete, every thread is done reading S
      if (tid == 0) {
        tc_fence_after();
        const uint32_t p = smem_u32(sP), v = smem_u32(sV);
#pragma unroll
        for (int kk = 0; kk < TS / 16; ++kk)
          umma_bf16(tmem + 128, desc_p_kmajor(p, kk), desc_mnmajor64(v, kk), idesc(128, 64, false, true), kk > 0);
        umma_commit(&bars[1]);
      }
      mbar_wait(&bars[1], ph_mma);
      ph_mma ^= 1;
      tc_fence_after();
      {
        uint32_t r0[32], r1[32];
        tmem_ld32(tmem + lane_base + 128, r0);
        tmem_ld32(tmem + lane_base + 160, r1);
        tmem_wait_ld();
#pragma unroll
        for (int d = 0; d < 32; ++d) {
          o[d] = fmaf(o[d], corr, __uint_as_float(r0[d]));
          o[32 + d] = fmaf(o[32 + d], corr, __uint_as_float(r1[d]));
        }
      }
      tc_fence_before();
      __syncthreads();  // K, V, P and both accumulators are free for the next key tile
    }
    if (i < S) {
      const float inv = 1.f / l;
      __nv_bfloat16* dst = a.o + (long long)(row0 + i) * a.ldo + h * DH;
#pragma unroll
      for (int d8 = 0; d8 < DH; d8 += 8) {
        uint4 u;
        __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) hh[e] = __floats2bfloat162_rn(o[d8 + 2 * e] * inv, o[d8 + 2 * e + 1] * inv);
        *reinterpret_cast<uint4*>(dst + d8) = u;
      }
      if (a.lse) a.lse[((long long)b * a.H + h) * S + i] = m * a.scale + __logf(l);
    }
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

bool svla_attn_flash_supported(int mode, int dtype, int S, int dh, long long ld, long long ldo, const void* q,
                               const void* k, const void* v, const void* o) {
  return dtype == SVLA_BF16 && dh == DH && S > 256 && mode == SVLA_ATTN_FULL && ld % 8 == 0 && ldo % 8 == 0 && al16(q) &&
         al16(k) && al16(v) && al16(o);
}

int svla_attn_flash_fwd(svla_ctx* ctx, const void* q, const void* k, const void* v, long long ld, void* o, long long ldo,
                        float* lse, int B, int S, int H, float scale, cudaStream_t st) {
  CUtensorMap mq, mk, mv;
  const long long rows = (long long)B * S;
  int rc;
  if ((rc = svla_make_tmap_bf16(ctx, q, (long long)H * DH, rows, ld, DH, TS, &mq))) return rc;
  if ((rc = svla_make_tmap_bf16(ctx, k, (long long)H * DH, rows, ld, DH, TS, &mk))) return rc;
  if ((rc = svla_make_tmap_bf16(ctx, v, (long long)H * DH, rows, ld, DH, TS, &mv))) return rc;
  AttnTcArgs a{};
  a.mode = SVLA_ATTN_FULL; a.B = B; a.S = S; a.H = H; a.scale = scale; a.lse = lse;
  a.o = reinterpret_cast<__nv_bfloat16*>(o); a.ldo = ldo;
  constexpr size_t smem = 81920 + 64 + 1024;
  static bool attr = false;
  if (!attr) {
    SVLA_CUDA(cudaFuncSetAttribute(attn_flash_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const long long items = (long long)B * H * ((S + TS - 1) / TS);
  const int grid = (int)std::min<long long>(items, 2LL * ctx->sm_count);
  attn_flash_fwd_kernel<<<grid, 128, smem, st>>>(mq, mk, mv, a);
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}
