// Device-side building blocks shared by the tcgen05 GEMM kernels (gemm_tc.cu: cta_group::1, gemm_tc2.cu:
// cta_group::2): PTX wrappers (mbarrier, TMA, tcgen05.mma / ld / commit), UMMA descriptors, and the epilogue
// (TMEM -> registers -> swizzled staging tile -> TMA store / coalesced stores).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int BM = 128, BK = 64, kThreads = 384;  // 4 control warps + 8 epilogue warps
constexpr int kEpiWarp0 = 4;


// ---------------------------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// UMMA shared-memory descriptor, 128B swizzle (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}

struct TcArgs {
  int M, N, K;
  int tiles_m, tiles_n, splits, kb_per_split;  // kb = K blocks of BK
  void* C; long long ldc; int dtypeC;
  const float* bias;
  const void* residual; long long ldr; int dtypeR;
  const void* aux; long long ldaux; int dtypeAux;
  int epilogue, accumulate;
  float alpha;
  float* ws;
  float* asum;     // bias gradient fused into a weight-gradient launch (gemm_tc2.cu), or NULL
  float* asum_ws;  // its split-K partials
  int tma_store;
  int dbg;  // SVLA_TC_DBG experiments: 1 = skip the epilogue entirely, 2 = TMEM loads only (no global stores)
};

__device__ __forceinline__ float ld_elem(const void* p, int dt, long long i) {
  return dt == SVLA_F32 ? __ldg(reinterpret_cast<const float*>(p) + i)
                        : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}

// loads 8 consecutive elements (16-byte aligned for bf16, 32-byte for f32)
__device__ __forceinline__ void ld8(const void* p, int dt, long long i, float* o) {
  if (dt == SVLA_F32) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i));
    const float4 b = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i + 4));
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  } else {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p) + i));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(h[j]);
      o[2 * j] = f.x; o[2 * j + 1] = f.y;
    }
  }
}
__device__ __forceinline__ void st8(void* p, int dt, long long i, const float* v) {
  if (dt == SVLA_F32) {
    float4* d = reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i);
    d[0] = make_float4(v[0], v[1], v[2], v[3]);
    d[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p) + i) = u;
  }
}

// ---- staged epilogue ------------------------------------------------------------------------------------
// Each epilogue warp owns a [32 rows x 32 columns] staging tile per chunk (128 B rows for fp32, 64 B rows for
// bf16), XOR-swizzled in 16-byte slots so both the per-lane row writes and the coalesced row reads are
// conflict-free up to the 4-wavefront minimum.  Every global access of the epilogue (output, residual, ReLU-mask
// operand, accumulate) is then a run of full 32-byte sectors along a row.
constexpr int kStgBytes = 32 * 128;  // per epilogue warp

template <bool F32> __device__ __forceinline__ int stg_off(int row, int slot) {
  return F32 ? row * 128 + ((slot ^ (row & 7)) << 4) : row * 64 + ((slot ^ ((row >> 1) & 3)) << 4);
}
template <bool F32>
__device__ __forceinline__ void stage_in(uint8_t* stg, const void* base, long long ld_bytes, long long col_bytes, int m0,
                                         int M, int lane) {
  constexpr int LPR = F32 ? 8 : 4, RPP = 32 / LPR;  // lanes per row, rows per pass
#pragma unroll
  for (int p = 0; p < 32 / RPP; ++p) {
    const int row = p * RPP + lane / LPR, slot = lane % LPR;
    if (m0 + row < M)
      *reinterpret_cast<uint4*>(stg + stg_off<F32>(row, slot)) = __ldg(reinterpret_cast<const uint4*>(
          reinterpret_cast<const uint8_t*>(base) + (long long)(m0 + row) * ld_bytes + col_bytes + slot * 16));
  }
  __syncwarp();
}
template <bool F32>
__device__ __forceinline__ void stage_out(const uint8_t* stg, void* base, long long ld_bytes, long long col_bytes, int m0,
                                          int M, int lane) {
  constexpr int LPR = F32 ? 8 : 4, RPP = 32 / LPR;
  __syncwarp();
#pragma unroll
  for (int p = 0; p < 32 / RPP; ++p) {
    const int row = p * RPP + lane / LPR, slot = lane % LPR;
    if (m0 + row < M)
      *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(base) + (long long)(m0 + row) * ld_bytes + col_bytes +
                                slot * 16) = *reinterpret_cast<const uint4*>(stg + stg_off<F32>(row, slot));
  }
  __syncwarp();
}
// 16-byte slot j of this lane's staged row -> floats (4 for fp32, 8 for bf16)
template <bool F32>
__device__ __forceinline__ void piece_load(const uint8_t* stg, int lane, int j, float* o) {
  const uint4 u = *reinterpret_cast<const uint4*>(stg + stg_off<F32>(lane, j));
  if (F32) {
    o[0] = __uint_as_float(u.x); o[1] = __uint_as_float(u.y); o[2] = __uint_as_float(u.z); o[3] = __uint_as_float(u.w);
  } else {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __bfloat1622float2(h[e]);
      o[2 * e] = f.x; o[2 * e + 1] = f.y;
    }
  }
}

// columns [c_begin, c_end) of the tile for the 32 rows starting at m0
template <bool F32>
__device__ __forceinline__ void epilogue_staged_t(const TcArgs& g, const CUtensorMap* mapC, uint8_t* stg0, uint32_t taddr,
                                                  int m0, int ntile0, int c_begin, int c_end, int sp, int lane) {
  constexpr int EP = F32 ? 4 : 8;    // elements per 16-byte slot
  constexpr int NS = F32 ? 8 : 4;    // slots per 32-column row
  constexpr int ES = F32 ? 4 : 2;
  const bool part = g.splits > 1;
  uint8_t* Cb = part ? reinterpret_cast<uint8_t*>(g.ws + (size_t)sp * g.M * g.N) : reinterpret_cast<uint8_t*>(g.C);
  const long long ldc_b = (part ? (long long)g.N : g.ldc) * ES;
  // bf16 tiles are 2 KB: two staging buffers per warp, so a TMA store can still be reading one while the next
  // chunk fills the other; fp32 tiles (4 KB) use the single buffer
  constexpr int kBufs = F32 ? 1 : 2;
  int chunk = 0;
#pragma unroll 1
  for (int c0 = c_begin; c0 < c_end; c0 += 32, ++chunk) {
    const int n0 = ntile0 + c0;
    if (n0 >= g.N) break;  // warp-uniform
    uint8_t* stg = stg0 + (kBufs == 2 ? (chunk & 1) * 2048 : 0);
    if (g.tma_store) {  // the bulk store that last read this buffer must have finished reading it
      if (lane == 0) {
        if (kBufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      __syncwarp();
    }
    float4 bias4[8];
    if (!part && g.bias) {
#pragma unroll
      for (int e = 0; e < 8; ++e) bias4[e] = __ldg(reinterpret_cast<const float4*>(g.bias + n0) + e);
    }
    uint32_t r[32];
    tmem_ld32(taddr + c0, r);
    tmem_wait_ld();
    if (!part) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float x0 = __uint_as_float(r[4 * e]) * g.alpha, x1 = __uint_as_float(r[4 * e + 1]) * g.alpha;
        float x2 = __uint_as_float(r[4 * e + 2]) * g.alpha, x3 = __uint_as_float(r[4 * e + 3]) * g.alpha;
        if (g.bias) { x0 += bias4[e].x; x1 += bias4[e].y; x2 += bias4[e].z; x3 += bias4[e].w; }
        if (g.epilogue == SVLA_EPI_RELU) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f); }
        else if (g.epilogue == SVLA_EPI_GELU) { x0 = gelu_erf(x0); x1 = gelu_erf(x1); x2 = gelu_erf(x2); x3 = gelu_erf(x3); }
        r[4 * e] = __float_as_uint(x0); r[4 * e + 1] = __float_as_uint(x1);
        r[4 * e + 2] = __float_as_uint(x2); r[4 * e + 3] = __float_as_uint(x3);
      }
      if (g.epilogue == SVLA_EPI_RELU_MASK) {
        stage_in<F32>(stg, g.aux, g.ldaux * ES, (long long)n0 * ES, m0, g.M, lane);
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          float a[EP];
          piece_load<F32>(stg, lane, j, a);
#pragma unroll
          for (int e = 0; e < EP; ++e)
            if (!(a[e] > 0.f)) r[j * EP + e] = 0u;
        }
        __syncwarp();
      }
      if (g.residual) {
        stage_in<F32>(stg, g.residual, g.ldr * ES, (long long)n0 * ES, m0, g.M, lane);
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          float a[EP];
          piece_load<F32>(stg, lane, j, a);
#pragma unroll
          for (int e = 0; e < EP; ++e) r[j * EP + e] = __float_as_uint(__uint_as_float(r[j * EP + e]) + a[e]);
        }
        __syncwarp();
      }
      if (g.accumulate) {
        stage_in<F32>(stg, g.C, ldc_b, (long long)n0 * ES, m0, g.M, lane);
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          float a[EP];
          piece_load<F32>(stg, lane, j, a);
#pragma unroll
          for (int e = 0; e < EP; ++e) r[j * EP + e] = __float_as_uint(__uint_as_float(r[j * EP + e]) + a[e]);
        }
        __syncwarp();
      }
    }
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      uint4 u;
      if (F32) {
        u = make_uint4(r[j * 4], r[j * 4 + 1], r[j * 4 + 2], r[j * 4 + 3]);
      } else {
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          h[e] = __floats2bfloat162_rn(__uint_as_float(r[j * 8 + 2 * e]), __uint_as_float(r[j * 8 + 2 * e + 1]));
      }
      *reinterpret_cast<uint4*>(stg + stg_off<F32>(lane, j)) = u;
    }
    if (g.tma_store) {
      if (g.dbg != 5) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0 && g.dbg != 4 && g.dbg != 5) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(mapC),
                     "r"(smem_u32(stg)), "r"(n0), "r"(m0)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    } else {
      stage_out<F32>(stg, Cb, ldc_b, (long long)n0 * ES, m0, g.M, lane);
    }
  }
  // no wait here: the staging buffers are private to this warp and re-checked at the top of the next chunk, so the
  // caller may release the TMEM accumulator while the last stores are still in flight
}

// generic per-thread epilogue (mixed residual / aux dtypes): thread = row, 16-byte accesses
__device__ __forceinline__ void epilogue_direct(const TcArgs& g, uint32_t taddr, int m, bool row_ok, int ntile0,
                                                int c_begin, int c_end, int sp) {
#pragma unroll 1
  for (int c = c_begin / 32; c < c_end / 32; ++c) {
    const int n0 = ntile0 + c * 32;
    if (n0 >= g.N) break;  // warp-uniform
    uint32_t r[32];
    tmem_ld32(taddr + c * 32, r);
    tmem_wait_ld();
    if (!row_ok) continue;
    if (g.splits > 1) {
      float* dst = g.ws + ((size_t)sp * g.M + m) * g.N + n0;
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                          __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
      continue;
    }
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[j + e]) * g.alpha;
      if (g.bias) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(g.bias + n0 + j));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(g.bias + n0 + j + 4));
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      }
      if (g.epilogue == SVLA_EPI_RELU) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
      } else if (g.epilogue == SVLA_EPI_GELU) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = gelu_erf(v[e]);
      } else if (g.epilogue == SVLA_EPI_RELU_MASK) {
        float a[8];
        ld8(g.aux, g.dtypeAux, (long long)m * g.ldaux + n0 + j, a);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = a[e] > 0.f ? v[e] : 0.f;
      }
      if (g.residual) {
        float a[8];
        ld8(g.residual, g.dtypeR, (long long)m * g.ldr + n0 + j, a);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] += a[e];
      }
      const long long ci = (long long)m * g.ldc + n0 + j;
      if (g.accumulate) {
        float a[8];
        ld8(g.C, g.dtypeC, ci, a);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] += a[e];
      }
      st8(g.C, g.dtypeC, ci, v);
    }
  }
}

}  // namespace
