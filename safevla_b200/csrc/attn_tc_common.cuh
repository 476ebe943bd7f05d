// Helpers shared by the tcgen05 attention kernels (attn_tc.cu: S <= 128, attn_tc2.cu: 128 < S <= 256).
#pragma once
#include "tc_common.cuh"

namespace {

constexpr int DH = 64, TS = 128;  // tile: 128 queries x 128 keys
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// instruction descriptor: D=f32, A=B=bf16
__host__ __device__ constexpr uint32_t idesc(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// [128 rows x 64] bf16 operand tile as TMA lands it (row = 128 B, 128B swizzle):
//   K-major view  (rows = M/N, 64 = K): k-step kk (16 elements) -> +32 B
//   MN-major view (rows = K, 64 = M/N): k-step kk (16 rows)     -> +2048 B
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t base, int kk) { return umma_desc(base + kk * 32, 16, 1024); }
__device__ __forceinline__ uint64_t desc_mnmajor64(uint32_t base, int kk) { return umma_desc(base + kk * 2048, 8192, 1024); }
// [128 rows x 128] bf16 P / dS tile written by the threads as two 64-column chunks of [128 rows x 128 B]:
//   K-major view  (rows = M queries, 128 = K keys): k-step kk -> chunk kk/4, +32 B * (kk%4)
//   MN-major view (rows = K queries, 128 = M keys, two 64-chunks 16384 B apart): k-step kk -> +2048 B
__device__ __forceinline__ uint64_t desc_p_kmajor(uint32_t base, int kk) {
  return umma_desc(base + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024);
}
__device__ __forceinline__ uint64_t desc_p_mnmajor(uint32_t base, int kk) { return umma_desc(base + kk * 2048, 16384, 1024); }

// write 8 consecutive bf16 (columns c8*8 .. +8 of row i) of a P / dS tile
__device__ __forceinline__ void store_p8(uint8_t* base, int i, int c8, const float* v) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
  const int chunk = c8 >> 3, cc = c8 & 7;
  *reinterpret_cast<uint4*>(base + chunk * 16384 + i * 128 + ((cc ^ (i & 7)) << 4)) = u;
}
// the same through an explicit st.shared (the backward kernels: measured faster there, slower in the forward ones,
// where the volatile asm keeps the compiler from interleaving the stores with the exponentials)
__device__ __forceinline__ void store_p8_sts(uint8_t* base, int i, int c8, const float* v) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
  const int chunk = c8 >> 3, cc = c8 & 7;
  sts128(smem_u32(base) + chunk * 16384 + i * 128 + ((cc ^ (i & 7)) << 4), u);
}

struct AttnTcArgs {
  int mode, B, S, H;
  float scale;
  const int64_t* traj;
  float* lse;
  __nv_bfloat16* o; long long ldo;
  const __nv_bfloat16* o_in; const __nv_bfloat16* d_o;
  __nv_bfloat16* dq; __nv_bfloat16* dk; __nv_bfloat16* dv; long long ldd;
  DropArgs drop;  // attn_tc2 (128 < S <= 256, mode FULL): dropout on P; mask row = row0 + item * 256 + query, group = key / 8
};

__device__ __forceinline__ void store_row64(__nv_bfloat16* dst, const uint32_t* r0, const uint32_t* r1, float mul) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const uint32_t* r = half ? r1 : r0;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      uint4 u;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e)
        h[e] = __floats2bfloat162_rn(__uint_as_float(r[j + 2 * e]) * mul, __uint_as_float(r[j + 2 * e + 1]) * mul);
      *reinterpret_cast<uint4*>(dst + half * 32 + j) = u;
    }
  }
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

// 32 consecutive fp32 accumulator columns -> 32 bf16 (64 B) of one output row
__device__ __forceinline__ void store_row32(__nv_bfloat16* dst, const uint32_t* r, float mul) {
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e)
      h[e] = __floats2bfloat162_rn(__uint_as_float(r[j + 2 * e]) * mul, __uint_as_float(r[j + 2 * e + 1]) * mul);
    *reinterpret_cast<uint4*>(dst + j) = u;
  }
}

}  // namespace
