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
  DropArgs drop;      // RELU_BITS: dropout after the ReLU (thr == 0: off)
  int dbg;  // SVLA_TC_DBG experiments: 1 = skip the epilogue entirely, 2 = TMEM loads only (no global stores),
            // 9 = 32-column chunks instead of the 64-column block epilogue (A/B switch used for the same-box comparison)
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

__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 u;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(saddr) : "memory");
  return u;
}
__device__ __forceinline__ void sts128(uint32_t saddr, const uint4& u) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}
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
      sts128(smem_u32(stg) + stg_off<F32>(row, slot), __ldg(reinterpret_cast<const uint4*>(
          reinterpret_cast<const uint8_t*>(base) + (long long)(m0 + row) * ld_bytes + col_bytes + slot * 16)));
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
                                slot * 16) = lds128(smem_u32(stg) + stg_off<F32>(row, slot));
  }
  __syncwarp();
}
// 16-byte slot j of this lane's staged row -> floats (4 for fp32, 8 for bf16)
template <bool F32>
__device__ __forceinline__ void piece_load(uint32_t stg_s, int lane, int j, float* o) {
  const uint4 u = lds128(stg_s + stg_off<F32>(lane, j));
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


// ---- side-operand prefetch ------------------------------------------------------------------------------
// The ReLU-mask operand / residual tile does not depend on the accumulator, so its global loads are issued a WHOLE
// TILE ahead of their use: while chunk c of this tile is processed, the registers it just released are refilled with
// chunk c of the tile this warp drains next (for the first tile: before the wait for the accumulator).  A K = 512
// tile takes ~0.6 us, i.e. about one DRAM round trip, whereas a chunk takes ~150 ns -- the one-chunk-ahead prefetch
// of round 1 still exposed most of the latency (FFN-down dgrad with the ReLU mask 0.69 PF/s, K = 512 residual
// launches 0.50 PF/s against 1.1-1.2 PF/s without a side operand).  The 64 registers per lane come from the control
// warps (setmaxnreg in the kernels).  Only the first side operand in application order (aux, residual) of bf16
// tiles is prefetched; anything else takes the synchronous stage_in path.
struct SidePre {
  uint4 v[4][4];  // [chunk][piece]: this lane's 16-byte pieces of the four 32-column chunks of the next tile
  int valid;      // bit c: chunk c is in registers
  float bias[4];  // bias of column (chunk c, lane) of the CURRENT tile (bias_prefetch)
  uint4 mb;       // MASK_BITS: the bit record of this lane's row over the warp's 128 columns (bits_prefetch)
};
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ int side_pick(const TcArgs& g, const void** base, long long* ld) {
  if (g.splits > 1 || g.dtypeC != SVLA_BF16) return 0;
  if (g.epilogue == SVLA_EPI_RELU_MASK && g.aux) { *base = g.aux; *ld = g.ldaux; return 1; }
  if (g.residual) { *base = g.residual; *ld = g.ldr; return 2; }
  return 0;
}
// bf16 chunk [32 rows x 64 B]: lane covers rows (lane / 4) + 8 p, 16-byte slot lane % 4
__device__ __forceinline__ void side_fetch(uint4* v, const void* base, long long ld_bytes, long long col_bytes, int m0,
                                           int M, int lane) {
  const uint8_t* p0 = reinterpret_cast<const uint8_t*>(base) + (long long)(m0 + (lane >> 2)) * ld_bytes + col_bytes +
                      (lane & 3) * 16;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    if (m0 + (lane >> 2) + 8 * p < M) v[p] = __ldg(reinterpret_cast<const uint4*>(p0 + (long long)(8 * p) * ld_bytes));
    else v[p] = make_uint4(0u, 0u, 0u, 0u);
  }
}
__device__ __forceinline__ void side_commit(uint32_t stg_s, const uint4* v, int lane) {
#pragma unroll
  for (int p = 0; p < 4; ++p) sts128(stg_s + stg_off<false>(p * 8 + (lane >> 2), lane & 3), v[p]);
  __syncwarp();
}
// The bias of the warp's columns, requested BEFORE the wait for the accumulator: with ~228 KB of shared memory carved
// out, L1 keeps almost nothing, so a per-chunk load next to its use is an L2 round trip on the critical path of every
// chunk (ncu source view of the K = 512 out-projection: 18 % of all stall samples sat on the bias broadcast shuffle).
__device__ __forceinline__ void bias_prefetch(SidePre& pre, const TcArgs& g, int n0, int lane, int nch) {
#pragma unroll
  for (int c = 0; c < 4; ++c)
    pre.bias[c] = (g.bias != nullptr && g.splits == 1 && c < nch && n0 + c * 32 < g.N) ? __ldg(g.bias + n0 + c * 32 + lane) : 0.f;
}
// Bit records (RELU_BITS / MASK_BITS) of a warp that drains 128 columns are moved as ONE 16-byte access per row and tile
// -- the record of a 256-column tile is 32 bytes, one sector -- instead of one 8-byte access per 64-column block.
__device__ __forceinline__ bool bits_wide(const TcArgs& g, int n_begin, int ncols) {
  return ncols == 128 && n_begin + 128 <= g.N && (g.ldaux & 3) == 0 && (reinterpret_cast<uintptr_t>(g.aux) & 15) == 0;
}
// the mask record of the backward launch does not depend on the accumulator either: requested before the wait
__device__ __forceinline__ void bits_prefetch(SidePre& pre, const TcArgs& g, int m0, int n_begin, int lane, int ncols) {
  if (g.epilogue != SVLA_EPI_MASK_BITS || g.splits > 1 || !bits_wide(g, n_begin, ncols)) return;
  pre.mb = make_uint4(0u, 0u, 0u, 0u);
  if (m0 + lane < g.M)
    pre.mb = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(g.aux) + (long long)(m0 + lane) * g.ldaux +
                                                  (n_begin >> 5)));
}
// every chunk of a warp's first tile, issued by the caller BEFORE it waits for the accumulator
__device__ __forceinline__ void side_prefetch_first(SidePre& pre, const TcArgs& g, int m0, int n0, int lane, int nch) {
  const void* base = nullptr;
  long long ld = 0;
  if (pre.valid || !side_pick(g, &base, &ld)) return;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    if (c < nch && n0 + c * 32 < g.N) {
      side_fetch(pre.v[c], base, ld * 2, (long long)(n0 + c * 32) * 2, m0, g.M, lane);
      pre.valid |= 1 << c;
    }
}

// columns [c_begin, c_end) of the tile for the 32 rows starting at m0
// (next_m0, next_ntile0): the tile this warp drains next (next_m0 < 0: none) -- its first side-operand chunk is
// requested while the last chunk of this tile is processed.
//
// The per-chunk instruction stream of the epilogue warps (LDTM -> math -> pack -> staging -> TMA store, four chunks
// per warp and tile) is what paces the K = 512 GEMMs, so the flag tests are folded at compile time: EPI >= 0
// instantiates exactly one combination (epilogue kind, bias, residual, accumulate); EPI = -1 is the run-time
// generic version for everything else.  For bf16 outputs ReLU and the ReLU mask act on the packed bf16 pairs
// (one HMNMX2 / HSET2 + LOP per two elements; identical results: rounding to bf16 preserves sign and zero).
template <bool F32, int EPI, bool BIAS, bool RES, bool ACC>
__device__ __forceinline__ void epilogue_staged_impl(const TcArgs& g, const CUtensorMap* mapC, uint8_t* stg0, uint32_t taddr,
                                                     int m0, int ntile0, int c_begin, int c_end, int sp, int lane,
                                                     SidePre& pre, int next_m0, int next_ntile0) {
  constexpr bool GEN = EPI < 0;
  constexpr int EP = F32 ? 4 : 8;    // elements per 16-byte slot
  constexpr int NS = F32 ? 8 : 4;    // slots per 32-column row
  constexpr int ES = F32 ? 4 : 2;
  const bool part = g.splits > 1;
  const int epi = GEN ? g.epilogue : EPI;
  const bool has_bias = GEN ? (g.bias != nullptr) : BIAS;
  const bool has_res = GEN ? (g.residual != nullptr) : RES;
  const bool has_acc = GEN ? (g.accumulate != 0) : ACC;
  uint8_t* Cb = part ? reinterpret_cast<uint8_t*>(g.ws + (size_t)sp * g.M * g.N) : reinterpret_cast<uint8_t*>(g.C);
  const long long ldc_b = (part ? (long long)g.N : g.ldc) * ES;
  // bf16 tiles are 2 KB: two staging buffers per warp, so a TMA store can still be reading one while the next
  // chunk fills the other; fp32 tiles (4 KB) use the single buffer
  constexpr int kBufs = F32 ? 1 : 2;
  int chunk = 0;
  const void* side_base = nullptr;
  long long side_ld = 0;
  int side = 0;
  if (!F32) {
    if (GEN) side = side_pick(g, &side_base, &side_ld);
    else if (EPI == SVLA_EPI_RELU_MASK) { side = 1; side_base = g.aux; side_ld = g.ldaux; }
    else if (RES) { side = 2; side_base = g.residual; side_ld = g.ldr; }
  }
#pragma unroll
  for (chunk = 0; chunk < 4; ++chunk) {  // unrolled: the prefetch registers are indexed by the chunk
    const int c0 = c_begin + chunk * 32;
    if (c0 >= c_end) break;  // warp-uniform (two chunks per warp for 128-wide tiles)
    const int n0 = ntile0 + c0;
    if (n0 >= g.N) break;    // warp-uniform
    uint8_t* stg = stg0 + (kBufs == 2 ? (chunk & 1) * 2048 : 0);
    const uint32_t stg_s = smem_u32(stg);
    if (g.tma_store) {  // the bulk store that last read this buffer must have finished reading it
      if (lane == 0) {
        if (kBufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      __syncwarp();
    }
    uint32_t r[32];
    tmem_ld32(taddr + c0, r);
    if (side) {  // this chunk's side operand: registers -> staging; then request the same chunk of the NEXT tile
      if (!((pre.valid >> chunk) & 1))
        side_fetch(pre.v[chunk], side_base, side_ld * ES, (long long)n0 * ES, m0, g.M, lane);  // cold start
      side_commit(stg_s, pre.v[chunk], lane);
      pre.valid &= ~(1 << chunk);
      if (next_m0 >= 0 && next_ntile0 + c0 < g.N) {
        side_fetch(pre.v[chunk], side_base, side_ld * ES, (long long)(next_ntile0 + c0) * ES, next_m0, g.M, lane);
        pre.valid |= 1 << chunk;
      }
    }
    // bias of the 32 columns: one coalesced load (lane l holds column l), broadcast by shuffles after the wait --
    // holding all 32 values per lane costs 31 more registers than the kernel has.  Loaded by bias_prefetch.
    const float bias_l = pre.bias[chunk];
    tmem_wait_ld();
    // which ops can run on the packed bf16 pairs after the conversion
    const bool relu_packed = !F32 && !part && epi == SVLA_EPI_RELU && !has_res && !has_acc;
    const bool mask_packed = !F32 && !part && epi == SVLA_EPI_RELU_MASK && !has_res && !has_acc;
    if (!part) {
      if (g.alpha != 1.f) {
#pragma unroll
        for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * g.alpha);
      }
      if (has_bias) {
#pragma unroll
        for (int e = 0; e < 32; ++e)
          r[e] = __float_as_uint(__uint_as_float(r[e]) + __shfl_sync(0xffffffffu, bias_l, e));
      }
      if (epi == SVLA_EPI_RELU && !relu_packed) {
#pragma unroll
        for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(fmaxf(__uint_as_float(r[e]), 0.f));
      } else if (epi == SVLA_EPI_GELU) {
#pragma unroll
        for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(gelu_erf(__uint_as_float(r[e])));
      } else if (epi == SVLA_EPI_RELU_MASK && !mask_packed) {
        if (side != 1) stage_in<F32>(stg, g.aux, g.ldaux * ES, (long long)n0 * ES, m0, g.M, lane);
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          float a[EP];
          piece_load<F32>(stg_s, lane, j, a);
#pragma unroll
          for (int e = 0; e < EP; ++e)
            if (!(a[e] > 0.f)) r[j * EP + e] = 0u;
        }
        __syncwarp();
      }
      if (has_res) {
        if (side != 2) stage_in<F32>(stg, g.residual, g.ldr * ES, (long long)n0 * ES, m0, g.M, lane);
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          float a[EP];
          piece_load<F32>(stg_s, lane, j, a);
#pragma unroll
          for (int e = 0; e < EP; ++e) r[j * EP + e] = __float_as_uint(__uint_as_float(r[j * EP + e]) + a[e]);
        }
        __syncwarp();
      }
      if (has_acc) {
        stage_in<F32>(stg, g.C, ldc_b, (long long)n0 * ES, m0, g.M, lane);
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          float a[EP];
          piece_load<F32>(stg_s, lane, j, a);
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
        if (relu_packed) {
          const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
          for (int e = 0; e < 4; ++e) h[e] = __hmax2(h[e], z);
        }
        if (mask_packed) {  // the side operand of this slot sits where the output goes: read it, then overwrite
          const uint4 a = lds128(stg_s + stg_off<false>(lane, j));
          const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
          u.x &= __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&a.x), z);
          u.y &= __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&a.y), z);
          u.z &= __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&a.z), z);
          u.w &= __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&a.w), z);
        }
      }
      sts128(stg_s + stg_off<F32>(lane, j), u);
    }
    if (g.tma_store) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(mapC),
                     "r"(stg_s), "r"(n0), "r"(m0)
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

// ---- 64-column blocks (bf16 outputs without a side operand: plain / bias / ReLU / GELU) -----------------------------
// The per-chunk fixed costs (store-buffer wait, warp syncs, the proxy fence with its MEMBAR, the TMA issue) are paid
// once per 64 columns instead of once per 32: two LDTMs are in flight together, the 32 x 128 B staging tile (128B
// swizzle, the warp's whole 4 KB) leaves through ONE TMA store of full 128-byte row segments.
template <int EPI, bool BIAS, bool DROP = false>
__device__ __forceinline__ void epilogue_block64(const TcArgs& g, const CUtensorMap* mapC2, uint8_t* stg0, uint32_t taddr,
                                                 int m0, int ntile0, int c_begin, int c_end, int lane, const SidePre& pre) {
  const uint32_t stg_s = smem_u32(stg0);
  constexpr bool kBits = EPI == SVLA_EPI_RELU_BITS || EPI == SVLA_EPI_MASK_BITS;
  const bool wide = kBits && bits_wide(g, ntile0 + c_begin, c_end - c_begin);
  uint4 rec = (EPI == SVLA_EPI_MASK_BITS) ? pre.mb : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
  for (int cb = c_begin; cb < c_end; cb += 64) {
    const int n0 = ntile0 + cb;
    if (n0 >= g.N) break;  // warp-uniform (N is a multiple of 64)
    uint32_t r0[32], r1[32];
    tmem_ld32(taddr + cb, r0);
    tmem_ld32(taddr + cb + 32, r1);
    const bool second = cb != c_begin;  // a warp drains one or two 64-column blocks per tile
    const float b0 = BIAS ? (second ? pre.bias[2] : pre.bias[0]) : 0.f;
    const float b1 = BIAS ? (second ? pre.bias[3] : pre.bias[1]) : 0.f;
    // bit record of this lane's row: two words (columns n0 .. n0 + 31, n0 + 32 .. n0 + 63)
    uint32_t* bits_p = reinterpret_cast<uint32_t*>(const_cast<void*>(g.aux)) + (long long)(m0 + lane) * g.ldaux + (n0 >> 5);
    uint2 mb = make_uint2(0u, 0u);
    uint32_t mbp0[4] = {0u, 0u, 0u, 0u}, mbp1[4] = {0u, 0u, 0u, 0u};
    if (EPI == SVLA_EPI_MASK_BITS) {
      if (wide) mb = second ? make_uint2(rec.z, rec.w) : make_uint2(rec.x, rec.y);
      else if (m0 + lane < g.M) mb = __ldg(reinterpret_cast<const uint2*>(bits_p));
    }
    tmem_wait_ld();
    if (g.alpha != 1.f) {
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        r0[e] = __float_as_uint(__uint_as_float(r0[e]) * g.alpha);
        r1[e] = __float_as_uint(__uint_as_float(r1[e]) * g.alpha);
      }
    }
    if (BIAS) {
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        r0[e] = __float_as_uint(__uint_as_float(r0[e]) + __shfl_sync(0xffffffffu, b0, e));
        r1[e] = __float_as_uint(__uint_as_float(r1[e]) + __shfl_sync(0xffffffffu, b1, e));
      }
    }
    if (EPI == SVLA_EPI_GELU) {
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        r0[e] = __float_as_uint(gelu_erf(__uint_as_float(r0[e])));
        r1[e] = __float_as_uint(gelu_erf(__uint_as_float(r1[e])));
      }
    }
    // the store that last read the staging tile must be done reading it
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t* r = j < 4 ? r0 : r1;
      const int jj = j & 3;
      uint4 u;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
      if (DROP) {
        const float dsc = g.drop.scale;  // 1 / (1 - p) of the fused dropout, applied in fp32
#pragma unroll
        for (int e = 0; e < 4; ++e)
          h[e] = __floats2bfloat162_rn(__uint_as_float(r[jj * 8 + 2 * e]) * dsc, __uint_as_float(r[jj * 8 + 2 * e + 1]) * dsc);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          h[e] = __floats2bfloat162_rn(__uint_as_float(r[jj * 8 + 2 * e]), __uint_as_float(r[jj * 8 + 2 * e + 1]));
      }
      if (EPI == SVLA_EPI_RELU || EPI == SVLA_EPI_RELU_BITS) {
        const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
        for (int e = 0; e < 4; ++e) h[e] = __hmax2(h[e], z);
      }
      if (EPI == SVLA_EPI_RELU_BITS && DROP) {
        // FFN dropout of the encoder layer: slot j holds columns n0 + 8 j .. + 8 of this lane's row = one Philox group
        const uint32_t keep = dropout_keep8(g.drop, g.drop.row0 + (uint32_t)(m0 + lane) * g.drop.row_stride,
                                            (uint32_t)((n0 >> 3) + j));
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          uint32_t& w = *reinterpret_cast<uint32_t*>(&h[e]);
          w &= (((keep >> (2 * e)) & 1u) * 0x0000FFFFu) | (((keep >> (2 * e + 1)) & 1u) * 0xFFFF0000u);
        }
      }
      if (EPI == SVLA_EPI_RELU_BITS) {
        // non-negative bf16 pairs: (w + 0x7FFF7FFF) has bit 15 / 31 set iff the low / high element is > 0; word wi of
        // the chunk (elements 2 wi, 2 wi + 1) lands at bits wi and 16 + wi, i.e. element e at (e >> 1) + 16 (e & 1).
        // Four independent partial records per chunk keep the dependency chains short (the epilogue paces K = 512).
        const uint32_t* w = reinterpret_cast<const uint32_t*>(&u);
        if (g.dbg == 6) {  // A/B switch: no packing at all (stores zeros)
        } else if (g.dbg == 7) {  // A/B switch: the serial shift-in chain
          uint32_t& acc = j < 4 ? mbp0[0] : mbp1[0];
#pragma unroll
          for (int e = 0; e < 4; ++e) acc = (acc >> 1) | ((w[e] + 0x7FFF7FFFu) & 0x80008000u);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int wi = jj * 4 + e;  // word index inside the 32-column chunk
            const uint32_t t = ((w[e] + 0x7FFF7FFFu) >> (15 - wi)) & (0x00010001u << wi);
            if (j < 4) mbp0[e] |= t; else mbp1[e] |= t;
          }
        }
      }
      if (EPI == SVLA_EPI_MASK_BITS) {
        const uint32_t word = j < 4 ? mb.x : mb.y;
        uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) w[e] &= ((word >> (jj * 4 + e)) & 0x00010001u) * 0xFFFFu;
      }
      sts128(stg_s + lane * 128 + ((j ^ (lane & 7)) << 4), u);
    }
    if (EPI == SVLA_EPI_RELU_BITS && g.dbg != 8) {
      const uint32_t w0 = (mbp0[0] | mbp0[1]) | (mbp0[2] | mbp0[3]), w1 = (mbp1[0] | mbp1[1]) | (mbp1[2] | mbp1[3]);
      if (!wide) {
        if (m0 + lane < g.M) *reinterpret_cast<uint2*>(bits_p) = make_uint2(w0, w1);
      } else if (!second) {
        rec.x = w0; rec.y = w1;
      } else {  // both blocks of the warp's 128 columns: one 16-byte store
        rec.z = w0; rec.w = w1;
        if (m0 + lane < g.M) *reinterpret_cast<uint4*>(bits_p - 2) = rec;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(mapC2), "r"(stg_s),
                   "r"(n0), "r"(m0)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
}

// dispatch on the (warp-uniform) flag combination; the listed ones are every combination the towers launch
template <bool F32>
__device__ __forceinline__ void epilogue_staged_t(const TcArgs& g, const CUtensorMap* mapC, uint8_t* stg0, uint32_t taddr,
                                                  int m0, int ntile0, int c_begin, int c_end, int sp, int lane,
                                                  SidePre& pre, int next_m0, int next_ntile0, const CUtensorMap* mapC2) {
#define SVLA_EPI_CALL(E, B, R, A) \
  epilogue_staged_impl<F32, E, B, R, A>(g, mapC, stg0, taddr, m0, ntile0, c_begin, c_end, sp, lane, pre, next_m0, next_ntile0)
  const bool b = g.bias != nullptr, rs = g.residual != nullptr, ac = g.accumulate != 0;
  const int e = g.epilogue;
  if (g.splits > 1) { SVLA_EPI_CALL(SVLA_EPI_NONE, false, false, false); return; }  // partial tiles: raw accumulators
  if (F32) {
    if (e == SVLA_EPI_NONE && !b && !rs && ac) SVLA_EPI_CALL(SVLA_EPI_NONE, false, false, true);
    else if (e == SVLA_EPI_NONE && !b && rs && !ac) SVLA_EPI_CALL(SVLA_EPI_NONE, false, true, false);
    else if (e == SVLA_EPI_NONE && !b && !rs && !ac) SVLA_EPI_CALL(SVLA_EPI_NONE, false, false, false);
    else SVLA_EPI_CALL(-1, false, false, false);
  } else {
    const bool blk = g.tma_store && !rs && !ac && e != SVLA_EPI_RELU_MASK && g.dbg != 9;  // 64-column blocks
    if (e == SVLA_EPI_RELU_BITS && g.drop.thr != 0u) {
      if (b) epilogue_block64<SVLA_EPI_RELU_BITS, true, true>(g, mapC2, stg0, taddr, m0, ntile0, c_begin, c_end, lane, pre);
      else epilogue_block64<SVLA_EPI_RELU_BITS, false, true>(g, mapC2, stg0, taddr, m0, ntile0, c_begin, c_end, lane, pre);
    }
    else if (e == SVLA_EPI_RELU_BITS && b) epilogue_block64<SVLA_EPI_RELU_BITS, true>(g, mapC2, stg0, taddr, m0, ntile0, c_begin, c_end, lane, pre);
    else if (e == SVLA_EPI_RELU_BITS) epilogue_block64<SVLA_EPI_RELU_BITS, false>(g, mapC2, stg0, taddr, m0, ntile0, c_begin, c_end, lane, pre);
    else if (e == SVLA_EPI_MASK_BITS) epilogue_block64<SVLA_EPI_MASK_BITS, false>(g, mapC2, stg0, taddr, m0, ntile0, c_begin, c_end, lane, pre);
    else if (ac) SVLA_EPI_CALL(-1, false, false, false);
    else if (blk && e == SVLA_EPI_NONE && b) epilogue_block64<SVLA_EPI_NONE, true>(g, mapC2, stg0, taddr, m0, ntile0, c_begin, c_end, lane, pre);
    else if (blk && e == SVLA_EPI_NONE && !b) epilogue_block64<SVLA_EPI_NONE, false>(g, mapC2, stg0, taddr, m0, ntile0, c_begin, c_end, lane, pre);
    else if (blk && e == SVLA_EPI_RELU && b) epilogue_block64<SVLA_EPI_RELU, true>(g, mapC2, stg0, taddr, m0, ntile0, c_begin, c_end, lane, pre);
    else if (blk && e == SVLA_EPI_GELU && b) epilogue_block64<SVLA_EPI_GELU, true>(g, mapC2, stg0, taddr, m0, ntile0, c_begin, c_end, lane, pre);
    else if (e == SVLA_EPI_NONE && b && !rs) SVLA_EPI_CALL(SVLA_EPI_NONE, true, false, false);
    else if (e == SVLA_EPI_RELU && b && !rs) SVLA_EPI_CALL(SVLA_EPI_RELU, true, false, false);
    else if (e == SVLA_EPI_NONE && b && rs) SVLA_EPI_CALL(SVLA_EPI_NONE, true, true, false);
    else if (e == SVLA_EPI_RELU_MASK && !b && !rs) SVLA_EPI_CALL(SVLA_EPI_RELU_MASK, false, false, false);
    else if (e == SVLA_EPI_NONE && !b && rs) SVLA_EPI_CALL(SVLA_EPI_NONE, false, true, false);
    else if (e == SVLA_EPI_NONE && !b && !rs) SVLA_EPI_CALL(SVLA_EPI_NONE, false, false, false);
    else if (e == SVLA_EPI_GELU && b && !rs) SVLA_EPI_CALL(SVLA_EPI_GELU, true, false, false);
    else SVLA_EPI_CALL(-1, false, false, false);
  }
#undef SVLA_EPI_CALL
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
