// Counter-based dropout masks (training-mode dropout of the fusion block, nn.TransformerEncoderLayer p = 0.1:
// allenact_dino_transformer.py:545-552).
//
// Philox4x32 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) with 7 rounds -- the
// smallest round count the paper reports as passing BigCrush -- keyed by the 64-bit seed; the 128-bit counter names the
// element group, so a mask is a pure function of (seed, step, site, row, column group): forward, backward, recompute
// and the mask dump used by the tests regenerate it independently, nothing is stored.
//
//   counter = (row, column / 8, site, step)        one call -> 4 x 32 bits -> eight 16-bit uniforms
//   element (row, column) is KEPT iff  u16[column % 8] >= round(p * 65536)          (kept values scale by 1 / (1 - p))
//
// `row` is the global row of the tensor the site acts on (GEMM row m; attention: item * 128 + query), `site` encodes
// (tower, layer, which of the four dropouts), `step` separates update repeats / rollouts.
#pragma once
#include <stdint.h>

struct DropArgs {
  uint32_t thr;        // round(p * 65536); 0 = dropout off
  float scale;         // 1 / (1 - p)
  uint32_t key0, key1; // seed
  uint32_t site, step;
  uint32_t row0, row_stride;  // global row of local row r: row0 + r * row_stride
};

__device__ __forceinline__ uint4 philox4x32_7(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return make_uint4(c0, c1, c2, c3);
}

// keep bits (bit e set = keep) of the eight elements of column group `cg` in row `row`
__device__ __forceinline__ uint32_t dropout_keep8(const DropArgs& d, uint32_t row, uint32_t cg) {
  const uint4 u = philox4x32_7(row, cg, d.site, d.step, d.key0, d.key1);
  uint32_t m = 0;
  m |= ((u.x & 0xFFFFu) >= d.thr) ? 1u : 0u;
  m |= ((u.x >> 16) >= d.thr) ? 2u : 0u;
  m |= ((u.y & 0xFFFFu) >= d.thr) ? 4u : 0u;
  m |= ((u.y >> 16) >= d.thr) ? 8u : 0u;
  m |= ((u.z & 0xFFFFu) >= d.thr) ? 16u : 0u;
  m |= ((u.z >> 16) >= d.thr) ? 32u : 0u;
  m |= ((u.w & 0xFFFFu) >= d.thr) ? 64u : 0u;
  m |= ((u.w >> 16) >= d.thr) ? 128u : 0u;
  return m;
}
