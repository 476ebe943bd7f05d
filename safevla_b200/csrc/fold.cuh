// Deterministic second stage of the two-stage column reductions (LayerNorm/RMSNorm parameter gradients, bias
// gradients): out[w] (+)= sum_b partial[b][w] for W columns and nb block partials.  Block = 32 columns x 32 row
// slices: every slice walks its rows with coalesced 128-byte reads, the slices are then folded through shared
// memory in a fixed order (no atomics; bit-stable run to run).
#pragma once
#include "common.cuh"

static __global__ void __launch_bounds__(1024) svla_fold_kernel(const float* __restrict__ partial, int nb, int W, int seg,
                                                         float* o0, float* o1, float* o2, int accumulate) {
  __shared__ float red[32][33];
  const int w = blockIdx.x * 32 + threadIdx.x, sl = threadIdx.y;
  float s = 0.f;
  if (w < W) {
    float s0 = 0.f, s1 = 0.f;
    int b = sl;
    for (; b + 32 < nb; b += 64) {
      s0 += partial[(size_t)b * W + w];
      s1 += partial[(size_t)(b + 32) * W + w];
    }
    if (b < nb) s0 += partial[(size_t)b * W + w];
    s = s0 + s1;
  }
  red[sl][threadIdx.x] = s;
  __syncthreads();
  if (sl == 0 && w < W) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += red[k][threadIdx.x];
    const int j = w / seg, d = w % seg;  // output vector j, element d
    float* o = j == 0 ? o0 : (j == 1 ? o1 : o2);
    if (o) o[d] = accumulate ? o[d] + t : t;
  }
}

static inline void svla_launch_fold(const float* partial, int nb, int W, int seg, float* o0, float* o1, float* o2,
                                    int accumulate, cudaStream_t st) {
  svla_fold_kernel<<<(W + 31) / 32, dim3(32, 32), 0, st>>>(partial, nb, W, seg, o0, o1, o2, accumulate);
}
