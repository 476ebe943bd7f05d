// HL-Gauss discrete-critic loss, forward + backward in one pass (reference utils/loss_functions.py:7-30 with the
// DiscreteCriticHead read-out of allenact_dino_transformer.py:743-766):
//   p_b   = (erf((s_{b+1} - y) / (sqrt(2) sigma)) - erf((s_b - y) / (sqrt(2) sigma))) / z      soft target, B bins
//   loss  = mean_r ( - sum_b p_b log_softmax(logits_r)_b )                                      F.cross_entropy
//   dlogits_r = (softmax(logits_r) * sum_b p_b - p_r) * grad_scale / R
//   value_r   = sum_b softmax(logits_r)_b (s_b + s_{b+1}) / 2                                   transform_from_probs
// One warp per row (lanes stride over the bins); the mean is a deterministic two-stage fold (fixed block partials,
// the last block to finish adds them in a fixed order).  `support` is the module's own torch.linspace buffer, so
// the bin edges are bit-identical to the reference's.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int kMaxPerLane = 32;  // bins <= 1024

__global__ void __launch_bounds__(256) hl_gauss_kernel(const float* __restrict__ logits, long long ldl,
                                                       const float* __restrict__ target, const float* __restrict__ support,
                                                       int B, float sigma, float inv_count, float grad_scale,
                                                       float* __restrict__ out_loss, float* __restrict__ dlogits,
                                                       float* __restrict__ values, long long R, float* partials,
                                                       unsigned int* ticket) {
  __shared__ float red[32];
  __shared__ unsigned int s_ticket;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float inv_s = 1.f / (sqrtf(2.f) * sigma);
  float acc = 0.f;
  for (long long r = (long long)blockIdx.x * nw + wib; r < R; r += (long long)gridDim.x * nw) {
    const float* lg = logits + r * ldl;
    const float y = target[r];
    const float c_lo = erff((support[0] - y) * inv_s), c_hi = erff((support[B] - y) * inv_s);
    const float inv_z = 1.f / (c_hi - c_lo);
    float l[kMaxPerLane], p[kMaxPerLane];
    float mx = -INFINITY, psum = 0.f;
#pragma unroll 4
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int b = i * 32 + lane;
      if (i * 32 >= B) break;
      if (b < B) {
        l[i] = lg[b];
        p[i] = (erff((support[b + 1] - y) * inv_s) - erff((support[b] - y) * inv_s)) * inv_z;
        mx = fmaxf(mx, l[i]);
        psum += p[i];
      }
    }
    mx = warp_max(mx);
    psum = warp_sum(psum);
    float se = 0.f;
#pragma unroll 4
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int b = i * 32 + lane;
      if (i * 32 >= B) break;
      if (b < B) se += expf(l[i] - mx);
    }
    se = warp_sum(se);
    const float lse = mx + logf(se), inv_se = 1.f / se;
    float ce = 0.f, val = 0.f;
#pragma unroll 4
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int b = i * 32 + lane;
      if (i * 32 >= B) break;
      if (b < B) {
        const float sm = expf(l[i] - mx) * inv_se;
        ce -= p[i] * (l[i] - lse);
        val += sm * 0.5f * (support[b] + support[b + 1]);
        if (dlogits) dlogits[r * ldl + b] = (sm * psum - p[i]) * inv_count * grad_scale;
      }
    }
    ce = warp_sum(ce);
    val = warp_sum(val);
    if (lane == 0) {
      acc += ce;
      if (values) values[r] = val;
    }
  }
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
  __threadfence();
  if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1u);
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  __threadfence();
  if (threadIdx.x < 32) {
    float t = 0.f;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += 32) t += partials[b];
    t = warp_sum(t);
    if (threadIdx.x == 0) {
      out_loss[0] = t * inv_count;
      *ticket = 0u;
    }
  }
}

}  // namespace

extern "C" int svla_hl_gauss_fwd_bwd(svla_ctx* ctx, const float* logits, long long ldl, const float* target,
                                     const float* support, int num_bins, float sigma, float grad_scale, float* out_loss,
                                     float* dlogits, float* values, long long R, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && logits && target && support && out_loss, "NULL argument");
  SVLA_CHECK_ARG(num_bins >= 1 && num_bins <= 32 * kMaxPerLane, "num_bins must be in [1, 1024]");
  SVLA_CHECK_ARG(R > 0 && ldl >= num_bins && sigma > 0.f, "bad shape / sigma");
  const int grid = (int)std::min<long long>((R + 7) / 8, std::min<long long>((long long)ctx->sm_count * 8, kMaxPartialBlocks));
  hl_gauss_kernel<<<grid, 256, 0, as_stream(stream)>>>(logits, ldl, target, support, num_bins, sigma, 1.f / (float)R,
                                                       grad_scale, out_loss, dlogits, values, R, ctx->partials,
                                                       ctx->tickets + 1);
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}
