// Fused PPO-Lagrangian loss forward + backward (reference
// training/online/loss/customized_loss.py:317-449).  One pass over the rollout:
//   read  A logits + action(8 B) + old_logp, adv, c_adv, values, returns  [+ c_values, c_returns]
//   write A dlogits + dvalues [+ dcvalues]
// i.e. 8A + 44 B per (t, n) in the SafePPOLogGrad configuration.  The mean reductions are a
// deterministic two-stage tree (fixed block partials, last block folds them in a fixed order);
// the gradient of each mean is closed-form, so no second pass is needed.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int kRows = 128;  // rows per tile == threads per block
constexpr int kMaxA = 64;
constexpr int kNQ = 10;     // reduced quantities

struct PpoArgs {
  const float* logits; const int64_t* actions; const float* old_logp; const float* adv; const float* c_adv;
  const float* values; const float* returns; const float* old_values;
  const float* c_values; const float* c_returns; const float* old_c_values;
  const float* lambda_dev;
  svla_ppo_hparams hp;
  float* out; float* dlogits; float* dvalues; float* dcvalues;
  long long R; int A;
  float* partials; unsigned int* ticket;
};

// value-loss term: returns 0.5*(.)^2-style contribution and d/dv (un-normalised)
__device__ __forceinline__ void value_term(float v, float ret, const float* oldp, long long i, float clip,
                                           int clipped, float& loss, float& dv) {
  const float e = v - ret;
  if (!clipped || oldp == nullptr) {
    loss = 0.5f * e * e;
    dv = e;
    return;
  }
  const float old = oldp[i];
  const float d = v - old;
  const float dc = fminf(fmaxf(d, -clip), clip);
  const float ec = old + dc - ret;
  const float l1 = e * e, l2 = ec * ec;
  const float inside = (d >= -clip && d <= clip) ? 1.f : 0.f;
  loss = 0.5f * fmaxf(l1, l2);
  if (l1 > l2) dv = e;
  else if (l1 < l2) dv = ec * inside;
  else dv = 0.5f * e + 0.5f * ec * inside;  // torch.max splits ties
}

// deterministic two-stage fold of the per-thread partial sums: block partials, then the last block to finish folds
// them in a fixed order
__device__ __forceinline__ void finish_reduction(const PpoArgs& a, float* acc, float* red, unsigned int* s_ticket_p,
                                                 float pen) {
  const int tid = threadIdx.x;
  // stage 1: block partials
#pragma unroll
  for (int j = 0; j < kNQ; ++j) {
    const float s = block_sum(acc[j], red);
    if (tid == 0) a.partials[(size_t)blockIdx.x * 16 + j] = s;
  }
  __threadfence();
  if (tid == 0) *s_ticket_p = atomicAdd(a.ticket, 1u);
  __syncthreads();
  if (*s_ticket_p != gridDim.x - 1) return;
  // stage 2: last block folds the partials in a fixed order (deterministic)
  __threadfence();
  if (tid < 32) {
    float tot[kNQ];
#pragma unroll
    for (int j = 0; j < kNQ; ++j) {
      float s = 0.f;
      for (int b = tid; b < (int)gridDim.x; b += 32) s += a.partials[(size_t)b * 16 + j];
      tot[j] = warp_sum(s);
    }
    if (tid == 0) {
      const float ic = a.hp.inv_count;
      const float action = tot[0] * ic, ent = tot[1] * ic, val = tot[2] * ic, cval = tot[3] * ic;
      a.out[0] = a.hp.w_value * val + a.hp.w_action * action + a.hp.w_entropy * ent + a.hp.w_cvalue * cval;
      a.out[1] = val;
      a.out[2] = action;
      a.out[3] = ent;
      a.out[4] = cval;
      a.out[5] = tot[4] * ic;
      a.out[6] = tot[5] * ic;
      a.out[7] = tot[6] * ic;
      a.out[8] = pen;
      a.out[9] = tot[7];
      a.out[10] = tot[8];  // number of rows whose action index was outside [0, A)
      for (int j = 11; j < SVLA_PPO_NSCALARS; ++j) a.out[j] = 0.f;
      *a.ticket = 0u;
    }
  }
}

__global__ void __launch_bounds__(kRows) ppo_lag_kernel(PpoArgs a) {
  extern __shared__ float tile[];  // [kRows][A+1]
  __shared__ float red[32];
  __shared__ unsigned int s_ticket;
  const int A = a.A, ldt = A + 1, tid = threadIdx.x;
  const bool has_pi = a.logits != nullptr, has_v = a.values != nullptr, has_cv = a.c_values != nullptr;
  const float pen = (a.hp.use_lagrangian && a.lambda_dev) ? *a.lambda_dev : 0.f;
  const float clip = a.hp.clip_param;
  const bool vec4 = (A & 3) == 0 && (reinterpret_cast<uintptr_t>(a.logits) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(a.dlogits) & 15) == 0;
  const float gs = a.hp.inv_count * a.hp.grad_scale;
  float acc[kNQ];
#pragma unroll
  for (int j = 0; j < kNQ; ++j) acc[j] = 0.f;

  const long long ntiles = (a.R + kRows - 1) / kRows;
  for (long long tile_i = blockIdx.x; tile_i < ntiles; tile_i += gridDim.x) {
    const long long r0 = tile_i * kRows;
    const int rows = (int)min((long long)kRows, a.R - r0);
    const long long i = r0 + tid;
    const bool live = tid < rows;
    if (has_pi) {
      // coalesced load of the [rows, A] logits tile (128-bit when A % 4 == 0), no per-element div/mod
      const float* src = a.logits + r0 * A;
      const int cnt = rows * A;
      if (vec4) {
        const int A4 = A >> 2, cnt4 = cnt >> 2, dq = kRows / A4, dr = kRows % A4;
        int row = tid / A4, c4 = tid % A4;
        for (int e = tid; e < cnt4; e += kRows) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(src) + e);
          float* d = tile + row * ldt + c4 * 4;
          d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
          c4 += dr; row += dq;
          if (c4 >= A4) { c4 -= A4; ++row; }
        }
      } else {
        const int dq = kRows / A, dr = kRows % A;
        int row = tid / A, col = tid % A;
        for (int e = tid; e < cnt; e += kRows) {
          tile[row * ldt + col] = __ldg(src + e);
          col += dr; row += dq;
          if (col >= A) { col -= A; ++row; }
        }
      }
      __syncthreads();
      if (live) {
        float* row = tile + tid * ldt;
        float mx = -INFINITY;
        for (int k = 0; k < A; ++k) mx = fmaxf(mx, row[k]);
        int act = (int)a.actions[i];
        if (act < 0 || act >= A) {  // corrupted / mis-shaped actions: counted (out[10]) and clamped, never read out of range
          acc[8] += 1.f;
          act = min(max(act, 0), A - 1);
        }
        const float l_act = row[act] - mx;
        // one exp per logit: e_k overwrites the tile; H = log(se) - sum(e_k (l_k - mx)) / se
        float se = 0.f, sel = 0.f;
        for (int k = 0; k < A; ++k) {
          const float d = row[k] - mx;
          const float e = expf(d);
          se += e;
          sel += (e > 0.f) ? e * d : 0.f;
          row[k] = e;
        }
        const float lse_rel = logf(se), inv_se = 1.f / se;
        const float logp_a = l_act - lse_rel;
        const float H = lse_rel - sel * inv_se;
        const float oldlp = a.old_logp[i];
        const float ratio = expf(logp_a - oldlp);
        const float clamped = fminf(fmaxf(ratio, 1.f - clip), 1.f + clip);
        const float x = a.adv[i] - (a.c_adv ? pen * a.c_adv[i] : 0.f);
        const float onep = 1.f + pen;
        const float surr1 = ratio * x / onep, surr2 = clamped * x / onep;
        const bool use_clamped = surr2 < surr1;
        const float aloss = -(use_clamped ? surr2 : surr1);
        acc[0] += aloss;
        acc[1] += -H;
        acc[4] += oldlp - logp_a;
        acc[5] += (fabsf(ratio - 1.f) > clip) ? 1.f : 0.f;
        acc[6] += ratio;
        acc[7] += x / onep;
        // d total / d logits
        const float g_lp = use_clamped ? 0.f : -(ratio * x / onep) * a.hp.w_action * gs;
        const float g_ent = a.hp.w_entropy * gs;
        for (int k = 0; k < A; ++k) {
          const float e = row[k];
          const float p = e * inv_se;
          float g = g_lp * ((k == act ? 1.f : 0.f) - p);
          if (g_ent != 0.f) g += g_ent * ((e > 0.f) ? p * (logf(e) - lse_rel + H) : 0.f);
          row[k] = g;
        }
      }
      __syncthreads();
      if (a.dlogits) {
        float* dst = a.dlogits + r0 * A;
        if (vec4) {
          const int A4 = A >> 2, cnt4 = cnt >> 2, dq = kRows / A4, dr = kRows % A4;
          int row = tid / A4, c4 = tid % A4;
          for (int e = tid; e < cnt4; e += kRows) {
            const float* d = tile + row * ldt + c4 * 4;
            reinterpret_cast<float4*>(dst)[e] = make_float4(d[0], d[1], d[2], d[3]);
            c4 += dr; row += dq;
            if (c4 >= A4) { c4 -= A4; ++row; }
          }
        } else {
          const int dq = kRows / A, dr = kRows % A;
          int row = tid / A, col = tid % A;
          for (int e = tid; e < cnt; e += kRows) {
            dst[e] = tile[row * ldt + col];
            col += dr; row += dq;
            if (col >= A) { col -= A; ++row; }
          }
        }
      }
      __syncthreads();
    }
    if (live && has_v) {
      float l, dv;
      value_term(a.values[i], a.returns[i], a.old_values, i, clip, a.hp.use_clipped_value_loss, l, dv);
      acc[2] += l;
      if (a.dvalues) a.dvalues[i] = dv * a.hp.w_value * gs;
    }
    if (live && has_cv) {
      float l, dv;
      value_term(a.c_values[i], a.c_returns[i], a.old_c_values, i, clip, a.hp.use_clipped_value_loss, l, dv);
      acc[3] += l;
      if (a.dcvalues) a.dcvalues[i] = dv * a.hp.w_cvalue * gs;
    }
  }
  finish_reduction(a, acc, red, &s_ticket, pen);
}


// ---- register-resident variant (A = 4*NV, 16-byte aligned logits / dlogits) ------------------------------------
// A warp owns 32 consecutive rows.  Their [32, A] logits slab is read with coalesced 128-bit loads, turned
// row-per-lane through a warp-private shared-memory slab (one STS.128 + one LDS.128 per 16 bytes, row stride
// odd in 16-byte units: conflict-free), processed entirely in registers, and the gradient slab leaves the same
// way.  No block-wide barrier inside the loop, so every warp keeps its own loads in flight; per-row arithmetic
// (order of every sum) is the same as ppo_lag_kernel's.
struct RowScalars {
  int act;
  float oldlp, adv, cadv, val, ret, cval, cret;
};

template <int NV>
__global__ void __launch_bounds__(256, 2) ppo_lag_vec_kernel(PpoArgs a) {
  constexpr int A = 4 * NV, STR = NV | 1, kWarps = 8;
  extern __shared__ float4 slab4[];
  __shared__ float red[32];
  __shared__ unsigned int s_ticket;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float4* ws = slab4 + warp * 32 * STR;
  const bool has_v = a.values != nullptr, has_cv = a.c_values != nullptr;
  const float pen = (a.hp.use_lagrangian && a.lambda_dev) ? *a.lambda_dev : 0.f;
  const float clip = a.hp.clip_param;
  const float gs = a.hp.inv_count * a.hp.grad_scale;
  const float onep = 1.f + pen;
  const float g_ent = a.hp.w_entropy * gs;
  float acc[kNQ];
#pragma unroll
  for (int j = 0; j < kNQ; ++j) acc[j] = 0.f;

  const long long ntiles = (a.R + 31) / 32;
  const long long stride = (long long)gridDim.x * kWarps;
  // loads of one 32-row tile: the logits slab (coalesced 128-bit) and this lane's per-row scalars
  float4 v[NV];
  RowScalars nx;
  auto fetch = [&](long long wt) {
    const long long r0 = wt * 32;
    const int rows = (int)min(32LL, a.R - r0);
    const float4* src4 = reinterpret_cast<const float4*>(a.logits + r0 * A);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int e = j * 32 + lane;
      if (e < rows * NV) v[j] = __ldg(src4 + e);
    }
    nx.act = 0;
    nx.oldlp = nx.adv = nx.cadv = nx.val = nx.ret = nx.cval = nx.cret = 0.f;
    if (lane < rows) {
      const long long i = r0 + lane;
      nx.act = (int)__ldg(a.actions + i);
      nx.oldlp = __ldg(a.old_logp + i);
      nx.adv = __ldg(a.adv + i);
      if (a.c_adv) nx.cadv = __ldg(a.c_adv + i);
      if (has_v) { nx.val = __ldg(a.values + i); nx.ret = __ldg(a.returns + i); }
      if (has_cv) { nx.cval = __ldg(a.c_values + i); nx.cret = __ldg(a.c_returns + i); }
    }
  };
  long long wt = (long long)blockIdx.x * kWarps + warp;
  if (wt < ntiles) fetch(wt);
  for (; wt < ntiles; wt += stride) {
    const long long r0 = wt * 32;
    const int rows = (int)min(32LL, a.R - r0);
    const long long i = r0 + lane;
    const bool live = lane < rows;
    const int cnt4 = rows * NV;
    const RowScalars c = nx;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int e = j * 32 + lane;
      if (e < cnt4) ws[(e / NV) * STR + (e % NV)] = v[j];
    }
    __syncwarp();
    float row[A];
    if (live) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 t = ws[lane * STR + j];
        row[4 * j] = t.x; row[4 * j + 1] = t.y; row[4 * j + 2] = t.z; row[4 * j + 3] = t.w;
      }
    }
    if (wt + stride < ntiles) fetch(wt + stride);  // next tile's loads fly during this tile's arithmetic
    if (live) {
      int act = c.act;
      if (act < 0 || act >= A) {  // counted (out[10]) and clamped, same rule as ppo_lag_kernel
        acc[8] += 1.f;
        act = min(max(act, 0), A - 1);
      }
      float mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < A; ++k) mx = fmaxf(mx, row[k]);
      float l_act = 0.f;
#pragma unroll
      for (int k = 0; k < A; ++k) l_act = (k == act) ? row[k] : l_act;
      l_act -= mx;
      float se = 0.f, sel = 0.f;
#pragma unroll
      for (int k = 0; k < A; ++k) {
        const float d = row[k] - mx;
        const float e = expf(d);
        se += e;
        sel += (e > 0.f) ? e * d : 0.f;
        row[k] = e;
      }
      const float lse_rel = logf(se), inv_se = 1.f / se;
      const float logp_a = l_act - lse_rel;
      const float H = lse_rel - sel * inv_se;
      const float ratio = expf(logp_a - c.oldlp);
      const float clamped = fminf(fmaxf(ratio, 1.f - clip), 1.f + clip);
      const float x = c.adv - (a.c_adv ? pen * c.cadv : 0.f);
      const float surr1 = ratio * x / onep, surr2 = clamped * x / onep;
      const bool use_clamped = surr2 < surr1;
      const float aloss = -(use_clamped ? surr2 : surr1);
      acc[0] += aloss;
      acc[1] += -H;
      acc[4] += c.oldlp - logp_a;
      acc[5] += (fabsf(ratio - 1.f) > clip) ? 1.f : 0.f;
      acc[6] += ratio;
      acc[7] += x / onep;
      const float g_lp = use_clamped ? 0.f : -(ratio * x / onep) * a.hp.w_action * gs;
#pragma unroll
      for (int k = 0; k < A; ++k) {
        const float e = row[k];
        const float p = e * inv_se;
        float g = g_lp * ((k == act ? 1.f : 0.f) - p);
        if (g_ent != 0.f) g += g_ent * ((e > 0.f) ? p * (logf(e) - lse_rel + H) : 0.f);
        row[k] = g;
      }
      if (a.dlogits) {
#pragma unroll
        for (int j = 0; j < NV; ++j)
          ws[lane * STR + j] = make_float4(row[4 * j], row[4 * j + 1], row[4 * j + 2], row[4 * j + 3]);
      }
    }
    __syncwarp();
    if (a.dlogits) {
      float4* dst4 = reinterpret_cast<float4*>(a.dlogits + r0 * A);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int e = j * 32 + lane;
        if (e < cnt4) dst4[e] = ws[(e / NV) * STR + (e % NV)];
      }
    }
    __syncwarp();
    if (live && has_v) {
      float l, dv;
      value_term(c.val, c.ret, a.old_values, i, clip, a.hp.use_clipped_value_loss, l, dv);
      acc[2] += l;
      if (a.dvalues) a.dvalues[i] = dv * a.hp.w_value * gs;
    }
    if (live && has_cv) {
      float l, dv;
      value_term(c.cval, c.cret, a.old_c_values, i, clip, a.hp.use_clipped_value_loss, l, dv);
      acc[3] += l;
      if (a.dcvalues) a.dcvalues[i] = dv * a.hp.w_cvalue * gs;
    }
  }
  finish_reduction(a, acc, red, &s_ticket, pen);
}

template <int NV>
void launch_vec(const PpoArgs& a, int sm_count, cudaStream_t st) {
  constexpr int STR = NV | 1;
  const long long ntiles = (a.R + 31) / 32;
  int grid = (int)std::min<long long>((ntiles + 7) / 8, (long long)sm_count * 2);
  if (grid > kMaxPartialBlocks) grid = kMaxPartialBlocks;
  ppo_lag_vec_kernel<NV><<<grid, 256, 8 * 32 * STR * sizeof(float4), st>>>(a);
}

}  // namespace

extern "C" int svla_ppo_lag_fwd_bwd(svla_ctx* ctx, const float* logits, const int64_t* actions, const float* old_logp,
                                    const float* adv, const float* c_adv, const float* values, const float* returns,
                                    const float* old_values, const float* c_values, const float* c_returns,
                                    const float* old_c_values, const float* lambda_dev, const svla_ppo_hparams* hp,
                                    float* out_scalars, float* dlogits, float* dvalues, float* dcvalues, long long R,
                                    int A, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && hp && out_scalars, "NULL ctx/hp/out_scalars");
  SVLA_CHECK_ARG(R > 0, "empty batch");
  if (logits) {
    SVLA_CHECK_ARG(A >= 1 && A <= kMaxA, "A must be in [1, 64]");
    SVLA_CHECK_ARG(actions && old_logp && adv, "policy term needs actions/old_logp/adv");
    SVLA_CHECK_ARG(!hp->use_lagrangian || (c_adv && lambda_dev), "lagrangian term needs c_adv and lambda");
  }
  SVLA_CHECK_ARG(!values || returns, "values without returns");
  SVLA_CHECK_ARG(!c_values || c_returns, "c_values without c_returns");
  PpoArgs a;
  a.logits = logits; a.actions = actions; a.old_logp = old_logp; a.adv = adv; a.c_adv = c_adv;
  a.values = values; a.returns = returns; a.old_values = old_values;
  a.c_values = c_values; a.c_returns = c_returns; a.old_c_values = old_c_values;
  a.lambda_dev = lambda_dev; a.hp = *hp;
  a.out = out_scalars; a.dlogits = dlogits; a.dvalues = dvalues; a.dcvalues = dcvalues;
  a.R = R; a.A = logits ? A : 1;
  a.partials = ctx->partials; a.ticket = ctx->tickets + 0;
  const long long ntiles = (R + kRows - 1) / kRows;
  int grid = (int)std::min<long long>(ntiles, (long long)ctx->sm_count * 8);
  if (grid > kMaxPartialBlocks) grid = kMaxPartialBlocks;
  const size_t smem = sizeof(float) * kRows * (a.A + 1);
  const bool vec = logits && (A % 4) == 0 && A <= 32 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0 &&
                   (!dlogits || (reinterpret_cast<uintptr_t>(dlogits) & 15) == 0) && !getenv("SVLA_PPO_GENERIC");
  if (vec) {
    switch (A / 4) {
      case 1: launch_vec<1>(a, ctx->sm_count, as_stream(stream)); break;
      case 2: launch_vec<2>(a, ctx->sm_count, as_stream(stream)); break;
      case 3: launch_vec<3>(a, ctx->sm_count, as_stream(stream)); break;
      case 4: launch_vec<4>(a, ctx->sm_count, as_stream(stream)); break;
      case 5: launch_vec<5>(a, ctx->sm_count, as_stream(stream)); break;
      case 6: launch_vec<6>(a, ctx->sm_count, as_stream(stream)); break;
      case 7: launch_vec<7>(a, ctx->sm_count, as_stream(stream)); break;
      default: launch_vec<8>(a, ctx->sm_count, as_stream(stream)); break;
    }
  } else {
    ppo_lag_kernel<<<grid, kRows, smem, as_stream(stream)>>>(a);
  }
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}
