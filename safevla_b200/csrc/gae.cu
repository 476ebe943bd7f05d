// GAE-lambda over the reward and cost streams in one launch.
//
//   delta_t = r_t + gamma * V_{t+1} * m_{t+1} - V_t
//   g_t     = delta_t + gamma*lam * m_{t+1} * g_{t+1}          (g_T = 0)
//   ret_t   = g_t + V_t ;  adv_t = ret_t - V_t ;  ret_T = V_T
//
// Algorithmic traffic: 20 B read + 16 B written per (t, n) for the stream pair (36 B).
// Two kernels:
//   * march: lanes = samplers (coalesced), warps = 16-step time chunks held in registers; chunks run their
//     dependent chains latest-first, handing g over through shared memory.  Same operation order and
//     roundings as the sequential recursion (explicit _rn intrinsics, no FMA contraction) -> bit-exact.
//   * warp scan: one warp per sampler, each lane owns a block of consecutive steps; the
//     recursion is the affine map g -> delta + a*g, composed with a warp suffix scan.
//     For small N (BASELINE shapes: 64 samplers), where the march has no parallelism.
#include <algorithm>

#include "common.cuh"

namespace {


struct GaeArgs {
  const float* r[2];
  const float* v[2];
  float* ret[2];
  float* adv[2];
  const float* m;
  int T, N, ns;
  float gamma, gl;
};

// Block = kNW*32 samplers (512 contiguous bytes of every row: DRAM-page friendly) x kW time-chunks of kL steps.
// Every thread first pulls its chunk into registers (all loads of the block in flight at once), computes what does
// not depend on g (delta_t, the coefficient gamma*lam*m_{t+1}) in parallel, then the chunks run their short
// dependent chains one after the other, latest first, handing g across through shared memory -- the exact
// operation order of the sequential recursion, so the result is bit-identical to it, while HBM sees one
// parallel pass (20 B read + 16 B written per (t, n)).
template <int kL, int kW, int kNW>
__global__ void __launch_bounds__(32 * kW * kNW) gae_march_kernel(GaeArgs a) {
  __shared__ float sG[2][32 * kNW];
  const int lane = (threadIdx.x & 31) + 32 * ((threadIdx.x >> 5) % kNW), warp = (threadIdx.x >> 5) / kNW;
  const int n = blockIdx.x * (32 * kNW) + lane;
  const bool live = n < a.N;
  const int T = a.T, N = a.N;
  if (warp == 0) {
    sG[0][lane] = 0.f;
    sG[1][lane] = 0.f;
    if (live) {
      a.ret[0][(size_t)T * N + n] = a.v[0][(size_t)T * N + n];
      if (a.ns > 1) a.ret[1][(size_t)T * N + n] = a.v[1][(size_t)T * N + n];
    }
  }
  const int span = kL * kW;
  for (int sup = (T + span - 1) / span - 1; sup >= 0; --sup) {
    const int t0 = sup * span + warp * kL;
    const int cnt = live ? max(0, min(kL, T - t0)) : 0;
    float rr[2][kL], vv[2][kL + 1], mm[kL];
#pragma unroll
    for (int j = 0; j < kL; ++j) {
      if (j < cnt) {
        const size_t off = (size_t)(t0 + j) * N + n;
        mm[j] = __ldg(a.m + off + N);  // m_{t+1}
#pragma unroll
        for (int s = 0; s < 2; ++s)
          if (s < a.ns) {
            rr[s][j] = __ldg(a.r[s] + off);
            vv[s][j] = __ldg(a.v[s] + off);
          }
      }
    }
#pragma unroll
    for (int j = 0; j < kL; ++j)  // V_{t+1} of the chunk's last step
      if (j == cnt - 1) {
#pragma unroll
        for (int s = 0; s < 2; ++s)
          if (s < a.ns) vv[s][j + 1] = __ldg(a.v[s] + (size_t)(t0 + j + 1) * N + n);
      }
    // everything that does not depend on g is done by all chunks in parallel: delta_t (in place of r_t) and
    // the recursion coefficient gamma*lam*m_{t+1} (in place of m_{t+1})
#pragma unroll
    for (int j = 0; j < kL; ++j)
      if (j < cnt) {
        const float m1 = mm[j];
#pragma unroll
        for (int s = 0; s < 2; ++s)
          if (s < a.ns)
            rr[s][j] = __fsub_rn(__fadd_rn(rr[s][j], __fmul_rn(__fmul_rn(a.gamma, vv[s][j + 1]), m1)), vv[s][j]);
        mm[j] = __fmul_rn(a.gl, m1);
      }
    __syncthreads();  // sG initialised / handed over from the previous super-chunk
    // the only serial part: g_t = delta_t + coef_t * g_{t+1}, two dependent roundings per step
#pragma unroll 1
    for (int c = kW - 1; c >= 0; --c) {
      if (warp == c && cnt > 0) {
        float g0 = sG[0][lane], g1 = sG[1][lane];
#pragma unroll
        for (int j = kL - 1; j >= 0; --j)
          if (j < cnt) {
            g0 = __fadd_rn(rr[0][j], __fmul_rn(mm[j], g0));
            rr[0][j] = g0;
            if (a.ns > 1) {
              g1 = __fadd_rn(rr[1][j], __fmul_rn(mm[j], g1));
              rr[1][j] = g1;
            }
          }
        sG[0][lane] = g0;
        sG[1][lane] = g1;
      }
      __syncthreads();
    }
    // returns / advantages from the g held in registers, all chunks in parallel again
#pragma unroll
    for (int j = 0; j < kL; ++j)
      if (j < cnt) {
        const size_t off = (size_t)(t0 + j) * N + n;
#pragma unroll
        for (int s = 0; s < 2; ++s)
          if (s < a.ns) {
            const float ret = __fadd_rn(rr[s][j], vv[s][j]);
            (s == 0 ? a.ret[0] : a.ret[1])[off] = ret;
            (s == 0 ? a.adv[0] : a.adv[1])[off] = __fsub_rn(ret, vv[s][j]);
          }
      }
  }
}

// ---- 128-bit variant of the march: every lane owns FOUR adjacent samplers --------------------------------------
// Same schedule and the same per-sampler operation order as gae_march_kernel (bit-exact), but every global access is
// a 16-byte vector (a warp covers 512 contiguous bytes of a row), a quarter of the load/store instructions and
// address arithmetic per byte.  Needs N % 4 == 0 and 16-byte aligned buffers.
struct F4 {
  float f[4];
};
__device__ __forceinline__ F4 ldg4(const float* p) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p));
  F4 o;
  o.f[0] = t.x; o.f[1] = t.y; o.f[2] = t.z; o.f[3] = t.w;
  return o;
}
__device__ __forceinline__ void stg4(float* p, const F4& v) {
  *reinterpret_cast<float4*>(p) = make_float4(v.f[0], v.f[1], v.f[2], v.f[3]);
}

template <int kL, int kW, int kNW>
__global__ void __launch_bounds__(32 * kW * kNW) gae_march4_kernel(GaeArgs a) {
  __shared__ float4 sG[2][32 * kNW];
  const int lane = (threadIdx.x & 31) + 32 * ((threadIdx.x >> 5) % kNW), warp = (threadIdx.x >> 5) / kNW;
  const int n = (blockIdx.x * (32 * kNW) + lane) * 4;
  const bool live = n < a.N;
  const int T = a.T, N = a.N;
  if (warp == 0) {
    sG[0][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
    sG[1][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
      stg4(a.ret[0] + (size_t)T * N + n, ldg4(a.v[0] + (size_t)T * N + n));
      if (a.ns > 1) stg4(a.ret[1] + (size_t)T * N + n, ldg4(a.v[1] + (size_t)T * N + n));
    }
  }
  const int span = kL * kW;
  for (int sup = (T + span - 1) / span - 1; sup >= 0; --sup) {
    const int t0 = sup * span + warp * kL;
    const int cnt = live ? max(0, min(kL, T - t0)) : 0;
    F4 rr[2][kL], vv[2][kL + 1], mm[kL];
#pragma unroll
    for (int j = 0; j < kL; ++j) {
      if (j < cnt) {
        const size_t off = (size_t)(t0 + j) * N + n;
        mm[j] = ldg4(a.m + off + N);  // m_{t+1}
#pragma unroll
        for (int s = 0; s < 2; ++s)
          if (s < a.ns) {
            rr[s][j] = ldg4(a.r[s] + off);
            vv[s][j] = ldg4(a.v[s] + off);
          }
      }
    }
#pragma unroll
    for (int j = 0; j < kL; ++j)  // V_{t+1} of the chunk's last step
      if (j == cnt - 1) {
#pragma unroll
        for (int s = 0; s < 2; ++s)
          if (s < a.ns) vv[s][j + 1] = ldg4(a.v[s] + (size_t)(t0 + j + 1) * N + n);
      }
#pragma unroll
    for (int j = 0; j < kL; ++j)
      if (j < cnt) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float m1 = mm[j].f[k];
#pragma unroll
          for (int s = 0; s < 2; ++s)
            if (s < a.ns)
              rr[s][j].f[k] = __fsub_rn(__fadd_rn(rr[s][j].f[k], __fmul_rn(__fmul_rn(a.gamma, vv[s][j + 1].f[k]), m1)),
                                        vv[s][j].f[k]);
          mm[j].f[k] = __fmul_rn(a.gl, m1);
        }
      }
    __syncthreads();
#pragma unroll 1
    for (int c = kW - 1; c >= 0; --c) {
      if (warp == c && cnt > 0) {
        const float4 t0v = sG[0][lane], t1v = sG[1][lane];
        float g0[4] = {t0v.x, t0v.y, t0v.z, t0v.w}, g1[4] = {t1v.x, t1v.y, t1v.z, t1v.w};
#pragma unroll
        for (int j = kL - 1; j >= 0; --j)
          if (j < cnt) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              g0[k] = __fadd_rn(rr[0][j].f[k], __fmul_rn(mm[j].f[k], g0[k]));
              rr[0][j].f[k] = g0[k];
              if (a.ns > 1) {
                g1[k] = __fadd_rn(rr[1][j].f[k], __fmul_rn(mm[j].f[k], g1[k]));
                rr[1][j].f[k] = g1[k];
              }
            }
          }
        sG[0][lane] = make_float4(g0[0], g0[1], g0[2], g0[3]);
        sG[1][lane] = make_float4(g1[0], g1[1], g1[2], g1[3]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < kL; ++j)
      if (j < cnt) {
        const size_t off = (size_t)(t0 + j) * N + n;
#pragma unroll
        for (int s = 0; s < 2; ++s)
          if (s < a.ns) {
            F4 ret, adv;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              ret.f[k] = __fadd_rn(rr[s][j].f[k], vv[s][j].f[k]);
              adv.f[k] = __fsub_rn(ret.f[k], vv[s][j].f[k]);
            }
            stg4((s == 0 ? a.ret[0] : a.ret[1]) + off, ret);
            stg4((s == 0 ? a.adv[0] : a.adv[1]) + off, adv);
          }
      }
  }
}

template <int kL, int kW, int kNW>
void launch_march4(const GaeArgs& a, cudaStream_t st) {
  gae_march4_kernel<kL, kW, kNW><<<(a.N / 4 + 32 * kNW - 1) / (32 * kNW), 32 * kW * kNW, 0, st>>>(a);
}

// warp per sampler; lane owns steps [lane*L, lane*L+L)
__global__ void __launch_bounds__(128) gae_warp_scan_kernel(GaeArgs a) {
  const int lane = threadIdx.x & 31;
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (n >= a.N) return;
  const int T = a.T, N = a.N;
  const int L = (T + 31) / 32;
  const int t0 = min(lane * L, T), t1 = min(t0 + L, T);
  float A[2] = {1.f, 1.f}, B[2] = {0.f, 0.f};
  for (int t = t1 - 1; t >= t0; --t) {  // chunk map g_{t1} -> g_{t0} = B + A * g_{t1}
    const size_t off = (size_t)t * N + n;
    const float m1 = a.m[off + N], al = a.gl * m1;
    for (int s = 0; s < a.ns; ++s) {
      const float delta = a.r[s][off] + a.gamma * a.v[s][off + N] * m1 - a.v[s][off];
      // apply step t after the already-composed later steps: new(x) = delta + al * (B + A x)
      B[s] = delta + al * B[s];
      A[s] = al * A[s];
    }
  }
  // inclusive suffix scan of compositions: F_l = f_l o f_{l+1} o ... o f_31
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const float A2 = __shfl_down_sync(0xffffffffu, A[s], o), B2 = __shfl_down_sync(0xffffffffu, B[s], o);
      if (lane + o < 32) {
        B[s] = B[s] + A[s] * B2;
        A[s] = A[s] * A2;
      }
    }
  }
  float g[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const float nxt = __shfl_down_sync(0xffffffffu, B[s], 1);  // F_{l+1}(0) = g entering lane l's block
    g[s] = (lane == 31) ? 0.f : nxt;
  }
  for (int t = t1 - 1; t >= t0; --t) {
    const size_t off = (size_t)t * N + n;
    const float m1 = a.m[off + N], al = a.gl * m1;
    for (int s = 0; s < a.ns; ++s) {
      const float v0 = a.v[s][off];
      const float delta = a.r[s][off] + a.gamma * a.v[s][off + N] * m1 - v0;
      g[s] = delta + al * g[s];
      const float ret = g[s] + v0;
      a.ret[s][off] = ret;
      a.adv[s][off] = ret - v0;
    }
  }
  if (lane == 0)
    for (int s = 0; s < a.ns; ++s) a.ret[s][(size_t)T * N + n] = a.v[s][(size_t)T * N + n];
}

// ---- advantage normalisation --------------------------------------------------------------
__global__ void __launch_bounds__(256) adv_stats_partial(const float* x, long long n, float* partials) {
  __shared__ float red[32];
  float s = 0.f, q = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    s += v;
    q += v * v;
  }
  s = block_sum(s, red);
  q = block_sum(q, red);
  if (threadIdx.x == 0) {
    partials[blockIdx.x * 2] = s;
    partials[blockIdx.x * 2 + 1] = q;
  }
}
__global__ void adv_stats_final(const float* partials, int nb, long long n, float* stats) {
  __shared__ float red[32];
  double s = 0.0, q = 0.0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < nb; ++i) {
      s += partials[2 * i];
      q += partials[2 * i + 1];
    }
    const double mean = s / (double)n;
    const double var = n > 1 ? (q - (double)n * mean * mean) / (double)(n - 1) : 0.0;  // torch.std: unbiased
    stats[0] = (float)mean;
    stats[1] = (float)sqrt(var > 0.0 ? var : 0.0);
  }
  (void)red;
}
// split form for data-parallel runs: sums[0..2] = {sum x, sum x^2, n}; the host all-reduces the three floats, then
// every rank derives the GLOBAL mean / unbiased std from them
__global__ void adv_sums_final(const float* partials, int nb, long long n, float* sums) {
  if (threadIdx.x == 0) {
    double s = 0.0, q = 0.0;
    for (int i = 0; i < nb; ++i) {
      s += partials[2 * i];
      q += partials[2 * i + 1];
    }
    sums[0] = (float)s;
    sums[1] = (float)q;
    sums[2] = (float)n;
  }
}
__global__ void adv_stats_from_sums(const float* sums, float* stats) {
  if (threadIdx.x == 0) {
    const double s = sums[0], q = sums[1], n = sums[2];
    const double mean = n > 0.0 ? s / n : 0.0;
    const double var = n > 1.0 ? (q - n * mean * mean) / (n - 1.0) : 0.0;
    stats[0] = (float)mean;
    stats[1] = (float)sqrt(var > 0.0 ? var : 0.0);
  }
}
__global__ void __launch_bounds__(256) adv_normalize(const float* x, float* y, long long n, const float* stats) {
  const float mean = stats[0], inv = 1.f / (stats[1] + 1e-5f);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = (x[i] - mean) * inv;
}


// ---- use_gae = False: plain discounted returns --------------------------------------------------------------
//   ret_T = V_T ;  ret_t = (ret_{t+1} * gamma) * m_{t+1} + r_t ;  adv_t = ret_t - V_t
// (upstream allenact RolloutBlockStorage.compute_returns, the `else` branch).  Lanes = samplers
// (coalesced rows), blockIdx.y = stream; each thread pulls a 16-step chunk of r / m / V into registers (all loads
// in flight at once), then runs the dependent chain: three roundings per step in the recursion's own order, no FMA.
constexpr int kDL = 16;
__global__ void __launch_bounds__(128) discounted_returns_kernel(GaeArgs a) {
  const int s = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.N) return;
  const float* r = a.r[s];
  const float* v = a.v[s];
  float* ret = a.ret[s];
  float* adv = a.adv[s];
  const int T = a.T, N = a.N;
  float g = v[(size_t)T * N + n];
  ret[(size_t)T * N + n] = g;
  for (int t1 = T; t1 > 0; t1 -= kDL) {
    const int t0 = max(0, t1 - kDL), cnt = t1 - t0;
    float rr[kDL], mm[kDL], vv[kDL];
#pragma unroll
    for (int j = 0; j < kDL; ++j)
      if (j < cnt) {
        const size_t off = (size_t)(t0 + j) * N + n;
        rr[j] = __ldg(r + off);
        mm[j] = __ldg(a.m + off + N);
        vv[j] = __ldg(v + off);
      }
#pragma unroll
    for (int j = kDL - 1; j >= 0; --j)
      if (j < cnt) {
        g = __fadd_rn(__fmul_rn(__fmul_rn(g, a.gamma), mm[j]), rr[j]);
        rr[j] = g;
      }
#pragma unroll
    for (int j = 0; j < kDL; ++j)
      if (j < cnt) {
        const size_t off = (size_t)(t0 + j) * N + n;
        ret[off] = rr[j];
        adv[off] = __fsub_rn(rr[j], vv[j]);
      }
  }
}

template <int kL, int kW, int kNW>
void launch_march(const GaeArgs& a, cudaStream_t st) {
  gae_march_kernel<kL, kW, kNW><<<(a.N + 32 * kNW - 1) / (32 * kNW), 32 * kW * kNW, 0, st>>>(a);
}

}  // namespace

extern "C" int svla_gae_dual(svla_ctx* ctx, const float* rewards, const float* costs, const float* value_preds,
                             const float* c_value_preds, const float* masks, float* returns, float* c_returns,
                             float* adv, float* c_adv, int T, int N, double gamma, double lam, int algo,
                             svla_stream stream) {
  SVLA_CHECK_ARG(ctx, "ctx is NULL");
  SVLA_CHECK_ARG(T >= 0 && N >= 0, "negative shape");
  SVLA_CHECK_ARG(rewards && value_preds && masks && returns && adv, "NULL reward-stream buffer");
  if (N == 0) return SVLA_OK;
  GaeArgs a;
  a.r[0] = rewards; a.v[0] = value_preds; a.ret[0] = returns; a.adv[0] = adv;
  a.r[1] = costs; a.v[1] = c_value_preds; a.ret[1] = c_returns; a.adv[1] = c_adv;
  a.ns = 1;
  if (costs) {
    SVLA_CHECK_ARG(c_value_preds && c_returns && c_adv, "NULL cost-stream buffer");
    a.ns = 2;
  }
  a.m = masks; a.T = T; a.N = N;
  a.gamma = (float)gamma;
  a.gl = (float)(gamma * lam);
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec_ok = N % 4 == 0 && al16(rewards) && al16(value_preds) && al16(masks) && al16(returns) && al16(adv) &&
                      (!costs || (al16(costs) && al16(c_value_preds) && al16(c_returns) && al16(c_adv)));
  if (algo == 0 || (algo >= 40 && algo < 50 && !vec_ok)) algo = 1;  // the chunked march is bit-exact and parallel over both N and T
  if (algo == 1) {
    // measured on B200 (tools/tune_gae.py): short register chunks win at large N (occupancy), long chunks at the
    // BASELINE shapes (64 samplers: one super-chunk, two blocks)
    if (N >= 8192 && vec_ok) launch_march4<1, 8, 1>(a, as_stream(stream));
    else if (N >= 2048) launch_march<2, 4, 4>(a, as_stream(stream));
    else launch_march<16, 8, 1>(a, as_stream(stream));
  } else if (algo == 11) {  // tuning variants of the same (bit-exact) kernel
    launch_march<16, 8, 1>(a, as_stream(stream));
  } else if (algo == 12) {
    launch_march<8, 4, 2>(a, as_stream(stream));
  } else if (algo == 13) {
    launch_march<16, 2, 4>(a, as_stream(stream));
  } else if (algo == 14) {
    launch_march<4, 4, 4>(a, as_stream(stream));
  } else if (algo == 15) {
    launch_march<8, 1, 4>(a, as_stream(stream));
  } else if (algo == 16) {
    launch_march<8, 2, 8>(a, as_stream(stream));
  } else if (algo == 17) {
    launch_march<4, 1, 4>(a, as_stream(stream));
  } else if (algo == 18) {
    launch_march<4, 2, 4>(a, as_stream(stream));
  } else if (algo == 19) {
    launch_march<4, 8, 4>(a, as_stream(stream));
  } else if (algo == 20) {
    launch_march<2, 4, 4>(a, as_stream(stream));
  } else if (algo == 21) {
    launch_march<2, 8, 4>(a, as_stream(stream));
  } else if (algo == 22) {
    launch_march<4, 4, 8>(a, as_stream(stream));
  } else if (algo == 23) {
    launch_march<4, 4, 2>(a, as_stream(stream));
  } else if (algo == 24) {
    launch_march<2, 2, 4>(a, as_stream(stream));
  } else if (algo == 25) {
    launch_march<2, 1, 4>(a, as_stream(stream));
  } else if (algo == 26) {
    launch_march<1, 4, 4>(a, as_stream(stream));
  } else if (algo >= 40 && algo < 50 && vec_ok) {
    switch (algo) {
      case 40: launch_march4<2, 4, 2>(a, as_stream(stream)); break;
      case 41: launch_march4<2, 4, 1>(a, as_stream(stream)); break;
      case 42: launch_march4<4, 4, 1>(a, as_stream(stream)); break;
      case 43: launch_march4<1, 16, 2>(a, as_stream(stream)); break;
      case 44: launch_march4<2, 8, 1>(a, as_stream(stream)); break;
      case 45: launch_march4<1, 8, 1>(a, as_stream(stream)); break;
      case 46: launch_march4<1, 8, 2>(a, as_stream(stream)); break;
      case 47: launch_march4<1, 16, 1>(a, as_stream(stream)); break;
      case 48: launch_march4<1, 4, 1>(a, as_stream(stream)); break;
      default: launch_march4<1, 4, 2>(a, as_stream(stream)); break;
    }
  } else if (algo == 2) {
    const long long threads = (long long)N * 32;
    gae_warp_scan_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, as_stream(stream)>>>(a);
  } else {
    svla_set_error("svla_gae_dual: bad algo %d", algo);
    return SVLA_ERR_BAD_ARG;
  }
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

extern "C" int svla_discounted_returns_dual(svla_ctx* ctx, const float* rewards, const float* costs,
                                            const float* value_preds, const float* c_value_preds, const float* masks,
                                            float* returns, float* c_returns, float* adv, float* c_adv, int T, int N,
                                            double gamma, svla_stream stream) {
  SVLA_CHECK_ARG(ctx, "ctx is NULL");
  SVLA_CHECK_ARG(T >= 0 && N >= 0, "negative shape");
  SVLA_CHECK_ARG(rewards && value_preds && masks && returns && adv, "NULL reward-stream buffer");
  if (N == 0) return SVLA_OK;
  GaeArgs a;
  a.r[0] = rewards; a.v[0] = value_preds; a.ret[0] = returns; a.adv[0] = adv;
  a.r[1] = costs; a.v[1] = c_value_preds; a.ret[1] = c_returns; a.adv[1] = c_adv;
  a.ns = 1;
  if (costs) {
    SVLA_CHECK_ARG(c_value_preds && c_returns && c_adv, "NULL cost-stream buffer");
    a.ns = 2;
  }
  a.m = masks; a.T = T; a.N = N;
  a.gamma = (float)gamma;
  a.gl = 0.f;
  discounted_returns_kernel<<<dim3((N + 127) / 128, a.ns), 128, 0, as_stream(stream)>>>(a);
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

extern "C" int svla_normalize_advantage(svla_ctx* ctx, const float* adv, float* norm_adv, float* stats,
                                        long long n, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && adv && norm_adv && stats, "NULL argument");
  if (n <= 0) return SVLA_OK;
  int nb = (int)((n + 255) / 256);
  if (nb > 1024) nb = 1024;
  adv_stats_partial<<<nb, 256, 0, as_stream(stream)>>>(adv, n, ctx->partials);
  SVLA_LAUNCH_CHECK();
  adv_stats_final<<<1, 32, 0, as_stream(stream)>>>(ctx->partials, nb, n, stats);
  SVLA_LAUNCH_CHECK();
  adv_normalize<<<nb, 256, 0, as_stream(stream)>>>(adv, norm_adv, n, stats);
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

extern "C" int svla_advantage_sums(svla_ctx* ctx, const float* adv, long long n, float* sums, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && adv && sums, "NULL argument");
  int nb = (int)((std::max<long long>(n, 1) + 255) / 256);
  if (nb > 1024) nb = 1024;
  adv_stats_partial<<<nb, 256, 0, as_stream(stream)>>>(adv, n, ctx->partials);
  SVLA_LAUNCH_CHECK();
  adv_sums_final<<<1, 32, 0, as_stream(stream)>>>(ctx->partials, nb, n, sums);
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

extern "C" int svla_normalize_advantage_from_sums(svla_ctx* ctx, const float* adv, float* norm_adv, const float* sums,
                                                  float* stats, long long n, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && adv && norm_adv && sums && stats, "NULL argument");
  adv_stats_from_sums<<<1, 32, 0, as_stream(stream)>>>(sums, stats);
  SVLA_LAUNCH_CHECK();
  if (n <= 0) return SVLA_OK;
  int nb = (int)((n + 255) / 256);
  if (nb > 1024) nb = 1024;
  adv_normalize<<<nb, 256, 0, as_stream(stream)>>>(adv, norm_adv, n, stats);
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}
