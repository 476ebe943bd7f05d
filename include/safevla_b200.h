/*
 * safevla_b200.h -- C ABI of libsafevla_b200.so (sm_100a only).
 *
 * The reference (PKU-Alignment/SafeVLA) is 100 % Python and has NO native boundary of its
 * own; every entry point below therefore replaces a *PyTorch-eager code region* of the
 * constrained-PPO update path, cited as path:line relative to the reference root next to
 * each declaration.  The host side (safevla_b200/*.py) binds this file with ctypes
 * (INTEGRATION.md shows the stub) and mirrors the allenact plugin classes on top of it.
 *
 * Conventions
 *   - every buffer is a caller-owned DEVICE pointer (PyTorch caching allocator); the library
 *     never frees caller memory and allocates only the small per-context scratch below;
 *   - all work is enqueued asynchronously on the caller's `stream` (graph-capturable: no
 *     host syncs, no allocation, no data-dependent host control flow);
 *   - every call returns int: 0 = OK, negative = svla_status, positive = cudaError_t;
 *     svla_last_error() gives a thread-local message; no C++ exception crosses the ABI;
 *   - the device must be compute capability 10.x: svla_ctx_create fails otherwise (there
 *     is no fallback path).
 */
#ifndef SAFEVLA_B200_H
#define SAFEVLA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct svla_ctx svla_ctx;
typedef void* svla_stream; /* cudaStream_t */

typedef enum {
  SVLA_OK = 0,
  SVLA_ERR_BAD_ARG = -1,
  SVLA_ERR_UNSUPPORTED_ARCH = -2,
  SVLA_ERR_BAD_SHAPE = -3,
  SVLA_ERR_NO_DEVICE = -4,
  SVLA_ERR_INTERNAL = -5
} svla_status;

typedef enum { SVLA_F32 = 0, SVLA_BF16 = 1 } svla_dtype;

/* ---- context -------------------------------------------------------------------------- */
int svla_ctx_create(int device, svla_ctx** out);
int svla_ctx_destroy(svla_ctx* ctx);
const char* svla_last_error(void);
int svla_version(void);
int svla_sm_count(svla_ctx* ctx);
/* number of kernels this library has launched in the calling process (monotonic) */
unsigned long long svla_launch_count(void);
/* the host replays a CUDA graph that was captured over n launches of this library: keep the counter honest */
int svla_launch_count_add(unsigned long long n);

/* ======================================================================================
 * Scan / elementwise path (HBM bound)
 * ====================================================================================== */

/* GAE-lambda over the reward AND the cost stream in one launch (the code is
 * in the un-vendored allenact fork -- call-site witness architecture/models/
 * allenact_transformer_models/inference_agent.py:255-267 -- replaces the fork's two
 * `compute_returns` Python loops).  rewards/costs [T,N]; value_preds/c_value_preds/masks
 * [T+1,N] with row T the bootstrap; outputs returns/c_returns [T+1,N] (row T = bootstrap),
 * adv/c_adv [T,N].  `algo`: 0 = auto, 1 = thread-per-sampler march (bit-exact with the
 * sequential recursion), 2 = warp affine scan (small N).  costs/c_* may be NULL (single stream). */
int svla_gae_dual(svla_ctx* ctx, const float* rewards, const float* costs, const float* value_preds,
                  const float* c_value_preds, const float* masks, float* returns, float* c_returns,
                  float* adv, float* c_adv, int T, int N, double gamma, double lam, int algo,
                  svla_stream stream);

/* The `use_gae = False` branch of the same `compute_returns` (not used by the shipped config, kept
 * for API completeness): ret_T = V_T, ret_t = ret_{t+1} * gamma * m_{t+1} + r_t, adv_t = ret_t - V_t, on both
 * streams in one launch; same shapes as svla_gae_dual, bit-exact with the sequential loop. */
int svla_discounted_returns_dual(svla_ctx* ctx, const float* rewards, const float* costs, const float* value_preds,
                                 const float* c_value_preds, const float* masks, float* returns, float* c_returns,
                                 float* adv, float* c_adv, int T, int N, double gamma, svla_stream stream);

/* mean / unbiased std of adv over all T*N elements -> stats[0..1] (device), then
 * norm_adv = (adv - mean) / (std + 1e-5)  (allenact `norm_adv_targ`; unused by the shipped
 * config: normalize_advantage=False, training/online/dinov2_vits_tsfm_base.py:321). */
int svla_normalize_advantage(svla_ctx* ctx, const float* adv, float* norm_adv, float* stats,
                             long long n, svla_stream stream);

/* The same normalisation in two phases for data-parallel runs: sums[0..2] = {sum adv,
 * sum adv^2, n} of the local shard; the host all-reduces the three floats; every rank then normalises with the GLOBAL
 * mean / unbiased std (written to stats[0..1]). */
int svla_advantage_sums(svla_ctx* ctx, const float* adv, long long n, float* sums, svla_stream stream);
int svla_normalize_advantage_from_sums(svla_ctx* ctx, const float* adv, float* norm_adv, const float* sums,
                                       float* stats, long long n, svla_stream stream);

typedef struct {
  float clip_param;     /* training/online/loss/customized_loss.py:342 */
  float w_action;       /* action_loss_schedule(step) :394 */
  float w_value;        /* value_loss_coef :399 (PPOValue: 1.0) */
  float w_entropy;      /* entropy_coef :401 */
  float w_cvalue;       /* SafePPOValue weight (stage 0), 0 otherwise */
  float inv_count;      /* 1 / (T*N) of the batch the means run over */
  float grad_scale;     /* extra factor on every gradient (data-parallel weighting), 1.0 */
  int use_clipped_value_loss; /* :374-380 */
  int use_lagrangian;   /* 1: SafePPOLogGrad :350-359, 0: PPOLogGrad :208-212 */
} svla_ppo_hparams;

#define SVLA_PPO_NSCALARS 16
/* out_scalars (device, float[16]): 0 total, 1 value_loss, 2 action_loss(mean), 3 entropy term
 * (= mean(-H), the sign the reference logs, :401), 4 cvalue_loss, 5 approx KL(old||new),
 * 6 clip fraction, 7 mean ratio, 8 penalty (lambda used), 9 sum adv_hat, 10 number of rows whose action index was
 * outside [0, A) (such rows are evaluated with the index clamped; callers treat a non-zero count as an error),
 * 11..15 reserved. */

/* Fused SafePPOLogGrad / PPOLogGrad / PPOValue / SafePPOValue forward AND backward
 * (customized_loss.py:317-449, :178-298).  logits [R,A] fp32; actions int64
 * [R]; the rest fp32 [R].  `lambda_dev` is read on the device (no host sync; :348-349).
 * Any of values/returns (value term), c_values/c_returns (cost-value term), logits (policy
 * term) may be NULL to drop that term.  dlogits [R,A], dvalues [R], dcvalues [R] receive
 * d total / d input (may be NULL when the term is absent). */
int svla_ppo_lag_fwd_bwd(svla_ctx* ctx, const float* logits, const int64_t* actions, const float* old_logp,
                         const float* adv, const float* c_adv, const float* values, const float* returns,
                         const float* old_values, const float* c_values, const float* c_returns,
                         const float* old_c_values, const float* lambda_dev, const svla_ppo_hparams* hp,
                         float* out_scalars, float* dlogits, float* dvalues, float* dcvalues,
                         long long R, int A, svla_stream stream);

/* Lagrange multiplier update, omnisafe 0.5.0 common/lagrange.py (imported at
 * customized_loss.py:14; cost_limit plumbed at training/online/allenact_trainer.py:22,71):
 * Jc = cost_sum_cnt[0] / max(cost_sum_cnt[1], 1); Adam step on loss -lambda*(Jc - limit);
 * lambda clamped to [0, upper_bound] (upper_bound < 0: unbounded).  state_dev = float[4] {m, v, step, last Jc}. */
int svla_lagrange_update(svla_ctx* ctx, float* lambda_dev, float* state_dev, const float* cost_sum_cnt_dev,
                         float cost_limit, float lr, float upper_bound, svla_stream stream);

/* sum of squares of a flat fp32 buffer -> out_dev[0] (deterministic two-stage reduction).
 * Replaces DinoLLAMATxNavActorCritic.compute_total_grad_norm, allenact_dino_transformer.py:289-297,
 * and the norm half of clip_grad_norm_. */
int svla_sq_norm(svla_ctx* ctx, const float* x, long long n, float* out_dev, svla_stream stream);

typedef struct {
  float lr, beta1, beta2, eps;
  float max_grad_norm; /* <= 0: no clipping */
  float grad_prescale; /* multiplies g before everything (e.g. 1/world_size after a sum all-reduce) */
  int step;            /* 1-based Adam step */
  int zero_grad;       /* write zeros back to g */
} svla_adam_hparams;

/* Fused global-norm clip + Adam over the flat arena (training/online/dinov2_vits_tsfm_base.py:331,334;
 * torch.nn.utils.clip_grad_norm_ + torch.optim.Adam in the fork's engine).  sq_norm_dev holds the
 * sum of squares of (grad_prescale * g) computed by svla_sq_norm (pass NULL when not clipping).
 * Optionally refreshes the bf16 shadow of the parameters used by the tensor-core path. */
int svla_clip_adam(svla_ctx* ctx, float* p, float* g, float* m, float* v, void* p_bf16, long long n,
                   const float* sq_norm_dev, const svla_adam_hparams* hp, svla_stream stream);

/* HL-Gauss discrete-critic loss fwd+bwd (utils/loss_functions.py:7-30; read-out of DiscreteCriticHead,
 * allenact_dino_transformer.py:743-766): logits [R, num_bins] (row stride ldl), target [R], support
 * [num_bins + 1] = the module's torch.linspace(min, max, num_bins + 1).  out_loss[0] = F.cross_entropy(logits,
 * transform_to_probs(target)); dlogits (or NULL) = d out_loss / d logits * grad_scale; values (or NULL) [R] =
 * transform_from_probs(softmax(logits)). */
int svla_hl_gauss_fwd_bwd(svla_ctx* ctx, const float* logits, long long ldl, const float* target, const float* support,
                          int num_bins, float sigma, float grad_scale, float* out_loss, float* dlogits, float* values,
                          long long R, svla_stream stream);

/* ======================================================================================
 * Dense path (tensor-core bound): building blocks of the three towers
 * ====================================================================================== */

/* Training-mode dropout (nn.TransformerEncoderLayer p = 0.1 in the fusion block, allenact_dino_transformer.py:545-552;
 * SURVEY.md fact 8).  Masks are counter-based (Philox4x32-7, csrc/philox.cuh): element (row, col) of a site is kept
 * iff the 16-bit uniform drawn from (seed, step, site, row0 + row, col / 8)[col % 8] >= round(p * 65536), kept values
 * are scaled by 1 / (1 - p).  Nothing is stored: forward, backward, recompute and svla_dropout_rows (below) regenerate
 * the same mask from the same five numbers. */
typedef struct {
  float p;                  /* drop probability, 0 <= p < 1; 0 = off */
  unsigned long long seed;
  unsigned int site;        /* which dropout of the network (tower, layer, position) */
  unsigned int step;        /* update repeat / rollout counter */
  unsigned int row0;        /* global row of the tensor's first row (row chunks of one logical tensor) */
  unsigned int row_stride;  /* global rows between consecutive rows of the tensor (0 = 1); the CLS-row tensors of the
                               last fusion layer address every S-th row of the full sequence tensor */
} svla_dropout;

/* out[r, c] = keep(r, c) ? x[r, c] / (1 - p) : 0 over [rows, cols] (row strides ldx / ldo; in place allowed).  The
 * stand-alone form of every fused dropout site: dropout1 / dropout2 of the encoder layer around the residual adds, the
 * gradient masks of the backward, and -- applied to ones -- the mask itself (tests). */
int svla_dropout_rows(svla_ctx* ctx, const void* x, int dtype_x, long long ldx, void* out, int dtype_out, long long ldo,
                      long long rows, int cols, const svla_dropout* drop, svla_stream stream);

typedef enum {
  SVLA_EPI_NONE = 0,
  SVLA_EPI_RELU = 1,        /* C = relu(acc + bias) */
  SVLA_EPI_RELU_MASK = 2,   /* C = (acc + bias) * (aux > 0): ReLU backward fused into a dgrad */
  SVLA_EPI_GELU = 3,        /* C = gelu_erf(acc + bias): DINOv2 MLP (vision preprocessor, forward only) */
  /* ReLU with a one-bit-per-element record instead of re-reading the activation in the backward: the K = 512 GEMMs
   * of the fusion block are HBM-bound, and the [M, N] bf16 mask operand of RELU_MASK is 44 % of the masked dgrad's
   * traffic; the bit record is 1/16 of it.  `aux` = uint32 [M, N / 32] (ldaux in words), bit ((e >> 1) + 16 (e & 1))
   * of word n / 32 for column n = 32 (n / 32) + e.  tcgen05 path only: bf16 C, N % 64 == 0, no residual / accumulate
   * (svla_gemm returns SVLA_ERR_BAD_SHAPE otherwise; callers fall back to RELU / RELU_MASK). */
  SVLA_EPI_RELU_BITS = 4,   /* C = relu(acc + bias), aux bit = (C > 0)   (written) */
  SVLA_EPI_MASK_BITS = 5    /* C = acc * aux bit                          (read)    */
} svla_epilogue;

typedef struct {
  int M, N, K;
  const void* A; long long lda; int transA; /* transA=0: A is [M,K] row-major; 1: A is [K,M] row-major */
  const void* B; long long ldb; int transB; /* transB=0: B is [K,N] row-major; 1: B is [N,K] row-major (nn.Linear weight) */
  void* C; long long ldc;
  int dtypeA, dtypeB, dtypeC;       /* svla_dtype; accumulation is always fp32 */
  const float* bias;                /* [N] or NULL */
  const void* residual; long long ldr; int dtypeR; /* added after bias/activation, or NULL */
  const void* aux; long long ldaux; int dtypeAux;  /* RELU_MASK operand */
  int epilogue;                     /* svla_epilogue */
  int accumulate;                   /* C += result (C must be fp32) */
  float alpha;                      /* scales acc before bias */
  int impl;                         /* 0 auto, 1 SIMT fp32-FMA, 2 tcgen05 (bf16 operands) */
  float* colsum_a;                  /* transA only (weight gradients, A = dY stored [K, M]): colsum_a[m] += sum_k A[k, m],
                                       the bias gradient of the same nn.Linear; NULL = not wanted */
  const svla_dropout* dropout;      /* RELU_BITS only: C = dropout(relu(acc + bias)) -- the FFN dropout of the encoder
                                       layer fused into linear1's epilogue; the bit record then carries relu AND keep,
                                       so the backward is MASK_BITS with alpha = 1 / (1 - p).  NULL = none */
} svla_gemm_desc;

/* C = epi(alpha * op(A) op(B) + bias) [+ residual] -- every nn.Linear / 1x1 Conv2d forward, dgrad
 * and wgrad of the towers (allenact_dino_transformer.py:509-513,532-552; llama/model.py:203-222,355-357,437). */
int svla_gemm(svla_ctx* ctx, const svla_gemm_desc* d, svla_stream stream);
/* which kernel svla_gemm would run for this descriptor: 1 = fp32-FMA, 2 = tcgen05 */
int svla_gemm_which(const svla_gemm_desc* d);

/* column sums: out[n] (+)= sum_m x[m,n]  -- bias gradients. */
int svla_colsum(svla_ctx* ctx, const void* x, int dtype, long long M, int N, long long ldx, float* out,
                int accumulate, svla_stream stream);

/* rows are mapped as dst_row = (row / group) * group_stride + group_offset + row % group
 * (group <= 0: identity) so adapters can write straight into the [R, S, 512] fusion sequence. */
typedef struct { int group; int group_stride; int group_offset; } svla_rowmap;

/* y = [relu](LN(x [+ res]) * gamma + beta) [+ token]; saves mean/rstd (fp32 [rows]).  D = 512.
 * nn.LayerNorm eps 1e-5 (allenact_dino_transformer.py:511,541; fusion norm1/norm2 SURVEY A.7). */
int svla_layernorm_fwd(svla_ctx* ctx, const void* x, const void* res, int dtype_in, const float* gamma,
                       const float* beta, const float* token, int relu, float eps, void* y, int dtype_out,
                       svla_rowmap ymap, float* mean, float* rstd, long long rows, int D, svla_stream stream);
/* dx = LN backward of (dy [masked by y > token-shifted relu]); dgamma/dbeta/dtoken accumulate (fp32). */
int svla_layernorm_bwd(svla_ctx* ctx, const void* dy, int dtype_dy, svla_rowmap dymap, const void* x,
                       const void* res, int dtype_in, const float* gamma, const float* beta, int relu,
                       const float* mean, const float* rstd, void* dx, int dtype_dx, float* dgamma, float* dbeta,
                       float* dtoken, long long rows, int D, svla_stream stream);

/* RMSNorm (llama/model.py:57,70-71 eps 1e-5; T5LayerNorm eps 1e-6): y = x * rsqrt(mean(x^2)+eps) * w */
int svla_rmsnorm_fwd(svla_ctx* ctx, const void* x, int dtype_in, const float* w, float eps, void* y, int dtype_out,
                     float* rstd, long long rows, int D, svla_stream stream);
int svla_rmsnorm_bwd(svla_ctx* ctx, const void* dy, int dtype_dy, const void* x, int dtype_in, const float* w,
                     const float* rstd, void* dx, int dtype_dx, int accumulate_dx, float* dw, long long rows, int D,
                     svla_stream stream);

typedef enum {
  SVLA_ATTN_FULL = 0,       /* fusion block: no mask, scale 1/sqrt(dh) (nn.MultiheadAttention) */
  SVLA_ATTN_TRAJ_CAUSAL = 1,/* decoder: (traj[i]==traj[j]) && j<=i from traj_index, allenact_dino_transformer.py:399-402 */
  SVLA_ATTN_T5_BIAS = 2     /* T5: unscaled scores + relative-position bias[H,S,S] + key padding mask */
} svla_attn_mode;

/* Multi-head attention over packed projections.  q,k,v: row (b*S + s), head h at column h*dh
 * of buffers with leading dimension ld (elements); o [B*S, H*dh] (ld = ldo).  lse [B,H,S] fp32 is
 * saved for the backward.  traj: int64 [B,S] (mode 1); bias fp32 [H,S,S] and keymask int64 [B,S] (mode 2). */
int svla_attn_fwd(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld, void* o,
                  long long ldo, int dtype, float* lse, const int64_t* traj, const float* bias,
                  const int64_t* keymask, int B, int S, int H, int dh, float scale, svla_stream stream);
/* the same with dropout on the attention probabilities (nn.MultiheadAttention dropout, applied after the softmax
 * normalisation): tcgen05 kernels only (bf16, S <= 256, mode FULL).  Mask rows are (b * H + h) * W + query with
 * W = 128 for S <= 128 and W = 256 for 128 < S <= 256 (the two-camera fusion block), columns the keys.  The backward
 * of the S > 128 kernels takes delta from dO . O and therefore needs the forward output o (ignored for S <= 128). */
int svla_attn_drop_fwd(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld, void* o,
                       long long ldo, float* lse, const int64_t* traj, int B, int S, int H, int dh, float scale,
                       const svla_dropout* drop, svla_stream stream);
int svla_attn_drop_bwd(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld,
                       const void* o, const void* d_o, long long ldo, void* dq, void* dk, void* dv, long long ldd,
                       const float* lse, const int64_t* traj, int B, int S, int H, int dh, float scale,
                       const svla_dropout* drop, svla_stream stream);
int svla_attn_bwd(svla_ctx* ctx, int mode, const void* q, const void* k, const void* v, long long ld, const void* o,
                  const void* d_o, long long ldo, void* dq, void* dk, void* dv, long long ldd, int dtype,
                  const float* lse, const int64_t* traj, int B, int S, int H, int dh, float scale,
                  svla_stream stream);

/* test hook: 0 = auto (tcgen05 kernel for bf16, S <= 128; CUDA-core kernel otherwise), 1 = CUDA-core, 2 = tcgen05 */
int svla_set_attn_impl(int impl);

/* Single-query attention of the LAST fusion layer: only output token 0 of the fusion transformer is
 * consumed (allenact_dino_transformer.py:708), so its query/out-proj/FFN run for that row alone.
 * q [B, H*dh] (ldq), k/v rows b*S+s (ldkv), o [B, H*dh]; lse [B,H]. */
/* `drop` (may be NULL): dropout on the attention probabilities of the single query row, bf16 only; mask row
 * (b * H + h) * 128, i.e. the CLS row of the full-sequence mask the svla_attn_drop_* kernels use. */
int svla_attn_cls_fwd(svla_ctx* ctx, const void* q, long long ldq, const void* k, const void* v, long long ldkv,
                      void* o, long long ldo, int dtype, float* lse, int B, int S, int H, int dh, float scale,
                      const svla_dropout* drop, svla_stream stream);
int svla_attn_cls_bwd(svla_ctx* ctx, const void* q, long long ldq, const void* k, const void* v, long long ldkv,
                      const void* o, const void* d_o, long long ldo, void* dq, long long lddq, void* dk, void* dv,
                      long long lddkv, int dtype, const float* lse, int B, int S, int H, int dh, float scale,
                      const svla_dropout* drop, svla_stream stream);

/* Single-step decoder attention against the KV cache (rollout-side T = 1 inference; llama/model.py:224-247,279-317,
 * episode-start mask allenact_dino_transformer.py:386-397).  q [N, H*dh] (ldq); cache_k / cache_v [N, cache_rows,
 * H*dh] (row stride ldc) already holding this step's K / V at row `pos`; query n uses cached sequence
 * n / q_per_cache and attends to its rows [max(pos - time_step[n], 0), pos] (time_step NULL: rows [0, pos]);
 * o [N, H*dh].  q_per_cache = S with pos = S - 1 is plain unmasked attention over S-token sequences (the fp32
 * parity path of sequences longer than the shared-memory kernels take). */
int svla_attn_decode(svla_ctx* ctx, const void* q, long long ldq, const void* cache_k, const void* cache_v,
                     long long cache_rows, long long ldc, const int64_t* time_step, int pos, void* o, long long ldo,
                     int dtype, int N, int H, int dh, float scale, int q_per_cache, svla_stream stream);

/* SwiGLU gate (llama/model.py:360): g = silu(a) * b, a|b packed as [rows, 2*F] (w1 | w3 outputs). */
int svla_swiglu_fwd(svla_ctx* ctx, const void* ab, void* g, int dtype, long long rows, int F, svla_stream stream);
int svla_swiglu_bwd(svla_ctx* ctx, const void* ab, const void* dg, void* dab, int dtype, long long rows, int F,
                    svla_stream stream);

/* x[t,n,:] = obs_embed + E_a[masks!=0 ? prev_action : A] + E_h[in_hand] + sincos(time_step * div_term)
 * written in decoder order [N,T,512] (allenact_dino_transformer.py:353-385;
 * architecture/models/transformer_models/text_cond_visual_encoder.py:263-283). */
int svla_embed_time_fwd(svla_ctx* ctx, const void* obs_embed, int dtype_in, const int64_t* prev_actions,
                        const float* masks, const int64_t* in_hand, const int64_t* time_step, const float* E_a,
                        const float* E_h, const float* div_term, float* x_out, int T, int N, int A, int D,
                        svla_stream stream);
/* dx [N,T,512] -> d obs_embed [T,N,512] plus scatter-add into dE_a [(A+2),512], dE_h [3,512]. */
int svla_embed_time_bwd(svla_ctx* ctx, const float* dx, const int64_t* prev_actions, const float* masks,
                        const int64_t* in_hand, void* d_obs_embed, int dtype_out, float* dE_a, float* dE_h,
                        int T, int N, int A, int D, svla_stream stream);

/* [R, C_in, P] fp32 (NCHW with P = H*W sites) -> token-major [R, P, C_in] in `dtype_out`
 * (the 1x1-conv compressor then is a plain GEMM; allenact_dino_transformer.py:663-667). */
int svla_nchw_to_tokens(svla_ctx* ctx, const float* x, void* y, int dtype_out, long long R, int C, int P,
                        svla_stream stream);

/* generic row gather / scatter with dtype conversion: dst[dmap(i), :] = src[idx ? idx[i] : smap(i), :] */
int svla_copy_rows(svla_ctx* ctx, const void* src, int dtype_src, long long lds, svla_rowmap smap,
                   const int64_t* idx, void* dst, int dtype_dst, long long ldd, svla_rowmap dmap, long long rows,
                   int D, int accumulate, svla_stream stream);
/* broadcast one fp32 vector into mapped rows (fusion token): dst[dmap(i), :] = vec */
int svla_fill_rows(svla_ctx* ctx, const float* vec, void* dst, int dtype_dst, long long ldd, svla_rowmap dmap,
                   long long rows, int D, svla_stream stream);

/* ---- rollout-side vision preprocessor (DINOv2 ViT-S/14; dino_preprocessors.py:20-38,119-125,224-239) ----
 * uint8 frames [N, H, W, 3] -> normalised, cropped, im2col'd patches [N*PH*PW, Kpad] (bf16 or fp32), the A operand of
 * the patch-embedding GEMM: column c*patch^2 + dy*patch + dx = (pixel / 255 - mean[c]) / std[c]; mean3 / std3 are HOST
 * pointers to three floats. */
int svla_patchify_u8(svla_ctx* ctx, const uint8_t* img, int N, int H, int W, int patch, int crop_left, int crop_right,
                     const float* mean3, const float* std3, void* out, int dtype, int Kpad, svla_stream stream);
/* x[n, 0] = cls + pos[0]; x[n, 1 + p] = patches[n*num_patches + p] + pos[1 + p]  (cls, pos fp32) */
int svla_vit_assemble(svla_ctx* ctx, const void* patches, const float* cls, const float* pos, void* x, int dtype, int N,
                      int num_patches, int D, svla_stream stream);
/* AdaptiveAvgPool2d((OH, OW)) over the patch tokens of x [N, 1 + PH*PW, D] -> out fp32 [N, D, OH, OW] */
int svla_tokens_pool(svla_ctx* ctx, const void* x, int dtype, float* out, int N, int PH, int PW, int D, int OH, int OW,
                     svla_stream stream);

/* x *= *scale_dev (upstream gradient of the scalar loss applied to the fused kernel's gradients) */
int svla_scale_by(svla_ctx* ctx, float* x, long long n, const float* scale_dev, svla_stream stream);

/* fp32 -> bf16 cast of a flat buffer (parameter shadow) */
int svla_cast_bf16(svla_ctx* ctx, const float* x, void* y, long long n, svla_stream stream);

/* Attention of the parity-grade tensor-core mode (S <= 128, head dim 64, modes FULL / TRAJ_CAUSAL): q / k / v (and
 * d_o) are (hi, lo) bf16 pairs of the fp32 tensors as svla_split_concat(axis 1, pattern {0, 1}) writes them -- the
 * `_hi` pointer addresses head 0 of the hi half, the lo half starts `lo_off` elements further on the same row (row
 * stride ld) -- every product runs as hi*hi + lo*hi + hi*lo on the tcgen05 kernels, P / dS are split in registers;
 * outputs (o, dq, dk, dv) and lse are fp32.  Same semantics as svla_attn_fwd / svla_attn_bwd otherwise. */
int svla_attn_split_fwd(svla_ctx* ctx, int mode, const void* q_hi, const void* k_hi, const void* v_hi, long long lo_off,
                        long long ld, float* o, long long ldo, float* lse, const int64_t* traj, int B, int S, int H,
                        int dh, float scale, svla_stream stream);
int svla_attn_split_bwd(svla_ctx* ctx, int mode, const void* q_hi, const void* k_hi, const void* v_hi, long long lo_off,
                        long long ld, const void* do_hi, long long do_lo_off, long long lddo, float* dq, float* dk,
                        float* dv, long long ldd, const float* lse, const int64_t* traj, int B, int S, int H, int dh,
                        float scale, svla_stream stream);

/* Split-operand staging of the parity-grade tensor-core mode (precision "bf16x3" / "bf16x6"): x fp32 [rows, cols]
 * (row stride ldx) is decomposed into bf16 parts p0 = bf16(x), p1 = bf16(x - p0), p2 = bf16(x - p0 - p1) and the parts
 * named by pattern[0..nprod) are concatenated along the contraction dimension of the GEMM operand the tensor will be:
 *   axis = 1: out[r, j * cols + c] = part_{pattern[j]}(x[r, c])      (K-major operand, out [rows, nprod * cols])
 *   axis = 0: out[j * rows + r, c] = part_{pattern[j]}(x[r, c])      (MN-major operand, out [nprod * rows, cols])
 * With A staged by (0,1,0) and B by (0,0,1) ONE svla_gemm launch on the bf16 tcgen05 kernels accumulates
 * p0 q0 + p1 q0 + p0 q1 in fp32 -- the fp32 product to ~2^-16 relative (six products with three parts: ~2^-23).
 * Replaces nothing in the reference (which multiplies in fp32 on the CPU/GPU library path): it is how BASELINE's
 * "within 1e-4 rel fp32" gate is met on the tensor cores. */
int svla_split_concat(svla_ctx* ctx, const float* x, long long ldx, long long rows, int cols, void* out, long long ldo,
                      int axis, int nprod, const int* pattern, svla_stream stream);

/* 64-bit hash of each goal-byte row (uint8 [R, L]) for prompt de-duplication; replaces the per-row
 * CPU decode loop at allenact_dino_transformer.py:591-598. */
int svla_hash_rows(svla_ctx* ctx, const uint8_t* rows, long long R, int L, uint64_t* out, svla_stream stream);

/* Storage-side episode-cost bookkeeping for one environment step (the Jc = mean finished-episode cost that the
 * Lagrange update consumes; the reference's engine derives it from the task metrics of SafeRLStepResult,
 * tasks/abstract_task.py:333,369-380): episode_cost[n] += costs[n]; every sampler whose episode ended at this step
 * (mask_next[n] == 0) adds its episode total to sum_cnt[0], increments sum_cnt[1] and restarts from zero.
 * Deterministic (one block, fixed order). */
int svla_episode_cost_step(svla_ctx* ctx, const float* costs, const float* mask_next, float* episode_cost,
                           float* sum_cnt, int N, svla_stream stream);

/* K cost channels (extension: the reference has one scalar cost, tasks/abstract_task.py:333; BASELINE config 5 asks
 * for two).  The constraint term of SafePPOLogGrad (customized_loss.py:350-359) generalises to
 *   (A - sum_k lambda_k A_c,k) / (1 + sum_k lambda_k),
 * which equals the single-channel form with lambda_eff = sum_k lambda_k and c_adv_eff = sum_k lambda_k A_c,k /
 * lambda_eff (0 when lambda_eff = 0): this entry point produces that pair on the device (multipliers are read on the
 * device) so svla_ppo_lag_fwd_bwd is used unchanged.  c_adv: channel-major [K, R] fp32, 1 <= K <= 8. */
int svla_combine_cost_advantages(svla_ctx* ctx, const float* c_adv, const float* lambdas_dev, int K, long long R,
                                 float* c_adv_eff, float* lambda_eff_dev, svla_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* SAFEVLA_B200_H */
