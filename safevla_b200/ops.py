"""Tensor-level wrappers over the C ABI: each function validates shapes, passes raw device
pointers + the current CUDA stream to libsafevla_b200 and returns/filles torch tensors.
No function here computes anything with PyTorch ops -- torch only owns the memory."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L
from ._lib import IDENT, check, dt, get_ctx, load_library, ptr, stream_ptr


def _lib():
    return load_library()


def _cuda(*ts):
    for t in ts:
        if t is not None:
            assert t.is_cuda, "safevla_b200 ops take CUDA tensors only (no CPU fallback)"
            assert t.is_contiguous(), "non-contiguous tensor passed to a safevla_b200 op"


# ---- scan / elementwise path -------------------------------------------------------------------
def gae_dual(rewards, costs, value_preds, c_value_preds, masks, gamma: float, lam: float, algo: int = 0,
             out=None):
    """rewards/costs [T,N(,1)], value_preds/c_value_preds/masks [T+1,N(,1)] fp32.
    Returns (returns, c_returns [T+1,...], adv, c_adv [T,...])."""
    _cuda(rewards, costs, value_preds, c_value_preds, masks)
    T = rewards.shape[0]
    N = rewards[0].numel()
    assert value_preds.shape[0] == T + 1 and masks.shape[0] == T + 1
    if out is None:
        returns, adv = torch.empty_like(value_preds), torch.empty_like(rewards)
        c_returns = torch.empty_like(c_value_preds) if costs is not None else None
        c_adv = torch.empty_like(costs) if costs is not None else None
    else:
        returns, c_returns, adv, c_adv = out
    check(_lib().svla_gae_dual(get_ctx(), ptr(rewards), ptr(costs), ptr(value_preds), ptr(c_value_preds), ptr(masks),
                               ptr(returns), ptr(c_returns), ptr(adv), ptr(c_adv), T, N, float(gamma), float(lam),
                               algo, stream_ptr()), "svla_gae_dual")
    return returns, c_returns, adv, c_adv


def discounted_returns_dual(rewards, costs, value_preds, c_value_preds, masks, gamma: float, out=None):
    """`use_gae=False`: ret_t = ret_{t+1} * gamma * m_{t+1} + r_t on both streams; same shapes / outputs as gae_dual."""
    _cuda(rewards, costs, value_preds, c_value_preds, masks)
    T = rewards.shape[0]
    N = rewards[0].numel()
    assert value_preds.shape[0] == T + 1 and masks.shape[0] == T + 1
    if out is None:
        returns, adv = torch.empty_like(value_preds), torch.empty_like(rewards)
        c_returns = torch.empty_like(c_value_preds) if costs is not None else None
        c_adv = torch.empty_like(costs) if costs is not None else None
    else:
        returns, c_returns, adv, c_adv = out
    check(_lib().svla_discounted_returns_dual(get_ctx(), ptr(rewards), ptr(costs), ptr(value_preds), ptr(c_value_preds),
                                              ptr(masks), ptr(returns), ptr(c_returns), ptr(adv), ptr(c_adv), T, N,
                                              float(gamma), stream_ptr()), "svla_discounted_returns_dual")
    return returns, c_returns, adv, c_adv


def normalize_advantage(adv: torch.Tensor):
    _cuda(adv)
    out = torch.empty_like(adv)
    stats = torch.empty(2, device=adv.device, dtype=torch.float32)
    check(_lib().svla_normalize_advantage(get_ctx(), ptr(adv), ptr(out), ptr(stats), adv.numel(), stream_ptr()),
          "svla_normalize_advantage")
    return out, stats


def advantage_sums(adv: torch.Tensor) -> torch.Tensor:
    """[sum adv, sum adv^2, n] of the local shard (fp32 device tensor) -- all-reduce it, then normalise with
    `normalize_advantage_from_sums` so every rank uses the global statistics."""
    _cuda(adv)
    sums = torch.empty(3, device=adv.device, dtype=torch.float32)
    check(_lib().svla_advantage_sums(get_ctx(), ptr(adv), adv.numel(), ptr(sums), stream_ptr()), "svla_advantage_sums")
    return sums


def normalize_advantage_from_sums(adv: torch.Tensor, sums: torch.Tensor):
    _cuda(adv, sums)
    out = torch.empty_like(adv)
    stats = torch.empty(2, device=adv.device, dtype=torch.float32)
    check(_lib().svla_normalize_advantage_from_sums(get_ctx(), ptr(adv), ptr(out), ptr(sums), ptr(stats), adv.numel(),
                                                    stream_ptr()), "svla_normalize_advantage_from_sums")
    return out, stats


def ppo_lag_fwd_bwd(logits, actions, old_logp, adv, c_adv, values, returns, c_values, c_returns, lambda_dev,
                    hp: L.PpoHparams, old_values=None, old_c_values=None, want_grads: bool = True):
    """Returns (scalars[16] device tensor, dlogits, dvalues, dcvalues)."""
    _cuda(logits, actions, old_logp, adv, c_adv, values, returns, c_values, c_returns, lambda_dev)
    some = logits if logits is not None else (values if values is not None else c_values)
    dev = some.device
    if logits is not None:
        A = logits.shape[-1]
        R = logits.numel() // A
        assert actions.dtype == torch.int64 and actions.numel() == R
    else:
        A, R = 1, some.numel()
    scal = torch.empty(L.PPO_NSCALARS, device=dev, dtype=torch.float32)
    dlogits = torch.empty_like(logits) if (want_grads and logits is not None) else None
    dvalues = torch.empty_like(values) if (want_grads and values is not None) else None
    dcvalues = torch.empty_like(c_values) if (want_grads and c_values is not None) else None
    check(_lib().svla_ppo_lag_fwd_bwd(get_ctx(), ptr(logits), ptr(actions), ptr(old_logp), ptr(adv), ptr(c_adv),
                                      ptr(values), ptr(returns), ptr(old_values), ptr(c_values), ptr(c_returns),
                                      ptr(old_c_values), ptr(lambda_dev), C.byref(hp), ptr(scal), ptr(dlogits),
                                      ptr(dvalues), ptr(dcvalues), R, A, stream_ptr()), "svla_ppo_lag_fwd_bwd")
    return scal, dlogits, dvalues, dcvalues


def lagrange_update(lambda_dev, state_dev, cost_sum_cnt, cost_limit: float, lr: float, upper_bound: float = -1.0):
    _cuda(lambda_dev, state_dev, cost_sum_cnt)
    check(_lib().svla_lagrange_update(get_ctx(), ptr(lambda_dev), ptr(state_dev), ptr(cost_sum_cnt), cost_limit, lr,
                                      upper_bound, stream_ptr()), "svla_lagrange_update")


def episode_cost_step(costs, mask_next, episode_cost, sum_cnt):
    """Per-step Jc bookkeeping: costs / mask_next / episode_cost fp32 [N], sum_cnt fp32 [2] (updated in place)."""
    _cuda(costs, mask_next, episode_cost, sum_cnt)
    check(_lib().svla_episode_cost_step(get_ctx(), ptr(costs), ptr(mask_next), ptr(episode_cost), ptr(sum_cnt),
                                        episode_cost.numel(), stream_ptr()), "svla_episode_cost_step")


def combine_cost_advantages(c_adv: torch.Tensor, lambdas: torch.Tensor):
    """c_adv fp32 [K, ...] (channel-major), lambdas fp32 [K] (device) -> (c_adv_eff [...], lambda_eff [1]) such that
    (A - sum_k l_k A_k) / (1 + sum_k l_k) == (A - l_eff A_eff) / (1 + l_eff)."""
    _cuda(c_adv, lambdas)
    K = c_adv.shape[0]
    assert lambdas.numel() == K and c_adv.dtype == torch.float32 and lambdas.dtype == torch.float32
    out = torch.empty(c_adv.shape[1:], device=c_adv.device, dtype=torch.float32)
    lam = torch.empty(1, device=c_adv.device, dtype=torch.float32)
    check(_lib().svla_combine_cost_advantages(get_ctx(), ptr(c_adv), ptr(lambdas), K, out.numel(), ptr(out), ptr(lam),
                                              stream_ptr()), "svla_combine_cost_advantages")
    return out, lam


def sq_norm(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _cuda(x)
    assert x.dtype == torch.float32
    if out is None:
        out = torch.empty(1, device=x.device, dtype=torch.float32)
    check(_lib().svla_sq_norm(get_ctx(), ptr(x), x.numel(), ptr(out), stream_ptr()), "svla_sq_norm")
    return out


def clip_adam(p, g, m, v, p_bf16, sq_norm_dev, hp: L.AdamHparams):
    _cuda(p, g, m, v, p_bf16, sq_norm_dev)
    check(_lib().svla_clip_adam(get_ctx(), ptr(p), ptr(g), ptr(m), ptr(v), ptr(p_bf16), p.numel(), ptr(sq_norm_dev),
                                C.byref(hp), stream_ptr()), "svla_clip_adam")


def scale_by(x: torch.Tensor, scale_dev: torch.Tensor):
    _cuda(x, scale_dev)
    assert x.dtype == torch.float32 and scale_dev.dtype == torch.float32
    check(_lib().svla_scale_by(get_ctx(), ptr(x), x.numel(), ptr(scale_dev), stream_ptr()), "svla_scale_by")
    return x


def cast_bf16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _cuda(x)
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    check(_lib().svla_cast_bf16(get_ctx(), ptr(x), ptr(out), x.numel(), stream_ptr()), "svla_cast_bf16")
    return out


# ---- dropout -----------------------------------------------------------------------------------
def dropout_spec(p: float, seed: int, site: int, step: int, row0: int = 0, row_stride: int = 1) -> L.Dropout:
    """svla_dropout: the mask of a site is a pure function of (seed, step, site, row0 + row * row_stride, column)."""
    return L.Dropout(float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, int(site) & 0xFFFFFFFF, int(step) & 0xFFFFFFFF,
                     int(row0) & 0xFFFFFFFF, int(row_stride) & 0xFFFFFFFF)


def dropout_rows(x: torch.Tensor, out: torch.Tensor, drop: L.Dropout, rows=None):
    """out = keep ? x / (1 - p) : 0 over the 2-D views x / out (row strides respected; in place allowed)."""
    assert x.dim() == 2 and out.dim() == 2 and x.stride(1) == 1 and out.stride(1) == 1 and x.shape[1] == out.shape[1]
    rows = x.shape[0] if rows is None else rows
    check(_lib().svla_dropout_rows(get_ctx(), ptr(x), dt(x), x.stride(0), ptr(out), dt(out), out.stride(0), rows,
                                   x.shape[1], C.byref(drop), stream_ptr()), "svla_dropout_rows")
    return out


# ---- dense path --------------------------------------------------------------------------------
# split-operand products (svla_split_concat): (parts of A, parts of B) per product, most significant first
SPLIT_PATTERNS = {3: ((0, 1, 0), (0, 0, 1)), 6: ((0, 0, 1, 1, 0, 2), (0, 1, 0, 1, 2, 0))}
_SPLIT_CACHE = None  # {key: staged operand} while a weight epoch is open (see split_cache_open)


def split_concat(x: torch.Tensor, ldx: int, rows: int, cols: int, axis: int, pattern) -> torch.Tensor:
    """fp32 [rows, cols] view (row stride ldx) -> bf16 operand with the parts `pattern` concatenated along the
    contraction dimension: [rows, P * cols] (axis 1) or [P * rows, cols] (axis 0)."""
    P = len(pattern)
    out = torch.empty((rows, P * cols) if axis == 1 else (P * rows, cols), device=x.device, dtype=torch.bfloat16)
    pat = (C.c_int * P)(*pattern)
    check(_lib().svla_split_concat(get_ctx(), ptr(x), ldx, rows, cols, ptr(out), out.stride(0), axis, P, pat,
                                   stream_ptr()), "svla_split_concat")
    return out


def split_cache_open():
    """Staged copies of tensors flagged `cache_b` (weights) are reused until split_cache_clear(): a weight is split
    once per optimizer step and operand layout instead of once per launch."""
    global _SPLIT_CACHE
    if _SPLIT_CACHE is None:
        _SPLIT_CACHE = {}


def split_cache_clear():
    if _SPLIT_CACHE is not None:
        _SPLIT_CACHE.clear()


def _split_ok(M, N, K, P, trans_a, trans_b, lda, ldb):
    """Shapes the tcgen05 kernels take (gemm_tc.cu: svla_gemm_tc_supported) once the operands are staged."""
    if M < 64 or N < 64 or N % 64 or K < 64 or K % 4:
        return False
    if (trans_a and M % 8) or (not trans_a and (P * K) % 8) or (not trans_b and N % 8):
        return False
    return lda % 4 == 0 and ldb % 4 == 0


def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, trans_a=False, trans_b=True, bias=None,
         residual=None, aux=None, epilogue=L.EPI_NONE, accumulate=False, alpha=1.0, impl=0,
         M=None, N=None, K=None, lda=None, ldb=None, ldc=None, colsum_a=None, split=0, cache_b=False, dropout=None):
    """out[M,N] = epi(alpha * op(a) op(b) + bias) [+ residual].  a/b/out are 2-D views whose last
    dim is contiguous (row stride = leading dimension).  trans_b=True is the nn.Linear layout.
    split = 3 / 6 (fp32 operands only): the product runs on the bf16 tcgen05 kernels as a sum of 3 / 6 split-operand
    products in one launch (parity-grade tensor-core mode); shapes the tensor-core kernels do not take stay on the
    fp32 FMA kernel."""
    for t in (a, b, out, residual, aux):
        if t is not None:
            assert t.is_cuda and t.stride(-1) == 1 and t.dim() == 2
    lda = a.stride(0) if lda is None else lda
    ldb = b.stride(0) if ldb is None else ldb
    ldc = out.stride(0) if ldc is None else ldc
    if M is None:
        M = a.shape[1] if trans_a else a.shape[0]
    if K is None:
        K = a.shape[0] if trans_a else a.shape[1]
    if N is None:
        N = b.shape[0] if trans_b else b.shape[1]
    kb = b.shape[1] if trans_b else b.shape[0]
    assert kb == K, f"inner dims differ: {K} vs {kb}"
    assert out.shape[0] >= M and out.shape[1] >= N
    if split and a.dtype == torch.float32 and b.dtype == torch.float32 and not (trans_a and trans_b) \
            and _split_ok(M, N, K, len(SPLIT_PATTERNS[split][0]), trans_a, trans_b, lda, ldb):
        pa, pb = SPLIT_PATTERNS[split]
        a2 = split_concat(a, lda, K, M, 0, pa) if trans_a else split_concat(a, lda, M, K, 1, pa)
        key = (b.data_ptr(), ldb, N, K, trans_b, split)
        b2 = _SPLIT_CACHE.get(key) if (cache_b and _SPLIT_CACHE is not None) else None
        if b2 is None:
            b2 = split_concat(b, ldb, N, K, 1, pb) if trans_b else split_concat(b, ldb, K, N, 0, pb)
            if cache_b and _SPLIT_CACHE is not None:
                _SPLIT_CACHE[key] = b2
        if colsum_a is not None:  # the staged operand holds every part more than once: sum the fp32 tensor itself
            assert trans_a
            colsum(a[:K, :M], colsum_a, accumulate=True)
        return gemm(a2, b2, out, trans_a=trans_a, trans_b=trans_b, bias=bias, residual=residual, aux=aux,
                    epilogue=epilogue, accumulate=accumulate, alpha=alpha, impl=impl, M=M, N=N, ldc=ldc)
    d = L.GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.A, d.lda, d.transA = ptr(a), lda, int(trans_a)
    d.B, d.ldb, d.transB = ptr(b), ldb, int(trans_b)
    d.C, d.ldc = ptr(out), ldc
    d.dtypeA, d.dtypeB, d.dtypeC = dt(a), dt(b), dt(out)
    d.bias = ptr(bias)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() >= N
    if residual is not None:
        d.residual, d.ldr, d.dtypeR = ptr(residual), residual.stride(0), dt(residual)
    if aux is not None and epilogue in (L.EPI_RELU_BITS, L.EPI_MASK_BITS):
        assert aux.dtype == torch.int32 and aux.shape[0] >= M and aux.shape[1] * 32 >= N  # one bit per element
        d.aux, d.ldaux, d.dtypeAux = ptr(aux), aux.stride(0), dt(out)
    elif aux is not None:
        d.aux, d.ldaux, d.dtypeAux = ptr(aux), aux.stride(0), dt(aux)
    d.epilogue, d.accumulate, d.alpha, d.impl = epilogue, int(accumulate), alpha, impl
    if colsum_a is not None:  # bias gradient of the same linear, accumulated (trans_a launches only)
        assert trans_a and colsum_a.dtype == torch.float32 and colsum_a.numel() >= M
        d.colsum_a = ptr(colsum_a)
    if dropout is not None:  # RELU_BITS only: dropout fused behind the ReLU
        assert epilogue == L.EPI_RELU_BITS
        d.dropout = C.pointer(dropout)
    if PROFILE is None:
        check(_lib().svla_gemm(get_ctx(), C.byref(d), stream_ptr()), "svla_gemm")
        return out
    # bench.py: CUDA-event timing of every GEMM launch on the launching stream
    which = "svla_gemm_tc_kernel" if _lib().svla_gemm_which(C.byref(d)) == 2 else "gemm_simt_kernel"
    _timed(which, 2.0 * M * N * K, (M, N, K, int(trans_a), int(trans_b), epilogue, int(accumulate), str(a.dtype)[6:],
                                    str(out.dtype)[6:], residual is not None, colsum_a is not None),
           lambda: check(_lib().svla_gemm(get_ctx(), C.byref(d), stream_ptr()), "svla_gemm"))
    return out


PROFILE = None  # set to {} by bench.py to collect per-launch GEMM timings


def profile_summary(prof):
    """{kernel: {flops, ms, n}} from the recorded events (call after a device synchronize)."""
    out = {}
    for k, recs in (prof or {}).items():
        out[k] = {"flops": sum(r[0] for r in recs), "ms": sum(r[1].elapsed_time(r[2]) for r in recs), "n": len(recs)}
    return out


def colsum(x: torch.Tensor, out: torch.Tensor, accumulate: bool = True):
    assert x.dim() == 2 and x.stride(1) == 1
    check(_lib().svla_colsum(get_ctx(), ptr(x), dt(x), x.shape[0], x.shape[1], x.stride(0), ptr(out),
                             int(accumulate), stream_ptr()), "svla_colsum")


def layernorm_fwd(x, gamma, beta, y, *, res=None, token=None, relu=False, eps=1e-5, ymap=IDENT, mean=None,
                  rstd=None, rows=None):
    D = x.shape[-1]
    rows = x.numel() // D if rows is None else rows
    check(_lib().svla_layernorm_fwd(get_ctx(), ptr(x), ptr(res), dt(x), ptr(gamma), ptr(beta), ptr(token), int(relu),
                                    eps, ptr(y), dt(y), ymap, ptr(mean), ptr(rstd), rows, D, stream_ptr()),
          "svla_layernorm_fwd")
    return y


def layernorm_bwd(dy, x, gamma, beta, mean, rstd, dx, dgamma, dbeta, *, res=None, relu=False, dymap=IDENT,
                  dtoken=None, rows=None):
    D = x.shape[-1]
    rows = x.numel() // D if rows is None else rows
    check(_lib().svla_layernorm_bwd(get_ctx(), ptr(dy), dt(dy), dymap, ptr(x), ptr(res), dt(x), ptr(gamma), ptr(beta),
                                    int(relu), ptr(mean), ptr(rstd), ptr(dx), dt(dx), ptr(dgamma), ptr(dbeta),
                                    ptr(dtoken), rows, D, stream_ptr()), "svla_layernorm_bwd")
    return dx


def rmsnorm_fwd(x, w, y, eps, rstd=None):
    D = x.shape[-1]
    check(_lib().svla_rmsnorm_fwd(get_ctx(), ptr(x), dt(x), ptr(w), eps, ptr(y), dt(y), ptr(rstd), x.numel() // D, D,
                                  stream_ptr()), "svla_rmsnorm_fwd")
    return y


def rmsnorm_bwd(dy, x, w, rstd, dx, dw, accumulate_dx=False):
    D = x.shape[-1]
    check(_lib().svla_rmsnorm_bwd(get_ctx(), ptr(dy), dt(dy), ptr(x), dt(x), ptr(w), ptr(rstd), ptr(dx), dt(dx),
                                  int(accumulate_dx), ptr(dw), x.numel() // D, D, stream_ptr()), "svla_rmsnorm_bwd")
    return dx


def _timed(which: str, flops: float, meta, launch):
    """bench.py: CUDA events on the launching stream around one C-ABI call + the FLOPs it executes."""
    if PROFILE is None:
        launch()
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launch()
    e1.record()
    PROFILE.setdefault(which, []).append((flops, e0, e1, meta))


def _split_qkv(q, k, v, H, dh):
    """(hi, lo) bf16 staging of fp32 q / k / v column views for svla_attn_split_*: returns (buffer, q_hi, k_hi, v_hi
    views' element offsets, lo_off, ld).  One staging launch when the three views are adjacent columns of one buffer."""
    D, rows, ld = H * dh, q.shape[0], q.stride(0)
    if k.data_ptr() == q.data_ptr() + 4 * D and v.data_ptr() == q.data_ptr() + 8 * D:
        buf = split_concat(q, ld, rows, 3 * D, 1, (0, 1))  # [rows, 6D]: hi q|k|v, lo q|k|v
        return buf, (0, D, 2 * D), 3 * D, 6 * D
    buf = torch.empty(rows, 6 * D, device=q.device, dtype=torch.bfloat16)
    pat = (C.c_int * 2)(0, 1)
    for j, x in enumerate((q, k, v)):  # [hi | lo] pairs side by side: lo_off = D
        check(_lib().svla_split_concat(get_ctx(), ptr(x), x.stride(0), rows, D, buf.data_ptr() + 2 * (2 * j * D),
                                       6 * D, 1, 2, pat, stream_ptr()), "svla_split_concat")
    return buf, (0, 2 * D, 4 * D), D, 6 * D


def _use_split_attn(split, q, mode, S, dh):
    return bool(split) and q.dtype == torch.float32 and S <= 128 and dh == 64 and mode in (L.ATTN_FULL, L.ATTN_TRAJ_CAUSAL)


def attn_fwd(mode, q, k, v, o, lse, B, S, H=8, dh=64, scale=0.125, traj=None, bias=None, keymask=None, split=0,
             drop=None):
    """q/k/v: 2-D views [B*S, >=H*dh] sharing one row stride (e.g. column slices of a packed qkv buffer).
    split != 0 with fp32 tensors (parity-grade tensor-core mode): the products run as split-bf16 sums on the tcgen05
    kernel (svla_attn_split_fwd) instead of the fp32 CUDA-core kernel."""
    assert q.stride(0) == k.stride(0) == v.stride(0) and q.dtype == k.dtype == v.dtype == o.dtype
    if drop is not None:  # dropout on the attention probabilities (warp-specialised bf16 kernels only)
        _timed("attn_fwd", 4.0 * B * H * S * S * dh, (mode, B, S, "bf16+drop"), lambda: check(
            _lib().svla_attn_drop_fwd(get_ctx(), mode, ptr(q), ptr(k), ptr(v), q.stride(0), ptr(o), o.stride(0), ptr(lse),
                                      ptr(traj), B, S, H, dh, scale, C.byref(drop), stream_ptr()), "svla_attn_drop_fwd"))
        return o
    if _use_split_attn(split, q, mode, S, dh):
        buf, (qo, ko, vo), lo_off, ld = _split_qkv(q, k, v, H, dh)
        base = buf.data_ptr()
        _timed("attn_fwd", 3 * 4.0 * B * H * S * S * dh, (mode, B, S, "f32x3"), lambda: check(
            _lib().svla_attn_split_fwd(get_ctx(), mode, base + 2 * qo, base + 2 * ko, base + 2 * vo, lo_off, ld, ptr(o),
                                       o.stride(0), ptr(lse), ptr(traj), B, S, H, dh, scale, stream_ptr()),
            "svla_attn_split_fwd"))
        return o
    _timed("attn_fwd", 4.0 * B * H * S * S * dh, (mode, B, S, str(q.dtype)[6:]), lambda: check(
        _lib().svla_attn_fwd(get_ctx(), mode, ptr(q), ptr(k), ptr(v), q.stride(0), ptr(o), o.stride(0), dt(q),
                             ptr(lse), ptr(traj), ptr(bias), ptr(keymask), B, S, H, dh, scale, stream_ptr()),
        "svla_attn_fwd"))
    return o


def attn_bwd(mode, q, k, v, o, d_o, dq, dk, dv, lse, B, S, H=8, dh=64, scale=0.125, traj=None, split=0, drop=None):
    assert q.stride(0) == k.stride(0) == v.stride(0) and dq.stride(0) == dk.stride(0) == dv.stride(0)
    assert o.stride(0) == d_o.stride(0)
    if drop is not None:
        _timed("attn_bwd", 10.0 * B * H * S * S * dh, (mode, B, S, "bf16+drop"), lambda: check(
            _lib().svla_attn_drop_bwd(get_ctx(), mode, ptr(q), ptr(k), ptr(v), q.stride(0), ptr(o), ptr(d_o), d_o.stride(0),
                                      ptr(dq), ptr(dk), ptr(dv), dq.stride(0), ptr(lse), ptr(traj), B, S, H, dh, scale,
                                      C.byref(drop), stream_ptr()), "svla_attn_drop_bwd"))
        return
    if _use_split_attn(split, q, mode, S, dh):
        buf, (qo, ko, vo), lo_off, ld = _split_qkv(q, k, v, H, dh)
        D = H * dh
        dbuf = split_concat(d_o, d_o.stride(0), d_o.shape[0], D, 1, (0, 1))  # [rows, 2D]: hi | lo
        base = buf.data_ptr()
        _timed("attn_bwd", 3 * 10.0 * B * H * S * S * dh, (mode, B, S, "f32x3"), lambda: check(
            _lib().svla_attn_split_bwd(get_ctx(), mode, base + 2 * qo, base + 2 * ko, base + 2 * vo, lo_off, ld,
                                       ptr(dbuf), D, 2 * D, ptr(dq), ptr(dk), ptr(dv), dq.stride(0), ptr(lse), ptr(traj),
                                       B, S, H, dh, scale, stream_ptr()), "svla_attn_split_bwd"))
        return
    # five S x S x dh products: the score recompute, dP, dV, dK, dQ
    _timed("attn_bwd", 10.0 * B * H * S * S * dh, (mode, B, S, str(q.dtype)[6:]), lambda: check(
        _lib().svla_attn_bwd(get_ctx(), mode, ptr(q), ptr(k), ptr(v), q.stride(0), ptr(o), ptr(d_o), o.stride(0),
                             ptr(dq), ptr(dk), ptr(dv), dq.stride(0), dt(q), ptr(lse), ptr(traj), B, S, H, dh,
                             scale, stream_ptr()), "svla_attn_bwd"))


def attn_cls_fwd(q0, k, v, o, lse, B, S, H=8, dh=64, scale=0.125, drop=None):
    assert k.stride(0) == v.stride(0)
    dp = C.cast(C.pointer(drop), C.c_void_p) if drop is not None else None
    check(_lib().svla_attn_cls_fwd(get_ctx(), ptr(q0), q0.stride(0), ptr(k), ptr(v), k.stride(0), ptr(o), o.stride(0),
                                   dt(q0), ptr(lse), B, S, H, dh, scale, dp, stream_ptr()), "svla_attn_cls_fwd")
    return o


def attn_cls_bwd(q0, k, v, o, d_o, dq, dk, dv, lse, B, S, H=8, dh=64, scale=0.125, drop=None):
    assert k.stride(0) == v.stride(0) and dk.stride(0) == dv.stride(0) and o.stride(0) == d_o.stride(0)
    dp = C.cast(C.pointer(drop), C.c_void_p) if drop is not None else None
    check(_lib().svla_attn_cls_bwd(get_ctx(), ptr(q0), q0.stride(0), ptr(k), ptr(v), k.stride(0), ptr(o), ptr(d_o),
                                   o.stride(0), ptr(dq), dq.stride(0), ptr(dk), ptr(dv), dk.stride(0), dt(q0),
                                   ptr(lse), B, S, H, dh, scale, dp, stream_ptr()), "svla_attn_cls_bwd")


def patchify_u8(img, out, patch, crop_left, crop_right, mean, std):
    """img uint8 [N, H, W, 3] (device) -> out [N*PH*PW, Kpad]; mean / std: 3 python floats."""
    N, H, W, _ = img.shape
    assert img.dtype == torch.uint8 and img.is_contiguous() and out.is_contiguous()
    m3, s3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    check(_lib().svla_patchify_u8(get_ctx(), ptr(img), N, H, W, patch, crop_left, crop_right, m3, s3, ptr(out), dt(out),
                                  out.shape[1], stream_ptr()), "svla_patchify_u8")
    return out


def vit_assemble(patches, cls, pos, x, N, num_patches):
    D = patches.shape[-1]
    check(_lib().svla_vit_assemble(get_ctx(), ptr(patches), ptr(cls), ptr(pos), ptr(x), dt(x), N, num_patches, D,
                                   stream_ptr()), "svla_vit_assemble")
    return x


def tokens_pool(x, out, N, PH, PW, OH, OW):
    D = x.shape[-1]
    check(_lib().svla_tokens_pool(get_ctx(), ptr(x), dt(x), ptr(out), N, PH, PW, D, OH, OW, stream_ptr()),
          "svla_tokens_pool")
    return out


def hl_gauss_fwd_bwd(logits, target, support, sigma, grad_scale=1.0, want_grad=True, want_values=False):
    """logits fp32 [R, B]; target fp32 [R]; support fp32 [B + 1] -> (loss [1], dlogits or None, values or None)."""
    R, B = logits.shape
    assert logits.dtype == torch.float32 and logits.is_contiguous() and support.numel() == B + 1
    loss = torch.empty(1, device=logits.device)
    dl = torch.empty_like(logits) if want_grad else None
    vals = torch.empty(R, device=logits.device) if want_values else None
    check(_lib().svla_hl_gauss_fwd_bwd(get_ctx(), ptr(logits), logits.stride(0), ptr(target.contiguous()), ptr(support), B,
                                       float(sigma), float(grad_scale), ptr(loss), ptr(dl), ptr(vals), R, stream_ptr()),
          "svla_hl_gauss_fwd_bwd")
    return loss, dl, vals


def attn_decode(q, cache_k, cache_v, time_step, pos, o, H=8, dh=64, scale=0.125):
    """q [N, H*dh]; cache_k / cache_v [N, rows, H*dh] holding this step's K / V at row `pos`; time_step int64 [N]."""
    N = q.shape[0]
    assert cache_k.dim() == 3 and cache_k.stride() == cache_v.stride() and cache_k.stride(2) == 1
    assert cache_k.stride(0) == cache_k.shape[1] * cache_k.stride(1)
    check(_lib().svla_attn_decode(get_ctx(), ptr(q), q.stride(0), ptr(cache_k), ptr(cache_v), cache_k.shape[1],
                                  cache_k.stride(1), ptr(time_step), int(pos), ptr(o), o.stride(0), dt(q), N, H, dh,
                                  scale, 1, stream_ptr()), "svla_attn_decode")
    return o


def swiglu_fwd(ab, g):
    F = g.shape[-1]
    check(_lib().svla_swiglu_fwd(get_ctx(), ptr(ab), ptr(g), dt(ab), g.numel() // F, F, stream_ptr()), "svla_swiglu_fwd")
    return g


def swiglu_bwd(ab, dg, dab):
    F = dg.shape[-1]
    check(_lib().svla_swiglu_bwd(get_ctx(), ptr(ab), ptr(dg), ptr(dab), dt(ab), dg.numel() // F, F, stream_ptr()),
          "svla_swiglu_bwd")
    return dab


def embed_time_fwd(obs_embed, prev_actions, masks, in_hand, time_step, E_a, E_h, div_term, x_out, T, N, A):
    D = obs_embed.shape[-1]
    check(_lib().svla_embed_time_fwd(get_ctx(), ptr(obs_embed), dt(obs_embed), ptr(prev_actions), ptr(masks),
                                     ptr(in_hand), ptr(time_step), ptr(E_a), ptr(E_h), ptr(div_term), ptr(x_out),
                                     T, N, A, D, stream_ptr()), "svla_embed_time_fwd")
    return x_out


def embed_time_bwd(dx, prev_actions, masks, in_hand, d_obs_embed, dE_a, dE_h, T, N, A):
    D = dx.shape[-1]
    check(_lib().svla_embed_time_bwd(get_ctx(), ptr(dx), ptr(prev_actions), ptr(masks), ptr(in_hand),
                                     ptr(d_obs_embed), dt(d_obs_embed), ptr(dE_a), ptr(dE_h), T, N, A, D,
                                     stream_ptr()), "svla_embed_time_bwd")


def nchw_to_tokens(x: torch.Tensor, out: torch.Tensor):
    """x [R, C, H, W] fp32 -> out [R, H*W, C]"""
    R, Cn = x.shape[0], x.shape[1]
    P = x[0, 0].numel()
    check(_lib().svla_nchw_to_tokens(get_ctx(), ptr(x), ptr(out), dt(out), R, Cn, P, stream_ptr()),
          "svla_nchw_to_tokens")
    return out


def copy_rows(src, dst, rows, D, *, lds=None, ldd=None, smap=IDENT, dmap=IDENT, idx=None, accumulate=False):
    lds = src.stride(-2) if lds is None else lds
    ldd = dst.stride(-2) if ldd is None else ldd
    check(_lib().svla_copy_rows(get_ctx(), ptr(src), dt(src), lds, smap, ptr(idx), ptr(dst), dt(dst), ldd, dmap, rows,
                                D, int(accumulate), stream_ptr()), "svla_copy_rows")
    return dst


def fill_rows(vec, dst, rows, D, *, ldd=None, dmap=IDENT):
    ldd = dst.stride(-2) if ldd is None else ldd
    check(_lib().svla_fill_rows(get_ctx(), ptr(vec), ptr(dst), dt(dst), ldd, dmap, rows, D, stream_ptr()),
          "svla_fill_rows")
    return dst


def hash_rows(rows_u8: torch.Tensor) -> torch.Tensor:
    assert rows_u8.dtype == torch.uint8 and rows_u8.dim() == 2 and rows_u8.is_contiguous()
    out = torch.empty(rows_u8.shape[0], device=rows_u8.device, dtype=torch.int64)
    check(_lib().svla_hash_rows(get_ctx(), ptr(rows_u8), rows_u8.shape[0], rows_u8.shape[1], ptr(out), stream_ptr()),
          "svla_hash_rows")
    return out
