"""Per-kernel parity of the C-ABI entry points against torch/oracle references (B200 only)."""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import torch_oracle as TO  # noqa: E402  (checker only)


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from safevla_b200 import _lib
    _lib.get_ctx()
    return torch.device("cuda:0")


def _ops():
    from safevla_b200 import ops
    return ops


def _L():
    from safevla_b200 import _lib
    return _lib


def relerr(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()


# ------------------------------------------------------------------------------------------ GAE
@pytest.mark.parametrize("T,N", [(16, 1), (128, 64), (5, 3), (1, 7), (257, 33), (128, 5000), (37, 8192), (128, 8200)])
@pytest.mark.parametrize("algo", [1, 2, 42])
def test_gae_dual(dev, T, N, algo):
    g = torch.Generator().manual_seed(T * 1000 + N)
    r = torch.randn(T, N, 1, generator=g)
    c = (torch.rand(T, N, 1, generator=g) < 0.2).float()
    v = torch.randn(T + 1, N, 1, generator=g)
    vc = torch.randn(T + 1, N, 1, generator=g)
    m = (torch.rand(T + 1, N, 1, generator=g) > 0.1).float()
    ret, adv = TO.gae_returns(r, v, m, 0.99, 0.95)
    cret, cadv = TO.gae_returns(c, vc, m, 0.99, 0.95)
    o = _ops().gae_dual(r.to(dev), c.to(dev), v.to(dev), vc.to(dev), m.to(dev), 0.99, 0.95, algo)
    got = [t.cpu() for t in o]
    if algo != 2:  # same operation order as the recursion: bit-exact (42 = 128-bit march when N % 4 == 0)
        assert torch.equal(got[0], ret) and torch.equal(got[1], cret)
        assert torch.equal(got[2], adv) and torch.equal(got[3], cadv)
    else:
        for a, b in zip(got, (ret, cret, adv, cadv)):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-5), (a - b).abs().max()


def test_gae_closed_form_and_mask_cut(dev):
    # all-ones masks, zero values: returns are plain discounted sums with factor gamma*lam
    T, N = 40, 4
    r = torch.ones(T, N, 1)
    z = torch.zeros(T + 1, N, 1)
    m = torch.ones(T + 1, N, 1)
    ret, _, adv, _ = _ops().gae_dual(r.to(dev), None, z.to(dev), None, m.to(dev), 0.99, 0.95, 1)
    gl = 0.99 * 0.95
    expect = torch.tensor([(1 - gl ** (T - t)) / (1 - gl) for t in range(T)])
    assert torch.allclose(ret[:T, 0, 0].cpu(), expect, rtol=1e-5)
    # a zero mask at t+1 cuts the recursion: step t only sees its own reward
    m[10] = 0
    ret2, _, _, _ = _ops().gae_dual(r.to(dev), None, z.to(dev), None, m.to(dev), 0.99, 0.95, 2)
    assert torch.allclose(ret2[9, :, 0].cpu(), torch.ones(N))


def test_normalize_advantage_two_phase_matches_global_statistics(dev):
    """Data-parallel form: per-shard {sum, sum^2, n} added up (what the all-reduce does), then every shard normalised
    with the global mean / unbiased std -- equals normalising the concatenated batch in one call."""
    g = torch.Generator().manual_seed(3)
    full = (torch.randn(128, 12, 1, generator=g) * 3 + 0.7).to(dev)
    ref, st_ref = _ops().normalize_advantage(full)
    shards = [full[:, :5].contiguous(), full[:, 5:].contiguous()]
    sums = sum(_ops().advantage_sums(s) for s in shards)
    assert abs(sums[2].item() - full.numel()) == 0
    outs = [_ops().normalize_advantage_from_sums(s, sums) for s in shards]
    got = torch.cat([o[0] for o in outs], 1)
    assert torch.allclose(got, ref, atol=2e-6, rtol=1e-5)
    assert torch.allclose(outs[0][1], st_ref, atol=1e-6, rtol=1e-5)
    exp = (full - full.mean()) / (full.std() + 1e-5)
    assert torch.allclose(got, exp, atol=1e-5, rtol=1e-4)


def test_normalize_advantage(dev):
    a = torch.randn(128, 64, 1) * 3 + 1
    out, stats = _ops().normalize_advantage(a.to(dev))
    assert torch.allclose(out.cpu(), TO.normalize_advantage(a), atol=1e-4)


# ------------------------------------------------------------------------------------------ loss
def _loss_inputs(R, A, seed):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(R, A, generator=g) * 2
    actions = torch.randint(0, A, (R,), generator=g)
    logp = torch.log_softmax(logits, -1).gather(-1, actions[:, None])[:, 0]
    old = logp + 0.2 * torch.randn(R, generator=g)
    adv, cadv = torch.randn(R, 1, generator=g), torch.randn(R, 1, generator=g)
    values, returns = torch.randn(R, 1, generator=g), torch.randn(R, 1, generator=g)
    oldv = values + 0.2 * torch.randn(R, 1, generator=g)
    return logits, actions, old, adv, cadv, values, returns, oldv


@pytest.mark.parametrize("R,A", [(16, 6), (8192, 20), (1000, 3), (77, 64), (100003, 20), (4099, 8), (333, 32)])
@pytest.mark.parametrize("lam,ent,clipv", [(0.37, 0.01, False), (0.0, 0.0, False), (1.5, 0.05, True)])
def test_ppo_lag_fused(dev, R, A, lam, ent, clipv):
    L = _L()
    logits, actions, old, adv, cadv, values, returns, oldv = _loss_inputs(R, A, R + A)
    lg = logits.clone().requires_grad_(True)
    vl = values.clone().requires_grad_(True)
    total, info = TO.safe_ppo_log_grad(lg.view(R, 1, A), actions.view(R, 1), old.view(R, 1), adv.view(R, 1, 1),
                                       cadv.view(R, 1, 1), vl.view(R, 1, 1), returns.view(R, 1, 1), lam,
                                       entropy_coef=ent, use_clipped_value_loss=clipv, old_values=oldv.view(R, 1, 1))
    total.backward()
    hp = L.PpoHparams(0.1, 1.0, 0.5, ent, 0.0, 1.0 / R, 1.0, int(clipv), 1)
    scal, dl, dv, _ = _ops().ppo_lag_fwd_bwd(logits.to(dev), actions.to(dev), old.to(dev), adv.to(dev), cadv.to(dev),
                                             values.to(dev), returns.to(dev), None, None,
                                             torch.tensor([lam], device=dev), hp, old_values=oldv.to(dev))
    s = scal.cpu()
    assert abs(s[0] - total.item()) <= 2e-5 * max(1, abs(total.item())), (s[0], total.item())
    assert abs(s[1] - info["value"].item()) <= 2e-5 * max(1, abs(info["value"].item()))
    assert abs(s[2] - info["action"].item()) <= 2e-5 * max(1, abs(info["action"].item()))
    assert abs(s[3] - info["entropy"].item()) <= 2e-5
    assert relerr(dl.cpu(), lg.grad) < 2e-4, relerr(dl.cpu(), lg.grad)
    assert relerr(dv.cpu(), vl.grad) < 1e-5


def test_ppo_lambda0_equals_ppologgrad_and_is_deterministic(dev):
    L = _L()
    R, A = 4096, 20
    logits, actions, old, adv, cadv, values, returns, _ = [t.to(dev) for t in _loss_inputs(R, A, 5)]
    hp1 = L.PpoHparams(0.1, 1.0, 0.5, 0.0, 0.0, 1.0 / R, 1.0, 0, 1)
    hp0 = L.PpoHparams(0.1, 1.0, 0.5, 0.0, 0.0, 1.0 / R, 1.0, 0, 0)
    a = _ops().ppo_lag_fwd_bwd(logits, actions, old, adv, cadv, values, returns, None, None,
                               torch.zeros(1, device=dev), hp1)
    b = _ops().ppo_lag_fwd_bwd(logits, actions, old, adv, None, values, returns, None, None, None, hp0)
    c = _ops().ppo_lag_fwd_bwd(logits, actions, old, adv, cadv, values, returns, None, None,
                               torch.zeros(1, device=dev), hp1)
    assert torch.equal(a[0][:8], b[0][:8]) and torch.equal(a[1], b[1])  # lambda = 0  ==  PPOLogGrad, bit for bit
    assert torch.equal(a[0], c[0]) and torch.equal(a[1], c[1])          # run-to-run bit stable


def test_stage0_value_losses(dev):
    L = _L()
    R = 2048
    g = torch.Generator().manual_seed(3)
    v, r, cv, cr = [torch.randn(R, 1, generator=g) for _ in range(4)]
    hp = L.PpoHparams(0.1, 0.0, 1.0, 0.0, 1.0, 1.0 / R, 1.0, 0, 0)
    scal, _, dv, dcv = _ops().ppo_lag_fwd_bwd(None, None, None, None, None, v.to(dev), r.to(dev), cv.to(dev), cr.to(dev),
                                              None, hp)
    exp = TO.ppo_value_loss(v, r) + TO.ppo_value_loss(cv, cr)
    assert abs(scal[0].item() - exp.item()) < 1e-5 * exp.item()
    assert torch.allclose(dv.cpu(), (v - r) / R, rtol=1e-5, atol=1e-9)
    assert torch.allclose(dcv.cpu(), (cv - cr) / R, rtol=1e-5, atol=1e-9)


def test_hl_gauss_fused_loss_matches_reference_golden(dev):
    """Discrete-critic HL-Gauss loss (utils/loss_functions.py:7-30): fused fwd+bwd kernel against the reference's
    own outputs (tests/golden/hl_gauss.pt) -- loss, d loss / d logits and the value read-out."""
    from oracle.make_golden_hlgauss import inputs
    from safevla_b200.losses import HLGaussLoss
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "hl_gauss.pt"), weights_only=False)
    for rec in gold:
        c = rec["case"]
        logits, target = inputs(c)
        m = HLGaussLoss(c["vmin"], c["vmax"], c["bins"], c["sigma"])
        lg = logits.to(dev).requires_grad_(True)
        loss = m(lg, target.to(dev))
        loss.backward()
        assert abs(loss.item() - rec["loss"].item()) < 1e-5 * max(1, abs(rec["loss"].item()))
        assert relerr(lg.grad.cpu(), rec["dlogits"]) < 1e-4
        vals = m.transform_from_probs(torch.softmax(logits.to(dev), -1))
        assert relerr(vals.cpu(), rec["values"]) < 1e-5
        assert relerr(m.values_from_logits(logits.to(dev)).cpu(), rec["values"]) < 1e-5
        assert relerr(m.transform_to_probs(target.to(dev)).cpu(), rec["probs"]) < 1e-4


# ------------------------------------------------------------------------------------------ optimizer / lagrange
def test_lagrange_update(dev):
    from safevla_b200.lagrange import Lagrange
    lag = Lagrange(2.31964, device=dev)
    orc = TO.LagrangeOracle(2.31964)
    # first Adam step moves lambda by exactly +-lr (sign only) -- SURVEY A.5 KAT
    lag.update_lagrange_multiplier(5.0)
    assert abs(lag.lagrangian_multiplier.item() - (0.001 + 0.035)) < 1e-6
    orc.update(5.0)
    for jc in [4.0, 1.0, 0.2, 0.0, 0.0, 0.0, 3.0, 0.1, 0.0, 0.0]:
        lag.update_lagrange_multiplier(jc)
        assert abs(lag.lagrangian_multiplier.item() - orc.update(jc)) < 1e-5
    # zero-episode rollout KAT: no finished episode -> no Jc estimate -> lambda AND its Adam state stay untouched
    lag2 = Lagrange(1.0, lagrangian_multiplier_init=0.01, device=dev)
    lag2.update_from_sum_count(torch.tensor([7.0, 3.0], device=dev))
    lam_before, st_before = lag2.lagrangian_multiplier.clone(), lag2.state.clone()
    lag2.update_from_sum_count(torch.tensor([0.0, 0.0], device=dev))
    assert torch.equal(lag2.lagrangian_multiplier, lam_before) and torch.equal(lag2.state, st_before)
    lag2.update_from_sum_count(torch.tensor([0.0, 2.0], device=dev))  # finished episodes with zero cost DO count
    assert lag2.state[0, 2].item() == 2.0 and lag2.state[0, 3].item() == 0.0  # Adam step advanced, Jc = 0 recorded
    assert not torch.equal(lag2.lagrangian_multiplier, lam_before)


@pytest.mark.parametrize("n", [64, 4096 * 33, 21_000_000])
def test_sq_norm_and_clip_adam(dev, n):
    L = _L()
    g = torch.Generator().manual_seed(n % 1000)
    p0, gr = torch.randn(n, generator=g), torch.randn(n, generator=g) * 0.01
    p = p0.clone().to(dev)
    grad = gr.clone().to(dev)
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    shadow = torch.zeros(n, device=dev, dtype=torch.bfloat16)
    ref = torch.nn.Parameter(p0.clone().to(dev))
    opt = torch.optim.Adam([ref], lr=2e-5)
    for step in (1, 2, 3):
        ref.grad = gr.clone().to(dev) * step
        tn = torch.nn.utils.clip_grad_norm_([ref], 0.5)
        opt.step()
        grad.copy_(gr.to(dev) * step)
        sq = _ops().sq_norm(grad)
        assert abs(sq.sqrt().item() - tn.item()) < 1e-4 * tn.item()
        hp = L.AdamHparams(2e-5, 0.9, 0.999, 1e-8, 0.5, 1.0, step, 1)
        _ops().clip_adam(p, grad, m, v, shadow, sq, hp)
        assert grad.abs().max().item() == 0.0  # fused zero_grad
    assert (p - ref.detach()).abs().max().item() < 1e-6  # 1-2 ulp at |p| ~ 4
    assert torch.equal(shadow, p.to(torch.bfloat16))


# ------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(16, 512, 384), (1872, 1536, 512), (130, 7, 512), (512, 2048, 9000), (1, 1, 1),
                                   (300, 20, 8192)])
@pytest.mark.parametrize("ta,tb", [(False, True), (False, False), (True, False), (True, True)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gemm_simt(dev, M, N, K, ta, tb, dtype):
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn((K, M) if ta else (M, K), generator=g).to(dev, dtype)
    B = torch.randn((N, K) if tb else (K, N), generator=g).to(dev, dtype)
    bias = torch.randn(N, generator=g).to(dev)
    res = torch.randn(M, N, generator=g).to(dev, dtype)
    out = torch.empty(M, N, device=dev, dtype=dtype)
    _ops().gemm(A, B, out, trans_a=ta, trans_b=tb, bias=bias, residual=res, epilogue=_L().EPI_RELU, impl=1)
    a = (A.t() if ta else A).double()
    b = (B.t() if tb else B).double()
    exp = torch.relu(a @ b + bias.double()) + res.double()
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert relerr(out, exp) < tol, relerr(out, exp)
    # accumulate into fp32 + relu-mask epilogue
    acc = torch.randn(M, N, generator=g).to(dev)
    acc0 = acc.clone()
    aux = torch.randn(M, N, generator=g).to(dev, dtype)
    _ops().gemm(A, B, acc, trans_a=ta, trans_b=tb, aux=aux, epilogue=_L().EPI_RELU_MASK, accumulate=True, impl=1)
    exp2 = acc0.double() + (a @ b) * (aux.double() > 0)
    assert relerr(acc, exp2) < (1e-5 if dtype == torch.float32 else 2e-5), relerr(acc, exp2)


def test_gemm_strided_views(dev):
    # column slices of packed buffers and the CLS-row view (lda = S*D)
    x = torch.randn(10, 117 * 512, device=dev)
    w = torch.randn(1536, 512, device=dev)
    out = torch.empty(10, 512, device=dev)
    _ops().gemm(x[:, :512], w[512:1024], out, trans_b=True, impl=1)
    assert relerr(out, x[:, :512].double() @ w[512:1024].double().t()) < 1e-5


def test_colsum(dev):
    for M, N, dt in [(1000, 512, torch.float32), (5000, 2048, torch.bfloat16), (333, 6, torch.float32), (64, 1, torch.float32)]:
        x = torch.randn(M, N, device=dev).to(dt)
        out = torch.ones(N, device=dev)
        _ops().colsum(x, out, accumulate=True)
        assert torch.allclose(out, 1 + x.double().sum(0).float(), rtol=1e-4, atol=1e-3)


# ------------------------------------------------------------------------------------------ norms
@pytest.mark.parametrize("dt_", [torch.float32, torch.bfloat16])
def test_layernorm_fwd_bwd(dev, dt_):
    L = _L()
    rows, D, G, S, off = 84 * 5, 512, 84, 117, 1
    g = torch.Generator().manual_seed(0)
    x = torch.randn(rows, D, generator=g).to(dev, dt_)
    gamma, beta, token = [torch.randn(D, generator=g).to(dev) for _ in range(3)]
    y = torch.zeros(5 * S, D, device=dev, dtype=dt_)
    mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
    _ops().layernorm_fwd(x, gamma, beta, y, token=token, relu=True, ymap=L.RowMap(G, S, off), mean=mean, rstd=rstd)
    xr = x.float().clone().requires_grad_(True)
    gr, br, tr = [t.clone().requires_grad_(True) for t in (gamma, beta, token)]
    yr = torch.relu(torch.nn.functional.layer_norm(xr, (D,), gr, br, 1e-5)) + tr
    got = y.view(5, S, D)[:, off:off + G].reshape(rows, D).float()
    assert relerr(got, yr) < (1e-5 if dt_ == torch.float32 else 1e-2)
    dy_full = torch.randn(5 * S, D, generator=g).to(dev, dt_)
    dyr = dy_full.view(5, S, D)[:, off:off + G].reshape(rows, D).float()
    yr.backward(dyr)
    dx = torch.empty(rows, D, device=dev, dtype=dt_)
    dg, db, dtok = [torch.zeros(D, device=dev) for _ in range(3)]
    _ops().layernorm_bwd(dy_full, x, gamma, beta, mean, rstd, dx, dg, db, relu=True, dymap=L.RowMap(G, S, off),
                         dtoken=dtok)
    tol = 1e-4 if dt_ == torch.float32 else 2e-2
    assert relerr(dx.float(), xr.grad) < tol
    assert relerr(dg, gr.grad) < tol and relerr(db, br.grad) < tol and relerr(dtok, tr.grad) < tol


def test_rmsnorm_fwd_bwd(dev):
    rows, D = 777, 512
    g = torch.Generator().manual_seed(1)
    x = torch.randn(rows, D, generator=g).to(dev)
    w = torch.randn(D, generator=g).to(dev)
    y, rstd = torch.empty_like(x), torch.empty(rows, device=dev)
    _ops().rmsnorm_fwd(x, w, y, 1e-5, rstd)
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    yr = TO.rms_norm(xr, wr, 1e-5)
    assert relerr(y, yr) < 1e-5
    dy = torch.randn(rows, D, generator=g).to(dev)
    yr.backward(dy)
    base = torch.randn(rows, D, generator=g).to(dev)
    dx, dw = base.clone(), torch.zeros(D, device=dev)
    _ops().rmsnorm_bwd(dy, x, w, rstd, dx, dw, accumulate_dx=True)
    assert relerr(dx - base, xr.grad) < 1e-4 and relerr(dw, wr.grad) < 1e-4


# ------------------------------------------------------------------------------------------ attention
def _attn_ref(q, k, v, B, S, H, scale, mask=None, bias=None):
    dh = 64
    qq = q.view(B, S, H, dh).transpose(1, 2)
    kk = k.view(B, S, H, dh).transpose(1, 2)
    vv = v.view(B, S, H, dh).transpose(1, 2)
    s = (qq @ kk.transpose(-1, -2)) * scale
    if bias is not None:
        s = s + bias
    if mask is not None:
        s = s.masked_fill(~mask, float("-inf"))
    return (torch.softmax(s, -1) @ vv).transpose(1, 2).reshape(B * S, H * dh)


@pytest.mark.parametrize("mode,S,B", [(0, 117, 3), (0, 201, 2), (1, 128, 4), (1, 16, 1), (2, 32, 5), (1, 256, 2), (0, 256, 1)])
@pytest.mark.parametrize("dt_", [torch.float32, torch.bfloat16])
def test_attention_fwd_bwd(dev, mode, S, B, dt_):
    H, D = 8, 512
    g = torch.Generator().manual_seed(S + mode)
    qkv = (torch.randn(B * S, 3 * D, generator=g) * 0.5).to(dev, dt_)
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    traj = bias = km = mask = None
    scale = 0.125
    if mode == 1:
        traj = torch.cumsum((torch.rand(B, S, generator=g) < 0.1).long(), 1).to(dev)
        mask = torch.tril(traj[:, :, None] == traj[:, None, :]).unsqueeze(1)
    if mode == 2:
        scale = 1.0
        bias = torch.randn(H, S, S, generator=g).to(dev)
        km = (torch.arange(S)[None, :] < torch.randint(S // 2, S + 1, (B, 1), generator=g)).long().to(dev)
        mask = km.bool()[:, None, None, :]
    o = torch.empty(B * S, D, device=dev, dtype=dt_)
    lse = torch.empty(B * H * S, device=dev)
    _ops().attn_fwd(mode, q, k, v, o, lse, B, S, scale=scale, traj=traj, bias=bias, keymask=km)
    qr, kr, vr = [t.float().clone().requires_grad_(True) for t in (q, k, v)]
    ref = _attn_ref(qr, kr, vr, B, S, H, scale, mask, bias)
    tol = 2e-5 if dt_ == torch.float32 else 2e-2
    assert relerr(o.float(), ref) < tol, relerr(o.float(), ref)
    if mode == 2:
        return
    do = torch.randn(B * S, D, generator=g).to(dev, dt_)
    ref.backward(do.float())
    dqkv = torch.empty(B * S, 3 * D, device=dev, dtype=dt_)
    _ops().attn_bwd(mode, q, k, v, o, do, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], lse, B, S, scale=scale,
                    traj=traj)
    tol = 1e-4 if dt_ == torch.float32 else 3e-2
    for got, exp in ((dqkv[:, :D], qr.grad), (dqkv[:, D:2 * D], kr.grad), (dqkv[:, 2 * D:], vr.grad)):
        assert relerr(got.float(), exp) < tol, relerr(got.float(), exp)


@pytest.mark.parametrize("dt_", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("S", [117, 201, 33, 256, 2])
def test_attention_cls(dev, S, dt_):
    B, H, D = 37, 8, 512
    g = torch.Generator().manual_seed(S)
    kv = (torch.randn(B * S, 2 * D, generator=g) * 0.5).to(dev, dt_)
    q0 = (torch.randn(B, D, generator=g) * 0.5).to(dev, dt_)
    o, lse = torch.empty(B, D, device=dev, dtype=dt_), torch.empty(B * H, device=dev)
    _ops().attn_cls_fwd(q0, kv[:, :D], kv[:, D:], o, lse, B, S)
    qr, kr, vr = [t.float().clone().requires_grad_(True) for t in (q0, kv[:, :D], kv[:, D:])]
    qq = qr.view(B, 1, H, 64).transpose(1, 2)
    kk = kr.view(B, S, H, 64).transpose(1, 2)
    vv = vr.view(B, S, H, 64).transpose(1, 2)
    sc = qq @ kk.transpose(-1, -2) * 0.125
    ref = (torch.softmax(sc, -1) @ vv).transpose(1, 2).reshape(B, D)
    f32 = dt_ == torch.float32
    assert relerr(o.float(), ref) < (2e-5 if f32 else 1e-2)
    assert (lse.view(B, H) - torch.logsumexp(sc, -1).view(B, H)).abs().max().item() < (1e-4 if f32 else 1e-2)
    do = torch.randn(B, D, generator=g).to(dev, dt_)
    ref.backward(do.float())
    dq, dkv = torch.empty(B, D, device=dev, dtype=dt_), torch.empty(B * S, 2 * D, device=dev, dtype=dt_)
    _ops().attn_cls_bwd(q0, kv[:, :D], kv[:, D:], o, do, dq, dkv[:, :D], dkv[:, D:], lse, B, S)
    tol = 1e-4 if f32 else 3e-2
    assert relerr(dq.float(), qr.grad) < tol and relerr(dkv[:, :D].float(), kr.grad) < tol
    assert relerr(dkv[:, D:].float(), vr.grad) < tol


@pytest.mark.parametrize("dt_", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("N,pos,rows", [(3, 0, 8), (5, 7, 8), (64, 130, 256), (2, 499, 500)])
def test_attention_decode_kv_cache(dev, N, pos, rows, dt_):
    """Single-step attention against the KV cache with the episode-start mask (llama/model.py:279-317,
    allenact_dino_transformer.py:386-397)."""
    H, D = 8, 512
    g = torch.Generator().manual_seed(N * 1000 + pos)
    ck = (torch.randn(N, rows, D, generator=g) * 0.5).to(dev, dt_)
    cv = (torch.randn(N, rows, D, generator=g) * 0.5).to(dev, dt_)
    q = (torch.randn(N, 3 * D, generator=g) * 0.5).to(dev, dt_)[:, :D]  # strided view like the packed QKV buffer
    ts = torch.randint(0, pos + 5, (1, N), generator=g).to(dev)
    o = torch.empty(N, D, device=dev, dtype=dt_)
    _ops().attn_decode(q, ck, cv, ts, pos, o)
    start = torch.clamp(pos - ts.view(N), min=0)
    ar = torch.arange(pos + 1, device=dev)
    mask = (start[:, None] <= ar[None, :])[:, None, None, :]
    qq = q.float().view(N, 1, H, 64).transpose(1, 2)
    kk = ck[:, :pos + 1].float().view(N, pos + 1, H, 64).transpose(1, 2)
    vv = cv[:, :pos + 1].float().view(N, pos + 1, H, 64).transpose(1, 2)
    sc = (qq @ kk.transpose(-1, -2) * 0.125).masked_fill(~mask, float("-inf"))
    ref = (torch.softmax(sc, -1) @ vv).transpose(1, 2).reshape(N, D)
    assert relerr(o.float(), ref) < (2e-5 if dt_ == torch.float32 else 1e-2)


@pytest.mark.parametrize("S,B,H", [(433, 3, 6), (257, 5, 6), (300, 2, 8), (1024, 1, 6)])
def test_attention_flash_long_sequences(dev, S, B, H):
    """S > 256 (vision preprocessor): tcgen05 flash forward (bf16) and the CUDA-core path (fp32) against torch."""
    D = H * 64
    g = torch.Generator().manual_seed(S)
    qkv32 = (torch.randn(B * S, 3 * D, generator=g) * 0.6).to(dev)
    for dt_, tol in ((torch.bfloat16, 2e-2), (torch.float32, 2e-5)):
        qkv = qkv32.to(dt_)
        o = torch.empty(B * S, D, device=dev, dtype=dt_)
        lse = torch.empty(B * H * S, device=dev) if dt_ == torch.bfloat16 else None
        _ops().attn_fwd(0, qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, lse, B, S, H=H)
        q, k, v = [qkv[:, i * D:(i + 1) * D].float().view(B, S, H, 64).transpose(1, 2) for i in range(3)]
        sc = q @ k.transpose(-1, -2) * 0.125
        ref = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(B * S, D)
        assert relerr(o.float(), ref) < tol, (dt_, relerr(o.float(), ref))
        if lse is not None:
            assert (lse.view(B, H, S) - torch.logsumexp(sc, -1)).abs().max().item() < 2e-2


def test_vit_glue_kernels(dev):
    """patchify (normalise + crop + im2col), token assembly, adaptive pooling, GELU epilogue."""
    g = torch.Generator().manual_seed(0)
    N, H, W, P = 3, 28, 48, 14
    img = torch.randint(0, 256, (N, H, W, 3), generator=g, dtype=torch.uint8).to(dev)
    mean, std = (0.48, 0.45, 0.40), (0.26, 0.27, 0.28)
    out = torch.full((N * 2 * 3, 592), 7.0, device=dev)
    _ops().patchify_u8(img, out, P, 3, 3, mean, std)
    x = (img.permute(0, 3, 1, 2).float() / 255 - torch.tensor(mean, device=dev).view(1, 3, 1, 1)) / torch.tensor(std, device=dev).view(1, 3, 1, 1)
    ref = torch.nn.functional.unfold(x[:, :, :, 3:-3], kernel_size=P, stride=P).transpose(1, 2).reshape(N * 6, 588)
    assert torch.allclose(out[:, :588], ref, rtol=1e-6, atol=1e-6) and out[:, 588:].abs().max() == 0
    D, Np = 384, 6
    patches, cls, pos = torch.randn(N * Np, D, device=dev), torch.randn(D, device=dev), torch.randn(Np + 1, D, device=dev)
    xt = _ops().vit_assemble(patches, cls, pos, torch.empty(N * (Np + 1), D, device=dev), N, Np)
    exp = torch.cat([cls.expand(N, 1, D), patches.view(N, Np, D)], 1) + pos
    assert torch.equal(xt.view(N, Np + 1, D), exp)
    PH, PW = 16, 27
    tok = torch.randn(N * (PH * PW + 1), D, device=dev)
    pooled = _ops().tokens_pool(tok, torch.empty(N, D, 7, 12, device=dev), N, PH, PW, 7, 12)
    grid = tok.view(N, PH * PW + 1, D)[:, 1:].permute(0, 2, 1).reshape(N, D, PH, PW)
    assert torch.allclose(pooled, torch.nn.functional.adaptive_avg_pool2d(grid, (7, 12)), rtol=1e-5, atol=1e-6)
    for dt_, impl in ((torch.float32, 1), (torch.bfloat16, 2)):
        A = (torch.randn(512, 384, generator=g) * 0.5).to(dev, dt_)
        Bm = (torch.randn(1536, 384, generator=g) * 0.1).to(dev, dt_)
        bias = torch.randn(1536, generator=g).to(dev)
        o = torch.empty(512, 1536, device=dev, dtype=dt_)
        _ops().gemm(A, Bm, o, trans_b=True, bias=bias, epilogue=_L().EPI_GELU, impl=impl)
        refg = torch.nn.functional.gelu(A.double() @ Bm.double().t() + bias.double())
        assert relerr(o, refg) < (1e-5 if dt_ == torch.float32 else 1e-2)


# ------------------------------------------------------------------------------------------ glue
def test_swiglu(dev):
    rows, F = 300, 1536
    ab = torch.randn(rows, 2 * F, device=dev)
    gbuf = torch.empty(rows, F, device=dev)
    _ops().swiglu_fwd(ab, gbuf)
    abr = ab.clone().requires_grad_(True)
    ref = torch.nn.functional.silu(abr[:, :F]) * abr[:, F:]
    assert relerr(gbuf, ref) < 1e-5
    dg = torch.randn(rows, F, device=dev)
    ref.backward(dg)
    dab = torch.empty_like(ab)
    _ops().swiglu_bwd(ab, dg, dab)
    assert relerr(dab, abr.grad) < 1e-4


def test_embed_time(dev):
    T, N, A, D = 9, 4, 6, 512
    g = torch.Generator().manual_seed(2)
    obs = torch.randn(T * N, D, generator=g).to(dev)
    prev = torch.randint(0, A, (T, N), generator=g).to(dev)
    masks = (torch.rand(T, N, generator=g) > 0.3).float().to(dev)
    hand = torch.randint(0, 2, (T, N), generator=g).to(dev)
    ts = torch.randint(0, 500, (T, N), generator=g).to(dev)
    Ea, Eh = torch.randn(A + 2, D, generator=g).to(dev), torch.randn(3, D, generator=g).to(dev)
    div = torch.exp(torch.arange(0, D, 2) * (-math.log(10000.0) / D)).to(dev)
    x = torch.empty(N * T, D, device=dev)
    _ops().embed_time_fwd(obs, prev, masks, hand, ts, Ea, Eh, div, x, T, N, A)
    idx = torch.where(masks != 0, prev, torch.full_like(prev, A))
    ref = obs.view(T, N, D) + Ea[idx] + Eh[hand] + TO.time_encoding(div.cpu(), ts.cpu()).to(dev)
    assert (x.view(N, T, D).permute(1, 0, 2) - ref).abs().max().item() < 2e-5
    dx = torch.randn(N * T, D, generator=g).to(dev)
    dobs, dEa, dEh = torch.empty(T * N, D, device=dev), torch.zeros(A + 2, D, device=dev), torch.zeros(3, D, device=dev)
    _ops().embed_time_bwd(dx, prev, masks, hand, dobs, dEa, dEh, T, N, A)
    d_tn = dx.view(N, T, D).permute(1, 0, 2)
    assert torch.equal(dobs.view(T, N, D), d_tn.contiguous())
    exp_a = torch.zeros(A + 2, D, device=dev).index_add_(0, idx.reshape(-1), d_tn.reshape(-1, D))
    exp_h = torch.zeros(3, D, device=dev).index_add_(0, hand.reshape(-1), d_tn.reshape(-1, D))
    assert relerr(dEa, exp_a) < 1e-5 and relerr(dEh, exp_h) < 1e-5


def test_nchw_tokens_copy_fill_hash(dev):
    L = _L()
    x = torch.randn(7, 384, 7, 12, device=dev)
    for dt_ in (torch.float32, torch.bfloat16):
        y = torch.empty(7 * 84, 384, device=dev, dtype=dt_)
        _ops().nchw_to_tokens(x, y)
        assert torch.equal(y.view(7, 84, 384), x.reshape(7, 384, 84).transpose(1, 2).to(dt_))
    src = torch.randn(50, 512, device=dev)
    idx = torch.randint(0, 50, (30,), device=dev)
    dst = torch.zeros(30 * 4, 512, device=dev, dtype=torch.bfloat16)
    _ops().copy_rows(src, dst, 30, 512, idx=idx, dmap=L.RowMap(1, 4, 2))
    assert torch.equal(dst.view(30, 4, 512)[:, 2], src[idx].to(torch.bfloat16))
    vec = torch.randn(512, device=dev)
    _ops().fill_rows(vec, dst, 30, 512, dmap=L.RowMap(1, 4, 0))
    assert torch.equal(dst.view(30, 4, 512)[:, 0], vec.to(torch.bfloat16).expand(30, 512))
    rows = torch.randint(0, 255, (64, 1000), dtype=torch.uint8, device=dev)
    rows[10] = rows[3]
    h = _ops().hash_rows(rows)
    assert h[10] == h[3] and torch.unique(h).numel() == 63
    rows[10, 999] ^= 1
    assert _ops().hash_rows(rows)[10] != h[3]


# ------------------------------------------------------------------------------------------ tcgen05 GEMM
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 512), (1872, 1536, 512), (1000, 384, 2048),
                                   (117 * 64, 2048, 512), (70, 128, 64), (512, 512, 117 * 256)])
@pytest.mark.parametrize("ta,tb", [(False, True), (False, False), (True, False)])
def test_gemm_tcgen05(dev, M, N, K, ta, tb):
    if ta and M % 8:
        pytest.skip("MN-major A needs a 16-byte aligned leading dimension (falls back to the FMA kernel)")
    g = torch.Generator().manual_seed(M + 3 * N + 7 * K)
    A = (torch.randn((K, M) if ta else (M, K), generator=g) * 0.5).to(dev, torch.bfloat16)
    B = (torch.randn((N, K) if tb else (K, N), generator=g) * 0.5).to(dev, torch.bfloat16)
    a = (A.t() if ta else A).double()
    b = (B.t() if tb else B).double()
    ref = a @ b
    scale = ref.abs().max().item()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    _ops().gemm(A, B, out, trans_a=ta, trans_b=tb, impl=2)
    assert (out.double() - ref).abs().max().item() < 1e-2 * scale
    # epilogues: bias + relu + residual, bf16 out
    bias = torch.randn(N, generator=g).to(dev)
    res = torch.randn(M, N, generator=g).to(dev, torch.bfloat16)
    _ops().gemm(A, B, out, trans_a=ta, trans_b=tb, bias=bias, residual=res, epilogue=_L().EPI_RELU, impl=2)
    exp = torch.relu(ref + bias.double()) + res.double()
    assert (out.double() - exp).abs().max().item() < 1e-2 * scale
    # relu-mask + accumulate into fp32 (exact fp32 accumulation of bf16 products up to summation order)
    aux = torch.randn(M, N, generator=g).to(dev, torch.bfloat16)
    acc = torch.randn(M, N, generator=g).to(dev)
    acc0 = acc.clone()
    _ops().gemm(A, B, acc, trans_a=ta, trans_b=tb, aux=aux, epilogue=_L().EPI_RELU_MASK, accumulate=True, impl=2)
    exp2 = acc0.double() + ref * (aux.double() > 0)
    assert (acc.double() - exp2).abs().max().item() < 2e-5 * max(scale, 1.0) * max(1.0, (K / 512) ** 0.5)


@pytest.mark.parametrize("M,N,K,tb", [(117 * 70, 1024, 512, True), (117 * 70, 512, 512, False), (1000, 256, 192, True),
                                      (130, 128, 64, True), (40000, 512, 128, False)])
def test_gemm_tcgen05_epilogue_combinations(dev, M, N, K, tb):
    """Every compile-time specialised epilogue (tc_common.cuh: epilogue kind x bias x residual x accumulate, bf16 and
    fp32 outputs) and the run-time generic one, on shapes with many tiles per CTA (side-operand prefetch across tile
    boundaries), ragged M and both kernels (pair / single CTA)."""
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) * 0.5).to(dev, torch.bfloat16)
    B = (torch.randn((N, K) if tb else (K, N), generator=g) * 0.5).to(dev, torch.bfloat16)
    ref = A.double() @ (B.t() if tb else B).double()
    scale = ref.abs().max().item()
    bias = torch.randn(N, generator=g).to(dev)
    res = torch.randn(M, N, generator=g).to(dev, torch.bfloat16)
    aux = torch.randn(M, N, generator=g).to(dev, torch.bfloat16)
    aux[::7] = 0.0  # exact zeros and negative zeros must mask like ReLU'(0) = 0
    aux[3::11] = -0.0
    E = _L()
    gelu = lambda x: 0.5 * x * (1.0 + torch.erf(x / 2 ** 0.5))  # noqa: E731
    cases = [  # (epilogue, bias, residual, alpha) -> expected
        (E.EPI_NONE, True, False, 1.0, ref + bias.double()),
        (E.EPI_RELU, True, False, 1.0, torch.relu(ref + bias.double())),
        (E.EPI_NONE, True, True, 1.0, ref + bias.double() + res.double()),
        (E.EPI_RELU_MASK, False, False, 1.0, ref * (aux.double() > 0)),
        (E.EPI_NONE, False, True, 1.0, ref + res.double()),
        (E.EPI_NONE, False, False, 1.0, ref),
        (E.EPI_GELU, True, False, 1.0, gelu(ref + bias.double())),
        (E.EPI_RELU_MASK, False, True, 1.0, ref * (aux.double() > 0) + res.double()),   # generic path
        (E.EPI_RELU, False, False, 0.5, torch.relu(0.5 * ref)),                         # generic path, alpha
    ]
    for epi, use_b, use_r, alpha, exp in cases:
        out = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
        _ops().gemm(A, B, out, trans_b=tb, bias=bias if use_b else None, residual=res if use_r else None,
                    aux=aux if epi == E.EPI_RELU_MASK else None, epilogue=epi, alpha=alpha, impl=2)
        err = (out.double() - exp).abs().max().item()
        assert err < 1e-2 * max(scale, 1.0), (epi, use_b, use_r, err)
        if epi == E.EPI_RELU_MASK and not use_r:
            assert (out[aux <= 0] == 0).all(), "masked entries must be exact zeros"
    # fp32 outputs: plain, + fp32 residual, accumulate
    res32 = torch.randn(M, N, generator=g).to(dev)
    for use_r, acc in ((False, False), (True, False), (False, True)):
        out = torch.randn(M, N, generator=g).to(dev)
        out0 = out.clone()
        _ops().gemm(A, B, out, trans_b=tb, residual=res32 if use_r else None, accumulate=acc, impl=2)
        exp = ref + (res32.double() if use_r else 0) + (out0.double() if acc else 0)
        assert (out.double() - exp).abs().max().item() < 2e-5 * max(scale, 1.0) * max(1.0, (K / 512) ** 0.5)


@pytest.mark.parametrize("M,N,K", [(512, 512, 117 * 256), (2048, 512, 117 * 64), (1536, 512, 4000), (256, 256, 64),
                                   (512, 384, 84 * 96), (300, 512, 1000), (128, 512, 2048), (20, 512, 8192)])
def test_gemm_wgrad_fused_bias_gradient(dev, M, N, K):
    """dW = dY^T X with the bias gradient colsum(dY) out of the same launch (ones-MMA in the pair kernel; the
    dispatcher runs a separate column-sum pass for shapes that kernel does not take)."""
    g = torch.Generator().manual_seed(M + N + K)
    dy = (torch.randn(K, M, generator=g) * 0.5).to(dev, torch.bfloat16)
    x = (torch.randn(K, N, generator=g) * 0.5).to(dev, torch.bfloat16)
    gw0, gb0 = torch.randn(M, N, generator=g).to(dev), torch.randn(M, generator=g).to(dev)
    gw, gb = gw0.clone(), gb0.clone()
    _ops().gemm(dy, x, gw, trans_a=True, trans_b=False, accumulate=True, colsum_a=gb)
    ref_w = gw0.double() + dy.double().t() @ x.double()
    ref_b = gb0.double() + dy.double().sum(0)
    assert (gw.double() - ref_w).abs().max().item() < 2e-5 * ref_w.abs().max().item() * max(1.0, (K / 512) ** 0.5)
    assert (gb.double() - ref_b).abs().max().item() < 2e-5 * max(ref_b.abs().max().item(), 1.0) * max(1.0, (K / 512) ** 0.5)
    # bit-stable run to run
    gw2, gb2 = gw0.clone(), gb0.clone()
    _ops().gemm(dy, x, gw2, trans_a=True, trans_b=False, accumulate=True, colsum_a=gb2)
    assert torch.equal(gw, gw2) and torch.equal(gb, gb2)


@pytest.mark.parametrize("mode,S,B", [(0, 117, 5), (0, 128, 2), (1, 128, 3), (1, 16, 1), (0, 33, 300), (0, 201, 7),
                                      (0, 256, 3), (1, 256, 2), (0, 129, 160), (1, 200, 5)])
def test_attention_tcgen05_vs_cuda_core(dev, mode, S, B):
    """tcgen05 attention (forced) against the CUDA-core kernel on identical bf16 inputs, fwd and bwd."""
    H, D = 8, 512
    lib = _L().load_library()
    g = torch.Generator().manual_seed(S * 7 + mode)
    qkv = (torch.randn(B * S, 3 * D, generator=g) * 0.7).to(dev, torch.bfloat16)
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    do = torch.randn(B * S, D, generator=g).to(dev, torch.bfloat16)
    traj = torch.cumsum((torch.rand(B, S, generator=g) < 0.1).long(), 1).to(dev) if mode == 1 else None
    res = {}
    try:
        for impl in (1, 2):
            lib.svla_set_attn_impl(impl)
            o = torch.zeros(B * S, D, device=dev, dtype=torch.bfloat16)
            lse = torch.zeros(B * H * S, device=dev)
            _ops().attn_fwd(mode, q, k, v, o, lse, B, S, traj=traj)
            dqkv = torch.zeros(B * S, 3 * D, device=dev, dtype=torch.bfloat16)
            _ops().attn_bwd(mode, q, k, v, o, do, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], lse, B, S, traj=traj)
            torch.cuda.synchronize()
            res[impl] = (o.float(), lse, dqkv.float())
    finally:
        lib.svla_set_attn_impl(0)
    assert relerr(res[2][0], res[1][0]) < 2e-2, relerr(res[2][0], res[1][0])
    assert (res[2][1] - res[1][1]).abs().max().item() < 2e-2
    for c in range(3):
        a, b = res[2][2][:, c * D:(c + 1) * D], res[1][2][:, c * D:(c + 1) * D]
        assert relerr(a, b) < 4e-2, (c, relerr(a, b))


def test_gemm_tcgen05_strided_and_repeat(dev):
    # CLS-row view (lda = S*D), column-sliced weight, repeated launches reuse cached tensor maps
    S, D, R = 117, 512, 300
    x = (torch.randn(R, S * D, device=dev) * 0.5).to(torch.bfloat16)
    w = (torch.randn(3 * D, D, device=dev) * 0.1).to(torch.bfloat16)
    out = torch.empty(R, D, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        _ops().gemm(x[:, :D], w[0:D], out, trans_b=True, impl=2)
    ref = x[:, :D].double() @ w[0:D].double().t()
    assert (out.double() - ref).abs().max().item() < 1e-2 * ref.abs().max().item()
    qkv = torch.empty(R, 3 * D, device=dev, dtype=torch.bfloat16)
    _ops().gemm(x[:, :D], w, qkv, trans_b=True, impl=2)
    ref = x[:, :D].double() @ w.double().t()
    assert (qkv.double() - ref).abs().max().item() < 1e-2 * ref.abs().max().item()


# ------------------------------------------------------------------------------------------ use_gae = False
@pytest.mark.parametrize("T,N", [(16, 1), (128, 64), (5, 3), (1, 7), (257, 33), (40, 5000)])
def test_discounted_returns_dual_bit_exact(dev, T, N):
    """The `use_gae=False` branch of compute_returns: bit-exact against the sequential loop on both
    streams, and against the plain-C restatement oracle/c/gae_ref.c::discounted_returns_f32."""
    import ctypes as C
    import numpy as np
    from oracle.c import build as oracle_c
    g = torch.Generator().manual_seed(T * 131 + N)
    r = torch.randn(T, N, 1, generator=g)
    c = (torch.rand(T, N, 1, generator=g) < 0.2).float()
    v = torch.randn(T + 1, N, 1, generator=g)
    vc = torch.randn(T + 1, N, 1, generator=g)
    m = (torch.rand(T + 1, N, 1, generator=g) > 0.1).float()
    ret, adv = TO.gae_returns(r, v, m, 0.99, 0.95, use_gae=False)
    cret, cadv = TO.gae_returns(c, vc, m, 0.99, 0.95, use_gae=False)
    got = [t.cpu() for t in _ops().discounted_returns_dual(r.to(dev), c.to(dev), v.to(dev), vc.to(dev), m.to(dev), 0.99)]
    assert torch.equal(got[0], ret) and torch.equal(got[1], cret)
    assert torch.equal(got[2], adv) and torch.equal(got[3], cadv)
    one = _ops().discounted_returns_dual(r.to(dev), None, v.to(dev), None, m.to(dev), 0.99)
    assert torch.equal(one[0].cpu(), ret) and one[1] is None and torch.equal(one[2].cpu(), adv) and one[3] is None
    lib = C.CDLL(oracle_c.build())
    out = np.zeros((T + 1, N), dtype=np.float32)
    fp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    rn, vn, mn = (np.ascontiguousarray(t.reshape(t.shape[0], N).numpy()) for t in (r, v, m))
    lib.discounted_returns_f32(fp(rn), fp(vn), fp(mn), fp(out), C.c_int(T), C.c_int(N), C.c_double(0.99))
    assert np.array_equal(out, ret.reshape(T + 1, N).numpy())


def test_storage_use_gae_false(dev):
    from safevla_b200.storage import B200RolloutStorage
    from safevla_b200.synthetic import RolloutSpec, make_rollout
    T, N = 12, 4
    ro = make_rollout(RolloutSpec(T, N, 6, 1, episode_end_prob=0.2, seed=3))
    g = torch.Generator().manual_seed(0)
    vp, cvp = torch.randn(T + 1, N, 1, generator=g), torch.randn(T + 1, N, 1, generator=g).abs()
    st = B200RolloutStorage(T, dev)
    st.load_rollout(ro, vp, cvp, torch.zeros(T, N))
    st.before_updates(next_value=vp[T], next_c_value=cvp[T], use_gae=False, gamma=0.97)
    ret, adv = TO.gae_returns(ro["rewards"], vp, ro["masks"], 0.97, 0.95, use_gae=False)
    cret, cadv = TO.gae_returns(ro["costs"], cvp, ro["masks"], 0.97, 0.95, use_gae=False)
    assert torch.equal(st.returns.cpu(), ret) and torch.equal(st.adv_targ.cpu(), adv)
    assert torch.equal(st.c_returns.cpu(), cret) and torch.equal(st.c_adv_targ.cpu(), cadv)


# ------------------------------------------------------------------------------------------ split-operand GEMM
@pytest.mark.parametrize("rows,cols,ld", [(64, 64, 64), (1000, 512, 512), (300, 128, 640), (5, 2048, 2048)])
def test_split_concat_parts(dev, rows, cols, ld):
    """p0 + p1 (+ p2) reproduces x to 2^-16 (2^-24) relative; layouts of both concatenation axes."""
    g = torch.Generator().manual_seed(rows + cols)
    base = (torch.randn(rows, ld, generator=g) * torch.exp(3 * torch.randn(rows, ld, generator=g))).to(dev)
    x = base[:, :cols]
    for axis in (1, 0):
        out = _ops().split_concat(x, ld, rows, cols, axis, (0, 1, 2, 1))
        parts = [out[:, j * cols:(j + 1) * cols] if axis == 1 else out[j * rows:(j + 1) * rows] for j in range(4)]
        p0, p1, p2 = parts[0].float(), parts[1].float(), parts[2].float()
        assert torch.equal(parts[3], parts[1])
        assert torch.equal(p0, x.bfloat16().float())
        assert torch.equal(p1, (x - p0).bfloat16().float())
        assert torch.equal(p2, (x - p0 - p1).bfloat16().float())
        assert ((p0 + p1 - x).abs() <= 2.0 ** -16 * x.abs()).all()
        assert ((p0 + p1 + p2 - x).abs() <= 2.0 ** -23 * x.abs() + 1e-38).all()


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (1872, 1536, 512), (1000, 384, 2048), (117 * 64, 512, 512),
                                   (300, 64, 100)])
@pytest.mark.parametrize("ta,tb", [(False, True), (False, False), (True, False)])
@pytest.mark.parametrize("split,tol", [(3, 3e-5), (6, 5e-6)])
def test_gemm_split_operand_matches_fp64(dev, M, N, K, ta, tb, split, tol):
    """fp32 operands through the bf16 tcgen05 kernels as 3 (6) split products in ONE launch, against fp64:
    error relative to sum_k |a||b| (the fp32-FMA kernel itself sits at ~1e-6 on this measure)."""
    L = _L()
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn((K, M) if ta else (M, K), generator=g).to(dev)
    b = torch.randn((N, K) if tb else (K, N), generator=g).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    res = torch.randn(M, N, generator=g).to(dev)
    out = torch.empty(M, N, device=dev)
    l0 = L.load_library().svla_launch_count()
    _ops().gemm(a, b, out, trans_a=ta, trans_b=tb, bias=bias, residual=res, epilogue=L.EPI_NONE, split=split)
    A64, B64 = (a.t() if ta else a).double(), (b.t() if tb else b).double()
    ref = A64 @ B64 + bias.double() + res.double()
    scale = (A64.abs() @ B64.abs()).max().item()
    err = (out.double() - ref).abs().max().item() / scale
    assert err < tol, err
    if K % 8 == 0 and M >= 64:  # two staging launches + ONE GEMM launch (+ the split-K fold of weight-gradient shapes)
        assert L.load_library().svla_launch_count() - l0 in (3, 4)
    # weight-gradient form: accumulate + fused bias gradient stays exact
    if ta and not tb:
        acc = torch.randn(M, N, generator=g).to(dev)
        out2, cs = acc.clone(), torch.zeros(M, device=dev)
        _ops().gemm(a, b, out2, trans_a=True, trans_b=False, accumulate=True, colsum_a=cs, split=split)
        assert ((out2.double() - (acc.double() + A64 @ B64)).abs().max().item() / scale) < tol
        assert torch.allclose(cs, a.sum(0), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("mode,S,B", [(0, 117, 5), (0, 128, 2), (1, 128, 3), (1, 16, 1), (0, 33, 40), (0, 1, 3)])
@pytest.mark.parametrize("adjacent", [True, False])
def test_attention_split_operand_matches_fp64(dev, mode, S, B, adjacent):
    """Parity-grade attention (fp32 q / k / v as (hi, lo) bf16 pairs, three tcgen05 products per matmul, P / dS split
    in registers) against an fp64 reference: forward, log-sum-exp and all three gradients to ~1e-5 -- three orders of
    magnitude tighter than the bf16 kernels, on the same tensor cores."""
    H, D = 8, 512
    g = torch.Generator().manual_seed(S * 11 + mode)
    if adjacent:
        qkv = (torch.randn(B * S, 3 * D, generator=g) * 0.7).to(dev)
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    else:
        q, k, v = [(torch.randn(B * S, D, generator=g) * 0.7).to(dev) for _ in range(3)]
    do = torch.randn(B * S, D, generator=g).to(dev)
    traj = mask = None
    if mode == 1:
        traj = torch.cumsum((torch.rand(B, S, generator=g) < 0.1).long(), 1).to(dev)
        mask = torch.tril(traj[:, :, None] == traj[:, None, :]).unsqueeze(1)
    o, lse = torch.empty(B * S, D, device=dev), torch.empty(B * H * S, device=dev)
    l0 = _L().load_library().svla_launch_count()
    _ops().attn_fwd(mode, q, k, v, o, lse, B, S, traj=traj, split=3)
    assert _L().load_library().svla_launch_count() - l0 == (2 if adjacent else 4)  # staging + ONE attention launch
    qr, kr, vr = [t.double().clone().requires_grad_(True) for t in (q, k, v)]
    ref = _attn_ref(qr, kr, vr, B, S, H, 0.125, mask, None)
    assert relerr(o, ref) < 2e-5, relerr(o, ref)
    ref.backward(do.double())
    dq, dk, dv = [torch.empty(B * S, D, device=dev) for _ in range(3)]
    _ops().attn_bwd(mode, q, k, v, o, do, dq, dk, dv, lse, B, S, traj=traj, split=3)
    for name, got, exp in (("dq", dq, qr.grad), ("dk", dk, kr.grad), ("dv", dv, vr.grad)):
        assert relerr(got, exp) < 3e-5, (name, relerr(got, exp))


@pytest.mark.parametrize("M,N,K", [(512, 2048, 512), (117 * 70, 512, 384), (300, 256, 64), (4096 + 77, 2048, 512),
                                   (600, 576, 512),    # 18 record words per row: the 8-byte record path, ragged last tile
                                   (1000, 640, 256)])  # 20 words: 16-byte records, last tile half empty
def test_gemm_relu_bit_record_epilogues(dev, M, N, K):
    """RELU_BITS writes relu(x W^T + b) and one bit per element; MASK_BITS applies that record in the dgrad: both
    bit-identical to the RELU / RELU_MASK epilogues they replace (the record is 1/16 of the mask operand's bytes)."""
    L = _L()
    g = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, K, generator=g).to(dev, torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev, torch.bfloat16)
    b = torch.randn(N, generator=g).to(dev)
    ref = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    _ops().gemm(x, w, ref, bias=b, epilogue=L.EPI_RELU)
    out = torch.empty_like(ref)
    bits = torch.full((M, N // 32), -1, device=dev, dtype=torch.int32)
    _ops().gemm(x, w, out, bias=b, epilogue=L.EPI_RELU_BITS, aux=bits)
    assert torch.equal(out, ref)
    # decode the record: element e of word n / 32 sits at bit (e >> 1) + 16 (e & 1)
    e = torch.arange(32, device=dev)
    pos = (e >> 1) + 16 * (e & 1)
    dec = ((bits.to(torch.int64).unsqueeze(-1) >> pos) & 1).reshape(M, N).bool()
    assert torch.equal(dec, ref > 0)
    dy = torch.randn(M, N, generator=g).to(dev, torch.bfloat16)   # a dgrad whose OUTPUT has the activation's shape
    w2 = (torch.randn(K, N, generator=g) / K ** 0.5).to(dev, torch.bfloat16)  # dX[M, N] = dY2[M, K] W2[K, N]
    dy2 = torch.randn(M, K, generator=g).to(dev, torch.bfloat16)
    d_ref, d_out = torch.empty(M, N, device=dev, dtype=torch.bfloat16), torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    _ops().gemm(dy2, w2, d_ref, trans_b=False, aux=ref, epilogue=L.EPI_RELU_MASK)
    _ops().gemm(dy2, w2, d_out, trans_b=False, aux=bits, epilogue=L.EPI_MASK_BITS)
    assert torch.equal(d_out, d_ref)
    del dy
    # shapes the tensor-core path does not take are refused loudly (callers fall back to RELU / RELU_MASK)
    small = torch.empty(8, N, device=dev, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        _ops().gemm(x[:8], w, small, bias=b, epilogue=L.EPI_RELU_BITS, aux=bits[:8].contiguous())
