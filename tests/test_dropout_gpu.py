"""Training-mode dropout of the fusion block (reference: nn.TransformerEncoderLayer p = 0.1 live in train mode,
allenact_dino_transformer.py:193,545-552): the counter-based Philox4x32-7 masks (csrc/philox.cuh)
checked statistically, every fused site against torch with the SAME mask (dumped through svla_dropout_rows), and the
whole three-tower forward / loss / backward against the CPU oracle with the masks injected.  B200 only."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import torch_oracle as TO  # noqa: E402  (checker only)
from safevla_b200.params import TOWERS, init_state_dict  # noqa: E402
from safevla_b200.synthetic import RolloutSpec, make_rollout, prev_actions_from  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def relerr(a, b):
    return ((a.double().cpu() - b.double().cpu()).abs().max() / (b.double().abs().max() + 1e-30)).item()


def _mask(dev, rows, cols, p, seed, site, step, row0=0):
    """The pre-scaled mask (keep / (1 - p)) of a site: the stand-alone kernel applied to ones."""
    from safevla_b200 import ops
    ones = torch.ones(rows, cols, device=dev)
    return ops.dropout_rows(ones, torch.empty_like(ones), ops.dropout_spec(p, seed, site, step, row0))


@pytest.mark.parametrize("p", [0.1, 0.5, 0.03])
def test_dropout_masks_statistics_and_determinism(dev, p):
    rows, cols = 4096, 2048
    n = rows * cols
    m = _mask(dev, rows, cols, p, seed=1234, site=7, step=3)
    keep = m > 0
    thr = round(p * 65536) / 65536  # the realised drop probability (16-bit threshold)
    rate = keep.float().mean().item()
    assert abs(rate - (1 - thr)) < 5 * math.sqrt(thr * (1 - thr) / n), rate           # keep-rate within 5 sigma
    assert torch.all((m == 0) | ((m - 1 / (1 - p)).abs() < 1e-6))                       # kept values = 1 / (1 - p)
    # mean / variance preservation on data: E[drop(x)] = x, Var adds x^2 p / (1 - p)
    x = torch.randn(rows, cols, device=dev)
    y = x * m
    assert abs((y - x).mean().item()) < 5 * math.sqrt(p / (1 - p) / n)
    assert abs((y * y).mean().item() / (x * x).mean().item() - 1 / (1 - p)) < 0.01
    # no structure: per-row / per-column keep rates, neighbour correlations, the eight lanes of a Philox group
    sig_r, sig_c = math.sqrt(thr * (1 - thr) / cols), math.sqrt(thr * (1 - thr) / rows)
    assert (keep.float().mean(1) - (1 - thr)).abs().max().item() < 6 * sig_r
    assert (keep.float().mean(0) - (1 - thr)).abs().max().item() < 6 * sig_c
    k = keep.float() - rate
    for a, b in ((k[:, :-1], k[:, 1:]), (k[:-1], k[1:]), (k[:, :-8], k[:, 8:])):
        corr = (a * b).mean().item() / (thr * (1 - thr))
        assert abs(corr) < 6 / math.sqrt(n), corr
    lanes = keep.view(rows, cols // 8, 8).float().mean((0, 1))
    assert (lanes - (1 - thr)).abs().max().item() < 6 * math.sqrt(thr * (1 - thr) / (n / 8))
    # determinism and sensitivity to every field of the counter / key
    assert torch.equal(m, _mask(dev, rows, cols, p, 1234, 7, 3))
    for other in (_mask(dev, rows, cols, p, 1235, 7, 3), _mask(dev, rows, cols, p, 1234, 8, 3),
                  _mask(dev, rows, cols, p, 1234, 7, 4), _mask(dev, rows, cols, p, 1234 + (1 << 32), 7, 3)):
        agree = ((other > 0) == keep).float().mean().item()
        assert abs(agree - (thr * thr + (1 - thr) ** 2)) < 0.002, agree                # independent masks
    # row chunks of one logical tensor address the same mask
    assert torch.equal(_mask(dev, 100, cols, p, 1234, 7, 3, row0=900), m[900:1000])


def test_dropout_rows_dtypes_strides_inplace(dev):
    from safevla_b200 import ops
    spec = ops.dropout_spec(0.1, 5, 1, 1, 64)
    x = torch.randn(300, 512, device=dev)
    ref = x * _mask(dev, 364, 512, 0.1, 5, 1, 1)[64:]
    out = ops.dropout_rows(x, torch.empty_like(x), spec)
    assert torch.equal(out, ref)
    xb = x.bfloat16()
    ob = ops.dropout_rows(xb, torch.empty_like(xb), spec)
    assert torch.equal(ob, (xb.float() * _mask(dev, 364, 512, 0.1, 5, 1, 1)[64:]).bfloat16())
    big = torch.randn(300, 3 * 512, device=dev).bfloat16()
    view = big[:, 512:1024]
    exp = (view.float() * _mask(dev, 364, 512, 0.1, 5, 1, 1)[64:]).bfloat16()
    ops.dropout_rows(view, view, spec)  # in place on a strided view
    assert torch.equal(big[:, 512:1024], exp)
    assert torch.equal(ops.dropout_rows(x, torch.empty_like(x), ops.dropout_spec(0.0, 5, 1, 1)), x)  # p = 0: identity


@pytest.mark.parametrize("M,N,K", [(117 * 8, 2048, 512), (4096 + 77, 2048, 512), (512, 512, 384)])
def test_gemm_relu_bits_fused_dropout(dev, M, N, K):
    """linear1's epilogue with the FFN dropout fused behind the ReLU: equals relu(x W^T + b) * mask for the mask
    svla_dropout_rows reports; the bit record marks exactly the surviving positive elements, so the masked dgrad with
    alpha = 1 / (1 - p) is the dropout + ReLU backward."""
    from safevla_b200 import _lib as L
    from safevla_b200 import ops
    p = 0.1
    g = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, K, generator=g).to(dev, torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev, torch.bfloat16)
    b = torch.randn(N, generator=g).to(dev)
    spec = ops.dropout_spec(p, 99, 2 * 8 + 2, 5, row0=1000)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    bits = torch.empty(M, N // 32, device=dev, dtype=torch.int32)
    ops.gemm(x, w, out, bias=b, epilogue=L.EPI_RELU_BITS, aux=bits, dropout=spec)
    mask = _mask(dev, M, N, p, 99, 2 * 8 + 2, 5, row0=1000)
    ref = torch.relu(x.float() @ w.float().t() + b) * mask
    assert relerr(out.float(), ref) < 1e-2
    pre = x.float() @ w.float().t() + b
    sure = pre.abs() > 1e-2  # away from the ReLU boundary (the two accumulation orders may disagree on the sign there)
    assert torch.equal((out == 0)[sure], ((pre <= 0) | (mask == 0))[sure])
    e = torch.arange(32, device=dev)
    dec = ((bits.to(torch.int64).unsqueeze(-1) >> ((e >> 1) + 16 * (e & 1))) & 1).reshape(M, N).bool()
    assert torch.equal(dec, out > 0)
    dy2 = torch.randn(M, K, generator=g).to(dev, torch.bfloat16)
    w2 = (torch.randn(K, N, generator=g) / K ** 0.5).to(dev, torch.bfloat16)
    dz = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(dy2, w2, dz, trans_b=False, aux=bits, epilogue=L.EPI_MASK_BITS, alpha=1 / (1 - p))
    dz_ref = (dy2.float() @ w2.float()) * (out > 0) / (1 - p)
    assert relerr(dz.float(), dz_ref) < 1e-2


@pytest.mark.parametrize("S,B", [(117, 5), (128, 2), (33, 9), (201, 3), (256, 2), (129, 2)])
def test_attention_dropout_matches_torch_with_same_mask(dev, S, B):
    """Dropout on the attention probabilities inside the tcgen05 kernels (forward and backward regenerate the mask)
    against torch autograd with the dumped mask: O = (softmax(S) * M) V.  S <= 128: warp-specialised kernels, mask
    rows 128 wide; 128 < S <= 256 (two-camera fusion block): the two-tile kernels, mask rows 256 wide."""
    from safevla_b200 import ops
    H, D, p = 8, 512, 0.1
    g = torch.Generator().manual_seed(S)
    qkv = (torch.randn(B * S, 3 * D, generator=g) * 0.6).to(dev, torch.bfloat16)
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    do = torch.randn(B * S, D, generator=g).to(dev, torch.bfloat16)
    W = 128 if S <= 128 else 256  # mask rows per (sequence, head) and mask width
    spec = ops.dropout_spec(p, 7, 1 * 64 + 0 * 8 + 0, 2, row0=3 * H * W)
    o, lse = torch.empty(B * S, D, device=dev, dtype=torch.bfloat16), torch.empty(B * H * S, device=dev)
    ops.attn_fwd(0, q, k, v, o, lse, B, S, drop=spec)
    m = _mask(dev, B * H * W, W, p, 7, 64, 2, row0=3 * H * W).view(B, H, W, W)[:, :, :S, :S]
    qr, kr, vr = [t.float().clone().requires_grad_(True) for t in (q, k, v)]
    sp = lambda t: t.view(B, S, H, 64).transpose(1, 2)  # noqa: E731
    P = torch.softmax(sp(qr) @ sp(kr).transpose(-1, -2) * 0.125, -1)
    ref = ((P * m) @ sp(vr)).transpose(1, 2).reshape(B * S, D)
    assert relerr(o.float(), ref) < 2e-2
    # the saved log-sum-exp is that of the UNDROPPED scores
    lse_ref = torch.logsumexp(sp(qr) @ sp(kr).transpose(-1, -2) * 0.125, -1)
    assert (lse.view(B, H, S) - lse_ref).abs().max().item() < 2e-2
    ref.backward(do.float())
    dqkv = torch.empty(B * S, 3 * D, device=dev, dtype=torch.bfloat16)
    ops.attn_bwd(0, q, k, v, o, do, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], lse, B, S, drop=spec)
    tol = 3e-2 if S <= 128 else 6e-2  # the two-tile backward takes delta from the bf16 O . dO
    for name, got, exp in (("dq", dqkv[:, :D], qr.grad), ("dk", dqkv[:, D:2 * D], kr.grad), ("dv", dqkv[:, 2 * D:], vr.grad)):
        assert relerr(got.float(), exp) < tol, (name, relerr(got.float(), exp))
    # p = 0 through the dropout entry point == the plain kernel
    o0, o1 = torch.empty_like(o), torch.empty_like(o)
    ops.attn_fwd(0, q, k, v, o0, lse, B, S)
    ops.attn_fwd(0, q, k, v, o1, lse, B, S, drop=ops.dropout_spec(0.0, 7, 64, 2))
    assert torch.equal(o0, o1)


def _oracle_masks(dev, model, R, S, step):
    """drop(prefix, layer, kind) -> the pre-scaled CPU mask of that site, dumped from the device generator."""
    p, seed, cache = model.dropout, model.dropout_seed, {}

    def drop(prefix, layer, kind):
        key = (prefix, layer, kind)
        if key not in cache:
            site = TOWERS.index(prefix) * 64 + layer * 8 + kind
            if kind == 0:
                W = 128 if S <= 128 else 256
                m = _mask(dev, R * 8 * W, W, p, seed, site, step).view(R, 8, W, W)[:, :, :S, :S]
            else:
                m = _mask(dev, R * S, 2048 if kind == 2 else 512, p, seed, site, step).view(R, S, -1)
            cache[key] = m.cpu()
        return cache[key]
    return drop


@pytest.mark.parametrize("chunk_rows,budget,C", [(1024, 100 << 30, 1), (6, 100 << 30, 1), (4, 0, 1), (1024, 100 << 30, 2),
                                                 (5, 0, 2)])
def test_model_with_dropout_matches_oracle_given_the_masks(dev, chunk_rows, budget, C):
    """Whole three-tower forward + SafePPOLogGrad + backward with dropout 0.1 in every fusion layer (all four dropout
    positions of nn.TransformerEncoderLayer) against the CPU oracle evaluated with the very masks the device generated
    -- single chunk, row-chunked (masks addressed by global rows) and recompute-in-backward (masks regenerated)."""
    from safevla_b200.losses import SafePPOLogGrad
    from safevla_b200.model import B200SafeActorCritic
    T, N, A = 6, 2, 6
    sd = init_state_dict(A, C, seed=17, actor_gain=1.0)
    ro = make_rollout(RolloutSpec(T, N, A, C, episode_end_prob=0.25, seed=3))
    obs = {k: v[:-1] for k, v in ro["observations"].items()}
    prev, masks = prev_actions_from(ro["actions"]), ro["masks"][:-1]
    model = B200SafeActorCritic(A, C, precision="bf16", state_dict=sd, device=dev, dropout=0.1, dropout_seed=77,
                                chunk_rows=chunk_rows, stash_budget_bytes=budget, extras="off")
    model.set_trainable_towers((0, 1))
    out, _ = model({k: v.to(dev) for k, v in obs.items()}, None, prev.to(dev), masks.to(dev))
    assert model.dropout_step == 1
    drop = _oracle_masks(dev, model, T * N, 1 + 84 * C + 32, step=1)
    leaf = {k: (v.clone().requires_grad_(True) if "text_encoder" not in k and not k.endswith("div_term") else v)
            for k, v in sd.items()}
    ref = TO.safe_model_forward(leaf, obs, prev, masks, A, C, drop=drop)
    for got, key in ((out.distributions.raw_logits, "logits"), (out.values, "values"), (out.c_values, "c_values")):
        scale = max(ref[key].abs().max().item(), 0.25)
        assert (got.detach().cpu() - ref[key].detach()).abs().max().item() < 4e-2 * scale, key
    # the masks matter: the undropped forward is far away from the dropped one
    ref0 = TO.safe_model_forward(sd, obs, prev, masks, A, C)
    assert (ref0["values"] - ref["values"].detach()).abs().max().item() > 5 * (
        out.values.detach().cpu() - ref["values"].detach()).abs().max().item()
    g = torch.Generator().manual_seed(1)
    vp, cvp = torch.randn(T + 1, N, 1, generator=g), torch.randn(T + 1, N, 1, generator=g).abs()
    ret, adv = TO.gae_returns(ro["rewards"], vp, ro["masks"], 0.99, 0.95)
    _, cadv = TO.gae_returns(ro["costs"], cvp, ro["masks"], 0.99, 0.95)
    old_logp = torch.log_softmax(ref["logits"].detach(), -1).gather(-1, ro["actions"].unsqueeze(-1)).squeeze(-1) + 0.1
    loss = SafePPOLogGrad(clip_param=0.1, value_loss_coef=0.5, entropy_coef=0.01, use_clipped_value_loss=False,
                          action_loss_schedule=None, discrete_critics=False, normalize_advantage=False)
    batch = {"actions": ro["actions"].to(dev), "old_action_log_probs": old_logp.to(dev), "adv_targ": adv.to(dev),
             "c_adv_targ": cadv.to(dev), "values": vp[:-1].to(dev), "returns": ret[:-1].to(dev)}
    total, _ = loss.loss(0, batch, out, lagrangian_multiplier=torch.tensor(0.3))
    total.backward()
    ref_total, _ = TO.safe_ppo_log_grad(ref["logits"], ro["actions"], old_logp, adv, cadv, ref["values"], ret[:-1], 0.3,
                                        entropy_coef=0.01)
    ref_total.backward()
    assert abs(total.item() - ref_total.item()) < 2e-2 * max(1.0, abs(ref_total.item()))
    worst, cos_min = ("", 0.0), ("", 1.0)
    for k, v in leaf.items():
        if not (torch.is_tensor(v) and v.requires_grad) or v.grad is None or k.startswith("c_critic_tsfm."):
            continue
        gm, gr = model.get_parameter(k).grad.detach().cpu().double().reshape(-1), v.grad.double().reshape(-1)
        if gr.norm() < 1e-12:
            continue
        err = abs(gm.norm().item() - gr.norm().item()) / gr.norm().item()
        cos = (gm @ gr / (gm.norm() * gr.norm())).item()
        if err > worst[1]:
            worst = (k, err)
        if cos < cos_min[1]:
            cos_min = (k, cos)
    # measured (tools/dropout_grad_probe.py, 12 rows): norm errors <= 0.5 %, worst cosine 0.992 (linear1 of one layer)
    # -- the same as the bf16 path WITHOUT dropout shows on this tiny batch (0.994: operand rounding + ReLU-boundary
    # flips); a wrong mask or a missing 1 / (1 - p) anywhere would be a 10 % norm error or a cosine far below 0.9
    # two cameras (S = 201): the two-tile attention backward takes delta from the bf16 O . dO, its bound is wider
    assert worst[1] < (2e-2 if C == 1 else 8e-2), worst
    assert cos_min[1] > (0.985 if C == 1 else 0.97), cos_min


def test_dropout_zero_is_the_parity_path_and_modes(dev):
    from safevla_b200.model import B200SafeActorCritic
    T, N, A, C = 6, 2, 6, 1
    sd = init_state_dict(A, C, seed=17, actor_gain=1.0)
    ro = make_rollout(RolloutSpec(T, N, A, C, episode_end_prob=0.25, seed=3))
    obs = {k: v[:-1].to(dev) for k, v in ro["observations"].items()}
    prev, masks = prev_actions_from(ro["actions"]).to(dev), ro["masks"][:-1].to(dev)
    base = B200SafeActorCritic(A, C, precision="bf16", state_dict=sd, device=dev, extras="off")
    zero = B200SafeActorCritic(A, C, precision="bf16", state_dict=sd, device=dev, extras="off", dropout=0.0)
    drop = B200SafeActorCritic(A, C, precision="bf16", state_dict=sd, device=dev, extras="off", dropout=0.1)
    o_b, _ = base(obs, None, prev, masks)
    o_z, _ = zero(obs, None, prev, masks)
    assert torch.equal(o_b.values, o_z.values) and torch.equal(o_b.distributions.raw_logits, o_z.distributions.raw_logits)
    o_1, _ = drop(obs, None, prev, masks)
    drop._ctx_cache = None
    o_2, _ = drop(obs, None, prev, masks)  # a new forward draws new masks (dropout_step advanced)
    assert not torch.equal(o_1.values, o_2.values) and drop.dropout_step == 2
    with torch.no_grad():  # collection / evaluation: no dropout, equals the parity path
        o_n, _ = drop(obs, None, prev, masks)
    assert torch.equal(o_n.values, o_b.values)
    drop.eval()
    o_e, _ = drop(obs, None, prev, masks)
    assert torch.equal(o_e.values.detach(), o_b.values.detach())
    with pytest.raises(NotImplementedError):
        B200SafeActorCritic(A, C, precision="fp32", state_dict=sd, device=dev, dropout=0.1)


@pytest.mark.parametrize("T,N,C", [(6, 2, 1), (32, 8, 1), (6, 2, 2)])
def test_cls_only_last_layer_under_dropout_equals_the_full_layer(dev, T, N, C):
    """The CLS-row shortcut of the last fusion layer (K/V for every token, everything else for row 0) addresses the
    full layer's masks (attention row h*128 of each sequence, sub-layer rows r*S): same outputs and gradients as
    running the full layer with the same seed.  (32, 8): 256 CLS rows, the fused bit-record FFN path; (6, 2): the
    small-chunk fallback (ReLU launch + row pass)."""
    from safevla_b200.model import B200SafeActorCritic
    A = 6
    sd = init_state_dict(A, C, seed=17, actor_gain=1.0)
    ro = make_rollout(RolloutSpec(T, N, A, C, episode_end_prob=0.1, seed=3))
    obs = {k: v[:-1].to(dev) for k, v in ro["observations"].items()}
    prev, masks = prev_actions_from(ro["actions"]).to(dev), ro["masks"][:-1].to(dev)
    res = []
    for cls_only in (True, False):
        m = B200SafeActorCritic(A, C, precision="bf16", state_dict=sd, device=dev, dropout=0.1, dropout_seed=5,
                                extras="off", cls_only_last_layer=cls_only)
        m.set_trainable_towers((0, 1))
        out, _ = m(obs, None, prev, masks)
        g = torch.Generator().manual_seed(2)
        w = torch.randn(T, N, A, generator=g).to(dev)
        ((out.distributions.raw_logits * w).sum() + out.values.sum() * 0.3).backward()
        res.append((out.distributions.raw_logits.detach().clone(), out.values.detach().clone(), m.grad_arena.clone()))
    (l1, v1, g1), (l0, v0, g0) = res
    assert (l1 - l0).abs().max().item() < 2e-2 * max(l0.abs().max().item(), 0.25)
    assert (v1 - v0).abs().max().item() < 2e-2 * max(v0.abs().max().item(), 0.25)
    cos = (g1.double() @ g0.double() / (g1.double().norm() * g0.double().norm())).item()
    assert cos > 0.995 and abs(g1.norm().item() / g0.norm().item() - 1) < 1e-2, (cos, g1.norm().item(), g0.norm().item())


def test_update_with_dropout_is_deterministic_given_the_seed(dev):
    from safevla_b200.model import B200SafeActorCritic
    from safevla_b200.storage import B200RolloutStorage
    from safevla_b200.updater import PPOLagConfig, PPOLagUpdater
    T, N, A, C = 8, 4, 6, 1
    sd = init_state_dict(A, C, seed=21, actor_gain=1.0)
    ro = make_rollout(RolloutSpec(T, N, A, C, episode_end_prob=0.2, seed=77))
    g = torch.Generator().manual_seed(5)
    vp, cvp = torch.randn(T + 1, N, 1, generator=g), torch.randn(T + 1, N, 1, generator=g).abs()
    logp = -1.7 + 0.1 * torch.randn(T, N, generator=g)
    finals = []
    for seed in (11, 11, 12):
        model = B200SafeActorCritic(A, C, precision="bf16", state_dict=sd, device=dev, dropout=0.1, dropout_seed=seed,
                                    extras="off")
        st = B200RolloutStorage(T, dev)
        st.load_rollout(ro, vp, cvp, logp)
        res = PPOLagUpdater(model, PPOLagConfig(update_repeats=2, lr=1e-3)).update(st)
        assert torch.isfinite(res["loss_scalars"]).all() and model.dropout_step == 2
        finals.append(model.param_arena.clone())
    assert torch.equal(finals[0], finals[1]) and not torch.equal(finals[0], finals[2])
