"""End-to-end parity of the CUDA three-tower model / loss / update against the reference's golden
vectors (tests/golden/*.pt, minted by oracle/make_golden.py from the unmodified reference code)
and against the CPU oracle on fresh seeds.  B200 only."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import torch_oracle as TO  # noqa: E402  (checker only)
from oracle.make_golden import CASES, GOLDEN_DIR, build_inputs  # noqa: E402
from safevla_b200.params import init_state_dict  # noqa: E402
from safevla_b200.synthetic import prev_actions_from  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def relerr(a, b):
    return ((a.double().cpu() - b.double().cpu()).abs().max() / (b.double().abs().max() + 1e-30)).item()


def _setup(name, dev, precision, **kw):
    from safevla_b200.model import B200SafeActorCritic
    gold = torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)
    case = gold["case"]
    sd = init_state_dict(case["A"], case["C"], case["wseed"], actor_gain=1.0)
    model = B200SafeActorCritic(case["A"], case["C"], precision=precision, state_dict=sd, device=dev, **kw)
    spec, ro, extra = build_inputs(case)
    obs = {k: v[:-1].to(dev) for k, v in ro["observations"].items()}
    return gold, case, model, ro, extra, obs


def _batch(gold, ro, extra, dev):
    ret, adv = TO.gae_returns(ro["rewards"], extra["value_preds"], ro["masks"], 0.99, 0.95)
    cret, cadv = TO.gae_returns(ro["costs"], extra["c_value_preds"], ro["masks"], 0.99, 0.95)
    return {"actions": ro["actions"].to(dev), "old_action_log_probs": gold["old_logp"].to(dev),
            "adv_targ": adv.to(dev), "c_adv_targ": cadv.to(dev), "values": extra["value_preds"][:-1].to(dev),
            "returns": ret[:-1].to(dev), "c_returns": cret[:-1].to(dev),
            "c_values": extra["c_value_preds"][:-1].to(dev)}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("cls_only", [True, False])
def test_golden_forward_loss_backward_fp32(dev, name, cls_only):
    """BASELINE gate: logits / losses within 1e-4 relative of the reference PyTorch path (fp32)."""
    from safevla_b200.losses import SafePPOLogGrad
    gold, case, model, ro, extra, obs = _setup(name, dev, "fp32", cls_only_last_layer=cls_only)
    out, _ = model(obs, None, prev_actions_from(ro["actions"]).to(dev), ro["masks"][:-1].to(dev))
    assert relerr(out.distributions.raw_logits, gold["logits"]) < 1e-4
    assert relerr(out.distributions.logits, gold["log_probs"]) < 1e-4
    assert relerr(out.values, gold["values"]) < 1e-4
    assert relerr(out.c_values, gold["c_values"]) < 1e-4
    loss = SafePPOLogGrad(clip_param=0.1, value_loss_coef=0.5, entropy_coef=0.01, use_clipped_value_loss=False,
                          action_loss_schedule=None, discrete_critics=False, normalize_advantage=False)
    batch = _batch(gold, ro, extra, dev)
    total, info = loss.loss(0, batch, out, lagrangian_multiplier=torch.tensor(case["lam"]))
    for k in ("ppo_total", "value", "action", "entropy"):
        assert abs(info[k] - gold["info"][k]) <= 1e-4 * max(1.0, abs(gold["info"][k])), (k, info[k], gold["info"][k])
    assert set(gold["info"]) <= set(info)  # same info-dict keys as the reference
    total.backward()
    worst = ("", 0.0)
    for k, gn in gold["grad_norms"].items():
        p = model.get_parameter(k)
        if gn is None:  # reference leaves these without a gradient (cost tower, unused heads)
            assert p.grad is None or p.grad.abs().max().item() == 0.0, k
            continue
        mine = p.grad.norm().item()
        err = abs(mine - gn) / max(gn, 1e-8)
        if err > worst[1]:
            worst = (k, err)
    assert worst[1] < 2e-3, worst
    for k, gref in gold["grads"].items():
        assert relerr(model.get_parameter(k).grad, gref) < 2e-3, k


def test_golden_stage0_value_losses(dev):
    from safevla_b200.losses import PPOValue, SafePPOValue
    name = "cfg1_T16_N1_A6_C1"
    gold, case, model, ro, extra, obs = _setup(name, dev, "fp32")
    model.set_trainable_towers((1, 2))
    out, _ = model(obs, None, prev_actions_from(ro["actions"]).to(dev), ro["masks"][:-1].to(dev))
    batch = _batch(gold, ro, extra, dev)
    l1, _ = PPOValue(clip_param=0.1, use_clipped_value_loss=False).loss(0, batch, out)
    l2, _ = SafePPOValue(clip_param=0.1, use_clipped_value_loss=False).loss(0, batch, out)
    total = l1 + l2
    assert abs(total.item() - gold["stage0_loss"].item()) < 1e-4 * gold["stage0_loss"].item()
    total.backward()
    for k, gn in gold["stage0_grad_norms"].items():
        p = model.get_parameter(k)
        if gn is None:
            assert p.grad is None or p.grad.abs().max().item() == 0.0, k
        else:
            assert abs(p.grad.norm().item() - gn) / max(gn, 1e-8) < 2e-3, k


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "bf16"])
def test_discrete_critic_matches_reference_golden(dev, precision):
    """critic_type="discrete" end to end: DiscreteCriticHead in every tower, extras["full_logits" | "loss_func" |
    "stop_grad_logits"] (allenact_dino_transformer.py:434-439), SafePPOLogGrad(discrete_critics=True)
    (customized_loss.py:364-370) -- forward, loss terms and every parameter-gradient norm against the reference."""
    from oracle.make_golden import DISCRETE_CASES
    from safevla_b200.losses import HLGaussLoss, SafePPOLogGrad
    from safevla_b200.model import B200SafeActorCritic
    (name, case), = DISCRETE_CASES.items()
    gold = torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)
    sd = init_state_dict(case["A"], case["C"], case["wseed"], actor_gain=1.0, critic_type="discrete")
    model = B200SafeActorCritic(case["A"], case["C"], precision=precision, state_dict=sd, device=dev, critic_type="discrete")
    assert set(model.state_dict()) == set(sd)
    spec, ro, extra = build_inputs(case)
    obs = {k: v[:-1].to(dev) for k, v in ro["observations"].items()}
    out, _ = model(obs, None, prev_actions_from(ro["actions"]).to(dev), ro["masks"][:-1].to(dev))
    tol = {"fp32": 1e-4, "bf16x3": 1e-4, "bf16": 4e-2}[precision]
    ex = out.extras
    assert isinstance(ex["loss_func"], HLGaussLoss) and torch.equal(ex["stop_grad_logits"], ex["full_logits"].detach())
    for got, key in ((out.distributions.raw_logits, "logits"), (out.values, "values"), (out.c_values, "c_values"),
                     (ex["full_logits"], "full_logits")):
        assert got.shape == gold[key].shape and relerr(got, gold[key]) < tol, (key, relerr(got, gold[key]))
    loss = SafePPOLogGrad(clip_param=0.1, value_loss_coef=0.5, entropy_coef=0.01, use_clipped_value_loss=False,
                          action_loss_schedule=None, discrete_critics=True, normalize_advantage=False)
    total, info = loss.loss(0, _batch(gold, ro, extra, dev), out, lagrangian_multiplier=torch.tensor(case["lam"]))
    ltol = 1e-4 if precision != "bf16" else 5e-3
    for k in ("ppo_total", "value", "action", "entropy"):
        assert abs(info[k] - gold["info"][k]) <= ltol * max(1.0, abs(gold["info"][k])), (k, info[k], gold["info"][k])
    total.backward()
    gtol = 2e-3 if precision != "bf16" else 5e-2
    for k, gn in gold["grad_norms"].items():
        p = model.get_parameter(k)
        if gn is None:  # incl. the whole reward-critic tower: the value term trains the COST tower's head
            assert p.grad is None or p.grad.abs().max().item() == 0.0, k
        else:
            assert abs(p.grad.norm().item() - gn) / max(gn, 1e-8) < gtol, (k, p.grad.norm().item(), gn)
    if precision == "fp32":
        for k, gref in gold["grads"].items():
            assert relerr(model.get_parameter(k).grad, gref) < 2e-3, k


def test_golden_bf16_mode(dev):
    """bf16 tensor-core mode: same path, looser tolerance (bf16 operands, fp32 accumulation)."""
    gold, case, model, ro, extra, obs = _setup("cfg1_T16_N1_A6_C1", dev, "bf16")
    with torch.no_grad():
        out, _ = model(obs, None, prev_actions_from(ro["actions"]).to(dev), ro["masks"][:-1].to(dev))
    assert relerr(out.distributions.raw_logits, gold["logits"]) < 5e-2
    assert relerr(out.values, gold["values"]) < 5e-2


def test_chunking_and_recompute_are_exact(dev):
    """Row-chunked + recompute-in-backward schedules reproduce the single-chunk stash schedule."""
    from safevla_b200.losses import SafePPOLogGrad
    name = "T12_N2_A20_C2"
    grads = []
    for kw in (dict(chunk_rows=1024), dict(chunk_rows=5), dict(chunk_rows=7, stash_budget_bytes=0)):
        gold, case, model, ro, extra, obs = _setup(name, dev, "fp32", **kw)
        out, _ = model(obs, None, prev_actions_from(ro["actions"]).to(dev), ro["masks"][:-1].to(dev))
        loss = SafePPOLogGrad(clip_param=0.1, value_loss_coef=0.5, entropy_coef=0.01, use_clipped_value_loss=False,
                              action_loss_schedule=None, discrete_critics=False, normalize_advantage=False)
        total, _ = loss.loss(0, _batch(gold, ro, extra, dev), out, lagrangian_multiplier=torch.tensor(case["lam"]))
        total.backward()
        grads.append(model.grad_arena.clone())
    assert relerr(grads[1], grads[0]) < 1e-5 and relerr(grads[2], grads[0]) < 1e-5


def test_updater_matches_oracle_update(dev):
    """Whole update (GAE -> repeats x [fwd, loss, bwd, clip, Adam] -> lambda) vs the CPU oracle."""
    from oracle.update_oracle import oracle_update
    from safevla_b200.model import B200SafeActorCritic
    from safevla_b200.storage import B200RolloutStorage
    from safevla_b200.synthetic import RolloutSpec, make_rollout
    from safevla_b200.updater import PPOLagConfig, PPOLagUpdater
    T, N, A, C = 8, 2, 6, 1
    sd = init_state_dict(A, C, seed=21, actor_gain=1.0)
    ro = make_rollout(RolloutSpec(T, N, A, C, episode_end_prob=0.2, seed=77))
    g = torch.Generator().manual_seed(5)
    vp, cvp = torch.randn(T + 1, N, 1, generator=g), torch.randn(T + 1, N, 1, generator=g).abs()
    logp = -1.7 + 0.1 * torch.randn(T, N, generator=g)
    # eps = 1e-4 keeps Adam from amplifying round-off-level gradients (e.g. the key-bias gradient, which is
    # exactly zero in exact arithmetic) into +-lr steps that differ between any two implementations
    cfg = PPOLagConfig(update_repeats=2, lr=1e-3, eps=1e-4)
    ref_sd, ref_lam, ref_info = oracle_update(sd, ro, vp, cvp, logp, cfg, A, C)
    model = B200SafeActorCritic(A, C, precision="fp32", state_dict=sd, device=dev)
    st = B200RolloutStorage(T, dev)
    st.load_rollout(ro, vp, cvp, logp)
    upd = PPOLagUpdater(model, cfg)
    res = upd.update(st)
    assert abs(res["lambda"].item() - ref_lam) < 1e-5
    assert abs(res["loss_scalars"][0].item() - ref_info["last_total"]) < 1e-3 * max(1, abs(ref_info["last_total"]))
    mine = model.state_dict()
    worst, moved = ("", 0.0), 0.0
    for k, v in ref_sd.items():
        if "text_encoder" in k:
            continue
        err = (mine[k].cpu() - v).abs().max().item()
        moved = max(moved, (v - sd[k]).abs().max().item())
        if err > worst[1]:
            worst = (k, err)
    assert moved > 5e-4  # two Adam steps of lr 1e-3 really moved the parameters
    assert worst[1] < 2e-5, worst


@pytest.mark.parametrize("stage", [1, 0])
def test_two_cost_channels_update_matches_oracle(dev, stage):
    """K = 2 cost channels (BASELINE config 5's "dual cost channels"; an extension beyond the reference's scalar cost):
    per-channel GAE, (A - sum_k lam_k A_c,k) / (1 + sum_k lam_k), a two-output cost critic and two independently
    projected multipliers -- the whole update against the CPU oracle."""
    from oracle.update_oracle import oracle_update
    from safevla_b200.model import B200SafeActorCritic
    from safevla_b200.storage import B200RolloutStorage
    from safevla_b200.synthetic import RolloutSpec, make_rollout
    from safevla_b200.updater import PPOLagConfig, PPOLagUpdater
    T, N, A, C, K = 8, 4, 6, 2, 2
    sd = init_state_dict(A, C, seed=23, actor_gain=1.0, num_cost_channels=K)
    ro = make_rollout(RolloutSpec(T, N, A, C, episode_end_prob=0.25, seed=78, num_cost_channels=K))
    assert ro["costs"].shape == (T, N, K) and ro["episode_cost_sum"].shape == (K,)
    g = torch.Generator().manual_seed(6)
    vp, cvp = torch.randn(T + 1, N, 1, generator=g), torch.randn(T + 1, N, K, generator=g).abs()
    logp = -1.7 + 0.1 * torch.randn(T, N, generator=g)
    # one limit below and one far above the observed episode costs: lambda_0 grows, lambda_1 shrinks
    cfg = PPOLagConfig(update_repeats=2, lr=1e-3, eps=1e-4, stage=stage, cost_limit=(0.05, 50.0), lambda_init=0.4)
    ref_sd, ref_lams, ref_info = oracle_update(sd, ro, vp, cvp, logp, cfg, A, C)
    model = B200SafeActorCritic(A, C, precision="fp32", state_dict=sd, device=dev, num_cost_channels=K)
    st = B200RolloutStorage(T, dev, num_cost_channels=K)
    st.load_rollout(ro, vp, cvp, logp)
    upd = PPOLagUpdater(model, cfg)
    res = upd.update(st)
    lam = res["lambda"].cpu()
    assert lam.shape == (K,) and abs(lam[0].item() - ref_lams[0]) < 1e-5 and abs(lam[1].item() - ref_lams[1]) < 1e-5
    assert lam[0].item() > 0.4 > lam[1].item()
    # per-channel GAE is the bit-exact sequential recursion
    for k in range(K):
        cret, cadv = TO.gae_returns(ro["costs"][..., k:k + 1], cvp[..., k:k + 1], ro["masks"], 0.99, 0.95)
        assert torch.equal(st.c_returns_k[k].cpu(), cret) and torch.equal(st.c_adv_targ_k[k].cpu(), cadv)
    mine = model.state_dict()
    assert mine["c_critic_tsfm.critic.fc.weight"].shape == (K, 512)
    worst, moved = ("", 0.0), 0.0
    for k, v in ref_sd.items():
        if "text_encoder" in k:
            continue
        err = (mine[k].cpu() - v).abs().max().item()
        moved = max(moved, (v - sd[k]).abs().max().item())
        if err > worst[1]:
            worst = (k, err)
    assert moved > 5e-4 and worst[1] < 2e-5, worst


def test_combine_cost_advantages_kernel(dev):
    from safevla_b200 import ops
    g = torch.Generator().manual_seed(0)
    for K, R in ((2, 4096), (3, 1028), (1, 77), (8, 64)):
        ca = torch.randn(K, R, generator=g).to(dev)
        for lam in (torch.rand(K, generator=g), torch.zeros(K)):
            out, le = ops.combine_cost_advantages(ca, lam.to(dev))
            L = lam.sum().item()
            exp = (lam.to(dev)[:, None] * ca).sum(0) / L if L > 0 else torch.zeros(R, device=dev)
            assert abs(le.item() - L) < 1e-6 and torch.allclose(out, exp, atol=1e-6, rtol=1e-5)


# ------------------------------------------------------------------------------------------ rollout-side T = 1
@pytest.mark.parametrize("name", ["step_N3_A6_C1", "step_N2_A20_C2"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_single_step_inference_matches_reference(dev, name, precision):
    """forward() with T = 1 (KV-cache decoder step, episode-start mask, position wrap, reset by an update-mode
    forward) against the reference's per-step golden outputs; sampled actions identical given the same seed."""
    from oracle.make_golden_step import build_inputs as step_inputs, schedule
    from safevla_b200.model import B200SafeActorCritic
    gold = torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)
    case = gold["case"]
    sd = init_state_dict(case["A"], case["C"], case["wseed"], actor_gain=1.0)
    model = B200SafeActorCritic(case["A"], case["C"], precision=precision, state_dict=sd, device=dev,
                                max_steps=case["max_steps"], extras="off")
    ro, prev = step_inputs(case)
    it = iter(gold["steps"])
    tol = 1e-4 if precision == "fp32" else 5e-2  # bf16 operands, fp32 accumulation
    for kind, t0, t1 in schedule(case):
        obs = {k: v[t0:t1].to(dev) for k, v in ro["observations"].items()}
        with torch.no_grad():
            out, _ = model(obs, None, prev[t0:t1].to(dev), ro["masks"][t0:t1].to(dev))
        if kind == "update":
            continue
        rec = next(it)
        assert out.distributions.raw_logits.shape == rec["logits"].shape
        assert relerr(out.distributions.raw_logits, rec["logits"]) < tol, (t0, "logits")
        for got, ref in ((out.values, rec["values"]), (out.c_values, rec["c_values"])):
            # N scalars per step: scale the bf16 tolerance by at least 0.25 so a near-zero value is not a 0/0 test
            scale = ref.abs().max().item() if precision == "fp32" else max(ref.abs().max().item(), 0.25)
            assert (got.cpu().double() - ref.double()).abs().max().item() < tol * scale, (t0, got, ref)
        if precision == "fp32":
            # action sampling: the stock multinomial on our logits == on the reference's, same seed
            torch.manual_seed(1234 + t0)
            mine = out.distributions.sample().cpu()
            torch.manual_seed(1234 + t0)
            ref = torch.distributions.Categorical(logits=rec["logits"].to(dev)).sample().cpu()
            assert torch.equal(mine, ref)
            assert torch.equal(out.distributions.mode().cpu(), rec["logits"].argmax(-1))


def test_sampler_select_keeps_cache_rows(dev):
    from oracle.make_golden_step import build_inputs as step_inputs
    from safevla_b200.model import B200SafeActorCritic
    gold = torch.load(os.path.join(GOLDEN_DIR, "step_N3_A6_C1.pt"), weights_only=False)
    case = gold["case"]
    sd = init_state_dict(case["A"], case["C"], case["wseed"], actor_gain=1.0)
    ro, prev = step_inputs(case)

    def run(keep_after, steps):
        model = B200SafeActorCritic(case["A"], case["C"], precision="fp32", state_dict=sd, device=dev, max_steps=16,
                                    extras="off")
        sel = list(range(case["N"]))
        outs = []
        for t in range(steps):
            if t == keep_after:
                sel = [0, 2]
                model.sampler_select(sel)
            obs = {k: v[t:t + 1, sel].to(dev) for k, v in ro["observations"].items()}
            with torch.no_grad():
                out, _ = model(obs, None, prev[t:t + 1, sel].to(dev), ro["masks"][t:t + 1, sel].to(dev))
            outs.append(out.distributions.raw_logits.cpu())
        return outs

    a = run(3, 6)            # drop sampler 1 after three steps
    b = run(10 ** 9, 6)      # never drop
    for t in range(3, 6):
        assert torch.allclose(a[t], b[t][:, [0, 2]], rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------ vision preprocessor
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_dino_preprocessor_matches_hf_golden(dev, precision):
    """uint8 frames -> [N, 384, 7, 12] DINOv2 ViT-S/14 features (dino_preprocessors.py:20-38,119-125,224-239) against
    the HuggingFace Dinov2Model golden (tests/golden/dinov2_vits14.pt); hub and HF weight layouts both load."""
    from oracle import vit_oracle as VO
    from oracle.make_golden_vit import frames
    from safevla_b200.vision import B200DinoViTPreprocessor
    for rec in torch.load(os.path.join(GOLDEN_DIR, "dinov2_vits14.pt"), weights_only=False):
        c = rec["case"]
        sd = VO.init_hub_state_dict(c["wseed"])
        if c["name"].endswith("224x224"):
            sd = VO.hub_to_hf(sd)  # exercise the HuggingFace layout too
        pre = B200DinoViTPreprocessor("raw_navigation_camera", sd, precision=precision, device=dev, crop=c["crop"])
        out = pre.process({"raw_navigation_camera": frames(c).to(dev)})
        assert out.shape == rec["out"].shape and out.dtype == torch.float32
        tol = 1e-4 if precision == "fp32" else 5e-2  # bf16 operands through 12 blocks, fp32 accumulation
        assert relerr(out, rec["out"]) < tol, relerr(out, rec["out"])
