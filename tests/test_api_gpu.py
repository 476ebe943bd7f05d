"""The allenact-facing surface around the update: `batched_experience_generator` / `agent_input_for_next_step`
(fork API witnessed at architecture/models/allenact_transformer_models/inference_agent.py:246-269; batch keys consumed
at training/online/loss/customized_loss.py:327,344,352,375-384), the engine-style loop INTEGRATION.md section 2
advertises (storage batch -> model.forward -> loss plugin -> backward -> clip + Adam) against PPOLagUpdater.update, the
logging extras of allenact_dino_transformer.py:431-455, updater resume state.  B200 only."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import torch_oracle as TO  # noqa: E402  (checker only)
from safevla_b200.params import init_state_dict  # noqa: E402
from safevla_b200.synthetic import RolloutSpec, make_rollout  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _filled(dev, T, N, A, C, K=1, seed=77, normalize=False):
    from safevla_b200.storage import B200RolloutStorage
    ro = make_rollout(RolloutSpec(T, N, A, C, episode_end_prob=0.2, seed=seed, num_cost_channels=K))
    g = torch.Generator().manual_seed(seed + 1)
    vp, cvp = torch.randn(T + 1, N, 1, generator=g), torch.randn(T + 1, N, K, generator=g).abs()
    logp = -1.7 + 0.1 * torch.randn(T, N, generator=g)
    st = B200RolloutStorage(T, dev, num_cost_channels=K)
    st.load_rollout(ro, vp, cvp, logp)
    st.before_updates(next_value=vp[T], next_c_value=cvp[T], use_gae=True, gamma=0.99, tau=0.95,
                      normalize_advantage=normalize)
    return st, ro, vp, cvp, logp


@pytest.mark.parametrize("num_mini_batch", [1, 2, 4])
def test_batched_experience_generator_contract(dev, num_mini_batch):
    """Every key the loss plugins and the model read, with the fork's shapes; mini-batches partition the SAMPLER axis
    (whole trajectories stay together: the decoder attends over time) and their union is the rollout."""
    T, N, A, C = 8, 4, 6, 1
    st, ro, vp, cvp, logp = _filled(dev, T, N, A, C, normalize=True)
    ret, adv = TO.gae_returns(ro["rewards"], vp, ro["masks"], 0.99, 0.95)
    cret, cadv = TO.gae_returns(ro["costs"], cvp, ro["masks"], 0.99, 0.95)
    batches = list(st.batched_experience_generator(num_mini_batch))
    assert len(batches) == num_mini_batch
    per = N // num_mini_batch
    for b, batch in enumerate(batches):
        sl = slice(b * per, (b + 1) * per)
        assert set(batch) >= {"observations", "memory", "prev_actions", "masks", "actions", "old_action_log_probs",
                              "values", "c_values", "returns", "c_returns", "adv_targ", "c_adv_targ", "norm_adv_targ",
                              "c_norm_adv_targ"}
        assert batch["memory"] is None
        for k, v in batch["observations"].items():
            assert torch.equal(v.cpu(), ro["observations"][k][:T, sl]), k
        assert batch["actions"].shape == (T, per) and torch.equal(batch["actions"].cpu(), ro["actions"][:, sl])
        prev = torch.cat([torch.zeros(1, N, dtype=torch.int64), ro["actions"][:-1]], 0)
        assert torch.equal(batch["prev_actions"].cpu(), prev[:, sl])
        assert batch["masks"].shape == (T, per, 1) and torch.equal(batch["masks"].cpu(), ro["masks"][:T, sl])
        assert batch["old_action_log_probs"].shape == (T, per)
        assert torch.equal(batch["old_action_log_probs"].cpu(), logp[:, sl])
        for key, ref in (("values", vp[:T]), ("c_values", cvp[:T]), ("returns", ret[:T]), ("c_returns", cret[:T]),
                         ("adv_targ", adv), ("c_adv_targ", cadv)):
            assert batch[key].shape == (T, per, 1), key
            assert torch.equal(batch[key].cpu(), ref[:, sl]), key
        # normalisation uses the statistics of the WHOLE rollout, not of the mini-batch
        na = (adv - adv.mean()) / (adv.std() + 1e-5)
        assert torch.allclose(batch["norm_adv_targ"].cpu(), na[:, sl], rtol=1e-5, atol=1e-6)
        nca = (cadv - cadv.mean()) / (cadv.std() + 1e-5)
        assert torch.allclose(batch["c_norm_adv_targ"].cpu(), nca[:, sl], rtol=1e-5, atol=1e-6)
        for v in batch.values():
            if torch.is_tensor(v):
                assert v.is_contiguous() and v.device.type == "cuda"


def test_batched_experience_generator_cost_channels(dev):
    """K = 2 (extension): the c_* entries are channel-major [K, T, n, 1], the normalised cost advantages included
    (they used to be channel 0 only), and the loss plugins consume them as yielded."""
    from safevla_b200.losses import SafePPOLogGrad, SafePPOValue
    from safevla_b200.model import B200SafeActorCritic
    from safevla_b200.updater import PPOLagConfig, PPOLagUpdater
    T, N, A, C, K = 6, 4, 6, 1, 2
    st, ro, vp, cvp, logp = _filled(dev, T, N, A, C, K=K, normalize=True)
    for nmb in (1, 2):
        per = N // nmb
        for b, batch in enumerate(st.batched_experience_generator(nmb)):
            sl = slice(b * per, (b + 1) * per)
            for key in ("c_values", "c_returns", "c_adv_targ", "c_norm_adv_targ"):
                assert batch[key].shape == (K, T, per, 1), (key, batch[key].shape)
            for k in range(K):
                cret, cadv = TO.gae_returns(ro["costs"][..., k:k + 1], cvp[..., k:k + 1], ro["masks"], 0.99, 0.95)
                assert torch.equal(batch["c_returns"][k].cpu(), cret[:T, sl])
                assert torch.equal(batch["c_adv_targ"][k].cpu(), cadv[:, sl])
                nca = (cadv - cadv.mean()) / (cadv.std() + 1e-5)
                assert torch.allclose(batch["c_norm_adv_targ"][k].cpu(), nca[:, sl], rtol=1e-5, atol=1e-6)
    # plugin path == updater path at K = 2 (stage 0: PPOValue + SafePPOValue; stage 1: SafePPOLogGrad, normalised)
    sd = init_state_dict(A, C, seed=4, actor_gain=1.0, num_cost_channels=K)
    for stage in (0, 1):
        cfg = PPOLagConfig(update_repeats=1, lr=1e-3, eps=1e-4, stage=stage, cost_limit=(0.05, 50.0), lambda_init=0.4,
                           normalize_advantage=True, max_grad_norm=0.0)
        m_upd = B200SafeActorCritic(A, C, precision="fp32", state_dict=sd, device=dev, num_cost_channels=K, extras="off")
        st_u, *_ = _filled(dev, T, N, A, C, K=K)
        upd = PPOLagUpdater(m_upd, cfg)
        res = upd.update(st_u)
        m_plug = B200SafeActorCritic(A, C, precision="fp32", state_dict=sd, device=dev, num_cost_channels=K, extras="off")
        m_plug.set_trainable_towers((1, 2) if stage == 0 else (0, 1))
        (batch,) = list(st.batched_experience_generator(1))
        out, _ = m_plug(batch["observations"], None, batch["prev_actions"], batch["masks"])
        if stage == 0:
            from safevla_b200.losses import PPOValue
            l1, i1 = PPOValue(clip_param=0.1, use_clipped_value_loss=False).loss(0, batch, out)
            l2, i2 = SafePPOValue(clip_param=0.1, use_clipped_value_loss=False).loss(0, batch, out)
            total = l1 + l2
            assert abs(i1["value"] - res["loss_scalars"][1].item()) < 1e-5 * max(1, abs(i1["value"]))
            assert abs(i2["value"] - res["loss_scalars"][4].item()) < 1e-5 * max(1, abs(i2["value"]))
        else:
            loss = SafePPOLogGrad(clip_param=0.1, value_loss_coef=0.5, entropy_coef=0.0, use_clipped_value_loss=False,
                                  action_loss_schedule=None, discrete_critics=False, normalize_advantage=True)
            total, info = loss.loss(0, batch, out, lagrangian_multiplier=torch.full((K,), 0.4))
            assert abs(info["ppo_total"] - res["loss_scalars"][0].item()) < 1e-5 * max(1, abs(info["ppo_total"]))
        total.backward()
        opt = torch.optim.Adam([p for p in m_plug.parameters() if p.requires_grad], lr=1e-3, eps=1e-4)
        opt.step()
        a, b_ = m_plug.param_arena, m_upd.param_arena
        assert (a - b_).abs().max().item() < 2e-6, (stage, (a - b_).abs().max().item())


def test_agent_input_for_next_step_and_add_cycle(dev):
    """initialize / add / agent_input_for_next_step / after_updates as inference_agent.py:172-175,246-269 drives them."""
    from safevla_b200.storage import B200RolloutStorage
    T, N, A, C = 5, 3, 6, 1
    ro = make_rollout(RolloutSpec(T, N, A, C, episode_end_prob=0.3, seed=9))
    st = B200RolloutStorage(T, dev)
    st.initialize(observations={k: v[0] for k, v in ro["observations"].items()}, num_samplers=N)
    for t in range(T):
        inp = st.agent_input_for_next_step()
        assert inp["memory"] is None and inp["prev_actions"].shape == (1, N) and inp["masks"].shape == (1, N, 1)
        for k, v in inp["observations"].items():
            assert v.shape[:2] == (1, N) and torch.equal(v[0].cpu(), ro["observations"][k][t]), k
        if t > 0:
            assert torch.equal(inp["prev_actions"][0].cpu(), ro["actions"][t - 1])
            assert torch.equal(inp["masks"][0].cpu(), ro["masks"][t])
        st.add(observations={k: v[t + 1] for k, v in ro["observations"].items()}, memory=None,
               actions=ro["actions"][t].view(1, N), action_log_probs=torch.zeros(1, N, 1), value_preds=torch.zeros(1, N, 1),
               rewards=ro["rewards"][t], costs=ro["costs"][t], c_value_preds=torch.zeros(1, N, 1), masks=ro["masks"][t + 1])
    with pytest.raises(AssertionError):
        st.add(observations={}, memory=None, actions=ro["actions"][0], action_log_probs=torch.zeros(N), value_preds=torch.zeros(N),
               rewards=ro["rewards"][0], masks=ro["masks"][1])
    assert abs(st.cost_sum_cnt[0].item() - float(ro["episode_cost_sum"])) < 1e-5
    assert st.cost_sum_cnt[1].item() == float(ro["episode_count"])
    st.after_updates()
    inp = st.agent_input_for_next_step()
    assert torch.equal(inp["prev_actions"][0].cpu(), ro["actions"][T - 1])
    assert torch.equal(inp["masks"][0].cpu(), ro["masks"][T])
    for k, v in inp["observations"].items():
        assert torch.equal(v[0].cpu(), ro["observations"][k][T])
    assert st.cost_sum_cnt.abs().sum().item() == 0.0 and st.step == 0


@pytest.mark.parametrize("num_mini_batch", [1, 2])
def test_engine_style_loop_equals_updater(dev, num_mini_batch):
    """The route INTEGRATION.md section 2 advertises for an existing allenact engine: batches from
    `batched_experience_generator` -> `model.forward` -> `SafePPOLogGrad.loss` -> `backward` -> `clip_grad_norm_` +
    `torch.optim.Adam` on the model's own parameters.  num_mini_batch = 1: the same arithmetic as
    `PPOLagUpdater.update` (one schedule of C-ABI launches) -- parameters and loss must agree.  num_mini_batch = 2: the
    engine takes one optimizer step per sampler half; the first step is checked against the CPU oracle's update of
    that half, the whole loop for run-to-run determinism."""
    from oracle.update_oracle import oracle_update
    from safevla_b200.losses import SafePPOLogGrad
    from safevla_b200.model import B200SafeActorCritic
    from safevla_b200.updater import PPOLagConfig, PPOLagUpdater
    T, N, A, C = 8, 4, 6, 1
    sd = init_state_dict(A, C, seed=21, actor_gain=1.0)
    st, ro, vp, cvp, logp = _filled(dev, T, N, A, C)
    cfg = PPOLagConfig(update_repeats=2, lr=1e-3, eps=1e-4)
    loss = SafePPOLogGrad(clip_param=cfg.clip_param, value_loss_coef=cfg.value_loss_coef, entropy_coef=cfg.entropy_coef,
                          use_clipped_value_loss=False, action_loss_schedule=None, discrete_critics=False,
                          normalize_advantage=False)
    lam = torch.tensor(cfg.lambda_init)

    def engine_run():
        model = B200SafeActorCritic(A, C, precision="fp32", state_dict=sd, device=dev, extras="off")
        model.set_trainable_towers((0, 1))
        params = [p for _, p in model.named_parameters() if p.requires_grad]
        opt = torch.optim.Adam(params, lr=cfg.lr, betas=cfg.betas, eps=cfg.eps)
        totals, first = [], None
        for rep in range(cfg.update_repeats):
            for batch in st.batched_experience_generator(num_mini_batch):
                opt.zero_grad(set_to_none=False)
                out, _ = model(batch["observations"], batch["memory"], batch["prev_actions"], batch["masks"])
                total, info = loss.loss(rep, batch, out, lagrangian_multiplier=lam)
                total.backward()
                torch.nn.utils.clip_grad_norm_(params, cfg.max_grad_norm)
                opt.step()
                totals.append(info["ppo_total"])
                if first is None:
                    first = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        return model, totals, first

    model, totals, first = engine_run()
    assert len(totals) == cfg.update_repeats * num_mini_batch
    init = B200SafeActorCritic(A, C, precision="fp32", state_dict=sd, device=dev, extras="off").param_arena
    assert (model.param_arena - init).abs().max().item() > 5e-4  # the optimizer really moved the arena-backed parameters
    if num_mini_batch == 1:
        m2 = B200SafeActorCritic(A, C, precision="fp32", state_dict=sd, device=dev, extras="off")
        st2, *_ = _filled(dev, T, N, A, C)
        res = PPOLagUpdater(m2, cfg).update(st2)
        assert abs(res["loss_scalars"][0].item() - totals[-1]) < 1e-5 * max(1.0, abs(totals[-1]))
        err = (m2.param_arena - model.param_arena).abs().max().item()
        assert err < 5e-6, err
    else:
        per = N // num_mini_batch
        half = {k: ({kk: vv[:, :per] for kk, vv in v.items()} if isinstance(v, dict) else
                    (v[:, :per] if torch.is_tensor(v) and v.dim() >= 2 else v)) for k, v in ro.items()}
        one = PPOLagConfig(update_repeats=1, lr=1e-3, eps=1e-4)
        ref_sd, _, ref_info = oracle_update(sd, half, vp[:, :per], cvp[:, :per], logp[:, :per], one, A, C)
        assert abs(totals[0] - ref_info["last_total"]) < 1e-4 * max(1.0, abs(ref_info["last_total"]))
        worst = max((first[k] - v).abs().max().item() for k, v in ref_sd.items() if "text_encoder" not in k)
        assert worst < 2e-5, worst
        model_b, totals_b, _ = engine_run()
        assert torch.equal(model_b.param_arena, model.param_arena) and totals_b == totals


def test_extras_values_match_reference_quantities(dev):
    """a14: the logging extras describe the COST tower (separate_actor_critic.py:35): total gradient norm of its
    parameters and weight / bias / weight-gradient norms of its critic head (allenact_dino_transformer.py:431-455), as
    1-element CPU tensors; checked against the same quantities computed from the parameters with torch."""
    from safevla_b200.losses import PPOValue, SafePPOValue
    from safevla_b200.model import B200SafeActorCritic
    T, N, A, C = 6, 2, 6, 1
    sd = init_state_dict(A, C, seed=3, actor_gain=1.0)
    st, ro, vp, cvp, logp = _filled(dev, T, N, A, C)
    model = B200SafeActorCritic(A, C, precision="fp32", state_dict=sd, device=dev)
    model.set_trainable_towers((1, 2))
    (batch,) = list(st.batched_experience_generator(1))
    out, _ = model(batch["observations"], None, batch["prev_actions"], batch["masks"])
    ex = out.extras
    assert set(ex) >= {"total_norm", "weight_norm", "bias_norm", "weight_grad_norm", "stop_grad_values"}
    for k in ("total_norm", "weight_norm", "bias_norm", "weight_grad_norm"):
        assert ex[k].shape == (1,) and ex[k].device.type == "cpu"
    w, b = sd["c_critic_tsfm.critic.fc.weight"], sd["c_critic_tsfm.critic.fc.bias"]
    assert abs(ex["weight_norm"].item() - w.norm().item()) < 1e-5 * max(1.0, w.norm().item())
    assert abs(ex["bias_norm"].item() - b.norm().item()) < 1e-6
    assert ex["total_norm"].item() == 0.0 and ex["weight_grad_norm"].item() == 0.0  # nothing back-propagated yet
    assert torch.equal(ex["stop_grad_values"], out.c_values.detach())
    l1, _ = PPOValue(clip_param=0.1, use_clipped_value_loss=False).loss(0, batch, out)
    l2, _ = SafePPOValue(clip_param=0.1, use_clipped_value_loss=False).loss(0, batch, out)
    (l1 + l2).backward()
    out2, _ = model(batch["observations"], None, batch["prev_actions"], batch["masks"])  # extras see the gradients now
    cost = [p for n, p in model.named_parameters() if n.startswith("c_critic_tsfm.") and "text_encoder" not in n]
    tn = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in cost)).item()
    assert tn > 0 and abs(out2.extras["total_norm"].item() - tn) < 1e-4 * tn
    wg = model.get_parameter("c_critic_tsfm.critic.fc.weight").grad.norm().item()
    assert abs(out2.extras["weight_grad_norm"].item() - wg) < 1e-5 * max(wg, 1e-6)


def test_updater_state_dict_resume(dev):
    """Optimizer / multiplier state round trip (SURVEY section 5 checkpoint row; dinov2_vits_tsfm_base.py:79,329):
    update -> save -> fresh objects -> load -> update equals two uninterrupted updates bit for bit."""
    from safevla_b200.model import B200SafeActorCritic
    from safevla_b200.updater import PPOLagConfig, PPOLagUpdater
    T, N, A, C = 6, 2, 6, 1
    sd = init_state_dict(A, C, seed=8, actor_gain=1.0)
    cfg = PPOLagConfig(update_repeats=2, lr=1e-3)

    def fresh(state):
        m = B200SafeActorCritic(A, C, precision="bf16", state_dict=state, device=dev, extras="off")
        return m, PPOLagUpdater(m, cfg)

    m_a, u_a = fresh(sd)
    st, *_ = _filled(dev, T, N, A, C, seed=5)
    u_a.update(st)
    saved = copy.deepcopy({"model": {k: v.cpu() for k, v in m_a.state_dict().items()}, "updater": u_a.state_dict()})
    st2, *_ = _filled(dev, T, N, A, C, seed=6)
    u_a.update(st2)
    m_b, u_b = fresh(saved["model"])
    u_b.load_state_dict(saved["updater"])
    st3, *_ = _filled(dev, T, N, A, C, seed=6)
    u_b.update(st3)
    assert torch.equal(m_a.param_arena, m_b.param_arena)
    assert torch.equal(u_a.exp_avg, u_b.exp_avg) and torch.equal(u_a.exp_avg_sq, u_b.exp_avg_sq)
    assert torch.equal(u_a.lagrange.lagrangian_multiplier, u_b.lagrange.lagrangian_multiplier)
    assert u_a.tower_steps == u_b.tower_steps == [4, 4, 0]
    # the optimizer state is addressable by parameter name (torch.optim.Adam-style exp_avg / exp_avg_sq / step)
    named = u_b.named_optimizer_state()
    k = "critic_tsfm.decoder.norm.weight"
    assert named[k]["exp_avg"].shape == m_b.get_parameter(k).shape and named[k]["step"] == 4


def test_cuda_graph_replay_is_bit_identical_to_eager_launches(dev):
    """PPOLagConfig(cuda_graphs=True): the forward + loss + backward of an update repeat is captured once and replayed
    (what data-parallel runs use, where enqueueing ~700 launches per repeat from Python is slower than the GPU).  Two
    whole updates on two different rollouts -- the second one only refreshes the graph's static rollout context --
    must leave exactly the parameters, optimizer state, multiplier and loss scalars of the eagerly launched updates,
    and the launch counter must count replayed launches."""
    from safevla_b200 import _lib as L
    from safevla_b200.model import B200SafeActorCritic
    from safevla_b200.storage import B200RolloutStorage
    from safevla_b200.updater import PPOLagConfig, PPOLagUpdater
    T, N, A, C = 8, 4, 6, 1
    sd = init_state_dict(A, C, seed=21, actor_gain=1.0)
    lib = L.load_library()
    results = {}
    for mode in (False, True):
        model = B200SafeActorCritic(A, C, precision="bf16", state_dict=sd, device=dev, extras="off")
        upd = PPOLagUpdater(model, PPOLagConfig(update_repeats=3, lr=1e-3, cuda_graphs=mode))
        st = B200RolloutStorage(T, dev)
        scal, counts = [], []
        for seed in (77, 78, 79):
            ro = make_rollout(RolloutSpec(T, N, A, C, episode_end_prob=0.2, seed=seed))
            g = torch.Generator().manual_seed(seed)
            vp, cvp = torch.randn(T + 1, N, 1, generator=g), torch.randn(T + 1, N, 1, generator=g).abs()
            st.load_rollout(ro, vp, cvp, -1.7 + 0.1 * torch.randn(T, N, generator=g))
            model._ctx_cache = None
            n0 = lib.svla_launch_count()
            res = upd.update(st)
            torch.cuda.synchronize()
            counts.append(lib.svla_launch_count() - n0)
            scal.append(res["loss_scalars"].clone())
            st.after_updates()
        results[mode] = (model.param_arena.clone(), upd.exp_avg.clone(), upd.lagrange.lagrangian_multiplier.clone(), scal,
                         counts, len(upd._graphs))
    eager, graph = results[False], results[True]
    assert torch.equal(eager[0], graph[0]) and torch.equal(eager[1], graph[1]) and torch.equal(eager[2], graph[2])
    for a, b in zip(eager[3], graph[3]):
        assert torch.equal(a, b)
    assert eager[5] == 0 and graph[5] == 1          # one graph, reused by all three updates
    assert graph[4][1] == eager[4][1] and graph[4][2] == eager[4][2]  # replayed launches are counted
