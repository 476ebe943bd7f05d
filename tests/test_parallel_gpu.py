"""Data-parallel update on the GPU: two ranks (both on cuda:0, gloo carrying the CUDA arena -- NCCL refuses two ranks
on one device; the collective call site is the same) each own half of the samplers and must end up with the parameters
and the multiplier of the single-process update of the whole rollout."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

T, N, A, C = 8, 4, 6, 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs():
    from safevla_b200.params import init_state_dict
    from safevla_b200.synthetic import RolloutSpec, make_rollout
    sd = init_state_dict(A, C, seed=41, actor_gain=1.0)
    ro = make_rollout(RolloutSpec(T, N, A, C, episode_end_prob=0.25, seed=9))
    g = torch.Generator().manual_seed(4)
    vp, cvp = torch.randn(T + 1, N, 1, generator=g), torch.randn(T + 1, N, 1, generator=g).abs()
    logp = -1.7 + 0.1 * torch.randn(T, N, generator=g)
    return sd, ro, vp, cvp, logp


def _slice(ro, lo, hi):
    out = {}
    for k, v in ro.items():
        if isinstance(v, dict):
            out[k] = {kk: vv[:, lo:hi].contiguous() for kk, vv in v.items()}
        elif torch.is_tensor(v) and v.dim() >= 2:
            out[k] = v[:, lo:hi].contiguous()
        else:
            out[k] = v
    # Jc bookkeeping of the shard: recompute the finished-episode cost sum / count for these samplers
    masks, costs = out["masks"], out["costs"]
    ep = torch.zeros(hi - lo)
    s, c = 0.0, 0
    for t in range(costs.shape[0]):
        ep += costs[t, :, 0]
        fin = masks[t + 1, :, 0] == 0
        s += float(ep[fin].sum())
        c += int(fin.sum())
        ep[fin] = 0.0
    out["episode_cost_sum"], out["episode_count"] = torch.tensor(s), torch.tensor(float(c))
    return out


def _run(rank, world, port, out, normalize=False):
    from safevla_b200.model import B200SafeActorCritic
    from safevla_b200.parallel import shard_samplers
    from safevla_b200.storage import B200RolloutStorage
    from safevla_b200.updater import PPOLagConfig, PPOLagUpdater
    torch.cuda.set_device(0)
    dev = torch.device("cuda:0")
    if world > 1:
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
    sd, ro, vp, cvp, logp = _inputs()
    lo, hi = shard_samplers(N, world, rank)
    model = B200SafeActorCritic(A, C, precision="fp32", state_dict=sd, device=dev, extras="off")
    st = B200RolloutStorage(T, dev)
    st.load_rollout(_slice(ro, lo, hi), vp[:, lo:hi].contiguous(), cvp[:, lo:hi].contiguous(), logp[:, lo:hi].contiguous())
    upd = PPOLagUpdater(model, PPOLagConfig(update_repeats=2, lr=1e-3, eps=1e-4, normalize_advantage=normalize))
    res = upd.update(st)
    torch.cuda.synchronize()
    out[rank] = (model.param_arena.cpu(), res["lambda"].cpu(), res["grad_sq_norm"].cpu())
    if world > 1:
        dist.destroy_process_group()


@pytest.mark.parametrize("normalize", [False, True])
def test_two_rank_update_equals_single_process_update(normalize):
    """normalize=True: advantage normalisation uses the all-reduced {sum, sum^2, n}, i.e. global-batch statistics."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    mgr = mp.Manager()
    one, two = mgr.dict(), mgr.dict()
    mp.spawn(_run, args=(1, 0, one, normalize), nprocs=1, join=True)
    mp.spawn(_run, args=(2, _free_port(), two, normalize), nprocs=2, join=True)
    p1, lam1, sq1 = one[0]
    (pa, lama, sqa), (pb, lamb, sqb) = two[0], two[1]
    assert torch.equal(pa, pb) and torch.equal(lama, lamb), "ranks diverged"  # identical without a broadcast
    sd, *_ = _inputs()
    assert (p1 - pa).abs().max().item() < 2e-5, (p1 - pa).abs().max().item()
    assert abs(lam1.item() - lama.item()) < 1e-6
    assert abs(sq1.item() - sqa.item()) < 1e-4 * max(sq1.item(), 1e-12) + 1e-12
