"""Synthetic rollouts with the observation contract of the SafeVLA update path.

Shapes/dtypes follow SURVEY.md section 8b/8d (sensor definitions in the reference:
environment/navigation_sensors.py:144-183, :985-1042, environment/manipulation_sensors.py:10-26,
architecture/allenact_preprocessors/dino_preprocessors.py:30-35).  Everything is produced on
the host from a seeded ``torch.Generator`` so the same rollout can be fed to the CUDA path,
to the CPU oracle and (in the build container) to the reference code itself.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import numpy as np
import torch

GOAL_BYTES = 1000  # TaskNaturalLanguageSpecSensor pads to 1000 bytes (navigation_sensors.py:144-183)


@dataclass
class RolloutSpec:
    num_steps: int  # T
    num_samplers: int  # N (per rank)
    num_actions: int = 20  # A
    num_cameras: int = 1  # C
    prompt_tokens: int = 32  # L - 1 ids + EOS
    episode_end_prob: float = 1.0 / 64.0
    seed: int = 1234
    num_cost_channels: int = 1  # K > 1: costs [T, N, K] (extension beyond the reference's single scalar cost)


def encode_goal_ids(ids) -> np.ndarray:
    """Synthetic goal codec: blank-separated decimal token ids, NUL padded to 1000 bytes."""
    s = " ".join(str(int(i)) for i in ids).encode()
    assert len(s) <= GOAL_BYTES
    out = np.zeros(GOAL_BYTES, dtype=np.uint8)
    out[: len(s)] = np.frombuffer(s, dtype=np.uint8)
    return out


def make_rollout(spec: RolloutSpec, rank: int = 0, pin: bool = False) -> Dict[str, torch.Tensor]:
    """Returns host tensors for T+1 observation steps and T transition steps.

    keys: observations/{rgb_dinov2[,manipulation_rgb_dinov2],natural_language_spec,time_step,
    traj_index[,an_object_is_in_hand]}  (leading dims [T+1, N]), masks [T+1,N,1], rewards, costs
    [T,N,1], actions [T,N] int64, episode_costs/episode_count scalars (sum of undiscounted cost
    of episodes completed inside the rollout, and their number -> Jc).
    """
    T, N, A, C = spec.num_steps, spec.num_samplers, spec.num_actions, spec.num_cameras
    g = torch.Generator().manual_seed(spec.seed + rank)
    obs: Dict[str, torch.Tensor] = {}
    obs["rgb_dinov2"] = torch.randn(T + 1, N, 384, 7, 12, generator=g)
    if C == 2:
        obs["manipulation_rgb_dinov2"] = torch.randn(T + 1, N, 384, 7, 12, generator=g)
        obs["an_object_is_in_hand"] = (torch.rand(T + 1, N, 1, generator=g) < 0.1).to(torch.int64)

    done = torch.rand(T + 1, N, generator=g) < spec.episode_end_prob  # episode boundary before step t
    done[0] = False
    masks = (~done).to(torch.float32).unsqueeze(-1)
    time_step = torch.zeros(T + 1, N, dtype=torch.int64)
    traj_index = torch.zeros(T + 1, N, dtype=torch.int64)
    start_traj = torch.randint(0, 2048, (N,), generator=g)
    goal = torch.zeros(T + 1, N, GOAL_BYTES, dtype=torch.uint8)
    nl = spec.prompt_tokens - 1
    for n in range(N):
        ts, tr = int(torch.randint(0, 50, (1,), generator=g)), int(start_traj[n])
        cur = torch.from_numpy(encode_goal_ids(torch.randint(3, 32100, (nl,), generator=g).tolist()))
        for t in range(T + 1):
            if done[t, n]:
                ts, tr = 0, (tr + 1) % 2048
                cur = torch.from_numpy(encode_goal_ids(torch.randint(3, 32100, (nl,), generator=g).tolist()))
            time_step[t, n], traj_index[t, n] = ts, tr
            goal[t, n] = cur
            ts += 1
    obs["time_step"], obs["traj_index"], obs["natural_language_spec"] = time_step, traj_index, goal

    terminal = done[1:]  # step t is terminal iff a new episode starts at t+1
    rewards = (10.0 * (torch.rand(T, N, generator=g) < 0.3).float() * terminal.float()).unsqueeze(-1)
    K = spec.num_cost_channels
    if K == 1:
        costs = (torch.rand(T, N, 5, generator=g) < 0.05).float().sum(-1, keepdim=True)
    else:
        costs = (torch.rand(T, N, K, 5, generator=g) < 0.05).float().sum(-1)
    actions = torch.randint(0, A, (T, N), generator=g)

    # Jc bookkeeping: undiscounted cost of episodes that *finish* inside the rollout (per cost channel)
    ep_cost = torch.zeros(N, K)
    cost_sum, ep_cnt = torch.zeros(K), 0
    for t in range(T):
        ep_cost += costs[t]
        fin = terminal[t]
        cost_sum += ep_cost[fin].sum(0)
        ep_cnt += int(fin.sum())
        ep_cost[fin] = 0.0
    cost_sum = float(cost_sum[0]) if K == 1 else cost_sum
    out = {
        "observations": obs,
        "masks": masks,
        "rewards": rewards,
        "costs": costs,
        "actions": actions,
        "episode_cost_sum": torch.as_tensor(cost_sum, dtype=torch.float32),
        "episode_count": torch.tensor(float(ep_cnt), dtype=torch.float32),
    }
    if pin and torch.cuda.is_available():
        out = _pin(out)
    return out


def _pin(x):
    if isinstance(x, dict):
        return {k: _pin(v) for k, v in x.items()}
    return x.pin_memory() if isinstance(x, torch.Tensor) and x.dim() > 0 else x


def prev_actions_from(actions: torch.Tensor) -> torch.Tensor:
    """prev_actions[t] = actions[t-1]; row 0 is 0 (masked to the null token when masks[0]==0,
    and otherwise refers to the action taken before the rollout window, synthetic 0)."""
    return torch.cat([torch.zeros_like(actions[:1]), actions[:-1]], dim=0)
