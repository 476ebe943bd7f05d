"""Rollout ingestion: environment step results -> pinned host staging -> HBM rollout arena.

The reference's environment workers hand the engine one `SafeRLStepResult(observation, reward, cost, done, info)` per
sampler and step (tasks/abstract_task.py:369-380; the type lives in the un-vendored allenact fork); the engine batches
them (`batch_observations`, witnessed at architecture/models/allenact_transformer_models/inference_agent.py:229-231),
runs the sensor-preprocessor graph (the DINOv2 ViT on the raw frames) and calls
`rollout_storage.add(observations=, memory=, actions=, action_log_probs=, value_preds=, rewards=, costs=,
c_value_preds=, masks=)` (:255-267).  The goal string travels as NUL-padded bytes
(`convert_string_to_byte`, utils/string_utils.py:11-12; sensor environment/navigation_sensors.py:174-183).

`RolloutIngestor` does that hand-over for `B200RolloutStorage` with one staging slab per stream:

  * the N per-sampler numpy observations of a step are packed into PINNED host slabs (two sets, alternating, so the
    host may pack step t+1 while step t is still being copied);
  * one asynchronous H2D copy per stream on a dedicated copy stream; the compute stream waits on its event;
  * raw uint8 camera frames are uploaded as bytes (258 KB per 224 x 384 frame instead of 129 KB of fp32 features that
    the host would first have to compute) and encoded ON THE DEVICE by `B200DinoViTPreprocessor`;
  * `done` becomes `masks = 1 - done`; reward / cost / mask go to the arena through `storage.add`, which also keeps the
    per-sampler episode-cost totals and the (sum, count) of finished episodes that the Lagrange update reads
    (`svla_episode_cost_step`).

No arithmetic of the update path happens here; torch is used for memory and streams only.
"""
from __future__ import annotations

from typing import Any, Dict, List, Mapping, NamedTuple, Optional, Sequence

import numpy as np
import torch

GOAL_BYTES = 1000  # TaskNaturalLanguageSpecSensor(str_max_len=1000), environment/navigation_sensors.py:148-152


class SafeRLStepResult(NamedTuple):
    """Field-for-field mirror of the allenact-fork type built at tasks/abstract_task.py:369-380."""
    observation: Optional[Mapping[str, Any]]
    reward: Optional[float]
    cost: Optional[float]
    done: Optional[bool]
    info: Optional[Dict[str, Any]]


def convert_string_to_byte(str_to_encode: str, max_len: int = GOAL_BYTES) -> np.ndarray:
    """utils/string_utils.py:11-12: the string as `max_len` NUL-padded bytes (truncated beyond), uint8 [max_len]."""
    raw = str_to_encode.encode()[:max_len]
    out = np.zeros(max_len, dtype=np.uint8)
    out[: len(raw)] = np.frombuffer(raw, dtype=np.uint8)
    return out


def convert_byte_to_string(bytes_to_decode: np.ndarray) -> str:
    """utils/string_utils.py:15-18."""
    return bytes(np.asarray(bytes_to_decode, dtype=np.uint8).reshape(-1)).rstrip(b"\x00").decode()


class _Slab:
    """One pinned host tensor [N, ...] per observation key plus the per-step scalars, and its device twin."""

    def __init__(self, shapes: Dict[str, tuple], dtypes: Dict[str, torch.dtype], n: int, dev: torch.device):
        self.host = {k: torch.empty((n, *shapes[k]), dtype=dtypes[k]).pin_memory() for k in shapes}
        self.dev = {k: torch.empty((n, *shapes[k]), dtype=dtypes[k], device=dev) for k in shapes}
        self.scal_host = torch.empty(3, n, dtype=torch.float32).pin_memory()  # reward, cost, mask
        self.scal_dev = torch.empty(3, n, dtype=torch.float32, device=dev)
        self.copied = torch.cuda.Event()
        self.consumed = torch.cuda.Event()
        self.consumed.record()

    def bytes(self) -> int:
        return sum(v.numel() * v.element_size() for v in self.host.values()) + self.scal_host.numel() * 4


class RolloutIngestor:
    """Feeds `B200RolloutStorage` from per-sampler step results.

    frame_encoders: {raw-frame observation key: (B200DinoViTPreprocessor, feature key)} -- e.g.
    {"rgb_raw": (nav_vit, "rgb_dinov2"), "manipulation_rgb_raw": (manip_vit, "manipulation_rgb_dinov2")}: those keys
    are uploaded as uint8 frames and replaced by the encoder's output before they reach the storage.
    """

    def __init__(self, storage, num_samplers: int, *, frame_encoders: Optional[Dict[str, tuple]] = None,
                 goal_key: str = "natural_language_spec", str_max_len: int = GOAL_BYTES):
        if not torch.cuda.is_available():
            raise RuntimeError("RolloutIngestor stages into HBM; no CUDA device is visible")
        self.storage, self.N = storage, num_samplers
        self.dev = storage.dev
        self.frame_encoders = dict(frame_encoders or {})
        self.goal_key, self.str_max_len = goal_key, str_max_len
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self._slabs: List[_Slab] = []
        self._turn = 0
        self._goal_cache: Dict[str, np.ndarray] = {}
        self.h2d_bytes_per_step = 0

    # ------------------------------------------------------------------ host-side packing
    def _goal_bytes(self, goal) -> np.ndarray:
        if isinstance(goal, str):  # tasks that hand the instruction over as text
            b = self._goal_cache.get(goal)
            if b is None:
                b = self._goal_cache[goal] = convert_string_to_byte(goal, self.str_max_len)
            return b
        return np.asarray(goal, dtype=np.uint8).reshape(-1)

    def _make_slabs(self, observations: Sequence[Mapping[str, Any]]):
        shapes, dtypes = {}, {}
        for k, v in observations[0].items():
            a = self._goal_bytes(v) if k == self.goal_key else np.asarray(v)
            shapes[k] = tuple(a.shape)
            dtypes[k] = torch.from_numpy(np.zeros(1, dtype=a.dtype)).dtype
        self._slabs = [_Slab(shapes, dtypes, self.N, self.dev) for _ in range(2)]
        self.h2d_bytes_per_step = self._slabs[0].bytes()

    def _pack(self, slab: _Slab, observations: Sequence[Mapping[str, Any]], rewards, costs, dones):
        assert len(observations) == self.N, f"expected {self.N} samplers, got {len(observations)}"
        slab.consumed.synchronize()  # the device is done with this slab's previous contents
        for k, h in slab.host.items():
            dst = h.numpy()
            for n, ob in enumerate(observations):
                src = self._goal_bytes(ob[k]) if k == self.goal_key else np.asarray(ob[k])
                dst[n, ...] = src.reshape(dst.shape[1:])
        s = slab.scal_host.numpy()
        s[0] = 0.0 if rewards is None else np.asarray(rewards, dtype=np.float32)
        s[1] = 0.0 if costs is None else np.asarray(costs, dtype=np.float32)
        s[2] = 1.0 if dones is None else 1.0 - np.asarray(dones, dtype=np.float32)

    def _upload(self, slab: _Slab) -> Dict[str, torch.Tensor]:
        """Asynchronous H2D of the whole slab on the copy stream; returns the observation dict for the storage
        (frames already encoded), valid on the current (compute) stream."""
        with torch.cuda.stream(self.copy_stream):
            for k in slab.host:
                slab.dev[k].copy_(slab.host[k], non_blocking=True)
            slab.scal_dev.copy_(slab.scal_host, non_blocking=True)
            slab.copied.record()
        torch.cuda.current_stream().wait_event(slab.copied)
        obs: Dict[str, torch.Tensor] = {}
        for k, v in slab.dev.items():
            if k in self.frame_encoders:
                enc, out_key = self.frame_encoders[k]
                obs[out_key] = enc.encode(v)  # uint8 [N, H, W, 3] -> fp32 [N, 384, 7, 12] on the device
            else:
                obs[k] = v
        return obs

    # ------------------------------------------------------------------ public surface
    def reset(self, observations: Sequence[Mapping[str, Any]]):
        """First observations of a run: allocates the staging slabs and `storage.initialize(...)`s the arena."""
        self._make_slabs(observations)
        slab = self._slabs[0]
        self._pack(slab, observations, None, None, None)
        obs = self._upload(slab)
        self.storage.initialize(observations=obs, num_samplers=self.N)
        self.storage.masks[0].zero_()  # every sampler starts a new episode
        slab.consumed.record()
        self._turn = 1

    def push(self, step_results: Sequence[SafeRLStepResult], *, actions: torch.Tensor, action_log_probs: torch.Tensor,
             value_preds: torch.Tensor, c_value_preds: Optional[torch.Tensor] = None, memory=None):
        """One environment step of all N samplers.  `step_results[n].observation` is the observation AFTER the step
        (the first one of the next episode when `done`), as allenact's vector sampler returns it; actions /
        log-probs / value predictions are the agent's outputs that produced the step (device or host tensors)."""
        slab = self._slabs[self._turn & 1]
        self._turn += 1
        self._pack(slab, [r.observation for r in step_results], [r.reward for r in step_results],
                   [0.0 if r.cost is None else r.cost for r in step_results], [bool(r.done) for r in step_results])
        obs = self._upload(slab)
        N = self.N
        self.storage.add(observations=obs, memory=memory, actions=actions, action_log_probs=action_log_probs,
                         value_preds=value_preds, rewards=slab.scal_dev[0].view(N, 1), costs=slab.scal_dev[1].view(N, 1),
                         c_value_preds=c_value_preds, masks=slab.scal_dev[2].view(N, 1))
        slab.consumed.record()
