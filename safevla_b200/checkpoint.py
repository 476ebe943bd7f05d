"""Checkpoint compatibility: the three on-disk formats the reference reads and the one
it writes, mapped onto B200SafeActorCritic's `state_dict` keys (which ARE the reference's keys, so a converted
checkpoint loads with `model.load_state_dict`).

  * PyTorch-Lightning imitation-learning checkpoint  {"state_dict": {"model.<k>": ...}}  with the IL head names
    `actor.weight / actor.bias`                          (training/offline/train_utils.py:6-68)
  * allenact RL checkpoint                             {"model_state_dict": {...}}
    (allenact_dino_transformer.py:177-191; the published safe_{objnav,pickup,fetch}.pt,
    scripts/download_aligned_ckpt.py:50-55)
  * a bare state dict whose keys start with visual_encoder. / actor. / decoder.
    (format auto-detection: architecture/models/allenact_transformer_models/inference_agent.py:127-160)

Pure dictionary logic: no device work, usable without a GPU.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Mapping, NamedTuple, Optional

import torch

TOWER_PREFIXES = ("", "critic_tsfm.", "c_critic_tsfm.")
_DINO = "visual_encoder.image_encoder.model"  # frozen DINOv2 weights of IL checkpoints: never part of the towers


class LoadReport(NamedTuple):
    loaded: List[str]
    missing_in_checkpoint: List[str]
    unexpected_in_checkpoint: List[str]


def _rename_il_head(k: str) -> str:
    # IL policy head is a bare nn.Linear called `actor`; the RL model wraps it as LinearActorHead.linear
    if k.endswith("actor.weight"):
        return k[: -len("actor.weight")] + "actor.linear.weight"
    if k.endswith("actor.bias"):
        return k[: -len("actor.bias")] + "actor.linear.bias"
    return k


def detect_format(ckpt: Mapping) -> str:
    """'lightning' | 'allenact' | 'bare'  (inference_agent.py:127-160); raises ValueError otherwise."""
    if "state_dict" in ckpt:
        return "lightning"
    if "model_state_dict" in ckpt:
        return "allenact"
    if any(str(k).startswith(("visual_encoder.", "actor.", "decoder.")) for k in ckpt.keys()):
        return "bare"
    raise ValueError("Unknown checkpoint format. Expected one of: 'state_dict' key (PyTorch Lightning), "
                     f"'model_state_dict' key (AllenAct), or a direct state dict; found keys {list(ckpt.keys())[:10]}")


def to_model_state_dict(ckpt: Mapping) -> Dict[str, torch.Tensor]:
    """Any of the three formats -> {reference model key: tensor} (inference_agent.py:127-160)."""
    fmt = detect_format(ckpt)
    if fmt == "lightning":
        out = {}
        for k, v in ckpt["state_dict"].items():
            nk = k.replace("model.", "", 1) if k.startswith("model.") else k
            if nk == "actor.weight":
                nk = "actor.linear.weight"
            elif nk == "actor.bias":
                nk = "actor.linear.bias"
            out[nk] = v
        return out
    if fmt == "allenact":
        return dict(ckpt["model_state_dict"])
    return dict(ckpt)


def merge_il_checkpoint(model_state: Mapping[str, torch.Tensor], ckpt: Mapping, ckpt_prefix: str = "model.",
                        towers: Iterable[str] = TOWER_PREFIXES):
    """`load_pl_ckpt_allenact` (training/offline/train_utils.py:6-68) for the three-tower model: every tower's
    constructor receives `prev_checkpoint`, so each tower takes the IL weights found under `ckpt_prefix + <key>`
    (allenact_dino_transformer.py:169-176; separate_actor_critic.py:8-11,23-25).  Keys absent from the checkpoint
    keep the model's current value.  Returns (new_state_dict, LoadReport) with report keys relative to a tower."""
    src = {_rename_il_head(k): v for k, v in ckpt["state_dict"].items()}
    new = dict(model_state)
    loaded, missing = [], []
    single = sorted({k[len(p):] for p in towers for k in model_state if _belongs(k, p, towers)})
    for p in towers:
        for k in single:
            full = p + k
            if full not in model_state:
                continue
            if ckpt_prefix + k in src:
                v = src[ckpt_prefix + k]
                if tuple(v.shape) != tuple(model_state[full].shape):
                    raise ValueError(f"shape mismatch for {full}: checkpoint {tuple(v.shape)} vs model "
                                     f"{tuple(model_state[full].shape)}")
                new[full] = v
                if p == "":
                    loaded.append(k)
            elif p == "":
                missing.append(k)
    unexpected = [k[len(ckpt_prefix):] for k in src
                  if k.startswith(ckpt_prefix) and k[len(ckpt_prefix):] not in single and _DINO not in k]
    return new, LoadReport(loaded, missing, unexpected)


def _belongs(key: str, prefix: str, towers: Iterable[str]) -> bool:
    """True when `key` is a parameter of the tower with `prefix` ('' = the un-prefixed actor tower)."""
    if prefix:
        return key.startswith(prefix)
    return not any(p and key.startswith(p) for p in towers)


def strip_critic_towers(state: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """`prev_rl_checkpoint` path (allenact_dino_transformer.py:177-191): keeps the policy tower only."""
    return {k: v for k, v in state.items() if "critic_tsfm" not in k}


def load_checkpoint(model, ckpt_or_path, strict: bool = False):
    """inference_agent.py:122-165: torch.load (if a path) -> format detection -> load_state_dict(strict=False)."""
    ckpt = torch.load(ckpt_or_path, map_location="cpu", weights_only=False) if isinstance(ckpt_or_path, str) else ckpt_or_path
    return model.load_state_dict(to_model_state_dict(ckpt), strict=strict)


def load_il_checkpoint(model, ckpt_or_path, ckpt_prefix: str = "model.") -> LoadReport:
    ckpt = torch.load(ckpt_or_path, map_location="cpu", weights_only=False) if isinstance(ckpt_or_path, str) else ckpt_or_path
    new, report = merge_il_checkpoint(model.state_dict(), ckpt, ckpt_prefix)
    model.load_state_dict(new, strict=True)
    return report


def allenact_checkpoint(model_state: Mapping[str, torch.Tensor], total_steps: int = 0, optimizer_state: Optional[dict] = None,
                        **extra) -> dict:
    """The dict layout the fork's engine saves and `prev_rl_checkpoint` / the inference agent read back."""
    ck = {"model_state_dict": {k: v.detach().cpu().clone() for k, v in model_state.items()}, "total_steps": total_steps}
    if optimizer_state is not None:
        ck["optimizer_state_dict"] = optimizer_state
    ck.update(extra)
    return ck
