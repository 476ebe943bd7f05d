"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the rollout-side vision preprocessor
(architecture/allenact_preprocessors/dino_preprocessors.py:20-38,119-125,224-239: /255, mean/std normalisation,
[:, :, :, 3:-3] crop, DINOv2 ViT-S/14 `forward_features(...)["x_norm_patchtokens"]`, reshape to [B, 384, 16, 27],
AdaptiveAvgPool2d((7, 12))).

The ViT is a third-party dependency of the reference (torch.hub `facebookresearch/dinov2`, un-pinned, absent from
/root/reference and not fetchable offline).  Its published algorithm is restated below on the hub state-dict layout
(`blocks.N.attn.qkv.weight`, `ls1.gamma`, ...) and pinned against HuggingFace transformers' `Dinov2Model`, an
independent implementation of the same architecture (oracle/make_golden_vit.py -> tests/golden/dinov2_vits14.pt).
**Parity against the hub code itself is unpinned** (no reference test, golden vector or weight file exists for it).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

DINO_RGB_MEANS = (0.48145466, 0.4578275, 0.40821073)
DINO_RGB_STDS = (0.26862954, 0.26130258, 0.27577711)
DIM, HEADS, DEPTH, MLP, PATCH, GRID = 384, 6, 12, 1536, 14, 37


from safevla_b200.vision import init_hub_state_dict  # noqa: E402,F401  (seeded synthetic weights live with the product)


def hub_to_hf(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """hub layout -> transformers.Dinov2Model layout (splits the packed qkv)."""
    out = {"embeddings.cls_token": sd["cls_token"], "embeddings.mask_token": sd["mask_token"],
           "embeddings.position_embeddings": sd["pos_embed"],
           "embeddings.patch_embeddings.projection.weight": sd["patch_embed.proj.weight"],
           "embeddings.patch_embeddings.projection.bias": sd["patch_embed.proj.bias"],
           "layernorm.weight": sd["norm.weight"], "layernorm.bias": sd["norm.bias"]}
    for i in range(DEPTH):
        q, p = f"blocks.{i}.", f"encoder.layer.{i}."
        w, b = sd[q + "attn.qkv.weight"], sd[q + "attn.qkv.bias"]
        for j, n in enumerate(("query", "key", "value")):
            out[p + f"attention.attention.{n}.weight"] = w[j * DIM:(j + 1) * DIM]
            out[p + f"attention.attention.{n}.bias"] = b[j * DIM:(j + 1) * DIM]
        out[p + "attention.output.dense.weight"], out[p + "attention.output.dense.bias"] = sd[q + "attn.proj.weight"], sd[q + "attn.proj.bias"]
        out[p + "layer_scale1.lambda1"], out[p + "layer_scale2.lambda1"] = sd[q + "ls1.gamma"], sd[q + "ls2.gamma"]
        for n in ("norm1", "norm2"):
            out[p + n + ".weight"], out[p + n + ".bias"] = sd[q + n + ".weight"], sd[q + n + ".bias"]
        for n in ("fc1", "fc2"):
            out[p + "mlp." + n + ".weight"], out[p + "mlp." + n + ".bias"] = sd[q + "mlp." + n + ".weight"], sd[q + "mlp." + n + ".bias"]
    return out


def normalize_frames(frames_u8: torch.Tensor) -> torch.Tensor:
    """DataAugmentationPreprocessor.process without augmentation (:231-237) + bhwc -> bchw (:120)."""
    x = frames_u8.permute(0, 3, 1, 2).float() / 255.0
    x = x - torch.tensor(DINO_RGB_MEANS).view(1, 3, 1, 1)
    x = x / torch.tensor(DINO_RGB_STDS).view(1, 3, 1, 1)
    return x


def vit_patch_tokens(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """DinoVisionTransformer.forward_features(x)["x_norm_patchtokens"] for a ViT-S/14 without register tokens."""
    B, _, H, W = x.shape
    ph, pw = H // PATCH, W // PATCH
    t = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=PATCH).flatten(2).transpose(1, 2)
    t = torch.cat([sd["cls_token"].expand(B, -1, -1), t], 1)
    pos = sd["pos_embed"]
    grid = pos[:, 1:].reshape(1, GRID, GRID, DIM).permute(0, 3, 1, 2)
    if (ph, pw) != (GRID, GRID):
        grid = F.interpolate(grid.float(), size=(ph, pw), mode="bicubic", align_corners=False)
    t = t + torch.cat([pos[:, :1], grid.permute(0, 2, 3, 1).reshape(1, ph * pw, DIM)], 1)
    S, dh = t.shape[1], DIM // HEADS
    for i in range(DEPTH):
        q = f"blocks.{i}."
        y = F.layer_norm(t, (DIM,), sd[q + "norm1.weight"], sd[q + "norm1.bias"], 1e-6)
        qkv = (y @ sd[q + "attn.qkv.weight"].T + sd[q + "attn.qkv.bias"]).view(B, S, 3, HEADS, dh).permute(2, 0, 3, 1, 4)
        a = torch.softmax(qkv[0] @ qkv[1].transpose(-1, -2) / math.sqrt(dh), -1) @ qkv[2]
        a = a.transpose(1, 2).reshape(B, S, DIM) @ sd[q + "attn.proj.weight"].T + sd[q + "attn.proj.bias"]
        t = t + sd[q + "ls1.gamma"] * a
        y = F.layer_norm(t, (DIM,), sd[q + "norm2.weight"], sd[q + "norm2.bias"], 1e-6)
        m = F.gelu(y @ sd[q + "mlp.fc1.weight"].T + sd[q + "mlp.fc1.bias"]) @ sd[q + "mlp.fc2.weight"].T + sd[q + "mlp.fc2.bias"]
        t = t + sd[q + "ls2.gamma"] * m
    return F.layer_norm(t, (DIM,), sd["norm.weight"], sd["norm.bias"], 1e-6)[:, 1:]


def dino_preprocess(sd: Dict[str, torch.Tensor], frames_u8: torch.Tensor, crop=(3, 3), pool=(7, 12)) -> torch.Tensor:
    """uint8 [N, H, W, 3] -> fp32 [N, 384, 7, 12]  (DinoViTEmbedder.forward, dino_preprocessors.py:28-36)."""
    x = normalize_frames(frames_u8)
    x = x[:, :, :, crop[0]: x.shape[-1] - crop[1]]
    tok = vit_patch_tokens(sd, x)
    B, _, D = tok.shape
    ph, pw = x.shape[-2] // PATCH, x.shape[-1] // PATCH
    return F.adaptive_avg_pool2d(tok.permute(0, 2, 1).reshape(B, D, ph, pw), pool)
