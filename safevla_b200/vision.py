"""B200DinoViTPreprocessor: the rollout-side vision encoder -- drop-in for
`DataAugmentationPreprocessor` (normalisation, augmentation off) + `DinoViTPreprocessor`
(architecture/allenact_preprocessors/dino_preprocessors.py:20-38,41-125,171-239; the IL-side twin is
architecture/models/transformer_models/image_encoders.py:56-72): uint8 camera frames [N, 224, 384, 3] ->
/255, mean/std normalisation, crop [:, :, :, 3:-3] -> DINOv2 ViT-S/14 forward_features ->
x_norm_patchtokens [N, 432, 384] -> [N, 384, 16, 27] -> AdaptiveAvgPool2d((7, 12)) -> fp32 [N, 384, 7, 12],
i.e. exactly the `rgb_dinov2` / `manipulation_rgb_dinov2` tensors the update path consumes.

Every arithmetic step is a libsafevla_b200 launch: the fused normalise+crop+im2col kernel feeds the
patch-embedding GEMM (K = 588 padded to 592), the 12 blocks run LayerNorm -> QKV GEMM -> tcgen05 flash attention
(433 tokens, 6 heads x 64) -> projection GEMM (+ residual) -> LayerNorm -> MLP GEMMs (exact-GELU epilogue,
+ residual), then the final LayerNorm and the pooling kernel.  LayerScale is folded into the projection / fc2
weights and biases when the (frozen) weights are loaded: ls * (W x + b) = (ls . W) x + ls . b.

The ViT itself is a third-party dependency of the reference (torch.hub `facebookresearch/dinov2`, un-pinned;
weights not available offline).  Its published architecture is restated here and in oracle/vit_oracle.py and pinned
against HuggingFace transformers' architecture-identical `Dinov2Model` (tests/golden/dinov2_*.pt); both the hub
(`blocks.N.attn.qkv...`) and the HF (`encoder.layer.N.attention...`) state-dict layouts load.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from . import ops
from ._lib import ATTN_FULL, EPI_GELU

DINO_RGB_MEANS = (0.48145466, 0.4578275, 0.40821073)   # dino_preprocessors.py:42-43
DINO_RGB_STDS = (0.26862954, 0.26130258, 0.27577711)
PATCH, DIM, HEADS, DEPTH, MLP = 14, 384, 6, 12, 1536
LN_EPS = 1e-6


def init_hub_state_dict(seed: int) -> Dict[str, torch.Tensor]:
    """Seeded random weights in the torch.hub dinov2_vits14 layout (deterministic CPU generator) -- the synthetic
    stand-in for the hub checkpoint that is not available offline; benchmarks, tools and the golden recipes use it."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s, std=1.0: torch.randn(*s, generator=g) * std  # noqa: E731
    sd = {"cls_token": rn(1, 1, DIM, std=0.5), "pos_embed": rn(1, 1 + 37 * 37, DIM, std=0.3),
          "mask_token": torch.zeros(1, DIM),
          "patch_embed.proj.weight": rn(DIM, 3, PATCH, PATCH, std=1.0 / math.sqrt(3 * PATCH * PATCH)),
          "patch_embed.proj.bias": rn(DIM, std=0.1)}
    for i in range(DEPTH):
        q = f"blocks.{i}."
        sd.update({
            q + "norm1.weight": 1 + rn(DIM, std=0.1), q + "norm1.bias": rn(DIM, std=0.1),
            q + "attn.qkv.weight": rn(3 * DIM, DIM, std=1.5 / math.sqrt(DIM)), q + "attn.qkv.bias": rn(3 * DIM, std=0.1),
            q + "attn.proj.weight": rn(DIM, DIM, std=1.0 / math.sqrt(DIM)), q + "attn.proj.bias": rn(DIM, std=0.1),
            q + "ls1.gamma": 0.5 + torch.rand(DIM, generator=g),
            q + "norm2.weight": 1 + rn(DIM, std=0.1), q + "norm2.bias": rn(DIM, std=0.1),
            q + "mlp.fc1.weight": rn(MLP, DIM, std=1.0 / math.sqrt(DIM)), q + "mlp.fc1.bias": rn(MLP, std=0.1),
            q + "mlp.fc2.weight": rn(DIM, MLP, std=1.0 / math.sqrt(MLP)), q + "mlp.fc2.bias": rn(DIM, std=0.1),
            q + "ls2.gamma": 0.5 + torch.rand(DIM, generator=g),
        })
    sd["norm.weight"], sd["norm.bias"] = 1 + rn(DIM, std=0.1), rn(DIM, std=0.1)
    return sd


def hub_to_canonical(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """facebookresearch/dinov2 (torch.hub) or HF Dinov2Model state dict -> one canonical layout."""
    sd = {k[len("model."):] if k.startswith("model.") else k: v for k, v in sd.items()}
    out: Dict[str, torch.Tensor] = {}
    if "embeddings.cls_token" in sd:  # HuggingFace transformers
        e = "embeddings."
        out["cls"], out["pos"] = sd[e + "cls_token"].reshape(-1), sd[e + "position_embeddings"].reshape(-1, DIM)
        out["patch_w"], out["patch_b"] = sd[e + "patch_embeddings.projection.weight"], sd[e + "patch_embeddings.projection.bias"]
        for i in range(DEPTH):
            p, q = f"encoder.layer.{i}.", f"blocks.{i}."
            a = p + "attention.attention."
            out[q + "qkv_w"] = torch.cat([sd[a + "query.weight"], sd[a + "key.weight"], sd[a + "value.weight"]], 0)
            out[q + "qkv_b"] = torch.cat([sd[a + "query.bias"], sd[a + "key.bias"], sd[a + "value.bias"]], 0)
            out[q + "proj_w"], out[q + "proj_b"] = sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"]
            out[q + "ls1"], out[q + "ls2"] = sd[p + "layer_scale1.lambda1"], sd[p + "layer_scale2.lambda1"]
            for n in ("norm1", "norm2"):
                out[q + n + "_w"], out[q + n + "_b"] = sd[p + n + ".weight"], sd[p + n + ".bias"]
            for n in ("fc1", "fc2"):
                out[q + n + "_w"], out[q + n + "_b"] = sd[p + "mlp." + n + ".weight"], sd[p + "mlp." + n + ".bias"]
        out["norm_w"], out["norm_b"] = sd["layernorm.weight"], sd["layernorm.bias"]
    else:  # torch.hub facebookresearch/dinov2
        out["cls"], out["pos"] = sd["cls_token"].reshape(-1), sd["pos_embed"].reshape(-1, DIM)
        out["patch_w"], out["patch_b"] = sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"]
        for i in range(DEPTH):
            q = f"blocks.{i}."
            out[q + "qkv_w"], out[q + "qkv_b"] = sd[q + "attn.qkv.weight"], sd[q + "attn.qkv.bias"]
            out[q + "proj_w"], out[q + "proj_b"] = sd[q + "attn.proj.weight"], sd[q + "attn.proj.bias"]
            out[q + "ls1"], out[q + "ls2"] = sd[q + "ls1.gamma"], sd[q + "ls2.gamma"]
            for n in ("norm1", "norm2"):
                out[q + n + "_w"], out[q + n + "_b"] = sd[q + n + ".weight"], sd[q + n + ".bias"]
            for n in ("fc1", "fc2"):
                out[q + n + "_w"], out[q + n + "_b"] = sd[q + "mlp." + n + ".weight"], sd[q + "mlp." + n + ".bias"]
        out["norm_w"], out["norm_b"] = sd["norm.weight"], sd["norm.bias"]
    return {k: v.detach().float() for k, v in out.items()}


def interpolate_pos_embed(pos: torch.Tensor, ph: int, pw: int) -> torch.Tensor:
    """[1 + s*s, D] -> [1 + ph*pw, D]: bicubic resize of the patch-position grid (DINOv2 interpolate_pos_encoding;
    a constant of the frozen weights and the frame size, built once on the host)."""
    n = pos.shape[0] - 1
    s = int(math.isqrt(n))
    assert s * s == n
    if (ph, pw) == (s, s):
        return pos.clone()
    grid = pos[1:].reshape(1, s, s, -1).permute(0, 3, 1, 2).float()
    grid = F.interpolate(grid, size=(ph, pw), mode="bicubic", align_corners=False)
    return torch.cat([pos[:1], grid.permute(0, 2, 3, 1).reshape(ph * pw, -1)], 0)


class B200DinoViTPreprocessor:
    """`process({uuid: uint8 [N, H, W, 3]}) -> fp32 [N, 384, 7, 12]`  (dino_preprocessors.py:119-125)."""

    def __init__(self, rgb_input_uuid: str, state_dict: Dict[str, torch.Tensor], *, precision: str = "bf16",
                 device: Optional[torch.device] = None, crop: Tuple[int, int] = (3, 3), pool: Tuple[int, int] = (7, 12),
                 mean=DINO_RGB_MEANS, stdev=DINO_RGB_STDS, chunk_frames: int = 256):
        if not torch.cuda.is_available():
            raise RuntimeError("B200DinoViTPreprocessor needs a CUDA device (sm_100a); there is no CPU fallback")
        assert precision in ("bf16", "fp32")
        self.input_uuids = [rgb_input_uuid]
        self.dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.adt = torch.bfloat16 if precision == "bf16" else torch.float32
        self.crop, self.pool, self.mean, self.std, self.chunk = crop, pool, tuple(mean), tuple(stdev), chunk_frames
        self.kpad = (3 * PATCH * PATCH + 7) // 8 * 8  # 588 -> 592: 16-byte rows for TMA
        c = hub_to_canonical(state_dict)
        f32 = lambda t: t.to(self.dev, torch.float32).contiguous()  # noqa: E731
        op = lambda t: t.to(self.dev, self.adt).contiguous()        # noqa: E731  GEMM operand copy
        pw = torch.zeros(DIM, self.kpad)
        pw[:, : 3 * PATCH * PATCH] = c["patch_w"].reshape(DIM, -1)
        self.w = {"patch_w": op(pw), "patch_b": f32(c["patch_b"]), "cls": f32(c["cls"]), "norm_w": f32(c["norm_w"]),
                  "norm_b": f32(c["norm_b"])}
        self.pos_raw = c["pos"]
        self._pos: Dict[Tuple[int, int], torch.Tensor] = {}
        for i in range(DEPTH):
            q = f"blocks.{i}."
            ls1, ls2 = c[q + "ls1"], c[q + "ls2"]
            self.w.update({
                q + "norm1_w": f32(c[q + "norm1_w"]), q + "norm1_b": f32(c[q + "norm1_b"]),
                q + "norm2_w": f32(c[q + "norm2_w"]), q + "norm2_b": f32(c[q + "norm2_b"]),
                q + "qkv_w": op(c[q + "qkv_w"]), q + "qkv_b": f32(c[q + "qkv_b"]),
                # LayerScale folded into the (frozen) projection / fc2: ls * (W x + b) = (ls . W) x + ls . b
                q + "proj_w": op(ls1[:, None] * c[q + "proj_w"]), q + "proj_b": f32(ls1 * c[q + "proj_b"]),
                q + "fc1_w": op(c[q + "fc1_w"]), q + "fc1_b": f32(c[q + "fc1_b"]),
                q + "fc2_w": op(ls2[:, None] * c[q + "fc2_w"]), q + "fc2_b": f32(ls2 * c[q + "fc2_b"]),
            })

    def to(self, device):  # allenact Preprocessor surface (:113-117)
        assert torch.device(device) == self.dev, "weights live on the device the preprocessor was built for"
        return self

    def _pos_embed(self, ph: int, pw: int) -> torch.Tensor:
        if (ph, pw) not in self._pos:
            self._pos[(ph, pw)] = interpolate_pos_embed(self.pos_raw, ph, pw).to(self.dev).contiguous()
        return self._pos[(ph, pw)]

    @torch.no_grad()
    def encode(self, frames: torch.Tensor) -> torch.Tensor:
        """uint8 [N, H, W, 3] -> fp32 [N, 384, OH, OW]."""
        assert frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[-1] == 3
        out = torch.empty(frames.shape[0], DIM, *self.pool, device=self.dev)
        for n0 in range(0, frames.shape[0], self.chunk):
            self._encode_chunk(frames[n0:n0 + self.chunk].to(self.dev, non_blocking=True).contiguous(),
                               out[n0:n0 + self.chunk])
        return out

    def _encode_chunk(self, img: torch.Tensor, out: torch.Tensor):
        N, H, W, _ = img.shape
        ph, pw = H // PATCH, (W - sum(self.crop)) // PATCH
        P, S = ph * pw, ph * pw + 1
        M = N * S
        w, adt, dev = self.w, self.adt, self.dev
        new = lambda *s: torch.empty(*s, device=dev, dtype=adt)  # noqa: E731
        patches = ops.patchify_u8(img, new(N * P, self.kpad), PATCH, self.crop[0], self.crop[1], self.mean, self.std)
        pe = ops.gemm(patches, w["patch_w"], new(N * P, DIM), trans_b=True, bias=w["patch_b"])
        x = ops.vit_assemble(pe, w["cls"], self._pos_embed(ph, pw), new(M, DIM), N, P)
        for i in range(DEPTH):
            q = f"blocks.{i}."
            y = ops.layernorm_fwd(x, w[q + "norm1_w"], w[q + "norm1_b"], new(M, DIM), eps=LN_EPS)
            qkv = ops.gemm(y, w[q + "qkv_w"], new(M, 3 * DIM), trans_b=True, bias=w[q + "qkv_b"])
            ao = ops.attn_fwd(ATTN_FULL, qkv[:, 0:DIM], qkv[:, DIM:2 * DIM], qkv[:, 2 * DIM:3 * DIM], new(M, DIM), None, N, S,
                              H=HEADS, scale=1.0 / math.sqrt(DIM // HEADS))
            x = ops.gemm(ao, w[q + "proj_w"], new(M, DIM), trans_b=True, bias=w[q + "proj_b"], residual=x)
            y = ops.layernorm_fwd(x, w[q + "norm2_w"], w[q + "norm2_b"], new(M, DIM), eps=LN_EPS)
            hdn = ops.gemm(y, w[q + "fc1_w"], new(M, MLP), trans_b=True, bias=w[q + "fc1_b"], epilogue=EPI_GELU)
            x = ops.gemm(hdn, w[q + "fc2_w"], new(M, DIM), trans_b=True, bias=w[q + "fc2_b"], residual=x)
        xn = ops.layernorm_fwd(x, w["norm_w"], w["norm_b"], new(M, DIM), eps=LN_EPS)
        ops.tokens_pool(xn, out, N, ph, pw, self.pool[0], self.pool[1])

    def process(self, obs: Dict[str, torch.Tensor], *args, **kwargs) -> torch.Tensor:
        return self.encode(obs[self.input_uuids[0]])
