"""Explicit forward / backward schedule of one SafeVLA tower on the C-ABI kernels.

A tower = DinoLLAMATxNavActorCritic (reference: architecture/models/allenact_transformer_models/
allenact_dino_transformer.py:326-475 forward, :655-717 goal encoder; decoder
training/online/third_party_models/llama/model.py:425-467).  Nothing here uses autograd or
torch math: every arithmetic step is a libsafevla_b200 launch; torch tensors are only the
buffers.  The encoder (row-independent, ~99 % of the FLOPs) runs over *row chunks* of the
flattened (t, n) axis with either stashed activations or recompute-in-backward, so BASELINE
config 2 (8 192 rows x 117 tokens) fits in HBM; the decoder (needs whole trajectories) runs once.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from . import ops
from ._lib import (ATTN_FULL, ATTN_T5_BIAS, ATTN_TRAJ_CAUSAL, EPI_MASK_BITS, EPI_NONE, EPI_RELU, EPI_RELU_BITS,
                   EPI_RELU_MASK, RowMap)
from .params import D, DC_BINS, DC_HIDDEN, DC_MAX, DC_MIN, DC_SIGMA, DEC_FF, FF, ParamLayout, T5Layout

H, DH = 8, 64
LN_EPS, RMS_EPS, T5_EPS = 1e-5, 1e-5, 1e-6
TOK = 84  # 7 x 12 DINO sites per camera


class TowerWeights:
    """Views of one tower's tensors: fp32 master (`p`), GEMM operand copy (`w`: fp32 master in
    parity mode, bf16 shadow in fast mode) and fp32 gradient (`g`), all slices of flat arenas."""

    def __init__(self, layout: ParamLayout, prefix: str, params: torch.Tensor, grads: torch.Tensor,
                 shadow: Optional[torch.Tensor]):
        self.layout, self.prefix = layout, prefix
        self.params, self.grads, self.shadow = params, grads, shadow

    def _slice(self, arena, name, shape=None, count=1):
        s = self.layout.slots[self.prefix + name]
        n = s.numel * count
        v = arena[s.offset: s.offset + n]
        return v.view(shape if shape is not None else s.shape)

    def p(self, name, shape=None, count=1):
        return self._slice(self.params, name, shape, count)

    def g(self, name, shape=None, count=1):
        return self._slice(self.grads, name, shape, count)

    def w(self, name, shape=None, count=1, dtype=None):
        """GEMM operand: from the bf16 shadow when the consumer runs in bf16."""
        if dtype == torch.bfloat16:
            assert self.shadow is not None
            return self._slice(self.shadow, name, shape, count)
        return self._slice(self.params, name, shape, count)


def _mat(name_shape):  # 2-D weight shape helper for conv weights [O, I, 1, 1]
    return (name_shape[0], int(torch.Size(name_shape[1:]).numel()))


@dataclass
class EncStash:
    """Activations one encoder chunk keeps for its backward."""
    t: Dict[str, torch.Tensor] = field(default_factory=dict)


class Tower:
    def __init__(self, weights: TowerWeights, num_actions: int, num_cameras: int, act_dtype: torch.dtype,
                 cls_only_last_layer: bool = True, split: int = 0):
        self.W = weights
        self.A, self.C = num_actions, num_cameras
        self.adt = act_dtype  # encoder activation dtype
        self.cls_only = cls_only_last_layer
        self.dev = weights.params.device
        self.critic_type = weights.layout.critic_type
        self.dc_support = (torch.linspace(DC_MIN, DC_MAX, DC_BINS + 1, device=self.dev, dtype=torch.float32)
                           if self.critic_type == "discrete" else None)
        # training-mode dropout of the fusion block (nn.TransformerEncoderLayer p = 0.1: attention probabilities,
        # dropout1 / dropout2 before the residual adds, the FFN hidden layer; allenact_dino_transformer.py:545-552).
        # (p, seed, step) are set per forward / backward by the model; masks are regenerated, never stored.
        self.drop_p, self.drop_seed, self.drop_step, self.tower_idx = 0.0, 0, 0, 0
        # parity-grade tensor-core mode: fp32 activations / weights, every tensor-core-shaped product evaluated as
        # 3 (or 6) split-bf16 products in one tcgen05 launch (ops.gemm split=); 0 = operands as they are
        self.split = split
        # residual add of the attention sub-layer inside LayerNorm instead of the out-projection epilogue (encoder_fwd)
        self.ln_res = os.environ.get("SVLA_LN_RES", "1") != "0"

    def _site(self, layer: int, kind: int, row0: int, row_stride: int = 1):
        """Dropout spec of site (tower, layer, kind): kind 0 attention probabilities, 1 dropout1, 2 FFN, 3 dropout2.
        Buffer row r uses mask row row0 + r * row_stride (the CLS-only last layer addresses the full-sequence mask)."""
        return ops.dropout_spec(self.drop_p, self.drop_seed, self.tower_idx * 64 + layer * 8 + kind, self.drop_step, row0,
                                row_stride)

    def _gemm(self, a, b, out, **kw):
        if self.split:
            kw.setdefault("split", self.split)
            kw.setdefault("cache_b", not kw.get("trans_a", False))  # B is a weight in forward / dgrad launches
        return ops.gemm(a, b, out, **kw)

    # ------------------------------------------------------------------ helpers
    def _new(self, *shape, dtype=None):
        return torch.empty(*shape, device=self.dev, dtype=dtype or self.adt)

    def _lin_fwd(self, x, wname, bname, out, *, wshape=None, epi=EPI_NONE, residual=None, count=1):
        w = self.W.w(wname, wshape, count, dtype=x.dtype)
        b = self.W.p(bname, None, count) if bname else None
        if b is not None:
            b = b.reshape(-1)
        return self._gemm(x, w, out, trans_b=True, bias=b, epilogue=epi, residual=residual)

    def _relu_fwd(self, x, wname, bname, out, *, wshape=None, keep=True, dropout=None):
        """out = relu(x W^T + b).  Returns (out, bits): on the bf16 tensor-core path the epilogue also records one bit
        per element (out > 0) for the backward -- the masked dgrad then reads 1/16 of the bytes it would re-read from
        `out` (those K = 512 launches are HBM-bound); bits is None where the plain ReLU epilogue ran."""
        M, N = x.shape[0], out.shape[1]
        bits_ok = x.dtype == torch.bfloat16 and M >= 256 and N % 64 == 0 and N >= 256
        if dropout is not None and not bits_ok:  # small chunks: ReLU launch, then the mask as a row pass
            out, _ = self._lin_fwd(x, wname, bname, out, wshape=wshape, epi=EPI_RELU), None
            return ops.dropout_rows(out, out, dropout), None
        if (keep or dropout is not None) and bits_ok:
            bits = torch.empty(M, N // 32, device=self.dev, dtype=torch.int32)
            w = self.W.w(wname, wshape, 1, dtype=x.dtype)
            self._gemm(x, w, out, trans_b=True, bias=self.W.p(bname).reshape(-1), epilogue=EPI_RELU_BITS, aux=bits,
                       dropout=dropout)
            return out, bits
        return self._lin_fwd(x, wname, bname, out, wshape=wshape, epi=EPI_RELU), None

    def _lin_bwd(self, dy, x, wname, bname, *, wshape=None, dx=None, aux=None, residual=None, count=1, bits=None,
                 alpha=1.0):
        """dW += dy^T x ; db += colsum(dy) ; dx = dy W [* relu'(aux)] [+ residual].  `bits`: the bit record written by
        `_relu_fwd` for the activation `aux` (used instead of re-reading it)."""
        gw = self.W.g(wname, wshape, count)
        if gw.dim() != 2:
            gw = gw.view(gw.shape[0], -1)
        gb = self.W.g(bname, None, count).reshape(-1) if bname else None
        self._gemm(dy, x, gw, trans_a=True, trans_b=False, accumulate=True, colsum_a=gb)  # dW and db in one launch
        if dx is not None:
            w = self.W.w(wname, wshape, count, dtype=dy.dtype)
            if w.dim() != 2:
                w = w.view(w.shape[0], -1)
            if bits is not None and residual is None:
                self._gemm(dy, w, dx, trans_b=False, aux=bits, epilogue=EPI_MASK_BITS, alpha=alpha)
            else:
                self._gemm(dy, w, dx, trans_b=False, aux=aux, epilogue=EPI_RELU_MASK if aux is not None else EPI_NONE,
                           residual=residual, alpha=alpha)
        return dx

    # ------------------------------------------------------------------ encoder
    def encoder_fwd(self, vis: List[torch.Tensor], text_hidden: torch.Tensor, L: int, keep: bool, row_off: int = 0):
        """vis[c]: [Rc*84, 384] token-major DINO features (adt); text_hidden [Rc*L, 512] (adt); row_off: index of the
        chunk's first (t, n) row in the rollout (dropout masks are addressed by global rows).
        Returns (cls [Rc, 512] adt, stash or None)."""
        W, adt = self.W, self.adt
        dp = self.drop_p > 0.0
        if dp and (adt != torch.bfloat16 or 1 + TOK * self.C + L > 256):
            raise NotImplementedError("dropout > 0 is built on the bf16 tensor-core kernels (S <= 256)")
        Rc = vis[0].shape[0] // TOK
        S = 1 + TOK * self.C + L
        ve = "visual_encoder."
        st = EncStash()
        seq = self._new(Rc * S, D)
        ops.fill_rows(W.p(ve + "fusion_token"), seq, Rc, D, dmap=RowMap(1, S, 0))
        cam_tokens = ["visual_sensor_token_raw_navigation_camera", "visual_sensor_token_raw_manipulation_camera"]
        for c in range(self.C):
            Mv = Rc * TOK
            c1, c1b = self._relu_fwd(vis[c], ve + "visual_compressor.0.weight", ve + "visual_compressor.0.bias",
                                     self._new(Mv, D), wshape=(D, 384), keep=keep)
            c2, c2b = self._relu_fwd(c1, ve + "visual_compressor.2.weight", ve + "visual_compressor.2.bias",
                                     self._new(Mv, D), wshape=(D, D), keep=keep)
            a1 = self._lin_fwd(c2, ve + "visual_adapter.0.weight", ve + "visual_adapter.0.bias", self._new(Mv, D))
            mean, rstd = self._new(Mv, dtype=torch.float32), self._new(Mv, dtype=torch.float32)
            ops.layernorm_fwd(a1, W.p(ve + "visual_adapter.1.weight"), W.p(ve + "visual_adapter.1.bias"), seq,
                              token=W.p(ve + cam_tokens[c]), relu=True, eps=LN_EPS,
                              ymap=RowMap(TOK, S, 1 + TOK * c), mean=mean, rstd=rstd, rows=Mv)
            if keep:
                st.t.update({f"c1_{c}": c1, f"c2_{c}": c2, f"a1_{c}": a1, f"vmean_{c}": mean, f"vrstd_{c}": rstd,
                             f"c1b_{c}": c1b, f"c2b_{c}": c2b})
        Mt = Rc * L
        t1 = self._lin_fwd(text_hidden, ve + "text_adapter.0.weight", ve + "text_adapter.0.bias", self._new(Mt, D))
        tmean, trstd = self._new(Mt, dtype=torch.float32), self._new(Mt, dtype=torch.float32)
        ops.layernorm_fwd(t1, W.p(ve + "text_adapter.1.weight"), W.p(ve + "text_adapter.1.bias"), seq, relu=True,
                          eps=LN_EPS, ymap=RowMap(L, S, 1 + TOK * self.C), mean=tmean, rstd=trstd, rows=Mt)
        if keep:
            st.t.update({"t1": t1, "tmean": tmean, "trstd": trstd, "x0": seq})
        x = seq
        Ms = Rc * S
        for l in range(3):
            p = ve + f"fusion_xformer.layers.{l}."
            last = (l == 2) and self.cls_only
            if not last:
                qkv = self._lin_fwd(x, p + "self_attn.in_proj_weight", p + "self_attn.in_proj_bias", self._new(Ms, 3 * D))
                ao, lse = self._new(Ms, D), self._new(Rc * H * S, dtype=torch.float32)
                ops.attn_fwd(ATTN_FULL, qkv[:, 0:D], qkv[:, D:2 * D], qkv[:, 2 * D:3 * D], ao, lse, Rc, S,
                             scale=1.0 / math.sqrt(DH), split=self.split,
                             drop=self._site(l, 0, row_off * H * (256 if S > 128 else 128)) if dp else None)
                m1, r1 = self._new(Ms, dtype=torch.float32), self._new(Ms, dtype=torch.float32)
                m2, r2 = self._new(Ms, dtype=torch.float32), self._new(Ms, dtype=torch.float32)
                if not dp:
                    # out-projection (K = N = 512): its epilogue with the residual operand runs at half the speed of the
                    # plain one, LayerNorm streams the extra operand at 5.8 TB/s -- the residual add lives there
                    # (self.ln_res; s1 then holds the sub-layer output WITHOUT the residual, as in the dropout branch)
                    s1 = self._lin_fwd(ao, p + "self_attn.out_proj.weight", p + "self_attn.out_proj.bias",
                                       self._new(Ms, D), residual=None if self.ln_res else x)
                    x1 = ops.layernorm_fwd(s1, W.p(p + "norm1.weight"), W.p(p + "norm1.bias"), self._new(Ms, D),
                                           res=x if self.ln_res else None, eps=LN_EPS, mean=m1, rstd=r1)
                    hf, hfb = self._relu_fwd(x1, p + "linear1.weight", p + "linear1.bias", self._new(Ms, FF), keep=keep)
                    s2 = self._lin_fwd(hf, p + "linear2.weight", p + "linear2.bias", self._new(Ms, D), residual=x1)
                    x2 = ops.layernorm_fwd(s2, W.p(p + "norm2.weight"), W.p(p + "norm2.bias"), self._new(Ms, D),
                                           eps=LN_EPS, mean=m2, rstd=r2)
                else:
                    # x1 = LN(x + dropout1(attn)) ; x2 = LN(x1 + dropout2(W2 dropout(relu(W1 x1)))): the sub-layer
                    # output is dropped in place and the residual is added inside the LayerNorm kernel; s1 / s2 hold
                    # the DROPPED sub-layer outputs (the LayerNorm backward re-adds the residual)
                    s1 = self._lin_fwd(ao, p + "self_attn.out_proj.weight", p + "self_attn.out_proj.bias", self._new(Ms, D))
                    ops.dropout_rows(s1, s1, self._site(l, 1, row_off * S))
                    x1 = ops.layernorm_fwd(s1, W.p(p + "norm1.weight"), W.p(p + "norm1.bias"), self._new(Ms, D),
                                           res=x, eps=LN_EPS, mean=m1, rstd=r1)
                    hf, hfb = self._relu_fwd(x1, p + "linear1.weight", p + "linear1.bias", self._new(Ms, FF), keep=keep,
                                             dropout=self._site(l, 2, row_off * S))
                    s2 = self._lin_fwd(hf, p + "linear2.weight", p + "linear2.bias", self._new(Ms, D))
                    ops.dropout_rows(s2, s2, self._site(l, 3, row_off * S))
                    x2 = ops.layernorm_fwd(s2, W.p(p + "norm2.weight"), W.p(p + "norm2.bias"), self._new(Ms, D),
                                           res=x1, eps=LN_EPS, mean=m2, rstd=r2)
                if keep:
                    st.t.update({f"x_{l}": x, f"qkv_{l}": qkv, f"ao_{l}": ao, f"lse_{l}": lse, f"s1_{l}": s1,
                                 f"m1_{l}": m1, f"r1_{l}": r1, f"x1_{l}": x1, f"hf_{l}": hf, f"s2_{l}": s2,
                                 f"m2_{l}": m2, f"r2_{l}": r2, f"hfb_{l}": hfb})
                x = x2
            else:
                # Only output token 0 is consumed (allenact_dino_transformer.py:708): K/V for every token,
                # everything else for the CLS row only.  Bit-identical to the full layer on row 0.
                wi = W.w(p + "self_attn.in_proj_weight", dtype=x.dtype)
                bi = W.p(p + "self_attn.in_proj_bias")
                kv = self._gemm(x, wi[D:3 * D], self._new(Ms, 2 * D), trans_b=True, bias=bi[D:3 * D])
                xc = x.view(Rc, S * D)[:, :D]  # CLS rows, leading dimension S*D
                q0 = self._gemm(xc, wi[0:D], self._new(Rc, D), trans_b=True, bias=bi[0:D])
                ao, lse = self._new(Rc, D), self._new(Rc * H, dtype=torch.float32)
                m1, r1 = self._new(Rc, dtype=torch.float32), self._new(Rc, dtype=torch.float32)
                m2, r2 = self._new(Rc, dtype=torch.float32), self._new(Rc, dtype=torch.float32)
                if not dp:
                    ops.attn_cls_fwd(q0, kv[:, 0:D], kv[:, D:2 * D], ao, lse, Rc, S, scale=1.0 / math.sqrt(DH))
                    s1 = self._lin_fwd(ao, p + "self_attn.out_proj.weight", p + "self_attn.out_proj.bias",
                                       self._new(Rc, D), residual=xc)
                    x1 = ops.layernorm_fwd(s1, W.p(p + "norm1.weight"), W.p(p + "norm1.bias"), self._new(Rc, D),
                                           eps=LN_EPS, mean=m1, rstd=r1)
                    hf, hfb = self._relu_fwd(x1, p + "linear1.weight", p + "linear1.bias", self._new(Rc, FF), keep=keep)
                    s2 = self._lin_fwd(hf, p + "linear2.weight", p + "linear2.bias", self._new(Rc, D), residual=x1)
                    x2 = ops.layernorm_fwd(s2, W.p(p + "norm2.weight"), W.p(p + "norm2.bias"), self._new(Rc, D),
                                           eps=LN_EPS, mean=m2, rstd=r2)
                else:
                    # the masks are those of the full layer's CLS rows: row (row_off + r) * S of each [rows * S] site
                    crow = lambda kind: self._site(l, kind, row_off * S, S)
                    xcc = self._new(Rc, D)  # contiguous CLS rows: the LayerNorm residual operand has no row stride
                    ops.copy_rows(x, xcc, Rc, D, smap=RowMap(1, S, 0))
                    ops.attn_cls_fwd(q0, kv[:, 0:D], kv[:, D:2 * D], ao, lse, Rc, S, scale=1.0 / math.sqrt(DH),
                                     drop=self._site(l, 0, row_off * H * (256 if S > 128 else 128)))
                    s1 = self._lin_fwd(ao, p + "self_attn.out_proj.weight", p + "self_attn.out_proj.bias", self._new(Rc, D))
                    ops.dropout_rows(s1, s1, crow(1))
                    x1 = ops.layernorm_fwd(s1, W.p(p + "norm1.weight"), W.p(p + "norm1.bias"), self._new(Rc, D),
                                           res=xcc, eps=LN_EPS, mean=m1, rstd=r1)
                    hf, hfb = self._relu_fwd(x1, p + "linear1.weight", p + "linear1.bias", self._new(Rc, FF), keep=keep,
                                             dropout=crow(2))
                    s2 = self._lin_fwd(hf, p + "linear2.weight", p + "linear2.bias", self._new(Rc, D))
                    ops.dropout_rows(s2, s2, crow(3))
                    x2 = ops.layernorm_fwd(s2, W.p(p + "norm2.weight"), W.p(p + "norm2.bias"), self._new(Rc, D),
                                           res=x1, eps=LN_EPS, mean=m2, rstd=r2)
                if keep:
                    st.t.update({f"x_{l}": x, f"kv_{l}": kv, f"q0_{l}": q0, f"ao_{l}": ao, f"lse_{l}": lse,
                                 f"s1_{l}": s1, f"m1_{l}": m1, f"r1_{l}": r1, f"x1_{l}": x1, f"hf_{l}": hf,
                                 f"s2_{l}": s2, f"m2_{l}": m2, f"r2_{l}": r2, f"hfb_{l}": hfb})
                    if dp:
                        st.t[f"xc_{l}"] = xcc
                return x2, (st if keep else None)
        cls = self._new(Rc, D)
        ops.copy_rows(x, cls, Rc, D, smap=RowMap(1, S, 0))
        return cls, (st if keep else None)

    def encoder_bwd(self, d_cls: torch.Tensor, vis: List[torch.Tensor], text_hidden: torch.Tensor, L: int,
                    st: EncStash, row_off: int = 0):
        """Accumulates every encoder weight gradient of this chunk into the grad arena."""
        W, adt, t = self.W, self.adt, st.t
        dp = self.drop_p > 0.0
        dsc = 1.0 / (1.0 - self.drop_p) if dp else 1.0
        Rc = d_cls.shape[0]
        S = 1 + TOK * self.C + L
        Ms = Rc * S
        ve = "visual_encoder."
        dx = None  # gradient wrt the current layer's output [Ms, D]
        for l in (2, 1, 0):
            p = ve + f"fusion_xformer.layers.{l}."
            last = (l == 2) and self.cls_only
            # mask rows of this layer's [rows, *] buffers: every sequence row, or the CLS row of each sequence
            site = (lambda kind: self._site(l, kind, row_off * S, S)) if last else \
                   (lambda kind: self._site(l, kind, row_off * S))
            if last:
                rows = Rc
                dy = d_cls
            else:
                rows = Ms
                if dx is None:  # full last layer: gradient only on CLS rows
                    dx = torch.zeros(Ms, D, device=self.dev, dtype=adt)
                    ops.copy_rows(d_cls, dx, Rc, D, dmap=RowMap(1, S, 0))
                dy = dx
            ds2 = ops.layernorm_bwd(dy, t[f"s2_{l}"], W.p(p + "norm2.weight"), W.p(p + "norm2.bias"), t[f"m2_{l}"],
                                    t[f"r2_{l}"], self._new(rows, D), W.g(p + "norm2.weight"), W.g(p + "norm2.bias"),
                                    res=t[f"x1_{l}"] if dp else None)
            # gradient of the (dropped) sub-layer output: the residual branch keeps ds2, the FFN branch sees the mask
            dy2 = ops.dropout_rows(ds2, self._new(rows, D), site(3)) if dp else ds2
            dhf = self._lin_bwd(dy2, t[f"hf_{l}"], p + "linear2.weight", p + "linear2.bias", dx=self._new(rows, FF),
                                aux=t[f"hf_{l}"], bits=t.get(f"hfb_{l}"), alpha=dsc)
            del dy2
            dx1 = self._lin_bwd(dhf, t[f"x1_{l}"], p + "linear1.weight", p + "linear1.bias", dx=self._new(rows, D),
                                residual=ds2)
            del dhf, ds2
            ds1 = ops.layernorm_bwd(dx1, t[f"s1_{l}"], W.p(p + "norm1.weight"), W.p(p + "norm1.bias"), t[f"m1_{l}"],
                                    t[f"r1_{l}"], self._new(rows, D), W.g(p + "norm1.weight"), W.g(p + "norm1.bias"),
                                    res=(t[f"xc_{l}"] if last else t[f"x_{l}"]) if dp else
                                    (t[f"x_{l}"] if (self.ln_res and not last) else None))
            del dx1
            dy1 = ops.dropout_rows(ds1, self._new(rows, D), site(1)) if dp else ds1
            dao = self._lin_bwd(dy1, t[f"ao_{l}"], p + "self_attn.out_proj.weight", p + "self_attn.out_proj.bias",
                                dx=self._new(rows, D))
            del dy1
            x = t[f"x_{l}"]
            wi = W.w(p + "self_attn.in_proj_weight", dtype=adt)
            gwi, gbi = W.g(p + "self_attn.in_proj_weight"), W.g(p + "self_attn.in_proj_bias")
            if last:
                kv, q0 = t[f"kv_{l}"], t[f"q0_{l}"]
                dq0, dkv = self._new(Rc, D), self._new(Ms, 2 * D)
                ops.attn_cls_bwd(q0, kv[:, 0:D], kv[:, D:2 * D], t[f"ao_{l}"], dao, dq0, dkv[:, 0:D], dkv[:, D:2 * D],
                                 t[f"lse_{l}"], Rc, S, scale=1.0 / math.sqrt(DH),
                                 drop=self._site(l, 0, row_off * H * (256 if S > 128 else 128)) if dp else None)
                xc = x.view(Rc, S * D)[:, :D]
                # K/V projections of every token
                self._gemm(dkv, x, gwi[D:3 * D], trans_a=True, trans_b=False, accumulate=True, colsum_a=gbi[D:3 * D])
                dxn = self._gemm(dkv, wi[D:3 * D], self._new(Ms, D), trans_b=False)
                # Q projection + residual of the CLS row
                self._gemm(dq0, xc, gwi[0:D], trans_a=True, trans_b=False, accumulate=True, colsum_a=gbi[0:D])
                dxc = self._gemm(dq0, wi[0:D], self._new(Rc, D), trans_b=False, residual=ds1)
                ops.copy_rows(dxc, dxn, Rc, D, dmap=RowMap(1, S, 0), accumulate=True)
                dx = dxn
            else:
                qkv = t[f"qkv_{l}"]
                dqkv = self._new(Ms, 3 * D)
                ops.attn_bwd(ATTN_FULL, qkv[:, 0:D], qkv[:, D:2 * D], qkv[:, 2 * D:3 * D], t[f"ao_{l}"], dao,
                             dqkv[:, 0:D], dqkv[:, D:2 * D], dqkv[:, 2 * D:3 * D], t[f"lse_{l}"], Rc, S,
                             scale=1.0 / math.sqrt(DH), split=self.split,
                             drop=self._site(l, 0, row_off * H * (256 if S > 128 else 128)) if dp else None)
                dx = self._lin_bwd(dqkv, x, p + "self_attn.in_proj_weight", p + "self_attn.in_proj_bias",
                                   dx=self._new(Ms, D), residual=ds1)
                del dqkv
            del dao, ds1
        # ---- input stage: dx is the gradient of the assembled sequence [Rc, S, D]
        ops.colsum(dx.view(Rc, S * D)[:, :D], W.g(ve + "fusion_token"), accumulate=True)
        cam_tokens = ["visual_sensor_token_raw_navigation_camera", "visual_sensor_token_raw_manipulation_camera"]
        for c in range(self.C):
            Mv = Rc * TOK
            da1 = ops.layernorm_bwd(dx, t[f"a1_{c}"], W.p(ve + "visual_adapter.1.weight"),
                                    W.p(ve + "visual_adapter.1.bias"), t[f"vmean_{c}"], t[f"vrstd_{c}"],
                                    self._new(Mv, D), W.g(ve + "visual_adapter.1.weight"),
                                    W.g(ve + "visual_adapter.1.bias"), relu=True, dymap=RowMap(TOK, S, 1 + TOK * c),
                                    dtoken=W.g(ve + cam_tokens[c]), rows=Mv)
            dc2 = self._lin_bwd(da1, t[f"c2_{c}"], ve + "visual_adapter.0.weight", ve + "visual_adapter.0.bias",
                                dx=self._new(Mv, D), aux=t[f"c2_{c}"], bits=t.get(f"c2b_{c}"))
            dc1 = self._lin_bwd(dc2, t[f"c1_{c}"], ve + "visual_compressor.2.weight", ve + "visual_compressor.2.bias",
                                wshape=(D, D), dx=self._new(Mv, D), aux=t[f"c1_{c}"], bits=t.get(f"c1b_{c}"))
            self._lin_bwd(dc1, vis[c], ve + "visual_compressor.0.weight", ve + "visual_compressor.0.bias",
                          wshape=(D, 384))
            del da1, dc2, dc1
        Mt = Rc * L
        dt1 = ops.layernorm_bwd(dx, t["t1"], W.p(ve + "text_adapter.1.weight"), W.p(ve + "text_adapter.1.bias"),
                                t["tmean"], t["trstd"], self._new(Mt, D), W.g(ve + "text_adapter.1.weight"),
                                W.g(ve + "text_adapter.1.bias"), relu=True, dymap=RowMap(L, S, 1 + TOK * self.C),
                                rows=Mt)
        self._lin_bwd(dt1, text_hidden, ve + "text_adapter.0.weight", ve + "text_adapter.0.bias")

    # ------------------------------------------------------------------ decoder + heads
    def decoder_fwd(self, obs_embed, prev_actions, masks, in_hand, time_step, traj_nt, perm_tn, T, N,
                    want_logits: bool, want_values: bool, keep: bool):
        """obs_embed [T*N, 512] (adt); index tensors in [T, N] order; traj_nt int64 [N, T];
        perm_tn[t*N+n] = n*T+t.  Returns dict(logits [T,N,A], values [T,N,1]) fp32 and a stash.
        The residual stream h stays fp32; GEMM operands (normed activations, q/k/v, gate) use the
        tower's activation dtype, so bf16 mode runs on the tcgen05 kernels."""
        W, f32, ddt = self.W, torch.float32, self.adt
        Md = T * N
        t: Dict[str, torch.Tensor] = {}
        x = self._new(Md, D, dtype=f32)  # [N, T, D]
        ops.embed_time_fwd(obs_embed, prev_actions, masks, in_hand, time_step, W.p("last_actions_embed.weight"),
                           W.p("object_in_hand_embed.weight") if in_hand is not None else None,
                           self.div_term, x, T, N, self.A)
        h = x
        for l in range(3):
            p = f"decoder.layers.{l}."
            r1 = self._new(Md, dtype=f32)
            y1 = ops.rmsnorm_fwd(h, W.p(p + "attention_norm.weight"), self._new(Md, D, dtype=ddt), RMS_EPS, r1)
            qkv = self._gemm(y1, W.w(p + "attention.wq.weight", (3 * D, D), 3, dtype=ddt), self._new(Md, 3 * D, dtype=ddt))
            ao, lse = self._new(Md, D, dtype=ddt), self._new(N * H * T, dtype=f32)
            ops.attn_fwd(ATTN_TRAJ_CAUSAL, qkv[:, 0:D], qkv[:, D:2 * D], qkv[:, 2 * D:3 * D], ao, lse, N, T,
                         scale=1.0 / math.sqrt(DH), traj=traj_nt, split=self.split)
            h2 = self._gemm(ao, W.w(p + "attention.wo.weight", dtype=ddt), self._new(Md, D, dtype=f32), residual=h)
            r2 = self._new(Md, dtype=f32)
            y2 = ops.rmsnorm_fwd(h2, W.p(p + "ffn_norm.weight"), self._new(Md, D, dtype=ddt), RMS_EPS, r2)
            ab = self._gemm(y2, W.w(p + "feed_forward.w1.weight", (2 * DEC_FF, D), 2, dtype=ddt),
                          self._new(Md, 2 * DEC_FF, dtype=ddt))
            g = ops.swiglu_fwd(ab, self._new(Md, DEC_FF, dtype=ddt))
            h3 = self._gemm(g, W.w(p + "feed_forward.w2.weight", dtype=ddt), self._new(Md, D, dtype=f32), residual=h2)
            if keep:
                t.update({f"h_{l}": h, f"r1_{l}": r1, f"y1_{l}": y1, f"qkv_{l}": qkv, f"ao_{l}": ao, f"lse_{l}": lse,
                          f"h2_{l}": h2, f"r2_{l}": r2, f"y2_{l}": y2, f"ab_{l}": ab, f"g_{l}": g})
            h = h3
        rf = self._new(Md, dtype=f32)
        yf = ops.rmsnorm_fwd(h, W.p("decoder.norm.weight"), self._new(Md, D, dtype=ddt), RMS_EPS, rf)
        b_nt = self._gemm(yf, W.w("decoder.output.weight", dtype=ddt), self._new(Md, D, dtype=f32))
        b_tn = ops.copy_rows(b_nt, self._new(Md, D, dtype=f32), Md, D, idx=perm_tn)
        out = {}
        if want_logits:
            out["logits"] = self._gemm(b_tn, W.p("actor.linear.weight"), self._new(Md, self.A, dtype=f32),
                                     bias=W.p("actor.linear.bias")).view(T, N, self.A)
        if want_values and self.critic_type == "discrete":
            # DiscreteCriticHead (allenact_dino_transformer.py:743-766): 512 -> 256 -> ReLU -> 101 bin logits, value =
            # HL-Gauss read-out of softmax(logits) (one fused launch)
            h1 = self._gemm(b_tn, W.p("critic.fc.0.weight"), self._new(Md, DC_HIDDEN, dtype=f32),
                            bias=W.p("critic.fc.0.bias"), epilogue=EPI_RELU)
            fl = self._gemm(h1, W.p("critic.fc.2.weight"), self._new(Md, DC_BINS, dtype=f32), bias=W.p("critic.fc.2.bias"))
            _, _, vals = ops.hl_gauss_fwd_bwd(fl, torch.zeros(Md, device=self.dev), self.dc_support, DC_SIGMA,
                                              want_grad=False, want_values=True)
            out["values"], out["full_logits"] = vals.view(T, N, 1), fl.view(T, N, DC_BINS)
            if keep:
                t.update({"dc_h1": h1})
        elif want_values:
            nv = W.p("critic.fc.weight").shape[0]  # 1, or K for the cost tower of the K-cost-channel extension
            out["values"] = self._gemm(b_tn, W.p("critic.fc.weight"), self._new(Md, nv, dtype=f32),
                                     bias=W.p("critic.fc.bias")).view(T, N, nv)
        if keep:
            t.update({"hf": h, "rf": rf, "yf": yf, "b_tn": b_tn})
        return out, t

    def decoder_step(self, obs_embed, prev_actions, masks, in_hand, time_step, cache, pos: int, N: int,
                     want_logits: bool, want_values: bool):
        """One rollout step (T = 1) of the KV-cache decoder (llama/model.py:224-247,279-317 driven by
        allenact_dino_transformer.py:376-406): this step's K / V rows go to row `pos` of the per-layer caches
        `cache[l] = (k, v)`, each [N, max_steps, 512] in the activation dtype, and sampler n attends to rows
        [max(pos - time_step[n], 0), pos] -- the episode-start mask of :386-397.  obs_embed [N, 512];
        prev_actions / time_step int64 [1, N]; masks fp32 [1, N].  Returns dict(logits [1,N,A], values [1,N,1])."""
        W, f32, ddt = self.W, torch.float32, self.adt
        x = self._new(N, D, dtype=f32)
        ops.embed_time_fwd(obs_embed, prev_actions, masks, in_hand, time_step, W.p("last_actions_embed.weight"),
                           W.p("object_in_hand_embed.weight") if in_hand is not None else None,
                           self.div_term, x, 1, N, self.A)
        h = x
        for l in range(3):
            p = f"decoder.layers.{l}."
            ck, cv = cache[l]
            y1 = ops.rmsnorm_fwd(h, W.p(p + "attention_norm.weight"), self._new(N, D, dtype=ddt), RMS_EPS)
            qkv = self._gemm(y1, W.w(p + "attention.wq.weight", (3 * D, D), 3, dtype=ddt), self._new(N, 3 * D, dtype=ddt))
            ops.copy_rows(qkv[:, D:2 * D], ck.view(-1, D), N, D, dmap=RowMap(1, ck.shape[1], pos))
            ops.copy_rows(qkv[:, 2 * D:3 * D], cv.view(-1, D), N, D, dmap=RowMap(1, cv.shape[1], pos))
            ao = ops.attn_decode(qkv[:, 0:D], ck, cv, time_step, pos, self._new(N, D, dtype=ddt),
                                 scale=1.0 / math.sqrt(DH))
            h2 = self._gemm(ao, W.w(p + "attention.wo.weight", dtype=ddt), self._new(N, D, dtype=f32), residual=h)
            y2 = ops.rmsnorm_fwd(h2, W.p(p + "ffn_norm.weight"), self._new(N, D, dtype=ddt), RMS_EPS)
            ab = self._gemm(y2, W.w(p + "feed_forward.w1.weight", (2 * DEC_FF, D), 2, dtype=ddt),
                          self._new(N, 2 * DEC_FF, dtype=ddt))
            g = ops.swiglu_fwd(ab, self._new(N, DEC_FF, dtype=ddt))
            h = self._gemm(g, W.w(p + "feed_forward.w2.weight", dtype=ddt), self._new(N, D, dtype=f32), residual=h2)
        yf = ops.rmsnorm_fwd(h, W.p("decoder.norm.weight"), self._new(N, D, dtype=ddt), RMS_EPS)
        b = self._gemm(yf, W.w("decoder.output.weight", dtype=ddt), self._new(N, D, dtype=f32))
        out = {}
        if want_logits:
            out["logits"] = self._gemm(b, W.p("actor.linear.weight"), self._new(N, self.A, dtype=f32),
                                     bias=W.p("actor.linear.bias")).view(1, N, self.A)
        if want_values and self.critic_type == "discrete":
            h1 = self._gemm(b, W.p("critic.fc.0.weight"), self._new(N, DC_HIDDEN, dtype=f32), bias=W.p("critic.fc.0.bias"),
                            epilogue=EPI_RELU)
            fl = self._gemm(h1, W.p("critic.fc.2.weight"), self._new(N, DC_BINS, dtype=f32), bias=W.p("critic.fc.2.bias"))
            _, _, vals = ops.hl_gauss_fwd_bwd(fl, torch.zeros(N, device=self.dev), self.dc_support, DC_SIGMA,
                                              want_grad=False, want_values=True)
            out["values"], out["full_logits"] = vals.view(1, N, 1), fl.view(1, N, DC_BINS)
        elif want_values:
            nv = W.p("critic.fc.weight").shape[0]
            out["values"] = self._gemm(b, W.p("critic.fc.weight"), self._new(N, nv, dtype=f32),
                                     bias=W.p("critic.fc.bias")).view(1, N, nv)
        return out

    def decoder_bwd(self, dlogits, dvalues, t, prev_actions, masks, in_hand, traj_nt, perm_nt, T, N, dfull=None):
        """Returns d obs_embed [T*N, 512] (adt).  perm_nt[n*T+t] = t*N+n.  dfull: gradient of the discrete critic's
        bin logits [T, N, 101] (the HL-Gauss loss differentiates those, customized_loss.py:364-370)."""
        W, f32, ddt = self.W, torch.float32, self.adt
        Md = T * N

        def operand(x):  # GEMM-operand copy of an fp32 gradient in the activation dtype
            return x if ddt == f32 else ops.copy_rows(x, self._new(Md, D, dtype=ddt), Md, D)

        db_tn = None
        if dlogits is not None:
            dl = dlogits.view(Md, self.A)
            self._gemm(dl, t["b_tn"], W.g("actor.linear.weight"), trans_a=True, trans_b=False, accumulate=True)
            ops.colsum(dl, W.g("actor.linear.bias"), accumulate=True)
            db_tn = self._gemm(dl, W.p("actor.linear.weight"), self._new(Md, D, dtype=f32), trans_b=False)
        if self.critic_type == "discrete":
            if dvalues is not None:
                raise NotImplementedError("discrete critic: gradients flow through the bin logits (extras['full_logits']), "
                                          "as in the reference's losses; the value read-out is not differentiated")
            if dfull is not None:
                dfl = dfull.reshape(Md, DC_BINS).contiguous()
                h1 = t["dc_h1"]
                self._gemm(dfl, h1, W.g("critic.fc.2.weight"), trans_a=True, trans_b=False, accumulate=True)
                ops.colsum(dfl, W.g("critic.fc.2.bias"), accumulate=True)
                dh1 = self._gemm(dfl, W.p("critic.fc.2.weight"), self._new(Md, DC_HIDDEN, dtype=f32), trans_b=False, aux=h1,
                                 epilogue=EPI_RELU_MASK)
                self._gemm(dh1, t["b_tn"], W.g("critic.fc.0.weight"), trans_a=True, trans_b=False, accumulate=True)
                ops.colsum(dh1, W.g("critic.fc.0.bias"), accumulate=True)
                db_tn = self._gemm(dh1, W.p("critic.fc.0.weight"), self._new(Md, D, dtype=f32), trans_b=False,
                                   residual=db_tn)
        elif dvalues is not None:
            dv = dvalues.view(Md, -1)
            self._gemm(dv, t["b_tn"], W.g("critic.fc.weight"), trans_a=True, trans_b=False, accumulate=True)
            ops.colsum(dv, W.g("critic.fc.bias"), accumulate=True)
            db_tn = self._gemm(dv, W.p("critic.fc.weight"), self._new(Md, D, dtype=f32), trans_b=False,
                             residual=db_tn)
        db_nt = ops.copy_rows(db_tn, self._new(Md, D, dtype=ddt), Md, D, idx=perm_nt)
        self._gemm(db_nt, t["yf"], W.g("decoder.output.weight"), trans_a=True, trans_b=False, accumulate=True)
        dyf = self._gemm(db_nt, W.w("decoder.output.weight", dtype=ddt), self._new(Md, D, dtype=ddt), trans_b=False)
        dh = ops.rmsnorm_bwd(dyf, t["hf"], W.p("decoder.norm.weight"), t["rf"], self._new(Md, D, dtype=f32),
                             W.g("decoder.norm.weight"))
        for l in (2, 1, 0):
            p = f"decoder.layers.{l}."
            dh_o = operand(dh)
            self._gemm(dh_o, t[f"g_{l}"], W.g(p + "feed_forward.w2.weight"), trans_a=True, trans_b=False, accumulate=True)
            dg = self._gemm(dh_o, W.w(p + "feed_forward.w2.weight", dtype=ddt), self._new(Md, DEC_FF, dtype=ddt),
                          trans_b=False)
            dab = ops.swiglu_bwd(t[f"ab_{l}"], dg, self._new(Md, 2 * DEC_FF, dtype=ddt))
            self._gemm(dab, t[f"y2_{l}"], W.g(p + "feed_forward.w1.weight", (2 * DEC_FF, D), 2), trans_a=True,
                     trans_b=False, accumulate=True)
            dy2 = self._gemm(dab, W.w(p + "feed_forward.w1.weight", (2 * DEC_FF, D), 2, dtype=ddt),
                           self._new(Md, D, dtype=ddt), trans_b=False)
            ops.rmsnorm_bwd(dy2, t[f"h2_{l}"], W.p(p + "ffn_norm.weight"), t[f"r2_{l}"], dh,
                            W.g(p + "ffn_norm.weight"), accumulate_dx=True)  # dh := d h2
            ao = t[f"ao_{l}"]
            dh_o = operand(dh)
            self._gemm(dh_o, ao, W.g(p + "attention.wo.weight"), trans_a=True, trans_b=False, accumulate=True)
            dao = self._gemm(dh_o, W.w(p + "attention.wo.weight", dtype=ddt), self._new(Md, D, dtype=ddt), trans_b=False)
            qkv = t[f"qkv_{l}"]
            dqkv = self._new(Md, 3 * D, dtype=ddt)
            ops.attn_bwd(ATTN_TRAJ_CAUSAL, qkv[:, 0:D], qkv[:, D:2 * D], qkv[:, 2 * D:3 * D], ao, dao, dqkv[:, 0:D],
                         dqkv[:, D:2 * D], dqkv[:, 2 * D:3 * D], t[f"lse_{l}"], N, T, scale=1.0 / math.sqrt(DH),
                         traj=traj_nt, split=self.split)
            self._gemm(dqkv, t[f"y1_{l}"], W.g(p + "attention.wq.weight", (3 * D, D), 3), trans_a=True, trans_b=False,
                     accumulate=True)
            dy1 = self._gemm(dqkv, W.w(p + "attention.wq.weight", (3 * D, D), 3, dtype=ddt), self._new(Md, D, dtype=ddt),
                           trans_b=False)
            ops.rmsnorm_bwd(dy1, t[f"h_{l}"], W.p(p + "attention_norm.weight"), t[f"r1_{l}"], dh,
                            W.g(p + "attention_norm.weight"), accumulate_dx=True)  # dh := d h
        d_obs = self._new(Md, D)
        ops.embed_time_bwd(dh, prev_actions, masks, in_hand, d_obs, W.g("last_actions_embed.weight"),
                           W.g("object_in_hand_embed.weight") if in_hand is not None else None, T, N, self.A)
        return d_obs


class T5Encoder:
    """Frozen T5-small encoder forward (HF modeling_t5.py) on the C-ABI kernels.  Runs on the
    *unique* prompts of a rollout, once per rollout, shared by the towers.  fp32 mode: fp32 everywhere (parity);
    bf16 mode: bf16 GEMM operands from a bf16 copy of the frozen weights on the tcgen05 kernels, fp32 residual
    stream and fp32 accumulation -- the same split as the decoder."""

    def __init__(self, layout: T5Layout, arena: torch.Tensor, act_dtype: torch.dtype = torch.float32, split: int = 0):
        self.layout, self.arena = layout, arena
        self.dev = arena.device
        self.adt = act_dtype
        self.split = split
        self.shadow: Optional[torch.Tensor] = None
        self._bias_cache: Dict[int, torch.Tensor] = {}
        self.refresh()

    _gemm = Tower._gemm

    def refresh(self):
        """Re-derive everything computed from the frozen weights (call after they are loaded)."""
        self._bias_cache.clear()
        if self.adt == torch.bfloat16:
            self.shadow = ops.cast_bf16(self.arena, self.shadow)

    def w(self, name, shape=None, count=1, operand: bool = False):
        """fp32 master view, or (operand=True, bf16 mode) the bf16 GEMM-operand copy."""
        s = self.layout.slots[name]
        src = self.shadow if (operand and self.shadow is not None) else self.arena
        v = src[s.offset: s.offset + s.numel * count]
        return v.view(shape if shape is not None else s.shape)

    def position_bias(self, L: int) -> torch.Tensor:
        """[H, L, L] relative-position bias table lookup (constant of the frozen weights; built once
        per prompt length with index ops, not part of the per-step compute)."""
        if L not in self._bias_cache:
            from .t5_buckets import relative_position_bucket
            bucket = relative_position_bucket(L).to(self.dev)
            table = self.w("encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight")
            self._bias_cache[L] = table[bucket].permute(2, 0, 1).contiguous()
        return self._bias_cache[L]

    def forward(self, input_ids: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
        """input_ids, attention_mask int64 [U, L] (device) -> last_hidden_state [U*L, 512] fp32."""
        U, L = input_ids.shape
        M = U * L
        f32, odt = torch.float32, self.adt
        new = lambda *s, dtype=f32: torch.empty(*s, device=self.dev, dtype=dtype)  # noqa: E731
        x = ops.copy_rows(self.w("shared.weight"), new(M, D), M, D, idx=input_ids.reshape(-1).contiguous())
        bias = self.position_bias(L)
        km = attention_mask.contiguous()
        for i in range(6):
            a = f"encoder.block.{i}.layer.0."
            y = ops.rmsnorm_fwd(x, self.w(a + "layer_norm.weight"), new(M, D, dtype=odt), T5_EPS)
            qkv = self._gemm(y, self.w(a + "SelfAttention.q.weight", (3 * D, D), 3, operand=True), new(M, 3 * D, dtype=odt))
            ao = ops.attn_fwd(ATTN_T5_BIAS, qkv[:, 0:D], qkv[:, D:2 * D], qkv[:, 2 * D:3 * D], new(M, D, dtype=odt), None,
                              U, L, scale=1.0, bias=bias, keymask=km)
            x = self._gemm(ao, self.w(a + "SelfAttention.o.weight", operand=True), new(M, D), residual=x)
            f = f"encoder.block.{i}.layer.1."
            y = ops.rmsnorm_fwd(x, self.w(f + "layer_norm.weight"), new(M, D, dtype=odt), T5_EPS)
            hf = self._gemm(y, self.w(f + "DenseReluDense.wi.weight", operand=True), new(M, FF, dtype=odt),
                          epilogue=EPI_RELU)
            x = self._gemm(hf, self.w(f + "DenseReluDense.wo.weight", operand=True), new(M, D), residual=x)
        return ops.rmsnorm_fwd(x, self.w("encoder.final_layer_norm.weight"), new(M, D), T5_EPS)
