"""Part-1 ViT building blocks on libvfuse kernels.

Drop-in for the reference's ``llm_quest/multimodal/vision_transformer/vit_transformer_block.py``:
``LayerNorm`` (:12-31, eps added to the *std*), ``GELU`` (:34-44, exact erf), ``FFN`` (:47-67),
``ViTTransformerBlock`` (:70-127). Same constructor arguments and ``state_dict`` keys
(``scale``/``shift``, ``layers.0``/``layers.2``, ``att.*``, ``ln_1``/``ln_2``). Dropout modules
are kept for signature parity; the path is forward/eval only, where they are the identity.
"""

from __future__ import annotations

import torch
import torch.nn as nn

from ... import _lib
from ..._lib import VF_EPI_BIAS_F32, VF_EPI_BIAS_RES_F32, VF_EPI_GELU_ERF_BF16
from ...qwen.qwen3_5.qwen3_5_vision_model import (LN_STATS_IN_CONSUMER_MAX_ROWS, _Packed, _as_2d_bf16, _f32, _fold_ln,
                                                    _forward_only_guard, _w_bf16)
from .vit_attention import ViTMultiHeadAttention


class LayerNorm(nn.Module):
    """y = scale * (x - mean) / (std_biased + 1e-5) + shift  — vf_layernorm variant 1."""

    def __init__(self, emb_dim):
        super().__init__()
        self.eps = 1e-5
        self.scale = nn.Parameter(torch.ones(emb_dim))
        self.shift = nn.Parameter(torch.zeros(emb_dim))
        self._packed = _Packed()

    def packed(self):
        return _f32(self._packed, "s", self.scale), _f32(self._packed, "b", self.shift)

    def forward(self, x):
        _forward_only_guard(self)
        w, b = self.packed()
        x2d = x.reshape(-1, x.shape[-1])
        if x2d.dtype not in (torch.float32, torch.bfloat16):
            x2d = x2d.float()
        out = torch.empty(x2d.shape, dtype=x2d.dtype, device=x.device)
        _lib.layernorm(x2d.contiguous(), w, b, out, self.eps, variant=1)
        return out.view(x.shape).to(x.dtype)


class GELU(nn.Module):
    """Exact (erf) GELU (reference :34-44): x * 0.5 * (1 + erf(x / sqrt 2)). Inside FFN it is the lin1 GEMM's epilogue
    (VF_EPI_GELU_ERF_BF16); called on its own it is the vf_gelu kernel (fp32 or bf16, erff accuracy)."""

    def __init__(self):
        super().__init__()

    def forward(self, x):
        _forward_only_guard(self)
        if x.dtype not in (torch.float32, torch.bfloat16):
            return _lib.gelu(x.float()).to(x.dtype)
        return _lib.gelu(x)


class FFN(nn.Module):
    """emb -> 4*emb -> erf-GELU -> emb; GELU fused into the first GEMM's epilogue."""

    def __init__(self, cfg):
        super().__init__()
        self.layers = nn.Sequential(
            nn.Linear(cfg["emb_dim"], 4 * cfg["emb_dim"]),
            GELU(),
            nn.Linear(4 * cfg["emb_dim"], cfg["emb_dim"]),
        )
        self._packed = _Packed()

    def packed(self):
        c = self._packed
        l0, l2 = self.layers[0], self.layers[2]
        return (_w_bf16(c, "w1", l0.weight), _f32(c, "b1", l0.bias), _w_bf16(c, "w2", l2.weight), _f32(c, "b2", l2.bias))

    def forward(self, x):
        _forward_only_guard(self)
        w1, b1, w2, b2 = self.packed()
        h = _as_2d_bf16(x)
        g = torch.empty((h.shape[0], w1.shape[0]), dtype=torch.bfloat16, device=x.device)
        _lib.gemm(h, w1, VF_EPI_GELU_ERF_BF16, g, bias=b1)
        out = torch.empty((h.shape[0], w2.shape[0]), dtype=torch.float32, device=x.device)
        _lib.gemm(g, w2, VF_EPI_BIAS_F32, out, bias=b2)
        return out.view(*x.shape[:-1], -1).to(x.dtype)


class ViTTransformerBlock(nn.Module):
    """Pre-LN encoder block: x += att(ln_1(x)); x += ffn(ln_2(x))."""

    def __init__(self, cfg):
        super().__init__()
        self.att = ViTMultiHeadAttention(
            d_in=cfg["emb_dim"], d_out=cfg["emb_dim"], dropout=cfg["drop_rate"], num_heads=cfg["n_heads"],
            qkv_bias=cfg["qkv_bias"],
        )
        self.ln_1 = LayerNorm(cfg["emb_dim"])
        self.ln_2 = LayerNorm(cfg["emb_dim"])
        self.ffn = FFN(cfg)
        self.dropout = nn.Dropout(cfg["drop_rate"])

    def run_(self, x2d, B, S, work, ln1_pending=False, emit_next=False):
        """In-place update of the fp32 residual stream x2d [B*S, D].

        With work["stat"] present the two LayerNorms (eps on the std: vf_ln_row_stats / ln_part_in variant 1) are folded
        into the GEMMs around them exactly as in the Qwen tower (qwen3_5_vision_model.Qwen3_5VisionTransformerBlock.run_):
        out_proj / ffn.layers.2 produce bf16(x - shift) and the partial row sums, the packed QKV GEMM / ffn.layers.0
        consume them; the first LayerNorm of the chain runs stand-alone on the fp32 stream and yields the first shift."""
        l1w, l1b = self.ln_1.packed()
        l2w, l2b = self.ln_2.packed()
        w1, b1, w2, b2 = self.ffn.packed()
        h, g, stat, rows, shift = work["h"], work["g"], work.get("stat"), work.get("rows"), work.get("shift")
        D = x2d.shape[1]
        fold = stat is not None
        small = x2d.shape[0] <= LN_STATS_IN_CONSUMER_MAX_ROWS
        c = self.att._packed

        def consumer(norm, colsum):
            if small:
                return (stat, colsum, norm.eps, 1, shift)
            _lib.ln_row_stats(stat, D, norm.eps, rows, shift, variant=1)
            return (rows, colsum)

        if fold and ln1_pending:
            ctx = self.att.attend(h, B, S, folded_norm=self.ln_1, ln_in_fn=consumer)
        else:
            _lib.layernorm(x2d, l1w, l1b, h, self.ln_1.eps, variant=1, mean_out=shift if fold else None)
            ctx = self.att.attend(h, B, S)
        wo, bo = self.att.packed_out()
        producer = (h, stat, shift) if fold else None
        _lib.gemm(ctx, wo, VF_EPI_BIAS_RES_F32, x2d, bias=bo, res=x2d, ln_out=producer)
        if fold:
            w1f, b1f, cs1 = _fold_ln(self.ffn._packed, "fold_l0", self.ffn.layers[0], self.ln_2)
            _lib.gemm(h, w1f, VF_EPI_GELU_ERF_BF16, g, bias=b1f, ln_in=consumer(self.ln_2, cs1))
        else:
            _lib.layernorm(x2d, l2w, l2b, h, self.ln_2.eps, variant=1)
            _lib.gemm(h, w1, VF_EPI_GELU_ERF_BF16, g, bias=b1)
        _lib.gemm(g, w2, VF_EPI_BIAS_RES_F32, x2d, bias=b2, res=x2d, ln_out=producer if emit_next else None)

    def forward(self, x):
        _forward_only_guard(self)
        b, s, d = x.shape
        x2d = _lib.to_f32(x.reshape(-1, d))
        if x2d.data_ptr() == x.data_ptr():
            x2d = x2d.clone()
        work = {
            "h": torch.empty((b * s, d), dtype=torch.bfloat16, device=x.device),
            "g": torch.empty((b * s, 4 * d), dtype=torch.bfloat16, device=x.device),
        }
        self.run_(x2d, b, s, work)
        return x2d.view(b, s, d).to(x.dtype)
