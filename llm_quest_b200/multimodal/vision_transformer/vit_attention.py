"""Part-1 ViT multi-head attention on libvfuse kernels.

Drop-in for ``ViTMultiHeadAttention`` of the reference's
``llm_quest/multimodal/vision_transformer/vit_attention.py`` (:8-91): separate ``w_queries`` /
``w_keys`` / ``w_values`` / ``out_proj`` Linear parameters (same ``state_dict`` keys), no mask,
softmax(QK^T * head_dim^-0.5) V. The three projection weights are packed once into one [3D, D]
bf16 matrix so a single GEMM feeds the fused attention kernel.
"""

from __future__ import annotations

import torch
import torch.nn as nn

from ... import _lib
from ..._lib import VF_EPI_BIAS_BF16, VF_EPI_BIAS_F32, VFuseError
from ...qwen.qwen3_5.qwen3_5_vision_model import _Packed, _as_2d_bf16, _f32, _forward_only_guard, _w_bf16


class ViTMultiHeadAttention(nn.Module):
    def __init__(self, d_in, d_out, dropout, num_heads, qkv_bias=False):
        super().__init__()
        if d_out % num_heads != 0:
            raise ValueError("d_out must be divisible by num_heads")
        self.d_out = d_out
        self.num_heads = num_heads
        self.head_dim = d_out // num_heads
        self.att_scaling = self.head_dim**-0.5
        self.w_queries = nn.Linear(d_in, d_out, bias=qkv_bias)
        self.w_keys = nn.Linear(d_in, d_out, bias=qkv_bias)
        self.w_values = nn.Linear(d_in, d_out, bias=qkv_bias)
        self.dropout = nn.Dropout(dropout)
        self.out_proj = nn.Linear(d_out, d_out)
        self._packed = _Packed()

    def packed_qkv(self):
        ws = [self.w_queries.weight, self.w_keys.weight, self.w_values.weight]
        bs = [self.w_queries.bias, self.w_keys.bias, self.w_values.bias]
        w = self._packed.get("wqkv", ws, lambda: torch.cat([t.detach() for t in ws], 0).to(torch.bfloat16).contiguous())
        b = None
        if bs[0] is not None:
            b = self._packed.get("bqkv", bs, lambda: torch.cat([t.detach() for t in bs], 0).float().contiguous())
        return w, b

    def packed_out(self):
        return _w_bf16(self._packed, "wo", self.out_proj.weight), _f32(self._packed, "bo", self.out_proj.bias)

    def packed_qkv_folded(self, norm):
        """Packed [3D, D] QKV weight with the LayerNorm in front of it folded in (see _fold_ln): (bf16 gamma-scaled weight,
        fp32 folded bias, fp32 colsum of the rounded weight)."""
        ws = [self.w_queries.weight, self.w_keys.weight, self.w_values.weight]
        bs = [self.w_queries.bias, self.w_keys.bias, self.w_values.bias]

        def build():
            w = torch.cat([t.detach().float() for t in ws], 0)
            wf = (w * norm.scale.detach().float()[None, :]).to(torch.bfloat16).contiguous()
            b = w @ norm.shift.detach().float()
            if bs[0] is not None:
                b = b + torch.cat([t.detach().float() for t in bs], 0)
            return wf, b.contiguous(), wf.float().sum(dim=1).contiguous()

        return self._packed.get("wqkv_folded", ws + [t for t in bs if t is not None] + [norm.scale, norm.shift], build)

    def attend(self, h2d, B, S, folded_norm=None, ln_in_fn=None):
        """h2d bf16 [B*S, d_in] -> context bf16 [B*S, d_out]. folded_norm / ln_in_fn: the LayerNorm in front is folded into
        the QKV GEMM (h2d is then the shifted bf16 copy of the un-normalised stream; ln_in_fn(norm, colsum) -> ln_in)."""
        if self.head_dim % 8 != 0 or self.head_dim > 128:
            raise VFuseError(f"vf_attention_fwd_hd takes head dims that are multiples of 8 up to 128, got {self.head_dim}")
        qkv = torch.empty((B * S, 3 * self.d_out), dtype=torch.bfloat16, device=h2d.device)
        if folded_norm is not None:
            wf, bf_, cs = self.packed_qkv_folded(folded_norm)
            _lib.gemm(h2d, wf, VF_EPI_BIAS_BF16, qkv, bias=bf_, ln_in=ln_in_fn(folded_norm, cs))
        else:
            w, b = self.packed_qkv()
            _lib.gemm(h2d, w, VF_EPI_BIAS_BF16, qkv, bias=b)
        ctx = torch.empty((B * S, self.d_out), dtype=torch.bfloat16, device=h2d.device)
        _lib.attention(qkv, ctx, B, S, self.num_heads, self.att_scaling, head_dim=self.head_dim)
        return ctx

    def forward(self, x):
        _forward_only_guard(self)
        b, seq_len, d_in = x.shape
        ctx = self.attend(_as_2d_bf16(x), b, seq_len)
        wo, bo = self.packed_out()
        out = torch.empty((b * seq_len, self.d_out), dtype=torch.float32, device=x.device)
        _lib.gemm(ctx, wo, VF_EPI_BIAS_F32, out, bias=bo)
        return out.view(b, seq_len, self.d_out).to(x.dtype)
