"""The first consumer of the fused embeddings and MRoPE-I position ids, on libvfuse kernels.

Drop-in for ``MRoPEGatedAttention`` of the reference's ``llm_quest/qwen/qwen3_5/qwen3_5_text_model.py``
(:194-267; base class ``GatedAttention`` in ``qwen3_next/qwen3_next_attention.py:162-201``) in **prefill**:
same constructor, parameter names/shapes (``w_queries_gate``, ``w_keys``, ``w_values``, ``q_norm.scale``,
``k_norm.scale``, ``out_proj``) and ``forward`` signature. The rest of the text model (GatedDeltaNet layers,
MoE/FFN, KV-cache decode, sampling) stays the reference's PyTorch code and is out of scope.

    x -> [one GEMM: w_queries_gate | w_keys | w_values]           -> token-major [B*S, 2*H*hd + 2*G*hd] bf16
      -> [vf_mrope_apply_strided, in place: q heads, k heads]     zero-centred RMSNorm + MRoPE-I, no transposes
      -> [vf_attention_gqa_fwd]  causal GQA, head_dim 256, epilogue multiplies by sigmoid(gate)
      -> [GEMM out_proj]
"""

from __future__ import annotations

import torch
import torch.nn as nn

from ... import _lib
from ..._lib import VF_EPI_BIAS_BF16, VF_EPI_BIAS_F32, VFuseError
from .qwen3_5_vision_model import _Packed, _forward_only_guard


class ZeroCenteredRMSNorm(nn.Module):
    """The reference's zero-centred RMSNorm (qwen3_next_attention.py:20-46): ``scale`` starts at zero and the forward
    multiplies by ``1 + scale``, all in fp32, then casts back. Inside MRoPEGatedAttention it runs fused with MRoPE-I
    (vf_mrope_apply_strided); ``forward`` is the stand-alone vf_rmsnorm_zc kernel."""

    def __init__(self, emb_dim, eps=1e-6, dtype=None):
        super().__init__()
        self.scale = nn.Parameter(torch.zeros(emb_dim, dtype=dtype))
        self.eps = eps
        self._packed = _Packed()

    def one_plus_scale(self):
        # (1.0 + scale) is evaluated in the parameter's dtype, as the reference does, before the fp32 multiply
        return self._packed.get("w", [self.scale], lambda: (1.0 + self.scale.detach()).float().contiguous())

    def forward(self, x):
        _forward_only_guard(self)
        if not x.is_cuda:
            raise VFuseError("ZeroCenteredRMSNorm (llm_quest_b200) runs on CUDA sm_100a only; got a CPU tensor")
        if x.dtype not in (torch.float32, torch.bfloat16):
            return _lib.rmsnorm_zc(x.float(), self.one_plus_scale(), self.eps).to(x.dtype)
        return _lib.rmsnorm_zc(x, self.one_plus_scale(), self.eps)


class MRoPEGatedAttention(nn.Module):
    def __init__(self, cfg, layer_idx=None):
        super().__init__()
        self.d_in = cfg["emb_dim"]
        self.num_heads = cfg["n_heads"]
        self.num_kv_groups = cfg["num_kv_groups"]
        assert self.num_heads % self.num_kv_groups == 0, "num_heads must be divisible by num_kv_groups"
        self.head_dim = cfg["head_dim"]
        self.d_out = self.num_heads * self.head_dim
        self.dtype = cfg["dtype"]
        self.num_repeat = self.num_heads // self.num_kv_groups
        self.layer_idx = layer_idx
        self.mrope_section = cfg["mrope_section"]
        kv = self.num_kv_groups * self.head_dim
        self.w_queries_gate = nn.Linear(self.d_in, self.d_out * 2, bias=False, dtype=self.dtype)
        self.w_keys = nn.Linear(self.d_in, kv, bias=False, dtype=self.dtype)
        self.w_values = nn.Linear(self.d_in, kv, bias=False, dtype=self.dtype)
        self.q_norm = ZeroCenteredRMSNorm(self.head_dim, dtype=self.dtype)
        self.k_norm = ZeroCenteredRMSNorm(self.head_dim, dtype=self.dtype)
        self.out_proj = nn.Linear(self.d_out, self.d_in, bias=False, dtype=self.dtype)
        self._packed = _Packed()

    def _weights(self):
        c = self._packed
        ws = [self.w_queries_gate.weight, self.w_keys.weight, self.w_values.weight]
        w_in = c.get("w_in", ws, lambda: torch.cat([w.detach().to(torch.bfloat16) for w in ws], dim=0).contiguous())
        w_out = c.get("w_out", [self.out_proj.weight], lambda: self.out_proj.weight.detach().to(torch.bfloat16).contiguous())
        return w_in, w_out, self.q_norm.one_plus_scale(), self.k_norm.one_plus_scale()

    def forward(self, x, mask=None, cos=None, sin=None, position_ids=None, attn_mask=None, cache=None):
        """x [b, seq, d_in]; cos/sin [ctx, rot] fp32 tables (GlobalBuffers.get_rope_params); position_ids [3, b, seq]
        (None: 0..seq-1 on all three axes, the text-only case). ``mask`` is the reference's causal mask buffer: the
        kernel applies causality itself and does not read it. Prefill only."""
        _forward_only_guard(self)
        if cache is not None or attn_mask is not None:
            raise VFuseError("MRoPEGatedAttention (llm_quest_b200) covers prefill without KV cache / padding mask; "
                             "decode and padded batches stay on the reference module")
        if self.head_dim != 256:
            raise VFuseError(f"the causal GQA kernel is built for head_dim 256, got {self.head_dim}")
        if not x.is_cuda:
            raise VFuseError("MRoPEGatedAttention (llm_quest_b200) runs on CUDA sm_100a only; got a CPU tensor")
        b, seq, d_in = x.shape
        H, G, hd = self.num_heads, self.num_kv_groups, self.head_dim
        w_in, w_out, qn, kn = self._weights()
        x2d = _lib.to_bf16(x.reshape(-1, d_in))
        n_q, n_kv = 2 * H * hd, G * hd
        proj = torch.empty((b * seq, n_q + 2 * n_kv), dtype=torch.bfloat16, device=x.device)
        _lib.gemm(x2d, w_in, VF_EPI_BIAS_BF16, proj)
        if position_ids is None:
            position_ids = torch.arange(seq, device=x.device).expand(3, b, seq)
        cos = cos.to(device=x.device, dtype=torch.float32).contiguous()
        sin = sin.to(device=x.device, dtype=torch.float32).contiguous()
        # q heads sit at columns h*2*hd (their gate right behind them), k heads at n_q + g*hd
        _lib.mrope_apply_heads_(proj, 0, 2 * hd, b, H, seq, cos, sin, position_ids, self.mrope_section, qn, self.q_norm.eps, hd)
        _lib.mrope_apply_heads_(proj, n_q, hd, b, G, seq, cos, sin, position_ids, self.mrope_section, kn, self.k_norm.eps, hd)
        ctx = torch.empty((b * seq, H * hd), dtype=torch.bfloat16, device=x.device)
        _lib.attention_gqa(proj, proj[:, n_q:n_q + n_kv], proj[:, n_q + n_kv:], ctx, b, seq, H, G, hd**-0.5, True,
                           q_col0=0, q_head_stride=2 * hd, gate2d=proj, gate_col0=hd, gate_head_stride=2 * hd)
        out_dtype = x.dtype if x.dtype in (torch.float32, torch.bfloat16) else torch.float32
        out = torch.empty((b * seq, d_in), dtype=out_dtype, device=x.device)
        _lib.gemm(ctx, w_out, VF_EPI_BIAS_F32 if out_dtype == torch.float32 else VF_EPI_BIAS_BF16, out)
        return out.view(b, seq, d_in)
