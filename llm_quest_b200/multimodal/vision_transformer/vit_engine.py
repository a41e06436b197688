"""ViT -> LLM adapter on libvfuse GEMMs.

Drop-in for ``ViTAdapter`` of the reference's
``llm_quest/multimodal/vision_transformer/vit_engine.py`` (:9-59): 'simple' = one Linear,
'ffn' = Linear -> GELU(erf) -> (Dropout) -> Linear, ``bias=False`` by default, same ``state_dict``
keys (``adapter.weight`` or ``adapter.{0,3}.weight``). The training/eval loops of that file
(:62-265) are out of scope.

``forward_into`` additionally writes the adapter rows straight into a pre-allocated fused
``[b, n_vision + n_text, d]`` buffer (the ``torch.cat([vision, text], dim=1)`` of
``multimodal/vlm_engine.py:114`` / ``vlm_generation.py:66``) through the GEMM's row remap.
"""

from __future__ import annotations

import torch

from ... import _lib
from ..._lib import VF_EPI_BIAS_F32, VF_EPI_GELU_ERF_BF16
from ...qwen.qwen3_5.qwen3_5_vision_model import _Packed, _as_2d_bf16, _f32, _forward_only_guard, _w_bf16


class ViTAdapter(torch.nn.Module):
    def __init__(self, vit_d_out, llm_d_in, adapter_type="simple", hidden_size_factor=4, bias=False, dropout=0.0,
                 dtype=torch.float32):
        super().__init__()
        if adapter_type == "simple":
            self.adapter = torch.nn.Linear(vit_d_out, llm_d_in, bias=bias, dtype=dtype)
        elif adapter_type == "ffn":
            self.adapter = torch.nn.Sequential(
                torch.nn.Linear(vit_d_out, vit_d_out * hidden_size_factor, bias=bias, dtype=dtype),
                torch.nn.GELU(),
                torch.nn.Dropout(dropout) if dropout > 0.0 else torch.nn.Identity(),
                torch.nn.Linear(vit_d_out * hidden_size_factor, llm_d_in, bias=bias, dtype=dtype),
            )
        else:
            raise ValueError(f"Invalid adapter type: {adapter_type}")
        self._packed = _Packed()

    def _project(self, h, out, **remap):
        c = self._packed
        if isinstance(self.adapter, torch.nn.Linear):
            _lib.gemm(h, _w_bf16(c, "w", self.adapter.weight), VF_EPI_BIAS_F32, out, bias=_f32(c, "b", self.adapter.bias), **remap)
            return out
        l0, l3 = self.adapter[0], self.adapter[3]
        w0 = _w_bf16(c, "w0", l0.weight)
        g = torch.empty((h.shape[0], w0.shape[0]), dtype=torch.bfloat16, device=h.device)
        _lib.gemm(h, w0, VF_EPI_GELU_ERF_BF16, g, bias=_f32(c, "b0", l0.bias))
        _lib.gemm(g, _w_bf16(c, "w3", l3.weight), VF_EPI_BIAS_F32, out, bias=_f32(c, "b3", l3.bias), **remap)
        return out

    def out_features(self):
        return (self.adapter if isinstance(self.adapter, torch.nn.Linear) else self.adapter[3]).out_features

    def forward(self, x):
        _forward_only_guard(self)
        h = _as_2d_bf16(x)
        out = torch.empty((h.shape[0], self.out_features()), dtype=torch.float32, device=x.device)
        self._project(h, out)
        return out.view(*x.shape[:-1], -1).to(x.dtype)

    def forward_into(self, x, fused, row_off=0):
        """x [b, n_vis, d_vit]; fused fp32 [b, n_total, d_llm]: writes adapter(x) to fused[:, row_off:row_off+n_vis]."""
        _forward_only_guard(self)
        b, n_vis, _ = x.shape
        assert fused.dtype == torch.float32 and fused.is_contiguous() and fused.shape[0] == b
        self._project(_as_2d_bf16(x), fused.view(-1, fused.shape[-1]), grp_rows=n_vis, grp_stride=fused.shape[1],
                      row_off=row_off)
        return fused
