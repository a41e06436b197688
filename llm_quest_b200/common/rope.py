"""Rotary embeddings of the vision-encode-and-fuse path, backed by libvfuse kernels.

Mirrors the public surface of the reference's ``llm_quest/common/rope.py`` that the path uses:
``VisionRoPE.compute_angles_2d`` (:400-482), ``VisionRoPE.apply`` (:485-500), ``RoPE.apply``
(:180-243), ``RoPE.interleave_mrope_coeffs`` (:246-294) and ``RoPE.apply_mrope`` (:297-358).
YaRN / NTK scaling (:32-94) is text-only context extension and is out of scope.

``apply`` / ``apply_mrope`` run the CUDA kernels in ``csrc/vf_rope.cu``; table construction is
init-time host math and stays in torch.
"""

from __future__ import annotations

import torch

from .. import _lib


class RoPE:
    @staticmethod
    def compute_angles(base, head_dim, ctx_len, smooth_scaling_cfg=None, ntk_aware_scaling=True,
                       rotation_factor=1.0, dtype=torch.float32):
        """cos/sin [ctx_len, rot] for 1-D RoPE with optional partial rotation (reference :97-166)."""
        assert head_dim % 2 == 0, "head dim must be divisible by 2 as we need pairs"
        if smooth_scaling_cfg is not None:
            raise NotImplementedError("YaRN/NTK wavelength scaling is outside the vision-encode-and-fuse path")
        rot = int(head_dim * rotation_factor) if rotation_factor != 1.0 else head_dim
        inv_freq = 1.0 / base ** (2 * torch.arange(0, rot // 2, dtype=dtype) / rot)
        ang = torch.outer(torch.arange(0, ctx_len, dtype=dtype), inv_freq)
        ang = torch.cat([ang, ang], dim=-1)
        return torch.cos(ang), torch.sin(ang)

    @staticmethod
    def apply(x, cos, sin, position_ids=None):
        """Rotate-half RoPE on x [b, heads, seq, head_dim]; rotates the first cos.shape[-1] dims."""
        b, n_head, seq_length, head_dim = x.shape
        assert head_dim % 2 == 0, "head dim must be divisible by 2 as we need pairs"
        cos = cos.to(device=x.device, dtype=torch.float32).contiguous()
        sin = sin.to(device=x.device, dtype=torch.float32).contiguous()
        return _lib.rope_apply(x, cos, sin, position_ids)

    @staticmethod
    def interleave_mrope_coeffs(cos, sin, mrope_section):
        """[3, b, s, half] per-axis coefficients -> [b, s, half] in THWTHW... order (reference :246-294)."""
        half = cos.shape[-1]
        slot = torch.arange(half, device=cos.device)
        axis = torch.zeros(half, dtype=torch.long, device=cos.device)
        axis[(slot % 3 == 1) & (slot < 3 * mrope_section[1])] = 1
        axis[(slot % 3 == 2) & (slot < 3 * mrope_section[2])] = 2
        pick = axis.view(1, 1, 1, half).expand(1, *cos.shape[1:])
        return cos.gather(0, pick)[0], sin.gather(0, pick)[0]

    @staticmethod
    def apply_mrope(x, cos, sin, position_ids, mrope_section, norm_weight=None, norm_eps=1e-6):
        """MRoPE-I on x [b, heads, seq, head_dim] with position_ids [3, b, seq].

        ``norm_weight`` (fp32 [head_dim], already ``1 + scale``) additionally fuses the zero-centred
        RMSNorm the text model applies right before (qwen3_5_text_model.py:227-233).
        """
        cos = cos.to(device=x.device, dtype=torch.float32).contiguous()
        sin = sin.to(device=x.device, dtype=torch.float32).contiguous()
        return _lib.mrope_apply(x, cos, sin, position_ids, mrope_section, norm_weight, norm_eps)


class VisionRoPE:
    @staticmethod
    def compute_angles_2d(base, head_dim, height_patches, width_patches, num_frames=1, dtype=torch.float32):
        """Axial 2-D tables [num_frames*H*W, head_dim]: first quarter of the dims rotates with the patch
        row, second quarter with the patch column, then both are duplicated (reference :400-482)."""
        assert head_dim % 4 == 0, "head_dim must be divisible by 4 for 2D RoPE"
        half = head_dim // 2
        inv_freq = 1.0 / (base ** (2 * torch.arange(0, half // 2, dtype=dtype) / half))
        r = torch.arange(height_patches, dtype=dtype).repeat_interleave(width_patches)
        c = torch.arange(width_patches, dtype=dtype).repeat(height_patches)
        ang = torch.cat([torch.outer(r, inv_freq), torch.outer(c, inv_freq)], dim=-1)
        if num_frames > 1:
            ang = ang.repeat(num_frames, 1)
        ang = torch.cat([ang, ang], dim=-1)
        return torch.cos(ang), torch.sin(ang)

    @staticmethod
    def apply(x, cos, sin, position_ids=None):
        return RoPE.apply(x, cos, sin, position_ids)
