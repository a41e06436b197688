"""uint8 images -> the 5-D pixel tensor ``PatchEmbedding3D`` consumes, on the GPU.

The reference prepares its input on the host, per image, with torchvision
(``qwen3_5_generate_multimodal.py:40-46``): resize -> ``to_tensor`` -> ``normalize`` -> repeat the frame
``temporal_patch_size`` times -> permute to (B, C, T, H, W); at more than 5 k images/s that loop, and the
2.4 MB/image bf16 upload that follows it, become the bottleneck. Here the (already resized) uint8 images are
uploaded as they are (0.6 MB/image) and one kernel, ``vf_preprocess_u8``, does the rest; with fp32 output the
result is bit-identical to torchvision's.
"""

from __future__ import annotations

import torch

from ... import _lib


def pixels_from_uint8(images_u8: torch.Tensor, image_mean, image_std, temporal_patch_size: int = 2,
                      dtype: torch.dtype = torch.bfloat16, device=None) -> torch.Tensor:
    """images_u8: uint8 [B, H, W, 3] (HWC as PIL/numpy give it; CPU or CUDA) -> [B, 3, T, H, W] on the GPU.

    ``image_mean`` / ``image_std``: the three per-channel constants of ``cfg["image_mean"]`` /
    ``cfg["image_std"]`` (config.py:414-415). Every temporal slot holds the same frame, as in the reference."""
    if images_u8.dim() == 3:
        images_u8 = images_u8.unsqueeze(0)
    assert images_u8.dtype == torch.uint8 and images_u8.shape[-1] == 3, "expected uint8 [B, H, W, 3]"
    if not images_u8.is_cuda:
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        images_u8 = images_u8.to(device, non_blocking=True)
    return _lib.preprocess_u8(images_u8.contiguous(), image_mean, image_std, temporal_patch_size, dtype)
