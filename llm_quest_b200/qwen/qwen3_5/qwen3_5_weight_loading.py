"""Hugging Face -> llm_quest_b200 name mapping for the Qwen3.5 vision tower.

Mirror of the vision half of the reference's ``llm_quest/qwen/qwen3_5/qwen3_5_weight_loading.py``
(``get_vision_remapping_rules`` :60-81, ``load_qwen3_5_vlm_weights`` :119-178) without the download:
there is no network on the build/bench machines, so the entry points take an already loaded HF
``state_dict`` (or a local ``.safetensors`` / ``.pt`` file). The text model is out of scope.
"""

from __future__ import annotations

from pathlib import Path

import torch

_HF_VISUAL = "model.visual."
# (HF substring, ours) applied in order to every key under ``model.visual.``; norm1/norm2 keep their names
_VISION_RULES = [
    ("model.visual.patch_embed.proj.", "patch_embed.conv_proj."),
    ("model.visual.pos_embed.", "pos_embed."),
    ("model.visual.blocks.", "blocks."),
    (".attn.qkv.", ".att.qkv."),
    (".attn.proj.", ".att.proj."),
    (".mlp.linear_fc1.", ".ffn.lin1."),
    (".mlp.linear_fc2.", ".ffn.lin2."),
    ("model.visual.merger.norm.", "merge_adapter.norm."),
    ("model.visual.merger.linear_fc1.", "merge_adapter.lin1."),
    ("model.visual.merger.linear_fc2.", "merge_adapter.lin2."),
]


def get_vision_remapping_rules():
    """Same (hf_name_part, our_name_part) pairs as the reference's function of this name."""
    return list(_VISION_RULES)


def remap_vision_key(hf_key: str) -> str | None:
    """Our parameter name for an HF key, or None when the key is not part of the vision tower."""
    if not hf_key.startswith(_HF_VISUAL):
        return None
    name = hf_key
    for src, dst in _VISION_RULES:
        name = name.replace(src, dst)
    return None if name.startswith(_HF_VISUAL) else name


def convert_vision_weights(hf_state_dict, model_state_dict):
    """{our key: tensor} for every HF vision tensor that has a same-shaped counterpart in the model.
    Raises on a shape mismatch (a silently skipped tensor would leave random weights in place)."""
    out = {}
    for k, v in hf_state_dict.items():
        name = remap_vision_key(k)
        if name is None or name not in model_state_dict:
            continue
        want = tuple(model_state_dict[name].shape)
        if tuple(v.shape) != want:
            raise ValueError(f"{k} -> {name}: shape {tuple(v.shape)} does not match the model's {want}")
        out[name] = v
    return out


def load_qwen3_5_vision_weights(vision_model, source):
    """Load HF vision weights into a ``Qwen3_5VisionModel``. ``source``: HF state_dict, or a path to a
    ``.safetensors`` / torch file. Returns (missing_keys, unexpected_keys) like ``load_state_dict``."""
    if isinstance(source, (str, Path)):
        path = Path(source)
        if path.suffix == ".safetensors":
            from safetensors.torch import load_file  # optional dependency, only for this branch

            source = load_file(str(path))
        else:
            source = torch.load(str(path), map_location="cpu", weights_only=True)
    converted = convert_vision_weights(source, vision_model.state_dict())
    with torch.no_grad():
        res = vision_model.load_state_dict(converted, strict=False)
    return list(res.missing_keys), list(res.unexpected_keys)
