"""Opt-in aliasing of the B200 modules over the reference's module paths.

    import llm_quest_b200.shim as shim; shim.install()

After this, `import llm_quest.qwen.qwen3_5.qwen3_5_vision_model` (and the other hot-path modules)
resolve to the libvfuse-backed classes, so the reference's own callers — `Qwen3_5VLM.__init__`
(qwen3_5_vlm_model.py:43), `qwen3_5_generate_multimodal.py`, `multimodal/vlm_engine.py` — pick them
up without any source change. Only the hot-path modules are aliased; everything else of `llm_quest`
stays the reference's.
"""

from __future__ import annotations

import importlib
import sys

ALIASES = {
    "llm_quest.qwen.qwen3_5.qwen3_5_vision_model": "llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model",
    "llm_quest.qwen.qwen3_5.qwen3_5_vlm_model": "llm_quest_b200.qwen.qwen3_5.qwen3_5_vlm_model",
    "llm_quest.multimodal.vision_transformer.vit_model": "llm_quest_b200.multimodal.vision_transformer.vit_model",
    "llm_quest.multimodal.vision_transformer.vit_attention": "llm_quest_b200.multimodal.vision_transformer.vit_attention",
    "llm_quest.multimodal.vision_transformer.vit_transformer_block": "llm_quest_b200.multimodal.vision_transformer.vit_transformer_block",
}


def install(include_adapter: bool = False) -> list[str]:
    """Alias the modules; returns the aliased names. `include_adapter` also replaces
    `vit_engine` (only its ViTAdapter is implemented — the training loops are out of scope)."""
    names = dict(ALIASES)
    if include_adapter:
        names["llm_quest.multimodal.vision_transformer.vit_engine"] = "llm_quest_b200.multimodal.vision_transformer.vit_engine"
    for ref_name, ours in names.items():
        sys.modules[ref_name] = importlib.import_module(ours)
    return sorted(names)
