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


_saved: dict[str, object] = {}


def install(include_adapter: bool = False, only: list[str] | None = None) -> list[str]:
    """Alias the modules; returns the aliased names. `include_adapter` also replaces
    `vit_engine` (only its ViTAdapter is implemented — the training loops are out of scope).
    `only`: restrict to these reference module names, e.g. just the vision tower, so that the reference's OWN
    `Qwen3_5VLM.forward` (embedding lookup, masked_scatter, position ids, text model) runs around the B200 tower:

        shim.install(only=["llm_quest.qwen.qwen3_5.qwen3_5_vision_model"])
        import llm_quest.qwen.qwen3_5.qwen3_5_vlm_model as ref_vlm      # (importlib.reload it if already imported)
        model = ref_vlm.Qwen3_5VLM(cfg)                                  # its vision_model is the libvfuse tower
    """
    names = dict(ALIASES)
    if include_adapter:
        names["llm_quest.multimodal.vision_transformer.vit_engine"] = "llm_quest_b200.multimodal.vision_transformer.vit_engine"
    if only is not None:
        unknown = [n for n in only if n not in names]
        if unknown:
            raise KeyError(f"not a hot-path module of the reference: {unknown}")
        names = {n: names[n] for n in only}
    for ref_name, ours in names.items():
        if ref_name not in _saved:
            _saved[ref_name] = sys.modules.get(ref_name)
        mod = importlib.import_module(ours)
        sys.modules[ref_name] = mod
        # `from llm_quest.qwen.qwen3_5 import qwen3_5_vision_model` resolves through the parent package's attribute
        parent, _, leaf = ref_name.rpartition(".")
        if parent in sys.modules:
            setattr(sys.modules[parent], leaf, mod)
    return sorted(names)


def uninstall() -> None:
    """Undo install(): the reference's own modules (or nothing) are back under their names."""
    for ref_name, mod in _saved.items():
        parent, _, leaf = ref_name.rpartition(".")
        if mod is None:
            sys.modules.pop(ref_name, None)
            if parent in sys.modules and hasattr(sys.modules[parent], leaf):
                delattr(sys.modules[parent], leaf)
        else:
            sys.modules[ref_name] = mod
            if parent in sys.modules:
                setattr(sys.modules[parent], leaf, mod)
    _saved.clear()
