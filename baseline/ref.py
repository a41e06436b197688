"""Locate and import the UNMODIFIED reference (casinca/LLM-quest) for the baseline arms of bench.py and the drop-in tests.

The reference is installed once, offline, into the git-ignored ``baseline/_ref/`` by ``baseline/install_ref.sh``
(``pip install --target``; it travels to the GPU box with the repository snapshot). Nothing of it is committed.
Only bench.py's baseline legs (``--impl reference``, ``cpu_baseline``, ``gpu_eager_baseline``) and tests/ import it;
the product (llm_quest_b200/) never does.

Search order: ``$LLMQ_REF`` (a directory holding ``llm_quest/`` and ``config.py``), then ``baseline/_ref``.
``/root/reference`` is deliberately not searched: it does not exist on the GPU box.
"""

from __future__ import annotations

import importlib
import os
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent


def reference_dir() -> Path | None:
    for cand in (os.environ.get("LLMQ_REF"), HERE / "_ref"):
        if cand and (Path(cand) / "llm_quest" / "qwen" / "qwen3_5" / "qwen3_5_vision_model.py").exists():
            return Path(cand)
    return None


def available() -> bool:
    return reference_dir() is not None


def import_reference():
    """Put the reference on sys.path (front) and return its ``config`` module; raises ImportError when absent."""
    d = reference_dir()
    if d is None:
        raise ImportError("reference not installed: run baseline/install_ref.sh where /root/reference exists")
    if str(d) not in sys.path:
        sys.path.insert(0, str(d))
    return importlib.import_module("config")


def qwen_vision_model(cfg_overrides: dict):
    """Reference ``Qwen3_5VisionModel`` (llm_quest/qwen/qwen3_5/qwen3_5_vision_model.py:241) at QWEN3_5_08B_CONFIG + overrides."""
    config = import_reference()
    from llm_quest.qwen.qwen3_5.qwen3_5_vision_model import Qwen3_5VisionModel

    return Qwen3_5VisionModel({**config.QWEN3_5_08B_CONFIG, **cfg_overrides})


def vit_model(cfg_overrides: dict | None = None):
    """Reference Part-1 ``ViTModel`` (llm_quest/multimodal/vision_transformer/vit_model.py:92) at VIT_BASE_CONFIG."""
    config = import_reference()
    from llm_quest.multimodal.vision_transformer.vit_model import ViTModel

    return ViTModel({**config.VIT_BASE_CONFIG, **(cfg_overrides or {})})


def qwen_vlm(cfg_overrides: dict):
    """Reference ``Qwen3_5VLM`` (qwen3_5_vlm_model.py:21)."""
    config = import_reference()
    from llm_quest.qwen.qwen3_5.qwen3_5_vlm_model import Qwen3_5VLM

    return Qwen3_5VLM({**config.QWEN3_5_08B_CONFIG, **cfg_overrides})
