"""CPU: the C-ABI library builds, loads and exports every symbol include/vfuse.h declares; the
drop-in modules keep the reference's state_dict keys; nothing falls back to CPU compute."""

import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from llm_quest_b200 import build

    path = build.build()
    assert path.exists()
    return ctypes.CDLL(str(path))


def test_header_symbols_exported(lib):
    header = (ROOT / "include" / "vfuse.h").read_text()
    declared = set(re.findall(r"\b(vf_[a-z0-9_]+)\s*\(", header))
    declared -= {"vf_epilogue"}
    assert len(declared) >= 16
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in vfuse.h but not exported by libvfuse.so"
    from llm_quest_b200 import _lib

    assert set(_lib.EXPORTS) == declared


def test_version_and_error_string(lib):
    lib.vf_version.restype = ctypes.c_int
    lib.vf_last_error.restype = ctypes.c_char_p
    assert lib.vf_version() == 100
    assert isinstance(lib.vf_last_error(), bytes)


def test_argument_validation_without_gpu(lib):
    """Pure argument checks run before any CUDA call, so they are testable on a CPU box."""
    lib.vf_last_error.restype = ctypes.c_char_p
    rc = lib.vf_attention_fwd(None, None, 1, 1, 1, ctypes.c_float(1.0), None)
    assert rc == -1 and b"null" in lib.vf_last_error()
    rc = lib.vf_layernorm(None, 0, 0, None, None, None, 0, 0, 0, ctypes.c_float(0), 0, 0, 0, 0, None)
    assert rc == -1


def test_only_sm100a_code_in_library(lib):
    import subprocess

    out = subprocess.run(["cuobjdump", "--list-elf", str(ROOT / "llm_quest_b200" / "libvfuse.so")],
                         capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_cpu_tensors_are_rejected_not_emulated():
    from llm_quest_b200 import _lib
    from llm_quest_b200.common.rope import VisionRoPE

    cos, sin = VisionRoPE.compute_angles_2d(10_000, 64, 2, 2)
    with pytest.raises(_lib.VFuseError):
        VisionRoPE.apply(torch.randn(1, 1, 4, 64), cos, sin)


def test_state_dict_keys_match_reference(golden_qwen, golden_vit):
    from llm_quest_b200.multimodal.vision_transformer.vit_engine import ViTAdapter
    from llm_quest_b200.multimodal.vision_transformer.vit_model import ViTModel
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model import Qwen3_5VisionModel

    m = Qwen3_5VisionModel(golden_qwen["cfg"])
    ref = golden_qwen["state_dict"]
    assert list(m.state_dict().keys()) == list(ref.keys())
    assert all(m.state_dict()[k].shape == ref[k].shape for k in ref)
    assert "cos" not in m.state_dict() and m.cos.shape == golden_qwen["cos"].shape
    assert torch.equal(m.cos, golden_qwen["cos"]) and torch.equal(m.sin, golden_qwen["sin"])

    v = ViTModel(golden_vit["cfg"])
    ref = golden_vit["state_dict"]
    assert list(v.state_dict().keys()) == list(ref.keys())
    assert all(v.state_dict()[k].shape == ref[k].shape for k in ref)

    a = ViTAdapter(128, 256, adapter_type="ffn", hidden_size_factor=2, bias=True)
    assert list(a.state_dict().keys()) == list(golden_vit["adapter_state_dict"].keys())
    with pytest.raises(ValueError):
        ViTAdapter(8, 8, adapter_type="nope")


def test_same_seed_same_init_as_reference(golden_qwen):
    """Parameters are created in the reference's order, so seed 123 reproduces its weights."""
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model import Qwen3_5VisionModel

    torch.manual_seed(123)
    m = Qwen3_5VisionModel(golden_qwen["cfg"])
    for k, v in m.state_dict().items():
        assert torch.equal(v.to(torch.bfloat16), golden_qwen["state_dict"][k]), k


def test_reference_assertions_kept(golden_qwen):
    from llm_quest_b200.multimodal.vision_transformer.vit_attention import ViTMultiHeadAttention
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model import PatchEmbedding3D, Qwen3_5VisionModel

    cfg = dict(golden_qwen["cfg"])
    with pytest.raises(AssertionError, match="not divisible by patch size"):
        Qwen3_5VisionModel({**cfg, "img_width": 70})
    with pytest.raises(AssertionError, match="too large for the number of position embeddings"):
        Qwen3_5VisionModel({**cfg, "num_position_embeddings": 4})
    with pytest.raises(ValueError, match="divisible by num_heads"):
        ViTMultiHeadAttention(64, 65, 0.0, 4)
    pe = PatchEmbedding3D(32, 32, 3, 128, 16, 2)
    with pytest.raises(AssertionError, match="does not match expected shape"):
        pe._check(torch.zeros(1, 3, 2, 48, 32))
    with pytest.raises(AssertionError, match="not divisible by temporal_patch_size"):
        pe._check(torch.zeros(1, 3, 3, 32, 32))


def test_product_never_imports_oracle():
    for p in (ROOT / "llm_quest_b200").rglob("*.py"):
        txt = p.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt, p


def test_shim_aliases_reference_module_paths():
    import sys

    import llm_quest_b200.shim as shim

    saved = {k: sys.modules.get(k) for k in shim.ALIASES}
    try:
        names = shim.install()
        import importlib

        mod = importlib.import_module("llm_quest.qwen.qwen3_5.qwen3_5_vision_model")
        assert mod.__name__ == "llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model" and hasattr(mod, "Qwen3_5VisionModel")
        assert len(names) == len(shim.ALIASES)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
