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
    assert lib.vf_version() == 200
    assert isinstance(lib.vf_last_error(), bytes)


def test_ctypes_struct_layout_matches_the_header(tmp_path):
    """vf_epilogue crosses the C-ABI by pointer: the ctypes mirror in _lib.py must have the header's size and field offsets (a C
    program that includes include/vfuse.h prints them; plain gcc, as a reference-side binding would compile it)."""
    import subprocess

    from llm_quest_b200 import _lib

    names = [f[0] for f in _lib.vf_epilogue._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "vfuse.h"\nint main(void) {\n  printf("%zu\\n", sizeof(vf_epilogue));\n'
                   + "".join(f'  printf("{n} %zu\\n", offsetof(vf_epilogue, {n}));\n' for n in names) + "  return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    lines = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    assert int(lines[0]) == ctypes.sizeof(_lib.vf_epilogue)
    for line in lines[1:]:
        if line:
            n, off = line.split()
            assert getattr(_lib.vf_epilogue, n).offset == int(off), n


def test_ctypes_signatures_match_the_header_prototypes(lib):
    """Every prototype in include/vfuse.h against the argtypes _lib.py binds: same number of parameters, and pointer / int32 / int64 /
    float in the same positions (a swapped pair of ints would otherwise only show up as a wrong result on the GPU)."""
    from llm_quest_b200 import _lib

    L = _lib.lib()
    header = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "vfuse.h").read_text(), flags=re.S)
    header = re.sub(r"//[^\n]*", "", header)
    protos = re.findall(r"\bint\s+(vf_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S)
    assert len(protos) >= 20

    def kind_c(arg):
        if "*" in arg:
            return "ptr"
        for t, k in (("int64_t", "i64"), ("int32_t", "i32"), ("float", "f32"), ("int", "i32")):
            if re.search(rf"\b{t}\b", arg):
                return k
        raise AssertionError(arg)

    def kind_py(t):
        if t is ctypes.c_void_p or t is ctypes.c_char_p or hasattr(t, "contents") or getattr(t, "_type_", None) is not None and not isinstance(getattr(t, "_type_"), str):
            return "ptr"
        return {ctypes.c_int32: "i32", ctypes.c_int64: "i64", ctypes.c_float: "f32", ctypes.c_int: "i32"}[t]

    checked = 0
    for name, args in protos:
        fn = getattr(L, name)
        if fn.argtypes is None:
            continue
        c_kinds = [kind_c(a) for a in args.split(",") if a.strip() and a.strip() != "void"]
        py_kinds = [kind_py(t) for t in fn.argtypes]
        assert c_kinds == py_kinds, (name, c_kinds, py_kinds)
        checked += 1
    assert checked >= 20


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


def test_hf_vision_weight_remap_round_trip():
    """HF names -> ours (reference qwen3_5_weight_loading.py:60-81); every tower parameter is
    reachable from an HF-named dict, shapes are checked, non-vision keys are ignored."""
    import torch

    from llm_quest_b200.qwen.qwen3_5 import qwen3_5_weight_loading as WL
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model import Qwen3_5VisionModel

    cfg = {"vision_emb_dim": 128, "vision_n_layers": 2, "vision_num_heads": 2, "vision_hidden_dim": 256,
           "vision_rope_base": 10_000, "llm_d_in": 64, "img_width": 64, "img_height": 64, "patch_size": 16,
           "in_channels": 3, "temporal_patch_size": 2, "spatial_merge_size": 2, "num_position_embeddings": 16}
    torch.manual_seed(0)
    src = Qwen3_5VisionModel(cfg)
    inv = [(b, a) for a, b in WL.get_vision_remapping_rules()]

    def to_hf(k):
        if k.startswith("blocks."):
            k = "model.visual." + k
            for ours, hf in ((".att.qkv.", ".attn.qkv."), (".att.proj.", ".attn.proj."), (".ffn.lin1.", ".mlp.linear_fc1."),
                             (".ffn.lin2.", ".mlp.linear_fc2.")):
                k = k.replace(ours, hf)
            return k
        for ours, hf in inv:
            if k.startswith(ours):
                return hf + k[len(ours):]
        raise AssertionError(k)

    hf = {to_hf(k): v.clone() for k, v in src.state_dict().items()}
    hf["model.language_model.layers.0.mlp.up_proj.weight"] = torch.zeros(3)
    hf["mtp.fc.weight"] = torch.zeros(3)
    assert all(k.startswith(("model.visual.", "model.language_model.", "mtp.")) for k in hf)
    torch.manual_seed(1)
    dst = Qwen3_5VisionModel(cfg)
    missing, unexpected = WL.load_qwen3_5_vision_weights(dst, hf)
    assert missing == [] and unexpected == []
    for k, v in src.state_dict().items():
        assert torch.equal(dst.state_dict()[k], v), k
    bad = dict(hf)
    bad["model.visual.merger.linear_fc2.weight"] = torch.zeros(2, 2)
    with pytest.raises(ValueError, match="shape"):
        WL.convert_vision_weights(bad, dst.state_dict())


def test_fold_layernorm_host_side_identity():
    """_fold_ln (host side of the folded LayerNorm): rstd * (x W'^T - mean * colsum) + b' == LN(x) W^T + b in fp32 up to
    the bf16 rounding of W' — and colsum is the sum of the ROUNDED weight, the one the tensor cores multiply by."""
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model import Qwen3_5VisionModel, _Packed, _fold_ln

    torch.manual_seed(3)
    d, n = 96, 40
    lin, norm = torch.nn.Linear(d, n), torch.nn.LayerNorm(d, eps=1e-6)
    with torch.no_grad():
        norm.weight.add_(0.3 * torch.randn(d))
        norm.bias.add_(0.2 * torch.randn(d))
    cache = _Packed()
    wf, bf_, cs = _fold_ln(cache, "k", lin, norm)
    assert wf.dtype == torch.bfloat16 and bf_.dtype == torch.float32 and cs.dtype == torch.float32
    assert torch.equal(cs, wf.float().sum(1))
    assert _fold_ln(cache, "k", lin, norm)[0] is wf, "folded weights are cached until a parameter changes"
    x = torch.randn(17, d) * 2 + 0.7
    mean, rstd = x.mean(1, keepdim=True), (x.var(1, unbiased=False, keepdim=True) + 1e-6).rsqrt()
    with torch.no_grad():
        folded = rstd * (x @ wf.float().t() - mean * cs[None]) + bf_[None]
        ref = lin(norm(x))
    assert (folded - ref).abs().max() / ref.abs().max() < 5e-3      # bf16 rounding of gamma . W only
    with torch.no_grad():
        norm.weight.mul_(1.5)                                       # in-place edit bumps _version: cache must rebuild
    assert _fold_ln(cache, "k", lin, norm)[0] is not wf
    assert Qwen3_5VisionModel.ln_fold == 2          # both LayerNorms of a block folded by default
    # the identity holds for any row shift s: x' = x - s, mean' = mean(x')
    xs = x - 5.0
    with torch.no_grad():
        shifted = rstd * (xs @ wf.float().t() - xs.mean(1, keepdim=True) * cs[None]) + bf_[None]
    assert (shifted - ref).abs().max() / ref.abs().max() < 5e-3
