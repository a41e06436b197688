"""CPU: the oracle restatement against the golden vectors captured from the live reference
(oracle/make_golden.py) and against the worked example in the reference docstring."""

import numpy as np
import pytest
import torch

from oracle import fusion_oracle as FO
from oracle import vision_oracle as VO

IMG = 248056


def test_position_ids_docstring_example():
    # llm_quest/qwen/qwen3_5/qwen3_5_vlm_model.py:96-101 — a 2x2 (merged) image after text token 5
    ids = [[0, 1, 2, 3, 4, 5, IMG, IMG, IMG, IMG, 7]]
    out = FO.mrope_position_ids(ids, [[1, 4, 4]])
    assert out[0, 0].tolist() == [0, 1, 2, 3, 4, 5, 6, 6, 6, 6, 8]
    assert out[1, 0].tolist() == [0, 1, 2, 3, 4, 5, 6, 6, 7, 7, 8]
    assert out[2, 0].tolist() == [0, 1, 2, 3, 4, 5, 6, 7, 6, 7, 8]


def test_position_ids_survey_goldens():
    I = IMG
    out = FO.mrope_position_ids([[1, I, I, I, I, 2, 3] + [I] * 12 + [4]], [[1, 4, 4], [2, 4, 6]])
    assert out[0, 0].tolist() == [0, 1, 1, 1, 1, 3, 4, 5, 5, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6, 8]
    assert out[1, 0].tolist() == [0, 1, 1, 2, 2, 3, 4, 5, 5, 5, 6, 6, 6, 5, 5, 5, 6, 6, 6, 8]
    assert out[2, 0].tolist() == [0, 1, 2, 1, 2, 3, 4, 5, 6, 7, 5, 6, 7, 5, 6, 7, 5, 6, 7, 8]
    out = FO.mrope_position_ids([[1, I, I, I, 2]], [[1, 4, 4]])  # too few placeholders
    assert all(out[a, 0].tolist() == [0, 1, 1, 1, 1] for a in range(3))


def test_position_ids_all_golden_cases(golden_fusion):
    for i, c in enumerate(golden_fusion["position_cases"]):
        feeds = None if c["feeds"] is None else c["feeds"].numpy()
        mask = None if c["mask"] is None else c["mask"].numpy()
        out = FO.mrope_position_ids(c["ids"].numpy(), feeds, mask, IMG, 2)
        assert np.array_equal(out, c["expected"].numpy()), f"case {i}"
        assert out.dtype == np.int64


def test_scatter_goldens(golden_fusion):
    for c in golden_fusion["scatter_cases"]:
        tok = c["image_token_id"]
        rows = FO.scatter_row_map(c["ids"].numpy(), None, tok)
        assert np.array_equal(rows, c["row_map"].numpy())
        out = FO.fuse_embeddings(c["ids"].numpy(), c["table"].view(torch.uint16).numpy(),
                                 c["vision"].to(torch.bfloat16).view(torch.uint16).numpy(), image_token_id=tok)
        assert np.array_equal(out, c["expected"].view(torch.uint16).numpy())


def test_scatter_too_few_vision_rows_raises():
    import pytest

    with pytest.raises(ValueError):
        FO.fuse_embeddings(np.array([[1, 1, 0]]), np.zeros((4, 8), np.uint16), np.zeros((1, 8), np.uint16), image_token_id=1)


def test_merge_index_golden(golden_rope):
    gi = VO.merge_gather_index(2, 4, 6, 2)
    assert torch.equal(gi, golden_rope["merge_index_2_4_6"])
    assert gi[:4].tolist() == [[0, 1, 6, 7], [2, 3, 8, 9], [4, 5, 10, 11], [12, 13, 18, 19]]
    assert gi[-1].tolist() == [40, 41, 46, 47]
    # closed form used by the CUDA kernel
    t, nh, nw = 2, 4, 6
    for tok in range(t * nh * nw):
        f, sp = divmod(tok, nh * nw)
        r, c = divmod(sp, nw)
        k = (f * (nh // 2) + r // 2) * (nw // 2) + c // 2
        s = (r % 2) * 2 + (c % 2)
        assert gi[k, s] == tok


def test_rope_goldens(golden_rope):
    g = golden_rope["rope2d"]
    assert torch.equal(VO.rotate_half_apply(g["x"], g["cos"], g["sin"]), g["expected"])
    cos, sin = VO.axial_rope_tables(10_000, 64, 3, 5)
    assert torch.equal(cos, g["cos"]) and torch.equal(sin, g["sin"])
    # layout facts from SURVEY.md §8a Q3: position (1,0) rotates slots 0-15 & 32-47, (0,1) 16-31 & 48-63
    nz = lambda row: set(torch.nonzero(sin[row]).flatten().tolist())
    assert nz(5) == set(range(0, 16)) | set(range(32, 48))
    assert nz(1) == set(range(16, 32)) | set(range(48, 64))


def test_mrope_goldens(golden_rope):
    g = golden_rope["mrope"]
    t = g["table"]
    cos, sin = VO.text_rope_tables(t["ctx"], t["base"], t["head_dim"], t["factor"])
    assert cos.shape == (8192, 64)
    out = VO.mrope_apply(g["x"], cos, sin, g["position_ids"], g["sections"])
    assert torch.equal(out, g["expected"])
    assert "".join("THW"[a] for a in VO.mrope_slot_axes(32, [11, 11, 10])) == "THW" * 10 + "TH"
    normed = VO.zero_centered_rmsnorm(g["x"], g["norm_scale"])
    assert torch.equal(VO.mrope_apply(normed, cos, sin, g["position_ids"], g["sections"]), g["expected_norm_mrope"])


def test_qwen_tower_golden(golden_qwen):
    sd = {k: v.float() for k, v in golden_qwen["state_dict"].items()}
    px = golden_qwen["pixels"].float()
    out = VO.qwen_vision_forward(sd, golden_qwen["cfg"], px)
    hid = VO.qwen_vision_forward(sd, golden_qwen["cfg"], px, return_hidden=True)
    assert out.shape == golden_qwen["out"].shape == (3, 2 * 6 * 4 // 4, 128)
    assert VO.max_norm_err(out, golden_qwen["out"]) < 2e-5
    assert VO.max_norm_err(hid, golden_qwen["hidden"]) < 2e-5


def test_vit_golden(golden_vit):
    sd = {k: v.float() for k, v in golden_vit["state_dict"].items()}
    img = golden_vit["images"].float()
    assert VO.max_norm_err(VO.vit_forward(sd, golden_vit["cfg"], img), golden_vit["logits"]) < 2e-5
    assert VO.max_norm_err(VO.vit_forward(sd, golden_vit["cfg"], img, True), golden_vit["hidden"]) < 2e-5
    asd = {k: v.float() for k, v in golden_vit["adapter_state_dict"].items()}
    assert VO.max_norm_err(VO.vit_adapter_forward(asd, golden_vit["hidden"]), golden_vit["adapter_out"]) < 2e-5


def test_patch_embed_matmul_equals_conv3d():
    torch.manual_seed(0)
    x = torch.randn(2, 3, 4, 32, 48)
    w = torch.randn(16, 3, 2, 16, 16) * 0.05
    b = torch.randn(16)
    ref = torch.nn.functional.conv3d(x, w, b, stride=(2, 16, 16)).flatten(2).transpose(1, 2)
    assert VO.max_norm_err(VO.patch_embed3d(x, w, b), ref) < 1e-5


def test_feeds_3d_shape():
    assert FO.feeds_3d_shape((2, 3, 8, 448, 448), 28, 28, 2).tolist() == [[4, 28, 28]]
    assert FO.feeds_3d_shape((2, 1568, 1536), 28, 28, 2).tolist() == [[2, 28, 28]]


def test_preprocess_oracle_equals_torchvision():
    """the oracle's uint8 -> pixel-tensor step is torchvision's to_tensor + normalize + the
    reference's temporal duplication (qwen3_5_generate_multimodal.py:40-46), bit for bit."""
    tv = pytest.importorskip("torchvision.transforms.functional")
    Image = pytest.importorskip("PIL.Image")
    import numpy as np

    rng = np.random.default_rng(5)
    arr = rng.integers(0, 256, size=(48, 64, 3), dtype=np.uint8)
    mean, std = [0.5, 0.5, 0.5], [0.5, 0.5, 0.5]          # config.py QWEN3_5 image_mean / image_std
    t = tv.normalize(tv.to_tensor(Image.fromarray(arr)), mean=mean, std=std)
    ref = t.unsqueeze(0).repeat(2, 1, 1, 1).unsqueeze(0).permute(0, 2, 1, 3, 4)
    got = VO.preprocess_u8(torch.from_numpy(arr)[None], mean, std, 2)
    assert got.shape == (1, 3, 2, 48, 64) and torch.equal(got, ref)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]   # dataset.py:342
    t = tv.normalize(tv.to_tensor(Image.fromarray(arr)), mean=mean, std=std)
    assert torch.equal(VO.preprocess_u8(torch.from_numpy(arr)[None], mean, std, 1)[0, :, 0], t)


def test_text_attention_golden(golden_text_attention):
    """the oracle's MRoPEGatedAttention prefill against the output of the live reference module."""
    g = golden_text_attention
    sd = {k: v.float() for k, v in g["state_dict"].items()}
    cfg, r = g["cfg"], g["rope"]
    cos, sin = VO.text_rope_tables(r["ctx"], r["base"], cfg["head_dim"], r["factor"])
    got = VO.mrope_gated_attention_forward(sd, cfg, g["x"].float(), cos, sin, g["position_ids"])
    assert VO.max_norm_err(got, g["expected"]) <= 2e-5
    got = VO.mrope_gated_attention_forward(sd, cfg, g["x"].float(), cos, sin, None)
    assert VO.max_norm_err(got, g["expected_1d"]) <= 2e-5


def test_part2_tiny_goldens(golden_part2):
    """Round-2 fixtures from the live reference: get_embeddings + concat (vlm_engine.py:5-20,114), a ViT at the
    TINY_VIT_CONFIG dims (4x4 patches, head_dim 32), GELU.forward and ZeroCenteredRMSNorm.forward."""
    g = golden_part2
    p2 = g["part2"]
    text = VO.get_embeddings(p2["ids"], p2["tok"], p2["pos"])
    assert torch.equal(text, p2["text"]) and torch.equal(torch.cat([p2["vision"], text], 1), p2["fused"])
    tv = g["tiny_vit"]
    sd = {k: v.float() for k, v in tv["state_dict"].items()}
    img = tv["images"].float()
    assert tv["cfg"]["patch_size"] == 4 and tv["cfg"]["emb_dim"] // tv["cfg"]["n_heads"] == 32
    assert VO.max_norm_err(VO.vit_forward(sd, tv["cfg"], img, output_hidden_states=True), tv["hidden"]) <= 2e-5
    assert VO.max_norm_err(VO.vit_forward(sd, tv["cfg"], img), tv["logits"]) <= 2e-5
    assert torch.equal(VO.gelu_erf(g["gelu"]["x"]), g["gelu"]["y"])
    r = g["rmsnorm"]
    assert torch.equal(VO.zero_centered_rmsnorm(r["x"], r["scale"]), r["y"])
    assert torch.equal(VO.zero_centered_rmsnorm(r["x"].to(torch.bfloat16), r["scale"]), r["y_bf16"])
