"""GPU parity of the drop-in modules (the reference-facing nn.Module surface) against the golden
outputs of the live reference (tests/golden/, tiny configs) and against the CPU oracle at the
BASELINE.json configs (batch reduced: no op mixes samples).

Tolerance (north_star): max|delta|/max|ref| <= 1e-2 and cosine >= 0.9999 vs the fp32 reference.
"""

import numpy as np
import pytest
import torch

from oracle import fusion_oracle as FO
from oracle import vision_oracle as VO

pytestmark = pytest.mark.gpu
IMG = 248056
TOL, COS = 1e-2, 0.9999


def check_close(got, ref, what, tol=TOL, cos=COS):
    got, ref = got.float().cpu(), ref.float().cpu()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    e, c = VO.max_norm_err(got, ref), VO.cosine(got, ref)
    print(f"{what}: max_norm_err={e:.3e} cosine={c:.7f}")
    assert e <= tol and c >= cos, f"{what}: err {e:.3e} cos {c:.6f}"


def qwen_cfg(px=448, **kw):
    cfg = {
        "vision_emb_dim": 768, "vision_n_layers": 12, "vision_num_heads": 12, "vision_hidden_dim": 3072,
        "vision_rope_base": 10_000, "llm_d_in": 1024, "img_width": px, "img_height": px, "patch_size": 16,
        "in_channels": 3, "temporal_patch_size": 2, "spatial_merge_size": 2, "num_position_embeddings": 2304,
        "image_token_id": IMG, "vocab_size": 248_320, "emb_dim": 1024, "dtype": torch.bfloat16,
    }
    cfg.update(kw)
    return cfg


def test_qwen_tower_tiny_golden(golden_qwen):
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model import Qwen3_5VisionModel

    m = Qwen3_5VisionModel(golden_qwen["cfg"])
    m.load_state_dict({k: v.float() for k, v in golden_qwen["state_dict"].items()})
    m = m.cuda().eval()
    with torch.inference_mode():
        out = m(golden_qwen["pixels"].cuda().float())
        hid, B, S = m.encode_hidden(golden_qwen["pixels"].cuda())
    assert out.dtype == torch.float32 and out.shape == golden_qwen["out"].shape
    check_close(hid.view(B, S, -1), golden_qwen["hidden"], "tiny tower hidden vs reference")
    check_close(out, golden_qwen["out"], "tiny tower out vs reference")


def test_qwen_submodules_tiny(golden_qwen):
    """Each reference-facing sub-module forward works on its own (drop-in at any granularity)."""
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model import Qwen3_5VisionModel

    cfg = golden_qwen["cfg"]
    sd = {k: v.float() for k, v in golden_qwen["state_dict"].items()}
    m = Qwen3_5VisionModel(cfg)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    px = golden_qwen["pixels"].float()
    with torch.inference_mode():
        pe = m.patch_embed(px.cuda())
        check_close(pe, VO.patch_embed3d(px, sd["patch_embed.conv_proj.weight"], sd["patch_embed.conv_proj.bias"]),
                    "PatchEmbedding3D", tol=3e-3)
        x = golden_qwen["hidden"]
        frames = x.shape[1] // m.n_spatial_patches
        cos, sin = m.cos.repeat(frames, 1), m.sin.repeat(frames, 1)
        blk_sd = {k[len("blocks.1."):]: v for k, v in sd.items() if k.startswith("blocks.1.")}
        # block forward on an arbitrary hidden state
        import torch.nn.functional as F

        D = cfg["vision_emb_dim"]
        h = F.layer_norm(x, (D,), blk_sd["norm1.weight"], blk_sd["norm1.bias"], 1e-6)
        qkv = h @ blk_sd["att.qkv.weight"].t() + blk_sd["att.qkv.bias"]
        B, S, _ = x.shape
        q, k, v = (t.view(B, S, 2, 64).transpose(1, 2) for t in qkv.chunk(3, -1))
        q, k = VO.rotate_half_apply(q, cos.cpu(), sin.cpu()), VO.rotate_half_apply(k, cos.cpu(), sin.cpu())
        att = (torch.softmax(q @ k.transpose(-1, -2) / 8, -1) @ v).transpose(1, 2).reshape(B, S, D)
        att = att @ blk_sd["att.proj.weight"].t() + blk_sd["att.proj.bias"]
        check_close(m.blocks[1].att(h.cuda(), cos, sin), att, "Qwen3_5VisionAttention", tol=5e-3)
        x1 = x + att
        h2 = F.layer_norm(x1, (D,), blk_sd["norm2.weight"], blk_sd["norm2.bias"], 1e-6)
        ffn = VO.gelu_tanh(h2 @ blk_sd["ffn.lin1.weight"].t() + blk_sd["ffn.lin1.bias"]) @ blk_sd["ffn.lin2.weight"].t() + blk_sd["ffn.lin2.bias"]
        check_close(m.blocks[1].ffn(h2.cuda()), ffn, "Qwen3_5VisionFFN", tol=5e-3)
        check_close(m.blocks[1](x.cuda(), cos, sin), x1 + ffn, "Qwen3_5VisionTransformerBlock", tol=5e-3)
        check_close(m.merge_adapter(x.cuda()), VO.merge_adapter_forward(sd, "merge_adapter.", x, m.n_height_patches, m.n_width_patches, 2),
                    "ViTMergeAdapter", tol=5e-3)


def test_vit_tiny_golden(golden_vit):
    from llm_quest_b200.multimodal.vision_transformer.vit_engine import ViTAdapter
    from llm_quest_b200.multimodal.vision_transformer.vit_model import ViTModel

    m = ViTModel(golden_vit["cfg"])
    m.load_state_dict({k: v.float() for k, v in golden_vit["state_dict"].items()})
    m = m.cuda().eval()
    img = golden_vit["images"].cuda().float()
    with torch.inference_mode():
        logits = m(img)
        hidden = m(img, output_hidden_states=True)
    check_close(hidden, golden_vit["hidden"], "tiny ViT hidden vs reference")
    check_close(logits, golden_vit["logits"], "tiny ViT logits vs reference")
    a = ViTAdapter(128, 256, adapter_type="ffn", hidden_size_factor=2, bias=True)
    a.load_state_dict({k: v.float() for k, v in golden_vit["adapter_state_dict"].items()})
    a = a.cuda().eval()
    with torch.inference_mode():
        out = a(golden_vit["hidden"].cuda())
        check_close(out, golden_vit["adapter_out"], "ViTAdapter vs reference", tol=5e-3)
        # Part-2 fusion: adapter rows written straight into the [vision | text] buffer (vlm_engine.py:114)
        text = torch.randn(3, 11, 256, generator=torch.Generator().manual_seed(5))
        fused = torch.zeros((3, 17 + 11, 256), device="cuda")
        fused[:, 17:] = text.cuda()
        a.forward_into(golden_vit["hidden"].cuda(), fused)
        check_close(fused, torch.cat([golden_vit["adapter_out"], text], dim=1), "Part-2 fused concat", tol=5e-3)
    with pytest.raises(RuntimeError, match="forward-only"):
        m.train()(img)


def _random_qwen_sd(model, seed=123):
    """Random-init weights at the module's own init scale, rounded to bf16-representable values."""
    torch.manual_seed(seed)
    sd = {}
    for k, v in model.state_dict().items():
        sd[k] = v.detach().to(torch.bfloat16).float()
    return sd


@pytest.mark.parametrize("px,T,B,ln_fold", [(448, 2, 2, 2), (224, 4, 1, 2), (448, 2, 2, 0), (448, 2, 2, 1)])
def test_qwen_tower_full_size_vs_oracle(px, T, B, ln_fold):
    """cfg-2 / cfg-4 shapes with the batch reduced to what the CPU oracle finishes in seconds. ln_fold: every
    LayerNorm placement (norm1 and norm2 folded into the GEMMs (default) / stand-alone kernels / norm1 folded only)."""
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model import Qwen3_5VisionModel

    cfg = qwen_cfg(px)
    torch.manual_seed(123)
    m = Qwen3_5VisionModel(cfg).eval()
    m.ln_fold = ln_fold
    sd = _random_qwen_sd(m)
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(1234)
    pixels = torch.randn(B, 3, T, px, px, generator=g).to(torch.bfloat16).float()
    torch.set_num_threads(max(1, torch.get_num_threads()))
    with torch.inference_mode():
        ref = VO.qwen_vision_forward(sd, cfg, pixels)
        out = m.cuda()(pixels.cuda())
    assert out.shape == (B, (T // 2) * (px // 32) ** 2, 1024)
    check_close(out, ref, f"Qwen3-ViT tower {px}px T={T} ln_fold={ln_fold} vs fp32 oracle")


def test_qwen_tower_unrounded_fp32_weights_and_pixels_vs_oracle():
    """The north-star tolerance is defined on the SAME random-init fp32 weights and fp32 pixels the reference sees:
    nothing is pre-rounded to bf16 here — the oracle computes in fp32 on the un-rounded tensors, the B200 path rounds
    its GEMM operands itself. cfg-2 shape, 12 layers, batch 2."""
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model import Qwen3_5VisionModel

    cfg = qwen_cfg(448)
    torch.manual_seed(123)
    m = Qwen3_5VisionModel(cfg).eval()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}          # fp32 as initialised
    pixels = torch.randn(2, 3, 2, 448, 448, generator=torch.Generator().manual_seed(1234))   # fp32 randn
    assert not torch.equal(pixels, pixels.to(torch.bfloat16).float())
    with torch.inference_mode():
        ref = VO.qwen_vision_forward(sd, cfg, pixels)
        out = m.cuda()(pixels.cuda())
    check_close(out, ref, "Qwen3-ViT tower 448px, un-rounded fp32 weights + pixels vs fp32 oracle")


def test_vit_b16_unrounded_fp32_weights_vs_oracle():
    from llm_quest_b200.multimodal.vision_transformer.vit_model import ViTModel

    cfg = {"img_width": 224, "img_height": 224, "patch_size": 16, "num_channels": 3, "emb_dim": 768, "n_layers": 12,
           "n_heads": 12, "drop_rate": 0.1, "qkv_bias": True, "num_classes": 100}
    torch.manual_seed(123)
    m = ViTModel(cfg).eval()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    img = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(1234))
    with torch.inference_mode():
        ref_h = VO.vit_forward(sd, cfg, img, output_hidden_states=True)
        ref_l = VO.vit_forward(sd, cfg, img)
        mc = m.cuda()
        check_close(mc(img.cuda(), output_hidden_states=True), ref_h, "ViT-B/16 hidden, un-rounded fp32 weights vs fp32 oracle")
        check_close(mc(img.cuda()), ref_l, "ViT-B/16 logits, un-rounded fp32 weights vs fp32 oracle")


@pytest.mark.parametrize("px,T,what", [
    (448, 8, "12 layers at S=3136 (cfg-3 forward()-native clip)"),
    (448, 16, "12 layers at S=6272 (cfg-4 video)"),
])
def test_qwen_tower_long_sequences_full_depth_vs_oracle(px, T, what):
    """Error accumulates over depth and the attention kernel keeps a stale row max (lazy rescale): the FULL 12-layer
    tower at the long-sequence shapes, batch 1, un-rounded fp32 weights, against the fp32 oracle (tens of seconds of CPU)."""
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model import Qwen3_5VisionModel

    cfg = qwen_cfg(px)
    torch.manual_seed(123)
    m = Qwen3_5VisionModel(cfg).eval()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    pixels = torch.randn(1, 3, T, px, px, generator=torch.Generator().manual_seed(1234))
    with torch.inference_mode():
        ref = VO.qwen_vision_forward(sd, cfg, pixels)
        out = m.cuda()(pixels.cuda())
    check_close(out, ref, what)


@pytest.mark.parametrize("offset", [10.0, 50.0])
def test_qwen_tower_rows_with_large_mean_folded_layernorm(offset):
    """Outlier rows: a position embedding that puts every token's mean `offset` standard deviations away from zero (real
    ViT checkpoints carry such rows). The folded LayerNorms must hold the tolerance without any switch: the first
    LayerNorm of the chain runs on the fp32 stream and every later producer subtracts the row's running mean before the
    bf16 rounding (vf_epilogue.ln_shift)."""
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model import Qwen3_5VisionModel

    cfg = qwen_cfg(224, vision_n_layers=4)
    torch.manual_seed(123)
    m = Qwen3_5VisionModel(cfg).eval()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(7)
    sd["pos_embed.weight"] = sd["pos_embed.weight"] + offset * (1.0 + 0.2 * torch.randn(sd["pos_embed.weight"].shape[0], 1, generator=g))
    m.load_state_dict(sd)
    pixels = torch.randn(2, 3, 2, 224, 224, generator=torch.Generator().manual_seed(1234))
    with torch.inference_mode():
        ref = VO.qwen_vision_forward(sd, cfg, pixels)
        mc = m.cuda()
        assert mc.ln_fold == 2
        out = mc(pixels.cuda())
        mc.ln_fold = 0
        out_plain = mc(pixels.cuda())
    check_close(out, ref, f"tower with row means ~{offset} sigma, folded LayerNorms")
    check_close(out_plain, ref, f"tower with row means ~{offset} sigma, stand-alone LayerNorms")


@pytest.mark.parametrize("fold,batch", [(True, 2), (False, 2), (True, 24)])
def test_vit_b16_vs_oracle(fold, batch):
    """cfg-1: Part-1 ViT-B/16, 224^2. fold: ln_1 / ln_2 folded into the GEMMs around them (default) or stand-alone kernels;
    batch 24 = 4728 rows takes the vf_ln_row_stats route, batch 2 the statistics-in-the-consumer route (same bits per row)."""
    from llm_quest_b200.multimodal.vision_transformer.vit_model import ViTModel

    cfg = {"img_width": 224, "img_height": 224, "patch_size": 16, "num_channels": 3, "emb_dim": 768, "n_layers": 12,
           "n_heads": 12, "drop_rate": 0.1, "qkv_bias": True, "num_classes": 100}
    torch.manual_seed(123)
    m = ViTModel(cfg).eval()
    m.ln_fold = fold
    sd = _random_qwen_sd(m)
    m.load_state_dict(sd)
    img = torch.randn(batch, 3, 224, 224, generator=torch.Generator().manual_seed(1234)).to(torch.bfloat16).float()
    with torch.inference_mode():
        ref_h = VO.vit_forward(sd, cfg, img[:2], output_hidden_states=True)
        ref_l = VO.vit_forward(sd, cfg, img[:2])
        mc = m.cuda()
        hid, logits = mc(img.cuda(), output_hidden_states=True), mc(img.cuda())
        check_close(hid[:2], ref_h, f"ViT-B/16 hidden vs fp32 oracle (fold={fold}, batch {batch})")
        check_close(logits[:2], ref_l, f"ViT-B/16 logits vs fp32 oracle (fold={fold}, batch {batch})")
        if batch > 2 and fold:       # a sample's result does not depend on the statistics route its batch size selects
            assert torch.equal(mc(img[:2].cuda(), output_hidden_states=True), hid[:2])


def test_vlm_encode_and_fuse_vs_oracle():
    """cfg-3 semantics at reduced batch: fused embeddings + MRoPE ids through Qwen3_5VLM.encode_and_fuse."""
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_vlm_model import Qwen3_5VLM

    px = 64
    cfg = qwen_cfg(px, vision_n_layers=2, vocab_size=3000, image_token_id=2999)
    torch.manual_seed(123)
    vlm = Qwen3_5VLM(cfg).eval()
    sd = _random_qwen_sd(vlm.vision_model)
    vlm.vision_model.load_state_dict(sd)
    vlm = vlm.cuda()
    tok = cfg["image_token_id"]
    b, T = 3, 4  # 2 merged frames of 2x2 merged patches = 8 vision tokens per sample
    n_vis_per = (T // 2) * (px // 32) ** 2
    g = torch.Generator().manual_seed(4321)
    ids = torch.randint(0, 1000, (b, 40), generator=g)
    for s in range(b):
        ids[s, 5 + s : 5 + s + n_vis_per] = tok
    pixels = torch.randn(b, 3, T, px, px, generator=g).to(torch.bfloat16).float()
    with torch.inference_mode():
        embs, pid, mask = vlm.encode_and_fuse(ids.cuda(), pixels.cuda())
        vis_ref = VO.qwen_vision_forward(sd, cfg, pixels)
    table = vlm.language_model.emb_dict.weight.detach().cpu()
    # placement and untouched rows: bit-exact
    assert torch.equal(mask.cpu(), ids == tok)
    got = embs.cpu().view(-1, 1024)
    flat = ids.view(-1)
    assert torch.equal(got[flat != tok].view(torch.uint16), table[flat[flat != tok]].view(torch.uint16))
    check_close(got[flat == tok], vis_ref.reshape(-1, 1024), "fused vision rows vs fp32 oracle")
    exp_pid = FO.mrope_position_ids(ids.numpy(), [[T // 2, px // 16, px // 16]], None, tok, 2)
    assert torch.equal(pid.cpu(), torch.from_numpy(exp_pid))
    # text-only path
    with torch.inference_mode():
        e2, p2, m2 = vlm.encode_and_fuse(ids.cuda())
    assert m2 is None and torch.equal(e2.cpu().view(torch.uint16), table[ids].view(torch.uint16))
    assert torch.equal(p2.cpu(), torch.arange(40).expand(3, b, 40))
    # too few vision rows -> same failure mode as masked_scatter
    ids_bad = ids.clone()
    ids_bad[:, 30:39] = tok
    with pytest.raises(RuntimeError, match="masked_scatter"):
        vlm.encode_and_fuse(ids_bad.cuda(), pixels.cuda())
    assert FO.feeds_3d_shape(tuple(pixels.shape), px // 16, px // 16, 2).tolist() == vlm.get_feeds_3d_shape(pixels).tolist()


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs 3-5 at the shapes that stress the kernels (depth reduced so that the CPU oracle
# finishes in seconds: every layer runs the same kernels) and size-independent properties at full size
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("px,T,layers,npos,what", [
    (448, 16, 1, 2304, "cfg-4 video: 16 frames -> T'=8, S=6272 (49 query tiles, 98 key tiles)"),
    (672, 2, 2, 7056, "cfg-5 sweep 672 px: S=1764, pos-embed capacity 7056"),
    (1344, 2, 1, 7056, "cfg-5 sweep 1344 px: S=7056 = the whole pos-embed table"),
    (224, 8, 2, 2304, "cfg-3 forward()-native clip: 8 frames 224 px, cross-frame attention, S=784"),
])
def test_qwen_tower_config_shapes_vs_oracle(px, T, layers, npos, what):
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model import Qwen3_5VisionModel

    cfg = qwen_cfg(px, vision_n_layers=layers, num_position_embeddings=npos)
    torch.manual_seed(123)
    m = Qwen3_5VisionModel(cfg).eval()
    sd = _random_qwen_sd(m)
    m.load_state_dict(sd)
    pixels = torch.randn(1, 3, T, px, px, generator=torch.Generator().manual_seed(1234)).to(torch.bfloat16).float()
    with torch.inference_mode():
        ref = VO.qwen_vision_forward(sd, cfg, pixels)
        out = m.cuda()(pixels.cuda())
    assert out.shape == (1, (T // 2) * (px // 32) ** 2, 1024)
    check_close(out, ref, what)


def test_full_batch_equals_single_samples_bit_exact():
    """cfg-2 at FULL size (64 x 448^2, 12 layers): no op mixes samples, so a sample's output must not depend
    on the batch it travels in — bit for bit (same tiles, same accumulation order), and a second run of the
    same batch is identical (no atomics, no race)."""
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model import Qwen3_5VisionModel

    torch.manual_seed(123)
    m = Qwen3_5VisionModel(qwen_cfg(448)).eval().cuda()
    x = torch.randn(64, 3, 2, 448, 448, generator=torch.Generator().manual_seed(1234)).to(torch.bfloat16).cuda()
    with torch.inference_mode():
        full = m(x)
        again = m(x)
        part = m(x[[5, 37, 63]].contiguous())
    assert torch.isfinite(full).all()
    assert torch.equal(full, again), "two runs of the same batch differ"
    assert torch.equal(full[[5, 37, 63]], part), "a sample's embedding depends on its batch neighbours"


def test_vlm_cfg3_full_size_placement_and_ids():
    """cfg-3 at FULL size (32 samples x (4 images + 2048 text tokens), seq 2832): scatter placement, untouched
    rows and MRoPE-I ids bit-exact against the numpy oracle; vision rows equal the tower's own output."""
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_vlm_model import Qwen3_5VLM

    cfg = qwen_cfg(448, vision_n_layers=1, vocab_size=4096)
    torch.manual_seed(123)
    vlm = Qwen3_5VLM(cfg).eval().cuda()
    g = torch.Generator().manual_seed(4321)
    b, n_img, per = 32, 4, 196
    chunks = [410, 410, 410, 410, 408]
    rows = []
    for _ in range(b):
        parts = []
        for i, c in enumerate(chunks):
            parts.append(torch.randint(0, 1000, (c,), generator=g))
            if i < n_img:
                parts.append(torch.full((per,), IMG, dtype=torch.int64))
        rows.append(torch.cat(parts))
    ids = torch.stack(rows)
    assert ids.shape == (b, 2832)
    pixels = torch.randn(b * n_img, 3, 2, 448, 448, generator=g).to(torch.bfloat16)
    feeds = torch.tensor([[1, 28, 28]] * n_img)
    with torch.inference_mode():
        embs, pid, mask = vlm.encode_and_fuse(ids.cuda(), pixels.cuda(), feeds)
        vis = vlm.vision_model(pixels.cuda())
    assert torch.equal(mask.cpu(), ids == IMG)
    exp_pid = FO.mrope_position_ids(ids.numpy(), feeds.tolist(), None, IMG, 2)
    assert torch.equal(pid.cpu(), torch.from_numpy(exp_pid)) and int(pid.max()) == 2103
    got = embs.view(-1, 1024)
    flat = ids.view(-1).cuda()
    table = vlm.language_model.emb_dict.weight.detach()
    assert torch.equal(got[flat != IMG].view(torch.uint16), table[flat[flat != IMG]].view(torch.uint16))
    # vision rows land in flat (b, seq) order; the scatter epilogue rounds fp32 -> bf16 once
    assert torch.equal(got[flat == IMG], vis.reshape(-1, 1024).to(torch.bfloat16))


def test_vision_feature_cache_decode_loop():
    """in a decode loop the image is encoded once; later steps only gather/scatter. The cached
    path must give the same embeddings bit for bit and launch a handful of kernels instead of the tower."""
    from llm_quest_b200 import _lib
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_vlm_model import Qwen3_5VLM

    px = 64
    cfg = qwen_cfg(px, vision_n_layers=2, vocab_size=3000, image_token_id=2999)
    torch.manual_seed(123)
    vlm = Qwen3_5VLM(cfg).eval().cuda()
    tok, n_vis = 2999, (px // 32) ** 2
    g = torch.Generator().manual_seed(7)
    ids = torch.randint(0, 1000, (2, 20), generator=g)
    ids[:, 3:3 + n_vis] = tok
    pixels = torch.randn(2, 3, 2, px, px, generator=g).to(torch.bfloat16).cuda()
    with torch.inference_mode():
        ref_e, ref_p, _ = vlm.encode_and_fuse(ids.cuda(), pixels)
        vlm.enable_vision_cache()
        _lib.reset_launch_count()
        e1, p1, _ = vlm.encode_and_fuse(ids.cuda(), pixels)
        first = _lib.launch_count()
        ids2 = torch.cat([ids, torch.randint(0, 1000, (2, 1), generator=g)], dim=1)   # one generated token later
        _lib.reset_launch_count()
        e2, p2, _ = vlm.encode_and_fuse(ids2.cuda(), pixels)
        second = _lib.launch_count()
    assert torch.equal(e1, ref_e) and torch.equal(p1, ref_p)
    assert torch.equal(e2[:, :20], ref_e) and torch.equal(p2[:, :, :20], ref_p)
    assert second <= 6 < first, (first, second)          # scan + gather/scatter + position ids vs the whole tower
    # a changed image (new tensor version) or changed weights must miss
    with torch.no_grad():
        pixels2 = pixels.clone()
        pixels2 += 1.0
        _lib.reset_launch_count()
        e3, _, _ = vlm.encode_and_fuse(ids.cuda(), pixels2)
        assert _lib.launch_count() > 6 and not torch.equal(e3, ref_e)
        vlm.vision_model.merge_adapter.lin2.bias.add_(1.0)
        _lib.reset_launch_count()
        vlm.encode_and_fuse(ids.cuda(), pixels2)
        assert _lib.launch_count() > 6


def test_uint8_images_through_streamed_encoder():
    """uint8 HWC uploads -> vf_preprocess_u8 -> tower, through the streaming runtime, equals the
    tower run on the torchvision-style pixel tensor (bit-exact: same bf16 pixels, same kernels)."""
    from llm_quest_b200.pipeline import StreamedEncoder
    from llm_quest_b200.qwen.qwen3_5.preprocess import pixels_from_uint8
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model import Qwen3_5VisionModel

    px = 64
    cfg = qwen_cfg(px, vision_n_layers=2)
    torch.manual_seed(123)
    m = Qwen3_5VisionModel(cfg).eval().cuda()
    mean, std = [0.5, 0.5, 0.5], [0.5, 0.5, 0.5]
    u8 = torch.randint(0, 256, (4, px, px, 3), generator=torch.Generator().manual_seed(3), dtype=torch.uint8).pin_memory()
    ref_pixels = VO.preprocess_u8(u8, mean, std, 2).to(torch.bfloat16)
    with torch.inference_mode():
        ref = m(ref_pixels.cuda())
    enc = StreamedEncoder(m, depth=2, pre_fn=lambda d: pixels_from_uint8(d, mean, std, 2))
    outs = []
    for _ in range(3):
        enc.submit(u8)
        outs += [o.clone() for o in enc.ready()]
    outs += [o.clone() for o in enc.drain()]
    assert len(outs) == 3
    for o in outs:
        assert torch.equal(o, ref.cpu())


def test_mrope_gated_attention_prefill_golden(golden_text_attention):
    """the drop-in MRoPEGatedAttention (prefill) against the output of the live reference module
    (fp32, CPU) on the committed fixture: multimodal position ids and the text-only (1-D) case."""
    from llm_quest_b200.common.rope import RoPE
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_text_model import MRoPEGatedAttention

    g = golden_text_attention
    cfg = dict(g["cfg"], dtype=torch.float32)
    att = MRoPEGatedAttention(cfg, layer_idx=0).eval()
    assert set(att.state_dict().keys()) == set(g["state_dict"].keys())
    att.load_state_dict({k: v.float() for k, v in g["state_dict"].items()})
    att = att.cuda()
    r = g["rope"]
    cos, sin = RoPE.compute_angles(r["base"], cfg["head_dim"], r["ctx"], rotation_factor=r["factor"])
    x = g["x"].float().cuda()
    with torch.inference_mode():
        out = att(x, None, cos, sin, position_ids=g["position_ids"].cuda())
        out1 = att(x, None, cos, sin)
    assert out.dtype == torch.float32 and out.shape == g["expected"].shape
    check_close(out, g["expected"], "MRoPEGatedAttention prefill (multimodal ids) vs reference")
    check_close(out1, g["expected_1d"], "MRoPEGatedAttention prefill (text-only ids) vs reference")
    with pytest.raises(Exception, match="prefill"):
        att(x, None, cos, sin, cache=object())


def test_graphed_encoder_matches_eager():
    """CUDA-graph replay of the tower (small-batch latency path): identical bits, new inputs honoured."""
    from llm_quest_b200.pipeline import GraphedEncoder
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model import Qwen3_5VisionModel

    torch.manual_seed(123)
    m = Qwen3_5VisionModel(qwen_cfg(224, vision_n_layers=3)).eval().cuda()
    g = torch.Generator().manual_seed(9)
    xs = [torch.randn(2, 3, 2, 224, 224, generator=g).to(torch.bfloat16).cuda() for _ in range(3)]
    with torch.inference_mode():
        refs = [m(x).clone() for x in xs]
    ge = GraphedEncoder(m, xs[0])
    for x, r in zip(xs + xs[:1], refs + refs[:1]):
        out = ge(x)
        torch.cuda.synchronize()
        assert torch.equal(out, r)
    with pytest.raises(ValueError):
        ge(xs[0][:1])


def test_vision_cache_keys_on_tensor_identity_not_storage_address():
    """a (data_ptr, shape, version) key lets the caching allocator hand the freed address of image A to image B
    and the cache serve A's embeddings for B. The cache keeps image A's tensor alive and compares objects; an explicit
    image_id overrides identity."""
    from llm_quest_b200 import _lib
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_vlm_model import Qwen3_5VLM

    px = 64
    cfg = qwen_cfg(px, vision_n_layers=1, vocab_size=3000, image_token_id=2999)
    torch.manual_seed(123)
    vlm = Qwen3_5VLM(cfg).eval().cuda().enable_vision_cache()
    ids = torch.randint(0, 1000, (1, 12))
    ids[:, 3:7] = 2999
    ids = ids.cuda()
    g = torch.Generator().manual_seed(11)
    with torch.inference_mode():
        a = torch.randn(1, 3, 2, px, px, generator=g).cuda()
        ea, _, _ = vlm.encode_and_fuse(ids, a)
        addr = a.data_ptr()
        del a                                            # without the cache's reference this block would be recycled
        b = torch.randn(1, 3, 2, px, px, generator=g).cuda()
        assert b.data_ptr() != addr, "the cached pixel tensor must stay alive"
        _lib.reset_launch_count()
        eb, _, _ = vlm.encode_and_fuse(ids, b)
        assert _lib.launch_count() > 6 and not torch.equal(ea, eb), "a new image must miss"
        _lib.reset_launch_count()
        eb2, _, _ = vlm.encode_and_fuse(ids, b)
        assert _lib.launch_count() <= 6 and torch.equal(eb, eb2), "the same tensor object must hit"
        # explicit identity: same id hits whatever tensor carries the pixels, a new id misses
        c = b.clone()
        e1, _, _ = vlm.encode_and_fuse(ids, c, image_id="img-1")
        _lib.reset_launch_count()
        e2, _, _ = vlm.encode_and_fuse(ids, c.clone(), image_id="img-1")
        assert _lib.launch_count() <= 6 and torch.equal(e1, e2)
        _lib.reset_launch_count()
        vlm.encode_and_fuse(ids, c, image_id="img-2")
        assert _lib.launch_count() > 6


def test_mrope_gated_attention_prefill_cfg3_sequence_vs_oracle():
    """SURVEY §8f-1 at the cfg-3 sequence length: 2 x 2832 tokens, 8 query / 2 kv heads of 256, multimodal position ids
    (the tiny fixture only covers seq 150). Oracle: fp32 restatement pinned to the reference module by the fixture."""
    from llm_quest_b200.common.rope import RoPE
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_text_model import MRoPEGatedAttention

    cfg = {"emb_dim": 1024, "n_heads": 8, "num_kv_groups": 2, "head_dim": 256, "dtype": torch.float32, "p_dropout": 0.0,
           "training": False, "mrope_section": [11, 11, 10]}
    torch.manual_seed(123)
    att = MRoPEGatedAttention(cfg, layer_idx=3).eval()
    with torch.no_grad():
        att.q_norm.scale.add_(0.1 * torch.randn(256))
        att.k_norm.scale.add_(0.1 * torch.randn(256))
    sd = {k: v.detach().clone() for k, v in att.state_dict().items()}
    b, seq = 2, 2832
    g = torch.Generator().manual_seed(99)
    x = torch.randn(b, seq, 1024, generator=g)
    ids = torch.randint(0, 1000, (b, seq), generator=g)
    for i in range(4):
        ids[:, 410 + i * 606: 410 + i * 606 + 196] = IMG
    pid = torch.from_numpy(FO.mrope_position_ids(ids.numpy(), [[1, 28, 28]] * 4, None, IMG, 2))
    cos, sin = RoPE.compute_angles(10_000_000, 256, 8192, rotation_factor=0.25)
    with torch.inference_mode():
        ref = VO.mrope_gated_attention_forward(sd, cfg, x, cos, sin, pid)
        out = att.cuda()(x.cuda(), None, cos, sin, position_ids=pid.cuda())
    check_close(out, ref, "MRoPEGatedAttention prefill 2 x 2832 tokens vs fp32 oracle")


def _reference():
    """The unmodified reference (baseline/_ref, travels to the GPU box) or None."""
    from baseline import ref

    return ref if ref.available() else None


def test_tiny_vit_config_vs_reference():
    """TINY_VIT_CONFIG (config.py:175-186): 4x4 patches of 32x32 images (S = 65), emb 256, 8 heads of 32, 12 layers, 10
    classes — the patch sizes / head dims the TMA gather GEMM and the tcgen05 attention are not built for run through
    vf_im2col_patches + GEMM and the CUDA-core attention kernel. Checked against the live reference module when
    baseline/_ref is present, else against the oracle restatement."""
    from llm_quest_b200.multimodal.vision_transformer.vit_model import ViTModel

    cfg = {"img_width": 32, "img_height": 32, "patch_size": 4, "num_channels": 3, "emb_dim": 256, "n_layers": 12, "n_heads": 8,
           "drop_rate": 0.3, "qkv_bias": True, "num_classes": 10}
    torch.manual_seed(123)
    m = ViTModel(cfg).eval()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    img = torch.randn(5, 3, 32, 32, generator=torch.Generator().manual_seed(1234))
    R = _reference()
    with torch.inference_mode():
        if R is not None:
            config = R.import_reference()
            assert {k: config.TINY_VIT_CONFIG[k] for k in cfg} == cfg
            rm = R.vit_model(config.TINY_VIT_CONFIG).eval()
            rm.load_state_dict(sd)
            ref_h, ref_l = rm(img, output_hidden_states=True), rm(img)
        else:
            ref_h, ref_l = VO.vit_forward(sd, cfg, img, output_hidden_states=True), VO.vit_forward(sd, cfg, img)
        mc = m.cuda()
        check_close(mc(img.cuda(), output_hidden_states=True), ref_h, "TINY_VIT_CONFIG hidden states")
        check_close(mc(img.cuda()), ref_l, "TINY_VIT_CONFIG logits")


def test_part2_text_half_and_concat_vs_reference():
    """Part-2 fusion (vlm_engine.py:100-119): adapter(vit_hidden) ‖ get_embeddings(ids) in ONE pre-allocated buffer —
    adapter rows through the GEMM's row remap, token + position embeddings by vf_embed_pos_concat. Reference:
    get_embeddings + torch.cat from baseline/_ref when present, else the same two lines in torch."""
    from llm_quest_b200.multimodal import vlm_engine as VE
    from llm_quest_b200.multimodal.vision_transformer.vit_engine import ViTAdapter

    torch.manual_seed(5)
    b, n_vis, d_vit, d, seq, vocab, ctx = 3, 197, 768, 768, 40, 5000, 64

    class GPT(torch.nn.Module):          # the two tables get_embeddings touches (gpt_model.py: emb_dict, pos_emb_dict)
        def __init__(self):
            super().__init__()
            self.emb_dict = torch.nn.Embedding(vocab, d)
            self.pos_emb_dict = torch.nn.Embedding(ctx, d)

    gpt = GPT().eval()
    ad = ViTAdapter(d_vit, d, adapter_type="ffn", hidden_size_factor=2).eval()
    hid = torch.randn(b, n_vis, d_vit)
    ids = torch.randint(0, vocab, (b, seq))
    with torch.inference_mode():
        R = _reference()
        if R is not None:
            R.import_reference()
            from llm_quest.multimodal.vlm_engine import get_embeddings as ref_get

            text_ref = ref_get(ids, gpt)
        else:
            text_ref = gpt.emb_dict(ids) + gpt.pos_emb_dict(torch.arange(seq))
        vis_ref = VO.vit_adapter_forward({k: v.detach() for k, v in ad.state_dict().items()}, hid)
        ref = torch.cat([vis_ref, text_ref], dim=1)
        gpt, ad = gpt.cuda(), ad.cuda()
        text = VE.get_embeddings(ids.cuda(), gpt)
        fused, n = VE.fuse_vision_text(ad, hid.cuda(), ids.cuda(), gpt)
    assert n == n_vis and fused.shape == (b, n_vis + seq, d)
    assert torch.equal(text.cpu(), text_ref), "token + position embeddings are exact fp32 adds"
    assert torch.equal(fused[:, n_vis:].cpu(), text_ref)
    check_close(fused[:, :n_vis], vis_ref, "adapter rows of the fused buffer", tol=5e-3)
    check_close(fused, ref, "Part-2 fused [vision | text] buffer", tol=5e-3)


def test_standalone_gelu_and_rmsnorm_modules():
    """GELU.forward (vit_transformer_block.py:43-44) and ZeroCenteredRMSNorm.forward (qwen3_next_attention.py:41-46)
    as modules of their own."""
    from llm_quest_b200.multimodal.vision_transformer.vit_transformer_block import GELU
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_text_model import ZeroCenteredRMSNorm

    x = torch.randn(7, 33, 300, generator=torch.Generator().manual_seed(2)) * 3
    with torch.inference_mode():
        y = GELU()(x.cuda())
        torch.testing.assert_close(y.cpu(), VO.gelu_erf(x), rtol=2e-6, atol=2e-6)
        yb = GELU()(x.to(torch.bfloat16).cuda())
        assert yb.dtype == torch.bfloat16
        torch.testing.assert_close(yb.float().cpu(), VO.gelu_erf(x.to(torch.bfloat16).float()).to(torch.bfloat16).float(), rtol=1e-2, atol=1e-2)
        for dt in (torch.float32, torch.bfloat16):
            n = ZeroCenteredRMSNorm(256, dtype=dt)
            with torch.no_grad():
                n.scale.add_((0.2 * torch.randn(256)).to(dt))
            xs = (torch.randn(4, 8, 50, 256, generator=torch.Generator().manual_seed(3)) * 2 + 0.3).to(dt)
            ref = VO.zero_centered_rmsnorm(xs, n.scale.detach())
            got = n.cuda()(xs.cuda())
            assert got.dtype == dt and got.shape == xs.shape
            if dt == torch.float32:
                torch.testing.assert_close(got.cpu(), ref, rtol=2e-6, atol=2e-6)
            else:   # one bf16 rounding of the same fp32 value: at most one ulp apart
                assert (got.float().cpu() - ref.float()).abs().max() <= 2.0 ** -7 * ref.float().abs().max()


def test_reference_vlm_forward_around_the_b200_tower():
    """End-to-end drop-in proof: the reference's OWN Qwen3_5VLM.forward — embedding lookup,
    masked_scatter, compute_3d_position_ids, text model — with only the vision tower replaced through shim.install(),
    on the GPU, against the unmodified reference on the CPU. The tensors handed to the text model (fused embeddings,
    position ids) are compared exactly / within the path tolerance; the logits loosely (the bf16 text model is the
    reference's own code on both sides, GPU vs CPU arithmetic)."""
    import importlib
    import sys

    import llm_quest_b200.shim as shim

    R = _reference()
    if R is None:
        pytest.skip("baseline/_ref not installed (baseline/install_ref.sh)")
    config = R.import_reference()
    px = 64
    over = {"n_layers": 4, "img_width": px, "img_height": px, "vision_n_layers": 2, "vocab_size": 3000, "image_token_id": 2999,
            "context_length": 256}
    cfg = {**config.QWEN3_5_08B_CONFIG, **over}
    import llm_quest.qwen.qwen3_5.qwen3_5_vlm_model as ref_vlm_mod

    ref_vlm_mod = importlib.reload(ref_vlm_mod)
    torch.manual_seed(123)
    ref_model = ref_vlm_mod.Qwen3_5VLM(dict(cfg)).eval()
    assert type(ref_model.vision_model).__module__.startswith("llm_quest.")
    sd = {k: v.detach().clone() for k, v in ref_model.state_dict().items()}
    g = torch.Generator().manual_seed(4321)
    b, T = 2, 4
    n_vis = (T // 2) * (px // 32) ** 2
    ids = torch.randint(0, 1000, (b, 40), generator=g)
    ids[:, 7:7 + n_vis] = 2999
    pixels = torch.randn(b, 3, T, px, px, generator=g)

    def run(model, dev):
        seen = {}
        hook = model.language_model.register_forward_pre_hook(
            lambda mod, args, kwargs: seen.update(embs=kwargs["inputs_embs"].detach().float().cpu(), pid=kwargs["position_ids"].detach().cpu()),
            with_kwargs=True)
        with torch.inference_mode():
            logits = model(ids.to(dev), image_pixels=pixels.to(dev))
        hook.remove()
        return logits.float().cpu(), seen

    ref_logits, ref_seen = run(ref_model, "cpu")
    try:
        shim.install(only=["llm_quest.qwen.qwen3_5.qwen3_5_vision_model"])
        mod = importlib.reload(ref_vlm_mod)                     # the reference's module, now importing the B200 tower
        torch.manual_seed(123)
        model = mod.Qwen3_5VLM(dict(cfg)).eval()
        assert type(model.vision_model).__module__.startswith("llm_quest_b200.")
        assert mod.Qwen3_5VLM.forward.__code__.co_filename.endswith("baseline/_ref/llm_quest/qwen/qwen3_5/qwen3_5_vlm_model.py")
        model.load_state_dict(sd)                               # same keys, same shapes: the reference's checkpoint loads
        got_logits, got_seen = run(model.cuda(), "cuda")
    finally:
        shim.uninstall()
        importlib.reload(ref_vlm_mod)
    assert torch.equal(got_seen["pid"], ref_seen["pid"]), "position ids differ"
    mask = (ids == 2999).view(-1)
    ge, re_ = got_seen["embs"].view(-1, cfg["emb_dim"]), ref_seen["embs"].view(-1, cfg["emb_dim"])
    assert torch.equal(ge[~mask], re_[~mask]), "text rows differ"
    check_close(ge[mask], re_[mask], "vision rows handed to the reference's text model")
    c = VO.cosine(got_logits, ref_logits)
    print(f"logits of the reference VLM around the B200 tower vs the CPU reference: cosine {c:.6f}")
    assert got_logits.shape == ref_logits.shape == (b, 40, 3000) and c >= 0.99


def test_hf_named_checkpoint_through_the_tower_vs_oracle():
    """SURVEY §8f-4 on the GPU: a checkpoint under Hugging Face names (model.visual.*, built with the inverse of the
    reference's own get_vision_remapping_rules when baseline/_ref is present) goes through load_qwen3_5_vision_weights
    into the B200 tower; the forward must match the fp32 oracle run on the original tensors. Also the 3-D pre-extracted
    patch branch of get_feeds_3d_shape (qwen3_5_vlm_model.py:77-81)."""
    from llm_quest_b200.qwen.qwen3_5 import qwen3_5_weight_loading as WL
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model import Qwen3_5VisionModel
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_vlm_model import EmbeddingOnlyLM, Qwen3_5VLM

    rules = WL.get_vision_remapping_rules()
    R = _reference()
    if R is not None:
        R.import_reference()
        from llm_quest.qwen.qwen3_5.qwen3_5_weight_loading import get_vision_remapping_rules as ref_rules

        assert rules == ref_rules(), "the remapping table must equal the reference's"
    cfg = qwen_cfg(224, vision_n_layers=3)
    torch.manual_seed(321)
    src = Qwen3_5VisionModel(cfg).eval()
    sd = {k: v.detach().clone() for k, v in src.state_dict().items()}

    def to_hf(k):
        for hf, ours in rules:
            if hf.startswith("model.visual.") and k.startswith(ours):
                k = hf + k[len(ours):]
                break
        for hf, ours in rules:
            if not hf.startswith("model.visual."):
                k = k.replace(ours, hf)
        return k

    hf = {to_hf(k): v for k, v in sd.items()}
    assert all(k.startswith("model.visual.") for k in hf) and "model.visual.blocks.2.mlp.linear_fc2.weight" in hf
    hf["model.language_model.embed_tokens.weight"] = torch.zeros(4, 4)          # non-vision keys are ignored
    torch.manual_seed(999)
    dst = Qwen3_5VisionModel(cfg).eval()
    missing, unexpected = WL.load_qwen3_5_vision_weights(dst, hf)
    assert missing == [] and unexpected == []
    pixels = torch.randn(2, 3, 2, 224, 224, generator=torch.Generator().manual_seed(1234))
    with torch.inference_mode():
        ref = VO.qwen_vision_forward(sd, cfg, pixels)
        out = dst.cuda()(pixels.cuda())
    check_close(out, ref, "tower loaded from an HF-named checkpoint vs fp32 oracle")
    vlm = Qwen3_5VLM(cfg, language_model=EmbeddingOnlyLM({**cfg, "vocab_size": 64}))
    assert vlm.get_feeds_3d_shape(torch.zeros(2, 3 * 196, 1536)).tolist() == [[3, 14, 14]]      # (b, num_patches, features)
    assert vlm.get_feeds_3d_shape(pixels).tolist() == [[1, 14, 14]]
