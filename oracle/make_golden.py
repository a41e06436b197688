"""Pin the oracle against the LIVE reference and write the golden fixtures under tests/golden/.

Run in the build container only (it imports the reference from /root/reference, which does not
exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

For every piece of the path it (1) runs the reference's own module / method on seeded inputs,
(2) checks the oracle restatement against it (bit-exact for integer work, <= 2e-5 tensor-normalised
for float work — different but equivalent ATen op sequences), and (3) stores inputs + reference
outputs as small fixtures that tests/ load on any machine.
"""

from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("LLMQ_REF", "/root/reference"))
sys.dont_write_bytecode = True
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(REF))

from oracle import fusion_oracle as FO  # noqa: E402
from oracle import vision_oracle as VO  # noqa: E402

GOLD = ROOT / "tests" / "golden"
IMG = 248056


def bf16_exact(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def round_module_(m: torch.nn.Module) -> None:
    """Make every parameter exactly bf16-representable so the bf16 weight cast of the CUDA path is lossless."""
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(bf16_exact(p))


def tiny_qwen_cfg():
    return {
        "vision_emb_dim": 128, "vision_n_layers": 2, "vision_num_heads": 2, "vision_hidden_dim": 256,
        "vision_rope_base": 10_000, "llm_d_in": 128, "img_width": 64, "img_height": 96, "patch_size": 16,
        "in_channels": 3, "temporal_patch_size": 2, "spatial_merge_size": 2, "num_position_embeddings": 32,
        "image_token_id": IMG,
    }


def make_qwen_tower():
    from llm_quest.qwen.qwen3_5.qwen3_5_vision_model import Qwen3_5VisionModel

    cfg = tiny_qwen_cfg()
    torch.manual_seed(123)
    ref = Qwen3_5VisionModel(cfg).eval()
    round_module_(ref)
    g = torch.Generator().manual_seed(1234)
    pixels = bf16_exact(torch.randn(3, 3, 4, cfg["img_height"], cfg["img_width"], generator=g))
    with torch.inference_mode():
        out = ref(pixels)
        # hidden states before the merger, via the reference's own sub-modules
        x = ref.patch_embed(pixels)
        n = ref.n_spatial_patches
        frames = x.shape[1] // n
        x = x + ref.pos_embed(torch.arange(n)).unsqueeze(0).repeat(1, frames, 1)
        cos, sin = ref.cos.repeat(frames, 1), ref.sin.repeat(frames, 1)
        for blk in ref.blocks:
            x = blk(x, cos, sin)
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    o_out = VO.qwen_vision_forward(sd, cfg, pixels)
    o_hid = VO.qwen_vision_forward(sd, cfg, pixels, return_hidden=True)
    e1, e2 = VO.max_norm_err(o_out, out), VO.max_norm_err(o_hid, x)
    print(f"qwen tower: oracle vs reference  out {e1:.2e}  hidden {e2:.2e}")
    assert e1 < 2e-5 and e2 < 2e-5
    oc, os_ = VO.axial_rope_tables(cfg["vision_rope_base"], 64, ref.n_height_patches, ref.n_width_patches)
    assert torch.equal(oc, ref.cos) and torch.equal(os_, ref.sin), "axial rope tables differ"
    torch.save(
        {"cfg": cfg, "state_dict": {k: v.to(torch.bfloat16) for k, v in sd.items()}, "pixels": pixels.to(torch.bfloat16),
         "out": out.clone(), "hidden": x.clone(), "cos": ref.cos.clone(), "sin": ref.sin.clone()},
        GOLD / "qwen_tower_tiny.pt",
    )


def make_vit():
    from llm_quest.multimodal.vision_transformer.vit_engine import ViTAdapter
    from llm_quest.multimodal.vision_transformer.vit_model import ViTModel

    cfg = {"img_width": 64, "img_height": 64, "patch_size": 16, "num_channels": 3, "emb_dim": 128, "n_layers": 2,
           "n_heads": 2, "drop_rate": 0.1, "qkv_bias": True, "num_classes": 10}
    torch.manual_seed(123)
    ref = ViTModel(cfg).eval()
    round_module_(ref)
    g = torch.Generator().manual_seed(1234)
    images = bf16_exact(torch.randn(3, 3, 64, 64, generator=g))
    with torch.inference_mode():
        logits = ref(images)
        hidden = ref(images, output_hidden_states=True)
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    e1 = VO.max_norm_err(VO.vit_forward(sd, cfg, images), logits)
    e2 = VO.max_norm_err(VO.vit_forward(sd, cfg, images, output_hidden_states=True), hidden)
    print(f"part-1 vit: oracle vs reference  logits {e1:.2e}  hidden {e2:.2e}")
    assert e1 < 2e-5 and e2 < 2e-5

    torch.manual_seed(7)
    ad = ViTAdapter(128, 256, adapter_type="ffn", hidden_size_factor=2, bias=True).eval()
    round_module_(ad)
    with torch.inference_mode():
        a_out = ad(hidden)
    asd = {k: v.detach().clone() for k, v in ad.state_dict().items()}
    e3 = VO.max_norm_err(VO.vit_adapter_forward(asd, hidden), a_out)
    print(f"vit adapter: oracle vs reference {e3:.2e}")
    assert e3 < 2e-5
    torch.save(
        {"cfg": cfg, "state_dict": {k: v.to(torch.bfloat16) for k, v in sd.items()}, "images": images.to(torch.bfloat16),
         "logits": logits.clone(), "hidden": hidden.clone(),
         "adapter_state_dict": {k: v.to(torch.bfloat16) for k, v in asd.items()}, "adapter_out": a_out.clone()},
        GOLD / "vit_tiny.pt",
    )


class _Stub:
    """Just enough of Qwen3_5VLM for its integer methods (no text model is built)."""

    def __init__(self, merge=2):
        self.image_token_id = IMG
        self.merge_size = merge


def make_fusion():
    from llm_quest.qwen.qwen3_5.qwen3_5_vlm_model import Qwen3_5VLM

    stub = _Stub()
    cases = []

    def add(ids, feeds, mask=None):
        ids_t = torch.tensor(ids, dtype=torch.long)
        feeds_t = None if feeds is None else torch.tensor(feeds, dtype=torch.long)
        mask_t = None if mask is None else torch.tensor(mask, dtype=torch.bool)
        ref = Qwen3_5VLM.compute_3d_position_ids(stub, ids_t, feeds_t, image_mask=mask_t)
        ora = FO.mrope_position_ids(ids_t.numpy(), None if feeds is None else feeds, mask, IMG, 2)
        assert np.array_equal(ora, ref.numpy()), f"position ids differ for case {len(cases)}"
        cases.append({"ids": ids_t, "feeds": feeds_t, "mask": mask_t, "expected": ref.clone()})

    I = IMG
    add([[1, 2, 3, 4, 5, 6, I, I, I, I, 7, 8]], [[1, 4, 4]])                       # SURVEY golden (i) / reference docstring :96-101
    add([[1, I, I, I, I, 2, 3] + [I] * 12 + [4]], [[1, 4, 4], [2, 4, 6]])          # golden (ii): two feeds, one of 2 frames
    add([[1, I, I, I, 2]], [[1, 4, 4]])                                            # golden (iii): too few placeholders -> break
    add([[5, 6, 7, 8]], None)                                                      # text-only
    add([[1, 2, 3, 4], [I, I, I, I]], [[1, 4, 4]])                                 # a sample without placeholders in a multimodal batch
    add([[I] * 4 + [9] + [I] * 4 + [I] * 3], [[1, 4, 4], [1, 4, 4], [1, 4, 4]])    # third feed does not fit
    add([[3, I, I, 4, I, I, 5]], [[1, 4, 4]], mask=[[False, True, True, False, True, True, False]])  # explicit mask
    rng = np.random.default_rng(4321)
    for _ in range(6):
        b, seq = int(rng.integers(1, 5)), int(rng.integers(40, 2200))
        feeds = [[int(rng.integers(1, 4)), 2 * int(rng.integers(1, 6)), 2 * int(rng.integers(1, 6))]
                 for _ in range(int(rng.integers(1, 5)))]
        ids = rng.integers(0, 1000, size=(b, seq))
        for s in range(b):
            pos = 0
            for t, h, w in feeds:
                nt = t * (h // 2) * (w // 2)
                pos += int(rng.integers(0, 12))
                if rng.random() < 0.15:
                    nt = max(1, nt - 1)  # occasionally starve a feed
                if pos + nt > seq:
                    break
                ids[s, pos : pos + nt] = I
                pos += nt
        add(ids.tolist(), feeds)
    # cfg-3 shape: 4 images of 448^2 per sample + 2048 text tokens, batch 2 (the full batch is 32x this row)
    chunks = [410, 410, 410, 410, 408]
    row = []
    for i, c in enumerate(chunks):
        row += list(rng.integers(0, 1000, size=c))
        if i < 4:
            row += [I] * 196
    add([row, row[::-1]], [[1, 28, 28]] * 4)
    assert int(cases[-1]["expected"].max()) == 2103, int(cases[-1]["expected"].max())
    print(f"position ids: {len(cases)} cases bit-exact (oracle == reference)")

    # masked_scatter placement (placeholder id 63 so that it is a valid row of the 64-row toy table,
    # as 248056 is a valid row of the real 248320-row table)
    sc = []
    TOK = 63
    for b, seq, D in [(2, 9, 8), (3, 50, 16)]:
        ids = rng.integers(0, 50, size=(b, seq))
        ids[rng.random(size=(b, seq)) < 0.3] = TOK
        ids_t = torch.tensor(ids, dtype=torch.long)
        n_true = int((ids == TOK).sum())
        tg = torch.Generator().manual_seed(1000 + seq)      # seeded: the fixture reproduces byte for byte
        table = torch.randn(64, D, generator=tg).to(torch.bfloat16)
        vis = torch.randn(n_true + 3, D, generator=tg)
        embs = table[ids_t]
        mask = ids_t == TOK
        ref = embs.masked_scatter(mask.unsqueeze(-1).expand_as(embs), vis.to(embs.dtype))
        ora = FO.fuse_embeddings(ids, table.view(torch.uint16).numpy(), vis.to(torch.bfloat16).view(torch.uint16).numpy(),
                                 image_token_id=TOK)
        assert np.array_equal(ora, ref.view(torch.uint16).numpy()), "fusion differs from masked_scatter"
        sc.append({"ids": ids_t, "table": table, "vision": vis, "expected": ref.clone(), "image_token_id": TOK,
                   "row_map": torch.tensor(FO.scatter_row_map(ids, None, TOK))})
    print("masked_scatter placement: oracle row map == reference")
    torch.save({"position_cases": cases, "scatter_cases": sc}, GOLD / "fusion.pt")


def make_rope_and_merge():
    from llm_quest.common.buffers import GlobalBuffers
    from llm_quest.common.rope import RoPE, VisionRoPE
    from llm_quest.qwen.qwen3_5.qwen3_5_vision_model import ViTMergeAdapter
    from llm_quest.qwen.qwen3_next.qwen3_next_attention import ZeroCenteredRMSNorm

    out = {}
    # merge-gather indices (SURVEY golden Q7: t=2, nh=4, nw=6)
    ad = ViTMergeAdapter(1, 1, 4, 6, 2)
    x = torch.arange(48, dtype=torch.float32).view(1, 48, 1)
    g = x.view(1, 2, 2, 2, 3, 2, 1).permute(0, 1, 2, 4, 3, 5, 6).contiguous().view(1, -1, 4)[0].long()
    assert ad.m == 2
    gi = VO.merge_gather_index(2, 4, 6, 2)
    assert torch.equal(gi, g)
    assert gi[:4].tolist() == [[0, 1, 6, 7], [2, 3, 8, 9], [4, 5, 10, 11], [12, 13, 18, 19]] and gi[-1].tolist() == [40, 41, 46, 47]
    out["merge_index_2_4_6"] = gi

    # axial rope apply
    cos, sin = VisionRoPE.compute_angles_2d(10_000, 64, 3, 5)
    oc, os_ = VO.axial_rope_tables(10_000, 64, 3, 5)
    assert torch.equal(cos, oc) and torch.equal(sin, os_)
    gq = torch.Generator().manual_seed(99)
    q = torch.randn(2, 3, 15, 64, generator=gq)
    ref = VisionRoPE.apply(q, cos, sin)
    assert torch.equal(VO.rotate_half_apply(q, cos, sin), ref)
    out["rope2d"] = {"x": q, "cos": cos, "sin": sin, "expected": ref}

    # MRoPE-I
    cos_t, sin_t = GlobalBuffers.get_rope_params(8192, 10_000_000, 256, rotation_factor=0.25)
    oc, os_ = VO.text_rope_tables(8192, 10_000_000, 256, 0.25)
    assert torch.equal(cos_t, oc) and torch.equal(sin_t, os_)
    pid = torch.stack([torch.randint(0, 3000, (2, 37), generator=gq) for _ in range(3)])
    xq = torch.randn(2, 4, 37, 256, generator=gq)
    ref = RoPE.apply_mrope(xq, cos_t, sin_t, pid, [11, 11, 10])
    assert torch.equal(VO.mrope_apply(xq, cos_t, sin_t, pid, [11, 11, 10]), ref)
    axes = VO.mrope_slot_axes(32, [11, 11, 10])
    assert "".join("THW"[a] for a in axes) == "THW" * 10 + "TH", axes
    norm = ZeroCenteredRMSNorm(256)
    with torch.no_grad():
        norm.scale.copy_(torch.randn(256, generator=gq) * 0.1)
    with torch.inference_mode():
        nref = norm(xq)
    assert torch.equal(VO.zero_centered_rmsnorm(xq, norm.scale.detach()), nref)
    nm = RoPE.apply_mrope(nref, cos_t, sin_t, pid, [11, 11, 10])
    out["mrope"] = {"x": xq, "position_ids": pid, "sections": [11, 11, 10], "expected": ref,
                    "norm_scale": norm.scale.detach().clone(), "expected_norm_mrope": nm,
                    "table": {"ctx": 8192, "base": 10_000_000, "head_dim": 256, "factor": 0.25}}
    print("rope / mrope / rmsnorm / merge index: oracle == reference (bit-exact)")
    torch.save(out, GOLD / "rope_merge.pt")


def make_text_attention():
    """First consumer of the ids and fused embeddings: MRoPEGatedAttention in prefill."""
    from llm_quest.common.buffers import GlobalBuffers
    from llm_quest.qwen.qwen3_5.qwen3_5_text_model import MRoPEGatedAttention

    cfg = {"emb_dim": 256, "n_heads": 4, "num_kv_groups": 2, "head_dim": 256, "dtype": torch.float32, "p_dropout": 0.0,
           "training": False, "mrope_section": [11, 11, 10]}
    torch.manual_seed(123)
    att = MRoPEGatedAttention(cfg, layer_idx=0).eval()
    with torch.no_grad():
        att.q_norm.scale.copy_(0.1 * torch.randn(256))
        att.k_norm.scale.copy_(0.1 * torch.randn(256))
    round_module_(att)
    b, seq = 2, 150
    g = torch.Generator().manual_seed(77)
    x = bf16_exact(torch.randn(b, seq, 256, generator=g))
    # multimodal-looking ids: text, a 2x(4x6) image block, text (axes differ inside the block)
    ids = torch.randint(0, 1000, (b, seq), generator=g)
    ids[:, 20:20 + 12] = IMG
    ids[1, 80:80 + 12] = IMG
    pid = torch.from_numpy(FO.mrope_position_ids(ids.numpy(), [[2, 4, 6], [2, 4, 6]], None, IMG, 2))
    cos, sin = GlobalBuffers.get_rope_params(512, 10_000_000, 256, rotation_factor=0.25)
    mask = ~GlobalBuffers.get_causal_mask(512)
    with torch.inference_mode():
        ref = att(x, mask, cos, sin, position_ids=pid)
        ref_1d = att(x, mask, cos, sin, position_ids=torch.arange(seq).expand(3, b, seq))
    sd = {k: v.detach().clone() for k, v in att.state_dict().items()}
    oc, os_ = VO.text_rope_tables(512, 10_000_000, 256, 0.25)
    assert torch.equal(oc, cos) and torch.equal(os_, sin)
    got = VO.mrope_gated_attention_forward(sd, cfg, x, cos, sin, pid)
    err = VO.max_norm_err(got, ref)
    assert err <= 2e-5, err
    assert VO.max_norm_err(VO.mrope_gated_attention_forward(sd, cfg, x, cos, sin, None), ref_1d) <= 2e-5
    print(f"MRoPEGatedAttention prefill: oracle vs reference max_norm_err={err:.2e}")
    # weights and x are bf16-exact: stored as bf16 to keep the fixture small (tests cast back to fp32)
    sd = {k: v.to(torch.bfloat16) for k, v in sd.items()}
    torch.save({"cfg": {k: v for k, v in cfg.items() if k != "dtype"}, "state_dict": sd, "x": x.to(torch.bfloat16), "position_ids": pid,
                "expected": ref.clone(), "expected_1d": ref_1d.clone(), "rope": {"ctx": 512, "base": 10_000_000, "factor": 0.25}},
               GOLD / "text_attention.pt")


def make_part2_and_tiny():
    """Round-2 surface: Part-2 get_embeddings + concat (multimodal/vlm_engine.py:5-20,114), TINY_VIT_CONFIG dims
    (config.py:175-186: 4x4 patches, head_dim 32; 2 of the 12 layers to keep the fixture small), stand-alone GELU
    (vit_transformer_block.py:43-44) and ZeroCenteredRMSNorm.forward (qwen3_next_attention.py:41-46)."""
    import config
    from llm_quest.multimodal.vision_transformer.vit_model import ViTModel
    from llm_quest.multimodal.vision_transformer.vit_transformer_block import GELU
    from llm_quest.multimodal.vlm_engine import get_embeddings
    from llm_quest.qwen.qwen3_next.qwen3_next_attention import ZeroCenteredRMSNorm

    g = torch.Generator().manual_seed(2024)

    class GPT(torch.nn.Module):      # the two tables get_embeddings reads (gpt/gpt_model.py: emb_dict, pos_emb_dict)
        def __init__(self):
            super().__init__()
            self.emb_dict = torch.nn.Embedding(500, 64)
            self.pos_emb_dict = torch.nn.Embedding(32, 64)

    torch.manual_seed(123)
    gpt = GPT().eval()
    ids = torch.randint(0, 500, (3, 20), generator=g)
    vis = torch.randn(3, 5, 64, generator=g)
    with torch.no_grad():
        text = get_embeddings(ids, gpt)
        fused = torch.cat([vis, text], dim=1)                                   # vlm_engine.py:114
    assert torch.equal(VO.get_embeddings(ids, gpt.emb_dict.weight, gpt.pos_emb_dict.weight), text)

    cfg = dict(config.TINY_VIT_CONFIG, n_layers=2)
    torch.manual_seed(123)
    vit = ViTModel(cfg).eval()
    round_module_(vit)
    img = bf16_exact(torch.randn(4, 3, 32, 32, generator=g))
    with torch.no_grad():
        hid, logits = vit(img, output_hidden_states=True), vit(img)
    sd = {k: v.detach().clone() for k, v in vit.state_dict().items()}
    e_h = VO.max_norm_err(VO.vit_forward(sd, cfg, img, output_hidden_states=True), hid)
    e_l = VO.max_norm_err(VO.vit_forward(sd, cfg, img), logits)
    assert e_h <= 2e-5 and e_l <= 2e-5, (e_h, e_l)
    print(f"TINY_VIT_CONFIG dims (2 layers): oracle vs reference hidden {e_h:.2e} logits {e_l:.2e}")

    x = torch.randn(6, 50, generator=g) * 3
    with torch.no_grad():
        y_gelu = GELU()(x)
    assert torch.equal(VO.gelu_erf(x), y_gelu)
    norm = ZeroCenteredRMSNorm(64)
    with torch.no_grad():
        norm.scale.add_(0.2 * torch.randn(64, generator=g))
        xn = torch.randn(5, 7, 64, generator=g) * 2 + 0.3
        y_norm = norm(xn)
        y_norm_bf16 = norm(xn.to(torch.bfloat16))
    assert torch.equal(VO.zero_centered_rmsnorm(xn, norm.scale.detach()), y_norm)
    print("get_embeddings / GELU / ZeroCenteredRMSNorm: oracle == reference (bit-exact)")
    torch.save({"part2": {"ids": ids, "tok": gpt.emb_dict.weight.detach().clone(), "pos": gpt.pos_emb_dict.weight.detach().clone(),
                          "vision": vis, "text": text, "fused": fused},
                "tiny_vit": {"cfg": cfg, "state_dict": {k: v.to(torch.bfloat16) for k, v in sd.items()}, "images": img.to(torch.bfloat16),
                             "hidden": hid.clone(), "logits": logits.clone()},
                "gelu": {"x": x, "y": y_gelu}, "rmsnorm": {"x": xn, "scale": norm.scale.detach().clone(), "y": y_norm, "y_bf16": y_norm_bf16}},
               GOLD / "part2_tiny.pt")


MAKERS = {"fusion": lambda: make_fusion(), "rope": lambda: make_rope_and_merge(), "qwen": lambda: make_qwen_tower(),
          "vit": lambda: make_vit(), "text_attention": lambda: make_text_attention(), "part2_tiny": lambda: make_part2_and_tiny()}

if __name__ == "__main__":
    GOLD.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    for name in (sys.argv[1:] or list(MAKERS)):      # no argument: every fixture
        MAKERS[name]()
    for f in sorted(GOLD.glob("*.pt")):
        print(f"{f.name}: {f.stat().st_size / 1024:.0f} KiB")
