"""CPU oracle for the floating-point half of the vision-encode-and-fuse path.

TEST INFRASTRUCTURE ONLY. Nothing under ``llm_quest_b200/`` may import this file; it is used by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` as the checker (and as the timed CPU baseline), never as the product path.

It restates, as plain functions over a ``state_dict`` (fp32, CPU, PyTorch ATen ops — the same
third-party arithmetic the reference itself runs on, torch 2.11), what the reference's modules
compute. Pinning: the reference ships no golden vectors or tests for this path, so
the oracle is pinned against the LIVE reference instead — ``oracle/make_golden.py`` imports
``/root/reference`` in the build container, checks every function here against the reference
modules on seeded inputs (exact equality for the conv-free parts, <=2e-6 for the rest) and writes
the fixtures under ``tests/golden/`` that travel to the GPU box.

Reference line citations are relative to ``/root/reference``.
"""

from __future__ import annotations

import math

import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------------
# rotary tables
# ------------------------------------------------------------------------------------------------


def axial_rope_tables(base: float, head_dim: int, nh: int, nw: int, frames: int = 1):
    """cos/sin [frames*nh*nw, head_dim] of the axial 2-D RoPE.

    llm_quest/common/rope.py:400-482 — theta_i = base^(-2i/(head_dim/2)), i < head_dim/4; a patch at
    (row r, col c) gets angles [r*theta | c*theta], duplicated to fill head_dim.
    """
    assert head_dim % 4 == 0
    quarter = head_dim // 4
    theta = 1.0 / (base ** (2 * torch.arange(0, quarter, dtype=torch.float32) / (head_dim // 2)))
    rows = torch.arange(nh, dtype=torch.float32).repeat_interleave(nw)
    cols = torch.arange(nw, dtype=torch.float32).repeat(nh)
    ang = torch.cat([rows[:, None] * theta[None, :], cols[:, None] * theta[None, :]], dim=1)
    if frames > 1:
        ang = ang.repeat(frames, 1)
    ang = torch.cat([ang, ang], dim=1)
    return ang.cos(), ang.sin()


def text_rope_tables(ctx_len: int, base: float, head_dim: int, rotation_factor: float = 1.0):
    """cos/sin [ctx_len, rot] of the 1-D text RoPE with partial rotation.

    llm_quest/common/rope.py:97-166 (no YaRN): rot = head_dim*rotation_factor,
    theta_j = base^(-2j/rot), angles duplicated.
    """
    rot = int(head_dim * rotation_factor)
    theta = 1.0 / base ** (2 * torch.arange(0, rot // 2, dtype=torch.float32) / rot)
    ang = torch.outer(torch.arange(0, ctx_len, dtype=torch.float32), theta)
    ang = torch.cat([ang, ang], dim=-1)
    return ang.cos(), ang.sin()


def rotate_half_apply(x, cos, sin, position_ids=None):
    """Rotate-half RoPE on x [b, h, s, hd]; llm_quest/common/rope.py:180-243."""
    hd = x.shape[-1]
    rot = cos.shape[-1]
    rest = None
    if rot < hd:
        rest = x[..., rot:]
        x = x[..., :rot]
    if position_ids is not None:
        c = cos[position_ids].unsqueeze(1).to(x.dtype)
        s = sin[position_ids].unsqueeze(1).to(x.dtype)
    else:
        c = cos[: x.shape[2]].to(x.dtype)
        s = sin[: x.shape[2]].to(x.dtype)
    half = rot // 2
    turned = torch.cat((-x[..., half:], x[..., :half]), dim=-1)
    y = c * x + s * turned
    return y if rest is None else torch.cat((y, rest), dim=-1)


def mrope_slot_axes(half: int, sections) -> list[int]:
    """Which position axis (0=T,1=H,2=W) half-dim slot j uses under MRoPE-I.

    llm_quest/common/rope.py:283-294: start from T everywhere, overwrite slice(1, 3*sec_h, 3) with H
    and slice(2, 3*sec_w, 3) with W.
    """
    axes = [0] * half
    for j in range(1, min(3 * sections[1], half), 3):
        axes[j] = 1
    for j in range(2, min(3 * sections[2], half), 3):
        axes[j] = 2
    return axes


def mrope_apply(x, cos, sin, position_ids, sections):
    """MRoPE-I on x [b, h, s, hd] with position_ids [3, b, s]; llm_quest/common/rope.py:297-358."""
    hd = x.shape[-1]
    rot = cos.shape[-1]
    half = rot // 2
    axes = torch.tensor(mrope_slot_axes(half, sections))
    slot = torch.arange(half)
    # pos_for_slot[b, s, j] = position_ids[axes[j], b, s]
    pos = position_ids[axes, :, :].permute(1, 2, 0)  # [b, s, half]
    c = cos[:, :half][pos, slot]  # [b, s, half]
    s_ = sin[:, :half][pos, slot]
    c = torch.cat([c, c], dim=-1).unsqueeze(1).to(x.dtype)
    s_ = torch.cat([s_, s_], dim=-1).unsqueeze(1).to(x.dtype)
    rest = None
    if rot < hd:
        rest = x[..., rot:]
        x = x[..., :rot]
    turned = torch.cat((-x[..., half:], x[..., :half]), dim=-1)
    y = c * x + s_ * turned
    return y if rest is None else torch.cat((y, rest), dim=-1)


def zero_centered_rmsnorm(x, scale, eps: float = 1e-6):
    """llm_quest/qwen/qwen3_next/qwen3_next_attention.py:41-46."""
    xf = x.to(torch.float32)
    r = torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)
    return (xf * r * (1.0 + scale)).to(x.dtype)


# ------------------------------------------------------------------------------------------------
# Qwen3.5 vision tower (llm_quest/qwen/qwen3_5/qwen3_5_vision_model.py)
# ------------------------------------------------------------------------------------------------


def gelu_tanh(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x.pow(3))))


def gelu_erf(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def patch_embed3d(x, weight, bias):
    """[B,C,T,H,W] -> [B, T'*nh*nw, D] via the non-overlapping conv written as a matmul.

    qwen3_5_vision_model.py:105-107. K order (c, dt, py, px) == weight.flatten(1).
    """
    B, Cc, T, H, W = x.shape
    D, _, tp, P, _ = weight.shape
    nh, nw, Tp = H // P, W // P, T // tp
    cols = x.reshape(B, Cc, Tp, tp, nh, P, nw, P).permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(B, Tp * nh * nw, -1)
    return cols @ weight.reshape(D, -1).t() + bias


def merge_gather_index(frames: int, nh: int, nw: int, m: int = 2) -> torch.Tensor:
    """[frames*(nh/m)*(nw/m), m*m] source-token index of ViTMergeAdapter's view/permute/view.

    qwen3_5_vision_model.py:425-427.
    """
    idx = torch.arange(frames * nh * nw).view(frames, nh // m, m, nw // m, m)
    return idx.permute(0, 1, 3, 2, 4).reshape(-1, m * m)


def qwen_vision_forward(sd: dict, cfg: dict, pixels: torch.Tensor, return_hidden: bool = False):
    """Qwen3_5VisionModel.forward (qwen3_5_vision_model.py:336-370) as one function over a state_dict."""
    D = cfg["vision_emb_dim"]
    Hh = cfg["vision_num_heads"]
    hd = D // Hh
    P = cfg["patch_size"]
    nh, nw = cfg["img_height"] // P, cfg["img_width"] // P
    n = nh * nw
    x = patch_embed3d(pixels, sd["patch_embed.conv_proj.weight"], sd["patch_embed.conv_proj.bias"])
    B, S, _ = x.shape
    frames = S // n
    x = x + sd["pos_embed.weight"][:n].repeat(frames, 1)[None, :S]
    cos, sin = axial_rope_tables(cfg["vision_rope_base"], hd, nh, nw)
    cos, sin = cos.repeat(frames, 1), sin.repeat(frames, 1)
    for i in range(cfg["vision_n_layers"]):
        p = f"blocks.{i}."
        h = F.layer_norm(x, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6)
        qkv = h @ sd[p + "att.qkv.weight"].t() + sd[p + "att.qkv.bias"]
        q, k, v = (t.view(B, S, Hh, hd).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
        q = rotate_half_apply(q, cos, sin)
        k = rotate_half_apply(k, cos, sin)
        att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1) @ v
        att = att.transpose(1, 2).reshape(B, S, D)
        x = x + (att @ sd[p + "att.proj.weight"].t() + sd[p + "att.proj.bias"])
        h = F.layer_norm(x, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6)
        h = gelu_tanh(h @ sd[p + "ffn.lin1.weight"].t() + sd[p + "ffn.lin1.bias"])
        x = x + (h @ sd[p + "ffn.lin2.weight"].t() + sd[p + "ffn.lin2.bias"])
    if return_hidden:
        return x
    return merge_adapter_forward(sd, "merge_adapter.", x, nh, nw, cfg["spatial_merge_size"])


def merge_adapter_forward(sd, prefix, x, nh, nw, m):
    """ViTMergeAdapter.forward (qwen3_5_vision_model.py:411-431)."""
    B, S, D = x.shape
    frames = S // (nh * nw)
    h = F.layer_norm(x, (D,), sd[prefix + "norm.weight"], sd[prefix + "norm.bias"], 1e-6)
    gi = merge_gather_index(frames, nh, nw, m)
    h = h[:, gi, :].reshape(B, gi.shape[0], m * m * D)
    h = gelu_erf(h @ sd[prefix + "lin1.weight"].t() + sd[prefix + "lin1.bias"])
    return h @ sd[prefix + "lin2.weight"].t() + sd[prefix + "lin2.bias"]


# ------------------------------------------------------------------------------------------------
# Part-1 ViT classifier (llm_quest/multimodal/vision_transformer/)
# ------------------------------------------------------------------------------------------------


def std_layernorm(x, scale, shift, eps: float = 1e-5):
    """vit_transformer_block.py:27-31 — eps is added to the (biased) std, not to the variance."""
    mean = x.mean(dim=-1, keepdim=True)
    std = x.std(dim=-1, keepdim=True, unbiased=False)
    return scale * ((x - mean) / (std + eps)) + shift


def vit_forward(sd: dict, cfg: dict, images: torch.Tensor, output_hidden_states: bool = False):
    """ViTModel.forward in eval mode (vit_model.py:134-160; blocks vit_transformer_block.py:102-127;
    attention vit_attention.py:43-91)."""
    D, Hh, P = cfg["emb_dim"], cfg["n_heads"], cfg["patch_size"]
    hd = D // Hh
    B = images.shape[0]
    w = sd["patch_embedding.conv_proj.weight"]
    x = patch_embed3d(images.unsqueeze(2), w.unsqueeze(2), sd["patch_embedding.conv_proj.bias"])
    x = torch.cat([sd["patch_embedding.cls_token"].expand(B, -1, -1), x], dim=1) + sd["pos_embedding"]
    S = x.shape[1]
    for i in range(cfg["n_layers"]):
        p = f"transformer_blocks.{i}."
        h = std_layernorm(x, sd[p + "ln_1.scale"], sd[p + "ln_1.shift"])

        def lin(t, name):
            y = t @ sd[p + name + ".weight"].t()
            b = sd.get(p + name + ".bias")
            return y if b is None else y + b

        q, k, v = (lin(h, nm).view(B, S, Hh, hd).transpose(1, 2) for nm in ("att.w_queries", "att.w_keys", "att.w_values"))
        a = torch.softmax((q @ k.transpose(-1, -2)) * hd**-0.5, dim=-1) @ v
        x = x + lin(a.transpose(1, 2).reshape(B, S, D), "att.out_proj")
        h = std_layernorm(x, sd[p + "ln_2.scale"], sd[p + "ln_2.shift"])
        h = gelu_erf(lin(h, "ffn.layers.0"))
        x = x + lin(h, "ffn.layers.2")
    x = std_layernorm(x, sd["final_ln.scale"], sd["final_ln.shift"])
    if output_hidden_states:
        return x
    return x[:, 0] @ sd["classifier.weight"].t() + sd["classifier.bias"]


def get_embeddings(text_input: torch.Tensor, tok_table: torch.Tensor, pos_table: torch.Tensor):
    """llm_quest/multimodal/vlm_engine.py:5-20 — emb_dict(ids) + pos_emb_dict(arange(seq)); the Part-2 fusion then
    concatenates [vision ‖ text] along dim 1 (vlm_engine.py:114, vlm_generation.py:66)."""
    return tok_table[text_input] + pos_table[torch.arange(text_input.shape[1])]


def vit_adapter_forward(sd: dict, x: torch.Tensor):
    """ViTAdapter.forward (vit_engine.py:44-59): 'simple' = one Linear, 'ffn' = Linear-GELU(erf)-Linear."""
    if "adapter.weight" in sd:
        y = x @ sd["adapter.weight"].t()
        return y + sd["adapter.bias"] if "adapter.bias" in sd else y
    h = x @ sd["adapter.0.weight"].t()
    if "adapter.0.bias" in sd:
        h = h + sd["adapter.0.bias"]
    y = gelu_erf(h) @ sd["adapter.3.weight"].t()
    return y + sd["adapter.3.bias"] if "adapter.3.bias" in sd else y


def mrope_gated_attention_forward(sd: dict, cfg: dict, x, cos, sin, position_ids=None):
    """MRoPEGatedAttention.forward in prefill (qwen3_5_text_model.py:206-267; no KV cache, no padding mask):
    fused q|gate projection split per head (:216-219), k / v projections, zero-centred RMSNorm on q and k
    (qwen3_next_attention.py:41-46), MRoPE-I (rope.py:297-358), causal grouped-query SDPA (:252-260), sigmoid
    gate (:262), out_proj. sd keys: w_queries_gate.weight, w_keys.weight, w_values.weight, q_norm.scale,
    k_norm.scale, out_proj.weight."""
    b, seq, _ = x.shape
    H, G, hd = cfg["n_heads"], cfg["num_kv_groups"], cfg["head_dim"]
    qg = (x @ sd["w_queries_gate.weight"].t()).view(b, seq, H, 2 * hd)
    q, gate = qg[..., :hd], qg[..., hd:]
    k = (x @ sd["w_keys.weight"].t()).view(b, seq, G, hd).transpose(1, 2)
    v = (x @ sd["w_values.weight"].t()).view(b, seq, G, hd).transpose(1, 2)
    q = zero_centered_rmsnorm(q.transpose(1, 2), sd["q_norm.scale"])
    k = zero_centered_rmsnorm(k, sd["k_norm.scale"])
    if position_ids is None:
        position_ids = torch.arange(seq).expand(3, b, seq)
    q = mrope_apply(q, cos, sin, position_ids, cfg["mrope_section"])
    k = mrope_apply(k, cos, sin, position_ids, cfg["mrope_section"])
    rep = H // G
    kk, vv = k.repeat_interleave(rep, dim=1), v.repeat_interleave(rep, dim=1)
    att = (q @ kk.transpose(-1, -2)) * hd**-0.5
    att = att.masked_fill(torch.triu(torch.ones(seq, seq, dtype=torch.bool), diagonal=1), float("-inf"))
    ctx = (torch.softmax(att, dim=-1) @ vv).transpose(1, 2).reshape(b, seq, H * hd)
    ctx = ctx * torch.sigmoid(gate.reshape(b, seq, H * hd))
    return ctx @ sd["out_proj.weight"].t()


def preprocess_u8(images_u8: torch.Tensor, mean, std, temporal_patch_size: int = 2) -> torch.Tensor:
    """uint8 [B, H, W, 3] -> fp32 [B, 3, T, H, W], the pre-processing that feeds PatchEmbedding3D
    (qwen3_5_generate_multimodal.py:40-46 after the resize): torchvision to_tensor (HWC -> CHW, / 255),
    normalize ((x - mean) / std), repeat along a new temporal axis, permute to (B, C, T, H, W).
    Pinned against torchvision itself in tests/test_oracle.py."""
    x = images_u8.permute(0, 3, 1, 2).contiguous().to(torch.float32).div(255)
    mean_t = torch.as_tensor(mean, dtype=torch.float32).view(1, 3, 1, 1)
    std_t = torch.as_tensor(std, dtype=torch.float32).view(1, 3, 1, 1)
    x = (x - mean_t) / std_t
    return x.unsqueeze(1).repeat(1, temporal_patch_size, 1, 1, 1).permute(0, 2, 1, 3, 4).contiguous()


# ------------------------------------------------------------------------------------------------
# error metrics (tensor-normalised max error, cosine)
# ------------------------------------------------------------------------------------------------


def max_norm_err(a: torch.Tensor, ref: torch.Tensor) -> float:
    a, ref = a.double().flatten(), ref.double().flatten()
    return float((a - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def cosine(a: torch.Tensor, ref: torch.Tensor) -> float:
    a, ref = a.double().flatten(), ref.double().flatten()
    return float((a @ ref) / (a.norm() * ref.norm()).clamp_min(1e-30))
