"""Part-2 VLM early fusion on libvfuse kernels: the encode + adapter + concat step.

Drop-in for the fusion step of the reference's ``llm_quest/multimodal/vlm_engine.py`` (``get_embeddings`` :5-20 and
the concat of :100-119) and ``vlm_generation.py:60-69``:

    vision_embeddings = adapter(vit_hidden_states)                       # [b, 197, d]
    text_embeddings   = get_embeddings(input_ids, vlm_model)             # tok_emb[ids] + pos_emb[arange(seq)]
    combined          = torch.cat([vision_embeddings, text_embeddings], dim=1)

Here the fused ``[b, 197 + seq, d]`` buffer is allocated once; the adapter's last GEMM writes its rows into
``[:, :197]`` through the row remap of its epilogue and ONE gather kernel (vf_embed_pos_concat) writes token + position
embeddings into ``[:, 197:]`` — no intermediate tensors, no cat. The training / generation loops, the loss and the GPT-2
model itself stay the reference's (out of scope); ``model`` only has to expose ``emb_dict`` and ``pos_emb_dict``.
"""

from __future__ import annotations

import torch

from .. import _lib
from .._lib import VFuseError


def _tables(model):
    tok, pos = model.emb_dict.weight.detach(), model.pos_emb_dict.weight.detach()
    if tok.dtype != pos.dtype or tok.dtype not in (torch.float32, torch.bfloat16):
        raise VFuseError(f"emb_dict / pos_emb_dict must share one of fp32 / bf16, got {tok.dtype} / {pos.dtype}")
    return tok.contiguous(), pos.contiguous()


def get_embeddings(text_input, model):
    """Token + positional embeddings of the text ids, [batch, seq, emb_dim] (reference vlm_engine.py:5-20)."""
    if not text_input.is_cuda:
        raise VFuseError("get_embeddings (llm_quest_b200) needs CUDA tensors; there is no CPU fallback")
    tok, pos = _tables(model)
    b, seq = text_input.shape
    out = torch.empty((b, seq, tok.shape[1]), dtype=torch.float32, device=text_input.device)
    _lib.embed_pos_concat(text_input, tok, pos, out, 0)
    return out if tok.dtype == torch.float32 else out.to(tok.dtype)


def fuse_vision_text(adapter, vit_hidden_states, input_ids, model):
    """[adapter(vit_hidden_states) ‖ get_embeddings(input_ids, model)] along dim 1, fp32 [b, n_vis + seq, d]
    (reference vlm_engine.py:105-114). Returns (combined_embeddings, num_vision_tokens)."""
    if not input_ids.is_cuda:
        raise VFuseError("fuse_vision_text (llm_quest_b200) needs CUDA tensors; there is no CPU fallback")
    tok, pos = _tables(model)
    b, n_vis, _ = vit_hidden_states.shape
    seq = input_ids.shape[1]
    d = tok.shape[1]
    if adapter.out_features() != d:
        raise VFuseError(f"adapter maps to {adapter.out_features()} features but the text embeddings have {d}")
    fused = torch.empty((b, n_vis + seq, d), dtype=torch.float32, device=input_ids.device)
    adapter.forward_into(vit_hidden_states, fused, row_off=0)
    _lib.embed_pos_concat(input_ids, tok, pos, fused, n_vis)
    return fused, n_vis
