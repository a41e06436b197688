"""Part-1 ViT classifier on libvfuse kernels.

Drop-in for ``PatchEmbedding2D`` (:19-89) and ``ViTModel`` (:92-160) of the reference's
``llm_quest/multimodal/vision_transformer/vit_model.py``: same constructor arguments, ``forward``
signatures, ``state_dict`` keys (``pos_embedding``, ``patch_embedding.cls_token``,
``patch_embedding.conv_proj.*``, ``transformer_blocks.{i}.*``, ``final_ln.*``, ``classifier.*``).

conv2d patchify runs as the im2col-free TMA-gather GEMM (the 3-D kernel with T = tp = 1) whose
epilogue adds bias and the positional embedding and writes token-major rows 1..N of every sample;
row 0 (class token + pos[0]) is written by vf_vit_cls_pos. Patch sizes other than 16 (TINY_VIT_CONFIG, config.py:175-186)
go through vf_im2col_patches + the plain GEMM; head dims other than 64 through the CUDA-core attention kernel.
"""

from __future__ import annotations

import torch
import torch.nn as nn

from ... import _lib
from ..._lib import VF_EPI_BIAS_F32, VF_EPI_BIAS_RES_F32, VFuseError
from ...qwen.qwen3_5.qwen3_5_vision_model import _Packed, _f32, _forward_only_guard, _w_bf16
from .vit_transformer_block import LayerNorm, ViTTransformerBlock


class PatchEmbedding2D(nn.Module):
    def __init__(self, img_width, img_height, patch_size, num_channels, emb_dim):
        super().__init__()
        assert img_width % patch_size == 0, f"Image width {img_width} not divisible by patch size {patch_size}"
        assert img_height % patch_size == 0, f"Image height {img_height} not divisible by patch size {patch_size}"
        self.img_width = img_width
        self.img_height = img_height
        self.patch_size = patch_size
        self.num_patches = (img_width * img_height) // patch_size**2
        k = (patch_size, patch_size)
        self.conv_proj = nn.Conv2d(num_channels, emb_dim, kernel_size=k, stride=k, padding=0, bias=True)
        self.cls_token = nn.Parameter(torch.randn(1, 1, emb_dim))
        self._packed = _Packed()

    def embed_into(self, x, pos=None):
        """fp32 [B*(N+1), D]: row 0 = cls (+pos[0]), rows 1.. = conv(x)+bias (+pos[1+p])."""
        assert x.shape[2] == self.img_width and x.shape[3] == self.img_height, (
            f"Input image shape {x.shape} does not match expected shape {self.img_width}x{self.img_height}"
        )
        if not x.is_cuda:
            raise VFuseError("PatchEmbedding2D (llm_quest_b200) runs on CUDA sm_100a only; got a CPU tensor")
        B = x.shape[0]
        D = self.conv_proj.out_channels
        S = self.num_patches + 1
        w = _w_bf16(self._packed, "w", self.conv_proj.weight, (D, -1))
        bias = _f32(self._packed, "b", self.conv_proj.bias)
        cls = _f32(self._packed, "cls", self.cls_token).view(-1)
        out = torch.empty((B * S, D), dtype=torch.float32, device=x.device)
        if pos is None:
            pos = self._packed.get("zero_pos", [self.cls_token], lambda: torch.zeros((S, D), device=x.device))
        if self.patch_size == 16:
            px = _lib.to_bf16(x).unsqueeze(2)  # [B, C, 1, H, W]
            _lib.patch_embed(px, w, bias, pos[1:], out, self.patch_size, 1, S, 1)
            _lib.vit_cls_pos(cls, pos[0], out, B, S, D)
        else:
            # other patch sizes (TINY_VIT_CONFIG: 4x4): patches unfolded to bf16 rows, every output row pre-filled with its
            # position embedding (+ class token on row 0), then one GEMM accumulates conv + bias onto rows 1.. of each sample
            xin = x.contiguous() if x.dtype in (torch.float32, torch.bfloat16) else x.float().contiguous()
            rows = _lib.im2col_patches(xin, self.patch_size)
            _lib.fill_rows(pos.contiguous(), out, B, S, D, add_row0=cls)
            _lib.gemm(rows, w, VF_EPI_BIAS_RES_F32, out, bias=bias, res=out, grp_rows=self.num_patches, grp_stride=S, row_off=1)
        return out, B, S

    def forward(self, x):
        _forward_only_guard(self)
        out, B, S = self.embed_into(x)
        return out.view(B, S, -1).to(self.conv_proj.weight.dtype)


class ViTModel(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.patch_embedding = PatchEmbedding2D(
            img_width=cfg["img_width"], img_height=cfg["img_height"], patch_size=cfg["patch_size"],
            num_channels=cfg["num_channels"], emb_dim=cfg["emb_dim"],
        )
        self.pos_embedding = nn.Parameter(torch.randn(1, self.patch_embedding.num_patches + 1, cfg["emb_dim"]))
        self.dropout = nn.Dropout(cfg["drop_rate"])
        self.transformer_blocks = nn.ModuleList([ViTTransformerBlock(cfg) for _ in range(cfg["n_layers"])])
        self.final_ln = LayerNorm(cfg["emb_dim"])
        self.classifier = nn.Linear(cfg["emb_dim"], cfg["num_classes"])
        self._packed = _Packed()

    ln_fold = True    # fold ln_1 / ln_2 of every block into the neighbouring GEMMs (plain attribute; False = stand-alone kernels)

    def forward(self, x, output_hidden_states=False):
        """x [b, C, H, W] -> logits [b, num_classes], or the final hidden states [b, N+1, D]."""
        _forward_only_guard(self)
        pos = _f32(self._packed, "pos", self.pos_embedding)[0]
        x2d, B, S = self.patch_embedding.embed_into(x, pos)
        D = x2d.shape[1]
        rows = B * S
        work = {
            "h": torch.empty((rows, D), dtype=torch.bfloat16, device=x.device),
            "g": torch.empty((rows, 4 * D), dtype=torch.bfloat16, device=x.device),
        }
        if self.ln_fold and D % 32 == 0:     # LayerNorms folded into the GEMMs around them (see ViTTransformerBlock.run_)
            work["stat"] = torch.empty((D // 32, rows, 2), dtype=torch.float32, device=x.device)
            work["rows"] = torch.empty((rows, 2), dtype=torch.float32, device=x.device)
            work["shift"] = torch.empty((rows,), dtype=torch.float32, device=x.device)
        last = len(self.transformer_blocks) - 1
        for i, block in enumerate(self.transformer_blocks):
            block.run_(x2d, B, S, work, ln1_pending=i > 0, emit_next=i < last)
        lw, lb = self.final_ln.packed()
        pdt = self.pos_embedding.dtype
        if output_hidden_states:
            hid = torch.empty((B * S, D), dtype=torch.float32, device=x.device)
            _lib.layernorm(x2d, lw, lb, hid, self.final_ln.eps, variant=1)
            hid = hid.view(B, S, D)
            return hid if pdt == torch.float32 else hid.to(pdt)
        # only the class-token rows (row 0 of every sample) feed the classifier
        cls_rows = x2d.view(B, S * D)[:, :D]
        h = torch.empty((B, D), dtype=torch.bfloat16, device=x.device)
        _lib.layernorm(cls_rows, lw, lb, h, self.final_ln.eps, variant=1)
        wc = _w_bf16(self._packed, "wc", self.classifier.weight)
        bc = _f32(self._packed, "bc", self.classifier.bias)
        logits = torch.empty((B, wc.shape[0]), dtype=torch.float32, device=x.device)
        _lib.gemm(h, wc, VF_EPI_BIAS_F32, logits, bias=bc)
        return logits if pdt == torch.float32 else logits.to(pdt)
