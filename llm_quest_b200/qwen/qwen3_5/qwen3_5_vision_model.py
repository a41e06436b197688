"""Qwen3.5 Qwen3-ViT tower + spatial-merge adapter on libvfuse (sm_100a) kernels.

Drop-in for the classes of the reference's ``llm_quest/qwen/qwen3_5/qwen3_5_vision_model.py``:
same class names, constructor arguments, ``forward`` signatures, ``state_dict`` keys/shapes and
assertion messages. Parameters live in the same ``nn`` containers, created in the
same order, so ``torch.manual_seed(s); Model(cfg)`` yields the same weights as the reference.

What differs is everything underneath ``forward``: activations stay token-major ``[B*S, D]`` for
the whole tower, the residual stream is fp32, GEMM/attention operands are bf16 and every op is one
hand-written kernel (csrc/):

    pixels -> [vf_patch_embed: 5-D TMA gather GEMM + bias + pos-embed]            -> x   fp32
    block 0:    x -> [vf_layernorm (+ row means = first shift)] -> h bf16
    per block:  h -> [vf_gemm QKV + bias + axial RoPE (+ folded norm1)]            -> qkv bf16
                qkv -> [vf_attention_fwd]                                          -> a   bf16
                a -> [vf_gemm proj + bias + residual (+ bf16 copy, row partial sums)]  -> x fp32, h bf16
                [vf_ln_row_stats] -> [vf_gemm lin1 + bias + tanh-GELU (+ folded norm2)] -> g bf16
                g -> [vf_gemm lin2 + bias + residual (+ bf16 copy, row partial sums)]  -> x fp32, h bf16 ; [vf_ln_row_stats]
    merger:     x -> [vf_layernorm + 2x2 merge gather] -> [lin1 + erf-GELU] -> [lin2 + bias] -> out

Forward only (the reference's callers run it under inference_mode/no_grad).
"""

from __future__ import annotations

import torch
import torch.nn as nn

from ... import _lib
from ...common.rope import VisionRoPE
from ..._lib import (VF_EPI_BIAS_BF16, VF_EPI_BIAS_F32, VF_EPI_BIAS_RES_F32, VF_EPI_GELU_ERF_BF16,
                     VF_EPI_GELU_TANH_BF16, VF_EPI_QKV_ROPE_BF16, VF_EPI_SCATTER_BF16, VFuseError)


def _forward_only_guard(module: nn.Module) -> None:
    if module.training and torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters()):
        raise RuntimeError(
            f"{type(module).__name__} (llm_quest_b200) is forward-only: call .eval() or run under "
            "torch.no_grad()/inference_mode(); the training loops of the reference are out of scope"
        )


class _Packed:
    """bf16 / fp32-contiguous copies of parameters for the kernels, rebuilt when a parameter changes
    (load_state_dict, .to(), in-place edits bump ``_version`` / move ``data_ptr``)."""

    def __init__(self):
        self._store = {}

    def get(self, key, params, build):
        # inference tensors carry no version counter (and cannot be modified in place anyway)
        sig = tuple((p.data_ptr(), 0 if p.is_inference() else p._version, p.device, p.dtype)
                    for p in params if p is not None)
        hit = self._store.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        with torch.no_grad():
            val = build()
        self._store[key] = (sig, val)
        return val


def _w_bf16(cache: _Packed, key, weight, shape2d=None):
    def build():
        w = weight.detach()
        if shape2d is not None:
            w = w.reshape(shape2d)
        return w.to(torch.bfloat16).contiguous()

    return cache.get(key, [weight], build)


def _f32(cache: _Packed, key, p):
    if p is None:
        return None
    return cache.get(key, [p], lambda: p.detach().to(torch.float32).contiguous())


def _fold_ln(cache: _Packed, key, lin: nn.Linear, norm):
    """LayerNorm folded into the Linear that follows it:
        LN(x) @ W^T + b = rstd * (x @ (gamma . W)^T - mean * colsum) + (b + W beta)
    Returns (bf16 gamma-scaled weight, fp32 folded bias, fp32 colsum of the ROUNDED weight — the GEMM multiplies by
    the rounded one, so the mean term must cancel against exactly that). `norm`: nn.LayerNorm (weight / bias) or the
    Part-1 LayerNorm (scale / shift)."""
    gamma = norm.weight if hasattr(norm, "weight") else norm.scale
    beta = norm.bias if hasattr(norm, "bias") else norm.shift

    def build():
        w = lin.weight.detach().to(torch.float32)
        wf = (w * gamma.detach().to(torch.float32)[None, :]).to(torch.bfloat16).contiguous()
        colsum = wf.to(torch.float32).sum(dim=1).contiguous()
        b = w @ beta.detach().to(torch.float32)
        if lin.bias is not None:
            b = b + lin.bias.detach().to(torch.float32)
        return wf, b.contiguous(), colsum

    return cache.get(key, [lin.weight, lin.bias, gamma, beta], build)


# Up to this many rows the consuming GEMM adds up the producer's partial row sums itself (vf_epilogue.ln_part_in) and the
# vf_ln_row_stats launch in between is dropped: at small batches the launch costs more than the re-read of the partials by
# every column tile (batch 1-4 of the Qwen tower, the Part-1 ViT at batch 8: 1576 rows). Same bits either way.
LN_STATS_IN_CONSUMER_MAX_ROWS = 4096


def _as_2d_bf16(x: torch.Tensor) -> torch.Tensor:
    return _lib.to_bf16(x.reshape(-1, x.shape[-1]))


class PatchEmbedding3D(nn.Module):
    """Video/image tensor -> patch embeddings (reference :47-109): non-overlapping conv3d expressed
    as an im2col-free GEMM whose A operand is gathered by TMA straight from the pixel tensor."""

    def __init__(self, img_width, img_height, num_channels, emb_dim, patch_size, temporal_patch_size):
        super().__init__()
        assert img_width % patch_size == 0, f"Image width {img_width} not divisible by patch size {patch_size}"
        assert img_height % patch_size == 0, f"Image height {img_height} not divisible by patch size {patch_size}"
        self.img_width = img_width
        self.img_height = img_height
        self.patch_size = patch_size
        self.temporal_patch_size = temporal_patch_size
        self.num_patches_per_image = (img_width * img_height) // patch_size**2
        k = (temporal_patch_size, patch_size, patch_size)
        self.conv_proj = nn.Conv3d(num_channels, emb_dim, kernel_size=k, stride=k, padding=0, bias=True)
        self._packed = _Packed()

    def _check(self, x):
        b, n_channels, time, img_h, img_w = x.shape
        assert img_h == self.img_height and img_w == self.img_width, (
            f"Input image shape {x.shape} does not match expected shape {self.img_height}x{self.img_width}"
        )
        assert time % self.temporal_patch_size == 0, (
            f"Input time shape {time} is not divisible by temporal_patch_size {self.temporal_patch_size}"
        )

    def embed_into(self, x, pos=None):
        """fp32 [B*S, D] = conv(x) + bias (+ pos[token % n]); the fused entry the tower uses."""
        self._check(x)
        B, _, T, H, W = x.shape
        P, tp = self.patch_size, self.temporal_patch_size
        if P != 16:
            raise VFuseError(f"PatchEmbedding3D (llm_quest_b200): the TMA patch gather is built for 16x16 patches, got {P}")
        S = (T // tp) * (H // P) * (W // P)
        D = self.conv_proj.out_channels
        w = _w_bf16(self._packed, "w", self.conv_proj.weight, (D, -1))
        bias = _f32(self._packed, "b", self.conv_proj.bias)
        out = torch.empty((B * S, D), dtype=torch.float32, device=x.device)
        _lib.patch_embed(_lib.to_bf16(x), w, bias, pos, out, P, tp, S, 0)
        return out, B, S

    def forward(self, x):
        _forward_only_guard(self)
        out, B, S = self.embed_into(x)
        return out.view(B, S, -1).to(self.conv_proj.weight.dtype)


class Qwen3_5VisionFFN(nn.Module):
    """lin1 -> tanh-GELU -> lin2 (reference :112-125); the GELU lives in lin1's GEMM epilogue."""

    def __init__(self, cfg):
        super().__init__()
        self.lin1 = nn.Linear(cfg["vision_emb_dim"], cfg["vision_hidden_dim"])
        self.lin2 = nn.Linear(cfg["vision_hidden_dim"], cfg["vision_emb_dim"])
        self.activ = nn.GELU(approximate="tanh")
        self._packed = _Packed()

    def packed(self):
        c = self._packed
        return (_w_bf16(c, "w1", self.lin1.weight), _f32(c, "b1", self.lin1.bias),
                _w_bf16(c, "w2", self.lin2.weight), _f32(c, "b2", self.lin2.bias))

    def forward(self, x):
        _forward_only_guard(self)
        w1, b1, w2, b2 = self.packed()
        h = _as_2d_bf16(x)
        g = torch.empty((h.shape[0], w1.shape[0]), dtype=torch.bfloat16, device=x.device)
        _lib.gemm(h, w1, VF_EPI_GELU_TANH_BF16, g, bias=b1)
        out = torch.empty((h.shape[0], w2.shape[0]), dtype=torch.float32, device=x.device)
        _lib.gemm(g, w2, VF_EPI_BIAS_F32, out, bias=b2)
        return out.view(*x.shape[:-1], -1).to(x.dtype)


class Qwen3_5VisionAttention(nn.Module):
    """Bidirectional MHA with axial 2-D RoPE (reference :128-192). QKV GEMM applies bias + RoPE in its
    epilogue; the attention kernel reads q/k/v in place from the token-major qkv buffer."""

    def __init__(self, cfg):
        super().__init__()
        self.d_in = cfg["vision_emb_dim"]
        self.num_heads = cfg["vision_num_heads"]
        self.head_dim = self.d_in // self.num_heads
        self.qkv = nn.Linear(self.d_in, self.d_in * 3, bias=True)
        self.proj = nn.Linear(self.d_in, self.d_in, bias=True)
        self._packed = _Packed()

    def packed(self):
        c = self._packed
        return (_w_bf16(c, "wqkv", self.qkv.weight), _f32(c, "bqkv", self.qkv.bias),
                _w_bf16(c, "wo", self.proj.weight), _f32(c, "bo", self.proj.bias))

    def _require_hd64(self):
        if self.head_dim != 64:
            raise VFuseError(f"the fused QKV+RoPE epilogue and the tcgen05 attention are built for head_dim 64, got {self.head_dim}")

    def attend(self, h2d, B, S, rope, folded=None, ln_in=None):
        """h2d bf16 [B*S, D] -> context bf16 [B*S, D]. rope = (cos_half, sin_half, period).
        folded/ln_in: norm1 folded into the QKV GEMM — h2d is then the bf16 copy of the shifted, un-normalised stream."""
        self._require_hd64()
        wqkv, bqkv, _, _ = self.packed()
        if folded is not None:
            wqkv, bqkv = folded
        D = self.d_in
        qkv = torch.empty((B * S, 3 * D), dtype=torch.bfloat16, device=h2d.device)
        _lib.gemm(h2d, wqkv, VF_EPI_QKV_ROPE_BF16, qkv, bias=bqkv, rope=(rope[0], rope[1], rope[2], 2 * D), ln_in=ln_in)
        ctx = torch.empty((B * S, D), dtype=torch.bfloat16, device=h2d.device)
        _lib.attention(qkv, ctx, B, S, self.num_heads, self.head_dim**-0.5)
        return ctx

    def forward(self, x, cos, sin):
        _forward_only_guard(self)
        b, seq_len, d_in = x.shape
        half = self.head_dim // 2
        cos_h = cos[:seq_len, :half].to(device=x.device, dtype=torch.float32).contiguous()
        sin_h = sin[:seq_len, :half].to(device=x.device, dtype=torch.float32).contiguous()
        ctx = self.attend(_as_2d_bf16(x), b, seq_len, (cos_h, sin_h, seq_len))
        _, _, wo, bo = self.packed()
        out = torch.empty((b * seq_len, d_in), dtype=torch.float32, device=x.device)
        _lib.gemm(ctx, wo, VF_EPI_BIAS_F32, out, bias=bo)
        return out.view(b, seq_len, d_in).to(x.dtype)


class Qwen3_5VisionTransformerBlock(nn.Module):
    """Pre-LN block: x += att(LN1(x)); x += ffn(LN2(x)) (reference :195-238)."""

    def __init__(self, cfg):
        super().__init__()
        self.norm1 = nn.LayerNorm(cfg["vision_emb_dim"], eps=1e-6)
        self.norm2 = nn.LayerNorm(cfg["vision_emb_dim"], eps=1e-6)
        self.att = Qwen3_5VisionAttention(cfg)
        self.ffn = Qwen3_5VisionFFN(cfg)
        self._packed = _Packed()

    def run_(self, x2d, B, S, rope, work, ln1_pending=False, emit_next=False):
        """In-place update of the fp32 residual stream x2d [B*S, D]; `work` holds reusable buffers.

        With work["stat"] present the LayerNorms are folded into the GEMMs around them (vf_epilogue.ln_*): the GEMM that
        produces x also writes bf16(x - shift) to work["h"] and per-row partial sums of the shifted row to work["stat"];
        vf_ln_row_stats turns them into (mean', rstd) in work["rows"] and advances work["shift"] to the row's mean; the
        GEMM that consumes LN(x) multiplies the bf16 copy by the gamma-scaled weight and normalises in its epilogue.
        ln1_pending: the previous block's lin2 left h/stat for this block's norm1; emit_next: this block's lin2 leaves
        them for the next block. The first LayerNorm of a chain is a stand-alone kernel on the fp32 stream (it also
        yields the first shift), so a row's mean never meets the bf16 rounding."""
        c = self._packed
        _, _, wo, bo = self.att.packed()
        w1, b1, w2, b2 = self.ffn.packed()
        h, g, stat, rows, shift = work["h"], work["g"], work.get("stat"), work.get("rows"), work.get("shift")
        D = x2d.shape[1]
        fold = stat is not None
        small = x2d.shape[0] <= LN_STATS_IN_CONSUMER_MAX_ROWS

        def consumer(eps, colsum):
            if small:
                return (stat, colsum, eps, 0, shift)
            _lib.ln_row_stats(stat, D, eps, rows, shift)
            return (rows, colsum)

        if fold and ln1_pending:
            wq, bq, csq = _fold_ln(c, "fold_qkv", self.att.qkv, self.norm1)
            ctx = self.att.attend(h, B, S, rope, folded=(wq, bq), ln_in=consumer(self.norm1.eps, csq))
        else:
            n1w, n1b = _f32(c, "n1w", self.norm1.weight), _f32(c, "n1b", self.norm1.bias)
            _lib.layernorm(x2d, n1w, n1b, h, self.norm1.eps, mean_out=shift if fold else None)
            ctx = self.att.attend(h, B, S, rope)
        producer = (h, stat, shift) if fold else None
        if fold and work.get("fold_norm2"):
            w1f, b1f, cs1 = _fold_ln(c, "fold_lin1", self.ffn.lin1, self.norm2)
            _lib.gemm(ctx, wo, VF_EPI_BIAS_RES_F32, x2d, bias=bo, res=x2d, ln_out=producer)
            _lib.gemm(h, w1f, VF_EPI_GELU_TANH_BF16, g, bias=b1f, ln_in=consumer(self.norm2.eps, cs1))
        else:
            n2w, n2b = _f32(c, "n2w", self.norm2.weight), _f32(c, "n2b", self.norm2.bias)
            _lib.gemm(ctx, wo, VF_EPI_BIAS_RES_F32, x2d, bias=bo, res=x2d)
            _lib.layernorm(x2d, n2w, n2b, h, self.norm2.eps, mean_out=shift if fold else None)
            _lib.gemm(h, w1, VF_EPI_GELU_TANH_BF16, g, bias=b1)
        _lib.gemm(g, w2, VF_EPI_BIAS_RES_F32, x2d, bias=b2, res=x2d, ln_out=producer if emit_next else None)

    def forward(self, x, cos, sin):
        _forward_only_guard(self)
        b, seq_len, d = x.shape
        half = self.att.head_dim // 2
        cos_h = cos[:seq_len, :half].to(device=x.device, dtype=torch.float32).contiguous()
        sin_h = sin[:seq_len, :half].to(device=x.device, dtype=torch.float32).contiguous()
        x2d = _lib.to_f32(x.reshape(-1, d)).clone() if x.dtype == torch.float32 else _lib.to_f32(x.reshape(-1, d))
        work = {
            "h": torch.empty((b * seq_len, d), dtype=torch.bfloat16, device=x.device),
            "g": torch.empty((b * seq_len, self.ffn.lin1.out_features), dtype=torch.bfloat16, device=x.device),
        }
        self.run_(x2d, b, seq_len, (cos_h, sin_h, seq_len), work)
        return x2d.view(b, seq_len, d).to(x.dtype)


class ViTMergeAdapter(nn.Module):
    """LayerNorm -> 2x2 spatial merge -> lin1 -> erf-GELU -> lin2 (reference :373-431). The merge
    gather is fused into the LayerNorm store; GELU and biases live in the GEMM epilogues."""

    def __init__(self, vit_d_out, llm_d_in, n_height_patches, n_width_patches, spatial_merge_size=2):
        super().__init__()
        self.m = spatial_merge_size
        self.n_h_patches = n_height_patches
        self.n_w_patches = n_width_patches
        self.merged_size = vit_d_out * (self.m**2)
        self.norm = nn.LayerNorm(vit_d_out, eps=1e-6)
        self.lin1 = nn.Linear(self.merged_size, self.merged_size)
        self.activ = nn.GELU()
        self.lin2 = nn.Linear(self.merged_size, llm_d_in)
        self._packed = _Packed()

    def merge_project(self, x2d, out=None, dst_rows=None, peer_ptrs=None, peer_multicast=False):
        """x2d fp32/bf16 [B*S, D] -> [B*S/m^2, llm_d_in]. With `dst_rows` (int32 per merged row) the
        last GEMM scatters bf16 rows straight into `out` (the fused text sequence); with `peer_ptrs` it stores
        them into the gathered buffer of every rank (fused all-gather, see parallel.FusedAllGather)."""
        c = self._packed
        nw_, nb_ = _f32(c, "nw", self.norm.weight), _f32(c, "nb", self.norm.bias)
        w1, b1 = _w_bf16(c, "w1", self.lin1.weight), _f32(c, "b1", self.lin1.bias)
        w2, b2 = _w_bf16(c, "w2", self.lin2.weight), _f32(c, "b2", self.lin2.bias)
        rows, D = x2d.shape
        mm = self.m * self.m
        assert rows % (self.n_h_patches * self.n_w_patches) == 0
        hm = torch.empty((rows // mm, mm * D), dtype=torch.bfloat16, device=x2d.device)
        _lib.layernorm(x2d, nw_, nb_, hm, self.norm.eps, 0, self.m, self.n_h_patches, self.n_w_patches)
        g = torch.empty_like(hm)
        _lib.gemm(hm, w1, VF_EPI_GELU_ERF_BF16, g, bias=b1)
        if dst_rows is not None:
            _lib.gemm(g, w2, VF_EPI_SCATTER_BF16, out, bias=b2, dst_rows=dst_rows)
            return out
        if out is None:
            out = torch.empty((rows // mm, w2.shape[0]), dtype=torch.float32, device=x2d.device)
        mode = VF_EPI_BIAS_F32 if out.dtype == torch.float32 else VF_EPI_BIAS_BF16
        _lib.gemm(g, w2, mode, out, bias=b2, peer_ptrs=peer_ptrs, peer_multicast=peer_multicast)
        return out

    def forward(self, x):
        _forward_only_guard(self)
        b, n_patches, vit_d_out = x.shape
        x2d = x.reshape(-1, vit_d_out)
        if x2d.dtype not in (torch.float32, torch.bfloat16):
            x2d = x2d.float()
        out = self.merge_project(x2d.contiguous())
        return out.view(b, n_patches // (self.m**2), -1).to(x.dtype)


class Qwen3_5VisionModel(nn.Module):
    """Complete vision tower (reference :241-370): patch embed + pos-embed, N pre-LN blocks with axial
    2-D RoPE attention, merge adapter. Output [B, T'*n/4, llm_d_in] in the parameters' dtype."""

    def __init__(self, cfg):
        super().__init__()
        emb_dim = cfg["vision_emb_dim"]
        n_layers = cfg["vision_n_layers"]
        num_heads = cfg["vision_num_heads"]
        rope_base = cfg["vision_rope_base"]
        llm_d_in = cfg["llm_d_in"]
        img_width = cfg["img_width"]
        img_height = cfg["img_height"]
        patch_size = cfg["patch_size"]

        assert img_width % patch_size == 0, f"Image width {img_width} not divisible by patch size {patch_size}"
        assert img_height % patch_size == 0, f"Image height {img_height} not divisible by patch size {patch_size}"
        self.n_width_patches = img_width // patch_size
        self.n_height_patches = img_height // patch_size
        self.n_spatial_patches = self.n_width_patches * self.n_height_patches
        assert self.n_spatial_patches <= cfg["num_position_embeddings"], (
            f"the image size {img_width}x{img_height} "
            f"is too large for the number of position embeddings {cfg['num_position_embeddings']}"
        )

        self.patch_embed = PatchEmbedding3D(
            img_width=img_width, img_height=img_height, num_channels=cfg["in_channels"], emb_dim=emb_dim,
            patch_size=patch_size, temporal_patch_size=cfg["temporal_patch_size"],
        )
        self.pos_embed = nn.Embedding(cfg["num_position_embeddings"], emb_dim)
        cos, sin = VisionRoPE.compute_angles_2d(
            base=rope_base, head_dim=emb_dim // num_heads, height_patches=self.n_height_patches,
            width_patches=self.n_width_patches, num_frames=1,
        )
        self.register_buffer("cos", cos, persistent=False)
        self.register_buffer("sin", sin, persistent=False)
        self.blocks = nn.ModuleList([Qwen3_5VisionTransformerBlock(cfg=cfg) for _ in range(n_layers)])
        self.merge_adapter = ViTMergeAdapter(
            vit_d_out=emb_dim, llm_d_in=llm_d_in, n_height_patches=self.n_height_patches,
            n_width_patches=self.n_width_patches, spatial_merge_size=cfg["spatial_merge_size"],
        )
        self._packed = _Packed()

    # -- fused pipeline ---------------------------------------------------------------------------
    def _rope_half(self, device):
        half = self.cos.shape[-1] // 2

        def build():
            return (self.cos[:, :half].to(device=device, dtype=torch.float32).contiguous(),
                    self.sin[:, :half].to(device=device, dtype=torch.float32).contiguous())

        return self._packed.get(("rope", str(device)), [self.cos, self.sin], build)

    # How much LayerNorm is folded into the GEMMs of a block: 2 = norm1 and norm2 (default), 1 = norm1 only, 0 = none
    # (stand-alone vf_layernorm launches). Plain attribute: set it on an instance (tests, A/B measurements).
    ln_fold = 2

    def encode_hidden(self, x):
        """pixels [B,C,T,H,W] -> (fp32 residual stream [B*S, D] after the last block, B, S)."""
        if not x.is_cuda:
            raise VFuseError("Qwen3_5VisionModel (llm_quest_b200) runs on CUDA sm_100a only; got a CPU tensor")
        pos = _f32(self._packed, "pos", self.pos_embed.weight)
        x2d, B, S = self.patch_embed.embed_into(x, pos)
        cos_h, sin_h = self._rope_half(x.device)
        rope = (cos_h, sin_h, self.n_spatial_patches)
        rows, D = x2d.shape
        work = {"h": torch.empty((rows, D), dtype=torch.bfloat16, device=x.device), "fold_norm2": self.ln_fold > 1}
        if self.ln_fold > 0 and len(self.blocks) > 0 and D % 32 == 0:
            work["stat"] = torch.empty((D // 32, rows, 2), dtype=torch.float32, device=x.device)
            work["rows"] = torch.empty((rows, 2), dtype=torch.float32, device=x.device)
            work["shift"] = torch.empty((rows,), dtype=torch.float32, device=x.device)
        if len(self.blocks):
            work["g"] = torch.empty((rows, self.blocks[0].ffn.lin1.out_features), dtype=torch.bfloat16, device=x.device)
        last = len(self.blocks) - 1
        for i, block in enumerate(self.blocks):
            block.run_(x2d, B, S, rope, work, ln1_pending=i > 0, emit_next=i < last)
        return x2d, B, S

    def forward(self, x, out=None, dst_rows=None, gather=None):
        """x: [B, C, T, H, W] pixels. Returns [B, num_merged_patches, llm_d_in].

        `out`/`dst_rows` (library extension, used by Qwen3_5VLM): scatter the merged rows as bf16
        directly into the fused text-embedding buffer instead of returning them.
        `gather` (library extension, sample-sharded multi-GPU runs): a parallel.FusedAllGather — the last GEMM stores
        this rank's rows into every rank's gathered buffer; returns [world*B, num_merged_patches, llm_d_in] bf16.
        """
        _forward_only_guard(self)
        x2d, B, S = self.encode_hidden(x)
        if gather is not None:
            slot = gather.next_slot()
            self.merge_adapter.merge_project(x2d, out=gather.local_rows(slot), peer_ptrs=gather.peer_ptrs(slot),
                                             peer_multicast=bool(gather.multicast_ptr))
            gather.barrier()
            return gather.gathered(slot).view(gather.world * B, -1, gather.cols)
        if dst_rows is not None:
            return self.merge_adapter.merge_project(x2d, out=out, dst_rows=dst_rows)
        merged = self.merge_adapter.merge_project(x2d)
        merged = merged.view(B, -1, merged.shape[-1])
        pd = self.pos_embed.weight.dtype
        return merged if pd == torch.float32 else merged.to(pd)
