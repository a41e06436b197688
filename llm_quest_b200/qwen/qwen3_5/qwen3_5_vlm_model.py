"""Qwen3.5 early-fusion VLM front end on libvfuse kernels.

Drop-in for ``Qwen3_5VLM`` of the reference's ``llm_quest/qwen/qwen3_5/qwen3_5_vlm_model.py``
(:21-228) for everything that happens BEFORE the text model is called (:198-218):

    emb_dict(input_ids)                       -> vf_embed_gather_scatter (skips placeholder rows)
    vision_model(image_pixels)                -> the tower in qwen3_5_vision_model.py
    masked_scatter(image_mask, vision_embeds) -> merger-lin2 GEMM epilogue scatters its rows straight
                                                 into inputs_embs (row map from vf_fuse_scan)
    compute_3d_position_ids                   -> vf_mrope_position_ids (one kernel, no host loop/sync)

The text model itself (``Qwen3_5TextModel``) is out of scope and stays the reference's PyTorch
module: pass it as ``language_model=`` or let the constructor import it from ``llm_quest`` when the
reference package is importable. Without one, an embedding-table-only stand-in is built so that the
encode-and-fuse path is usable on its own (``encode_and_fuse``).
"""

from __future__ import annotations

import torch
import torch.nn as nn

from ... import _lib
from ..._lib import VFuseError
from .qwen3_5_vision_model import Qwen3_5VisionModel, _forward_only_guard


class EmbeddingOnlyLM(nn.Module):
    """Stand-in for the out-of-scope text model: owns ``emb_dict`` (same key/shape/dtype/init as
    qwen3_5_text_model.py:352,372) and nothing else."""

    def __init__(self, cfg):
        super().__init__()
        self.emb_dict = nn.Embedding(cfg["vocab_size"], cfg["emb_dim"], dtype=cfg.get("dtype", torch.bfloat16))
        nn.init.xavier_uniform_(self.emb_dict.weight)

    def forward(self, *args, **kwargs):
        raise VFuseError(
            "no text model attached: Qwen3_5VLM(cfg, language_model=...) needs the reference's "
            "Qwen3_5TextModel to produce logits; use encode_and_fuse() for the vision-encode-and-fuse path"
        )


class Qwen3_5VLM(nn.Module):
    def __init__(self, cfg, language_model: nn.Module | None = None):
        super().__init__()
        self.image_token_id = cfg.get("image_token_id", 248056)
        self.merge_size = cfg["spatial_merge_size"]
        self.cfg = cfg
        self.vision_model = Qwen3_5VisionModel(self.cfg)
        if language_model is None:
            try:  # drop-in use inside the reference repo
                from llm_quest.qwen.qwen3_5.qwen3_5_text_model import Qwen3_5TextModel  # type: ignore

                language_model = Qwen3_5TextModel(self.cfg)
            except ImportError:
                language_model = EmbeddingOnlyLM(self.cfg)
        self.language_model = language_model
        # vision-feature cache: the reference's generate loop re-encodes the same image for
        # every new token (qwen3_5_generate_multimodal.py:107-123); with the cache on, the tower runs once per
        # pixel tensor and later calls only redo the (cheap) gather/scatter + position ids
        self._vision_cache_on = False
        self._vision_cache = None   # dict: tower parameter versions, image identity, merged rows bf16 [n_vis, D]

    # -- reference API ----------------------------------------------------------------------------
    def get_feeds_3d_shape(self, image_pixels):
        """CPU int64 [[frames, nh, nw]] for 5-D pixels or 3-D pre-extracted patches (reference :46-83)."""
        nh = self.vision_model.n_height_patches
        nw = self.vision_model.n_width_patches
        if image_pixels.dim() == 5:
            frames = image_pixels.shape[2] // self.cfg["temporal_patch_size"]
        else:
            frames = image_pixels.shape[1] // (nh * nw)
        return torch.tensor([[frames, nh, nw]])

    def compute_3d_position_ids(self, input_ids, feeds_3d_shape=None, image_mask=None):
        """[3, b, seq] int64 MRoPE position ids, bit-identical to the reference (:85-176)."""
        b, seq_len = input_ids.shape
        if not input_ids.is_cuda:
            raise VFuseError("compute_3d_position_ids (llm_quest_b200) needs CUDA tensors; there is no CPU fallback")
        if feeds_3d_shape is None:
            # text-only: the same kernel with an all-False mask yields arange on all three axes
            feeds = torch.zeros((0, 3), dtype=torch.int64)
            mask = torch.zeros((b, seq_len), dtype=torch.uint8, device=input_ids.device)
            return _lib.mrope_position_ids(input_ids, mask, self.image_token_id, feeds, self.merge_size)
        return _lib.mrope_position_ids(input_ids, image_mask, self.image_token_id, feeds_3d_shape, self.merge_size)

    # -- vision-feature cache ------------------------------------------------------------------------
    def enable_vision_cache(self, on: bool = True):
        """Keep the merged vision embeddings of the last ``image_pixels`` tensor and reuse them while the SAME tensor
        object is passed again (or the same explicit ``image_id``) and the tower's weights are unchanged.
        Off by default: the reference recomputes every call.

        The key is object identity, not storage identity: the cache holds a reference to the pixel tensor, so its
        memory cannot be freed and handed to the next image by the caching allocator (a (data_ptr, shape) key would
        then silently serve the previous image's embeddings). Content changed IN PLACE through torch ops bumps
        ``_version`` and invalidates the entry; writes made by raw-pointer kernels (libvfuse's own, e.g.
        vf_preprocess_u8 into a caller-owned buffer) do not — pass a fresh ``image_id`` or call clear_vision_cache()."""
        self._vision_cache_on = bool(on)
        if not on:
            self._vision_cache = None
        return self

    def clear_vision_cache(self):
        self._vision_cache = None

    def _cached_vision_rows(self, image_pixels, n_vis, D, image_id=None):
        ver = lambda t: 0 if t.is_inference() else t._version
        params = tuple((prm.data_ptr(), ver(prm)) for prm in self.vision_model.parameters())
        hit = self._vision_cache
        if hit is not None and hit["params"] == params and hit["rows"].shape == (n_vis, D):
            if image_id is not None:
                same = hit["image_id"] == image_id
            else:
                same = hit["image_id"] is None and hit["pixels"] is image_pixels and hit["version"] == ver(image_pixels)
            if same:
                return hit["rows"]
        rows = torch.empty((n_vis, D), dtype=torch.bfloat16, device=image_pixels.device)
        vm = self.vision_model
        x2d, _, _ = vm.encode_hidden(image_pixels)
        vm.merge_adapter.merge_project(x2d, out=rows)
        self._vision_cache = {"params": params, "image_id": image_id, "pixels": image_pixels, "version": ver(image_pixels),
                              "rows": rows}
        return rows

    # -- the fused path ---------------------------------------------------------------------------
    def encode_and_fuse(self, input_ids, image_pixels=None, feeds_3d_shape=None, check=True, image_id=None):
        """Everything of ``forward`` before the text model: returns (inputs_embs bf16 [b, seq, D],
        position_ids int64 [3, b, seq], image_mask bool [b, seq] or None).

        check=True validates, like masked_scatter does, that the vision tower produced at least as
        many rows as there are placeholders (one 4-byte device->host read); check=False skips it.
        image_id (any hashable, optional): explicit identity of the image for the vision-feature cache.
        """
        table = self.language_model.emb_dict.weight
        if not input_ids.is_cuda:
            raise VFuseError("encode_and_fuse needs CUDA tensors (sm_100a); there is no CPU fallback")
        if table.dtype != torch.bfloat16:
            raise VFuseError(f"emb_dict must be bfloat16 (cfg['dtype']), got {table.dtype}")
        b, seq = input_ids.shape
        D = table.shape[1]
        inputs_embs = torch.empty((b, seq, D), dtype=torch.bfloat16, device=input_ids.device)
        image_mask = None
        if image_pixels is None:
            _lib.embed_gather_scatter(input_ids, table.detach(), None, None, inputs_embs)
        else:
            vm = self.vision_model
            Bv, _, T, H, W = image_pixels.shape
            n_vis = Bv * (T // self.cfg["temporal_patch_size"]) * vm.n_spatial_patches // (self.merge_size**2)
            # rank of every placeholder in flat (b, seq) order + its inverse (vision row -> token row)
            row_map, n_ph, dst = _lib.fuse_scan(input_ids, None, self.image_token_id, inv_cap=n_vis)
            if check:
                n = int(n_ph.item())
                if n > n_vis:
                    raise RuntimeError(
                        f"masked_scatter: {n} image placeholders in input_ids but the vision tower yields only {n_vis} rows"
                    )
            if self._vision_cache_on:
                # one kernel moves text rows from the table and vision rows from the cached embeddings
                rows = self._cached_vision_rows(image_pixels, n_vis, D, image_id)
                _lib.embed_gather_scatter(input_ids, table.detach(), rows, row_map, inputs_embs)
            else:
                # text rows: gathered from the table; placeholder rows are written by the merger GEMM epilogue
                _lib.embed_gather_scatter(input_ids, table.detach(), None, row_map, inputs_embs, skip_vision=True, n_vis=n_vis)
                vm(image_pixels, out=inputs_embs.view(b * seq, D), dst_rows=dst)
            image_mask = (row_map >= 0).view(b, seq)
            if feeds_3d_shape is None:
                feeds_3d_shape = self.get_feeds_3d_shape(image_pixels)
        # the kernel derives the mask from input_ids == image_token_id itself (same as reference :206)
        position_ids = self.compute_3d_position_ids(input_ids, feeds_3d_shape, image_mask=None)
        return inputs_embs, position_ids, image_mask

    def forward(self, input_ids, image_pixels=None, feeds_3d_shape=None, attn_mask=None):
        _forward_only_guard(self.vision_model)
        # like the reference (:215), a multimodal call always derives the single feed from the pixels
        feeds = feeds_3d_shape if image_pixels is None else None
        inputs_embs, position_ids, _ = self.encode_and_fuse(input_ids, image_pixels, feeds)
        return self.language_model(inputs_embs=inputs_embs, position_ids=position_ids, attn_mask=attn_mask)
