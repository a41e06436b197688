"""CPU oracle (numpy, integer/byte work) for early fusion: MRoPE position ids, scatter placement.

TEST INFRASTRUCTURE ONLY — see the header of ``oracle/vision_oracle.py`` for who may import this.
Every function is pinned against the live reference by ``oracle/make_golden.py`` (bit-exact) and
against the worked example in the reference docstring
(llm_quest/qwen/qwen3_5/qwen3_5_vlm_model.py:96-101) in ``tests/test_oracle.py``.
"""

from __future__ import annotations

import numpy as np


def mrope_position_ids(input_ids, feeds=None, image_mask=None, image_token_id=248056, merge=2):
    """[3, b, seq] int64 position ids.

    llm_quest/qwen/qwen3_5/qwen3_5_vlm_model.py:85-176. Per sample the placeholders are consumed by
    the feeds in order, t*(h/m)*(w/m) at a time, stopping at the first feed that does not fit
    (:148-149). A placeholder's id is the running text position plus its (frame,row,col) offset in
    its feed; the last placeholder of a feed advances the running position by max(t, h/m, w/m) (:154),
    every text token by 1, every other placeholder by 0. The running position is the exclusive
    cumulative sum of those increments (:171).
    """
    ids = np.asarray(input_ids, dtype=np.int64)
    b, seq = ids.shape
    if feeds is None:
        return np.broadcast_to(np.arange(seq, dtype=np.int64), (3, b, seq)).copy()
    feeds = np.asarray(feeds, dtype=np.int64).reshape(-1, 3)
    mask = (ids == image_token_id) if image_mask is None else np.asarray(image_mask, dtype=bool)
    inc = (~mask).astype(np.int64)
    local = np.zeros((3, b, seq), dtype=np.int64)
    for s in range(b):
        where = np.flatnonzero(mask[s])
        if where.size == 0:
            continue
        used = 0
        for t, h, w in feeds:
            hm, wm = h // merge, w // merge
            count = t * hm * wm
            if used + count > where.size:
                break
            spots = where[used : used + count]
            k = np.arange(count, dtype=np.int64)
            inc[s, spots[-1]] = max(t, hm, wm)
            local[0, s, spots] = k // (hm * wm)
            local[1, s, spots] = (k % (hm * wm)) // wm
            local[2, s, spots] = k % wm
            used += count
    running = np.cumsum(inc, axis=1) - inc
    return running[None] + local


def scatter_row_map(input_ids, image_mask=None, image_token_id=248056):
    """Flat [b*seq] int32 map: j for the j-th placeholder in row-major (b, seq) order, -1 elsewhere.

    This is the placement torch.Tensor.masked_scatter performs at vlm_model.py:209-211 (the source
    rows are consumed in flat order of the True mask positions).
    """
    ids = np.asarray(input_ids, dtype=np.int64)
    mask = (ids == image_token_id) if image_mask is None else np.asarray(image_mask, dtype=bool)
    flat = mask.reshape(-1)
    out = np.full(flat.shape, -1, dtype=np.int32)
    out[flat] = np.arange(int(flat.sum()), dtype=np.int32)
    return out


def fuse_embeddings(input_ids, table_u16, vision_u16, image_mask=None, image_token_id=248056):
    """Early fusion on raw bf16 bit patterns (uint16): gather table rows, overwrite placeholder rows.

    vlm_model.py:198-211. Returns [b, seq, D] uint16.
    """
    ids = np.asarray(input_ids, dtype=np.int64)
    b, seq = ids.shape
    rows = scatter_row_map(ids, image_mask, image_token_id)
    vis = np.asarray(vision_u16).reshape(-1, table_u16.shape[1])
    n_true = int((rows >= 0).sum())
    if n_true > vis.shape[0]:
        raise ValueError(f"masked_scatter: {n_true} placeholders but only {vis.shape[0]} vision rows")
    out = table_u16[ids.reshape(-1)].copy()
    sel = rows >= 0
    out[sel] = vis[rows[sel]]
    return out.reshape(b, seq, -1)


def feeds_3d_shape(pixel_shape, n_height_patches, n_width_patches, temporal_patch_size):
    """[[frames, nh, nw]] — vlm_model.py:46-83 (5-D pixels or 3-D pre-extracted patches)."""
    if len(pixel_shape) == 5:
        frames = pixel_shape[2] // temporal_patch_size
    else:
        frames = pixel_shape[1] // (n_height_patches * n_width_patches)
    return np.array([[frames, n_height_patches, n_width_patches]], dtype=np.int64)
