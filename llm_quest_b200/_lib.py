"""ctypes binding of libvfuse.so (include/vfuse.h) plus thin tensor-level wrappers.

The library is the product; this file only marshals ``torch.Tensor`` storage pointers, shapes and
the current CUDA stream into the C ABI. There is deliberately NO fallback: if the shared library is
missing, or a call fails, an exception is raised (``VFuseError``) — nothing silently reverts to
PyTorch ops.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
_LIB_PATH = Path(os.environ.get("VFUSE_LIB", _PKG / "libvfuse.so"))

VF_EPI_BIAS_BF16 = 0
VF_EPI_BIAS_F32 = 1
VF_EPI_BIAS_RES_F32 = 2
VF_EPI_GELU_TANH_BF16 = 3
VF_EPI_GELU_ERF_BF16 = 4
VF_EPI_QKV_ROPE_BF16 = 5
VF_EPI_SCATTER_BF16 = 6

EXPORTS = [
    "vf_version", "vf_last_error", "vf_launch_count", "vf_launch_count_reset", "vf_gemm_bf16", "vf_gemm_set_debug",
    "vf_patch_embed", "vf_attention_fwd", "vf_attention_fwd_hd", "vf_attention_set_trace", "vf_attention_gqa_fwd", "vf_layernorm", "vf_ln_row_stats", "vf_vit_cls_pos", "vf_rope_apply",
    "vf_mrope_apply", "vf_mrope_apply_strided", "vf_mrope_position_ids", "vf_fuse_scan", "vf_embed_gather_scatter",
    "vf_cast_f32_to_bf16", "vf_cast_bf16_to_f32", "vf_preprocess_u8",
    "vf_gelu", "vf_rmsnorm_zc", "vf_embed_pos_concat", "vf_im2col_patches", "vf_fill_rows_f32",
]


class VFuseError(RuntimeError):
    """A libvfuse entry point returned a non-zero status (or the library is missing)."""


class vf_epilogue(C.Structure):
    _fields_ = [
        ("mode", C.c_int32),
        ("bias", C.c_void_p),
        ("out", C.c_void_p),
        ("ldo", C.c_int64),
        ("res", C.c_void_p),
        ("ldr", C.c_int64),
        ("grp_rows", C.c_int32),
        ("grp_stride", C.c_int64),
        ("row_off", C.c_int64),
        ("rope_cos", C.c_void_p),
        ("rope_sin", C.c_void_p),
        ("rope_period", C.c_int32),
        ("rope_cols", C.c_int32),
        ("dst_rows", C.c_void_p),
        ("n_peers", C.c_int32),
        ("peer_out", C.c_void_p * 8),
        ("peer_multicast", C.c_int32),
        ("ln_xb_out", C.c_void_p),
        ("ln_ldxb", C.c_int64),
        ("ln_stat_out", C.c_void_p),
        ("ln_stat_ld", C.c_int64),
        ("ln_shift", C.c_void_p),
        ("ln_row_stats", C.c_void_p),
        ("ln_colsum", C.c_void_p),
        ("ln_part_in", C.c_void_p),
        ("ln_shift_update", C.c_void_p),
        ("ln_eps", C.c_float),
        ("ln_variant", C.c_int32),
    ]


_lib = None


def lib() -> C.CDLL:
    """Load libvfuse.so once; raise VFuseError (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise VFuseError(
            f"{_LIB_PATH} not found: build it with `python -m llm_quest_b200.build` "
            "(there is no PyTorch/CPU fallback for the vision-encode-and-fuse kernels)"
        )
    L = C.CDLL(str(_LIB_PATH))
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    L.vf_version.restype = C.c_int
    L.vf_last_error.restype = C.c_char_p
    L.vf_launch_count.restype = C.c_int64
    L.vf_launch_count_reset.restype = None
    sigs = {
        "vf_gemm_bf16": [vp, i64, vp, i64, i32, i32, i32, C.POINTER(vf_epilogue), vp],
        "vf_gemm_set_debug": [vp],
        "vf_patch_embed": [vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, i64, i32, vp, i64, i64, i64, vp],
        "vf_attention_fwd": [vp, vp, i32, i32, i32, f32, vp],
        "vf_attention_fwd_hd": [vp, vp, i32, i32, i32, i32, f32, vp],
        "vf_attention_set_trace": [vp, i32, i32],
        "vf_attention_gqa_fwd": [vp, i64, i32, i32, vp, i64, vp, i64, vp, i64, vp, i64, i32, i32, i32, i32, i32, i32, i32,
                                 f32, i32, vp],
        "vf_layernorm": [vp, i32, i64, vp, vp, vp, i32, i64, i32, f32, i32, i32, i32, i32, vp, vp],
        "vf_ln_row_stats": [vp, i32, i64, i64, i32, f32, i32, vp, vp, vp],
        "vf_vit_cls_pos": [vp, vp, vp, i32, i64, i32, vp],
        "vf_rope_apply": [vp, vp, i32, i32, i32, i32, i32, vp, vp, i32, i64, vp, vp],
        "vf_mrope_apply": [vp, vp, i32, i32, i32, i32, i32, vp, vp, i32, i64, vp, i32, i32, i32, vp, f32, vp],
        "vf_mrope_apply_strided": [vp, vp, i32, i32, i32, i32, i32, C.POINTER(C.c_int64), C.POINTER(C.c_int64), vp, vp, i32,
                                   i64, vp, i32, i32, i32, vp, f32, vp],
        "vf_mrope_position_ids": [vp, vp, i64, vp, i32, i32, i32, i32, vp, vp],
        "vf_fuse_scan": [vp, vp, i64, i64, vp, vp, vp, i64, vp, vp],
        "vf_embed_gather_scatter": [vp, vp, i64, i32, vp, i32, i64, vp, vp, i64, i32, vp],
        "vf_cast_f32_to_bf16": [vp, vp, i64, vp],
        "vf_cast_bf16_to_f32": [vp, vp, i64, vp],
        "vf_preprocess_u8": [vp, i32, i32, i32, i32, C.POINTER(C.c_float), C.POINTER(C.c_float), vp, i32, vp],
        "vf_gelu": [vp, vp, i32, i64, i32, vp],
        "vf_rmsnorm_zc": [vp, i64, vp, vp, i64, i32, i64, i32, f32, vp],
        "vf_embed_pos_concat": [vp, vp, i64, vp, i64, i32, vp, i32, i32, i32, i64, i64, vp],
        "vf_im2col_patches": [vp, i32, i32, i32, i32, i32, i32, vp, i64, vp],
        "vf_fill_rows_f32": [vp, vp, vp, i32, i64, i32, vp],
    }
    for name, args in sigs.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().vf_last_error().decode(errors="replace")
        raise VFuseError(f"{what} failed with status {rc}: {msg}")


def launch_count() -> int:
    return int(lib().vf_launch_count())


def reset_launch_count() -> None:
    lib().vf_launch_count_reset()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t) -> int | None:
    return None if t is None else t.data_ptr()


def _require_cuda(*ts) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise VFuseError(
                "libvfuse kernels need CUDA tensors (sm_100a); got a CPU tensor and there is no CPU fallback"
            )


_DT = {torch.float32: 0, torch.bfloat16: 1}


class KernelTimer:
    """Optional per-launch CUDA-event timing (bench.py): `with KernelTimer() as kt: ...` records an
    event pair on the launching stream around every libvfuse launch made inside the block."""

    active = None

    def __init__(self, external: bool = False):
        # external=True: events that may be recorded inside a CUDA-graph capture (they become event-record nodes; after a
        # replay elapsed_time() gives the kernel's duration INSIDE the back-to-back step, at the step's clocks and cache state)
        self.external = external
        self.records = []  # (kernel family, meta dict, start event, end event)

    def __enter__(self):
        KernelTimer.active = self
        return self

    def __exit__(self, *exc):
        KernelTimer.active = None

    def summary(self):
        """{family: {"launches", "ms_total", "ms_avg", "flops", "bytes"}} — call after a synchronize."""
        out = {}
        for fam, meta, e0, e1 in self.records:
            d = out.setdefault(fam, {"launches": 0, "ms_total": 0.0, "flops": 0.0, "bytes": 0.0})
            d["launches"] += 1
            d["ms_total"] += e0.elapsed_time(e1)
            d["flops"] += meta.get("flops", 0.0)
            d["bytes"] += meta.get("bytes", 0.0)
        for d in out.values():
            d["ms_avg"] = d["ms_total"] / d["launches"]
        return out


class _timed:
    def __init__(self, family, **meta):
        self.family, self.meta = family, meta
        self.kt = KernelTimer.active

    def __enter__(self):
        if self.kt is not None:
            self.e0 = torch.cuda.Event(enable_timing=True, external=self.kt.external)
            self.e0.record()

    def __exit__(self, *exc):
        if self.kt is not None:
            e1 = torch.cuda.Event(enable_timing=True, external=self.kt.external)
            e1.record()
            self.kt.records.append((self.family, self.meta, self.e0, e1))


_EPI_NAMES = {0: "bias_bf16", 1: "bias_f32", 2: "bias_res_f32", 3: "gelu_tanh_bf16", 4: "gelu_erf_bf16",
              5: "qkv_rope_bf16", 6: "scatter_bf16"}


# ------------------------------------------------------------------------------------------------
# tensor-level wrappers
# ------------------------------------------------------------------------------------------------
def gemm(a, w, mode, out, bias=None, res=None, rope=None, dst_rows=None, grp_rows=0, grp_stride=0, row_off=0,
         peer_ptrs=None, peer_multicast=False, ln_out=None, ln_in=None):
    """out = epilogue(a @ w.T). a [M,K] bf16 (row stride free), w [N,K] bf16, see vf_epilogue_mode.
    peer_ptrs: device pointers (ints) of up to 8 destination buffers shaped like `out` (fused all-gather: the rows
    are stored to every one of them, `out` only provides dtype and row pitch); peer_multicast: the single pointer is an
    NVSwitch multicast mapping (written with multimem.st).
    ln_out = (xb bf16 [rows, N], stat fp32 [N/32, rows, 2][, shift fp32 [rows]]): LayerNorm producer side (bias_res_f32
    only): bf16 copy and partial sums of the rows minus their shift.
    ln_in = (row_stats fp32 [M, 2] (mean', rstd) from ln_row_stats(), colsum fp32 [N]): LayerNorm consumer side
    (GELU / QKV+RoPE epilogues); `w` and `bias` must be the folded ones (qwen3_5_vision_model._fold_ln). Small problems:
    ln_in = (partials fp32 [K/32, rows, 2], colsum, eps, variant, shift or None) — the launch adds the partials up itself."""
    _require_cuda(a, w, out, bias, res, dst_rows)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.dim() == 2 and w.dim() == 2
    assert a.stride(1) == 1 and w.stride(1) == 1 and out.stride(-1) == 1
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K
    ep = vf_epilogue()
    ep.mode = mode
    ep.bias = _p(bias)
    ep.out = out.data_ptr()
    ep.ldo = out.stride(-2) if out.dim() >= 2 else N
    if res is not None:
        assert res.dtype == torch.float32 and res.stride(-1) == 1
        ep.res = res.data_ptr()
        ep.ldr = res.stride(-2)
    ep.grp_rows, ep.grp_stride, ep.row_off = grp_rows, grp_stride, row_off
    if rope is not None:
        cos_h, sin_h, period, cols = rope
        assert cos_h.dtype == torch.float32 and cos_h.shape[-1] == 32 and cos_h.is_contiguous()
        ep.rope_cos, ep.rope_sin, ep.rope_period, ep.rope_cols = cos_h.data_ptr(), sin_h.data_ptr(), period, cols
    ep.dst_rows = _p(dst_rows)
    if peer_ptrs:
        assert len(peer_ptrs) <= 8
        ep.n_peers = len(peer_ptrs)
        ep.peer_multicast = int(bool(peer_multicast))
        for i, ptr in enumerate(peer_ptrs):
            ep.peer_out[i] = int(ptr)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == N
    if ln_out is not None:
        xb, stat = ln_out[:2]
        _require_cuda(xb, stat)
        assert xb.dtype == torch.bfloat16 and xb.stride(-1) == 1 and stat.dtype == torch.float32 and stat.is_contiguous()
        assert stat.dim() == 3 and stat.shape[0] == N // 32 and stat.shape[2] == 2
        ep.ln_xb_out, ep.ln_ldxb, ep.ln_stat_out, ep.ln_stat_ld = xb.data_ptr(), xb.stride(-2), stat.data_ptr(), stat.shape[1]
        if len(ln_out) > 2 and ln_out[2] is not None:
            shift = ln_out[2]
            _require_cuda(shift)
            assert shift.dtype == torch.float32 and shift.is_contiguous() and shift.numel() >= M
            ep.ln_shift = shift.data_ptr()
    if ln_in is not None:
        colsum = ln_in[1]
        _require_cuda(colsum)
        assert colsum.dtype == torch.float32 and colsum.is_contiguous() and colsum.numel() == N
        ep.ln_colsum = colsum.data_ptr()
        if len(ln_in) == 2:      # (row_stats, colsum): statistics finished by ln_row_stats()
            row_stats = ln_in[0]
            _require_cuda(row_stats)
            assert row_stats.dtype == torch.float32 and row_stats.is_contiguous() and row_stats.shape == (M, 2)
            ep.ln_row_stats = row_stats.data_ptr()
        else:                    # (partials, colsum, eps, variant, shift or None): small problems, no launch in between
            part, _, eps, variant, shift = ln_in
            _require_cuda(part, shift)
            assert part.dtype == torch.float32 and part.is_contiguous() and part.dim() == 3 and part.shape[0] == K // 32 and part.shape[2] == 2
            ep.ln_part_in, ep.ln_stat_ld, ep.ln_eps, ep.ln_variant = part.data_ptr(), part.shape[1], float(eps), int(variant)
            if shift is not None:
                assert shift.dtype == torch.float32 and shift.is_contiguous() and shift.numel() >= M
                ep.ln_shift_update = shift.data_ptr()
    with _timed("gemm_" + _EPI_NAMES.get(mode, str(mode)), flops=2.0 * M * N * K):
        check(lib().vf_gemm_bf16(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), M, N, K, C.byref(ep),
                                 _stream()), "vf_gemm_bf16")
    return out


def ln_row_stats(stat, D, eps, out, shift=None, variant=0):
    """partials fp32 [parts, rows, 2] (a producer's ln_out stat) -> out fp32 [rows, 2] = (mean of the shifted row, rstd);
    shift fp32 [rows] (the producer's ln_out shift) is advanced to the row's true mean."""
    _require_cuda(stat, out, shift)
    assert stat.dtype == torch.float32 and stat.is_contiguous() and stat.dim() == 3 and stat.shape[2] == 2
    assert out.dtype == torch.float32 and out.is_contiguous() and out.shape == (stat.shape[1], 2)
    if shift is not None:
        assert shift.dtype == torch.float32 and shift.is_contiguous() and shift.numel() >= stat.shape[1]
    with _timed("ln_row_stats", bytes=stat.numel() * 4 + out.numel() * 4):
        check(lib().vf_ln_row_stats(stat.data_ptr(), stat.shape[0], stat.shape[1], stat.shape[1], D, float(eps), int(variant),
                                    out.data_ptr(), _p(shift), _stream()), "vf_ln_row_stats")
    return out


def patch_embed(pixels, weight2d, bias, pos, out, P, tp, out_rows_per_sample, out_row_off):
    """pixels bf16 [B,C,T,H,W]; weight2d bf16 [N, C*tp*P*P]; out fp32 [rows, N] (see vfuse.h)."""
    _require_cuda(pixels, weight2d, out)
    assert pixels.dtype == torch.bfloat16 and pixels.is_contiguous() and pixels.dim() == 5
    B, Cc, T, H, W = pixels.shape
    N = weight2d.shape[0]
    with _timed("patch_embed", flops=2.0 * B * (T // tp) * (H // P) * (W // P) * N * weight2d.shape[1]):
        check(
            lib().vf_patch_embed(pixels.data_ptr(), B, Cc, T, H, W, P, tp, weight2d.data_ptr(), _p(bias), _p(pos),
                                 pos.stride(0) if pos is not None else 0, N, out.data_ptr(), out.stride(-2),
                                 out_rows_per_sample, out_row_off, _stream()),
            "vf_patch_embed",
        )
    return out


def attention(qkv, out, B, S, H, scale, head_dim=64):
    _require_cuda(qkv, out)
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and out.dtype == torch.bfloat16 and out.is_contiguous()
    assert qkv.numel() == B * S * 3 * H * head_dim and out.numel() == B * S * H * head_dim
    with _timed("attention", flops=4.0 * B * H * S * S * head_dim):
        check(lib().vf_attention_fwd_hd(qkv.data_ptr(), out.data_ptr(), B, S, H, head_dim, float(scale), _stream()),
              "vf_attention_fwd_hd")
    return out


def attention_gqa(q2d, k2d, v2d, out2d, B, S, Hq, Hkv, scale, causal=True, q_col0=0, q_head_stride=256, gate2d=None,
                  gate_col0=0, gate_head_stride=256):
    """Causal GQA attention, head_dim 256, on token-major 2-D views (row stride free, unit column stride)."""
    _require_cuda(q2d, k2d, v2d, out2d, gate2d)
    for t in (q2d, k2d, v2d, out2d) + ((gate2d,) if gate2d is not None else ()):
        assert t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1 and t.shape[0] == B * S
    with _timed("attention_gqa", flops=4.0 * B * Hq * S * S * 256 * (0.5 if causal else 1.0)):
        check(lib().vf_attention_gqa_fwd(q2d.data_ptr(), q2d.stride(0), q_col0, q_head_stride, k2d.data_ptr(), k2d.stride(0),
                                         v2d.data_ptr(), v2d.stride(0), out2d.data_ptr(), out2d.stride(0), _p(gate2d),
                                         gate2d.stride(0) if gate2d is not None else 0, gate_col0, gate_head_stride,
                                         B, S, Hq, Hkv, 256, float(scale), int(bool(causal)), _stream()),
              "vf_attention_gqa_fwd")
    return out2d


def layernorm(x2d, w, b, out, eps, variant=0, merge=1, nh=0, nw=0, mean_out=None):
    _require_cuda(x2d, w, b, out, mean_out)
    rows, D = x2d.shape
    assert x2d.stride(1) == 1 and out.is_contiguous()
    if mean_out is not None:
        assert mean_out.dtype == torch.float32 and mean_out.is_contiguous() and mean_out.numel() >= rows
    with _timed("layernorm", bytes=float(rows * D * (x2d.element_size() + out.element_size()))):
        check(
            lib().vf_layernorm(x2d.data_ptr(), _DT[x2d.dtype], x2d.stride(0), w.data_ptr(), b.data_ptr(),
                               out.data_ptr(), _DT[out.dtype], rows, D, float(eps), variant, merge, nh, nw, _p(mean_out),
                               _stream()),
            "vf_layernorm",
        )
    return out


def vit_cls_pos(cls, pos, out, B, rows_per_sample, D):
    _require_cuda(cls, pos, out)
    check(lib().vf_vit_cls_pos(cls.data_ptr(), pos.data_ptr(), out.data_ptr(), B, rows_per_sample, D, _stream()),
          "vf_vit_cls_pos")


def rope_apply(x, cos, sin, position_ids=None):
    _require_cuda(x, cos, sin, position_ids)
    B, H, S, hd = x.shape
    xc = x.contiguous()
    out = torch.empty_like(xc)
    pid = None if position_ids is None else position_ids.to(torch.int64).contiguous()
    with _timed("rope_apply", bytes=2.0 * xc.numel() * xc.element_size()):
        check(
            lib().vf_rope_apply(xc.data_ptr(), out.data_ptr(), _DT[x.dtype], B, H, S, hd, cos.data_ptr(), sin.data_ptr(),
                                cos.shape[-1], cos.shape[0], _p(pid), _stream()),
            "vf_rope_apply",
        )
    return out


def mrope_apply(x, cos, sin, position_ids, mrope_section, norm_weight=None, norm_eps=1e-6):
    _require_cuda(x, cos, sin, position_ids, norm_weight)
    B, H, S, hd = x.shape
    xc = x.contiguous()
    out = torch.empty_like(xc)
    pid = position_ids.to(torch.int64).contiguous()
    st, sh, sw = (int(v) for v in mrope_section)
    with _timed("mrope_apply", bytes=2.0 * xc.numel() * xc.element_size() + 8.0 * pid.numel()):
        check(
            lib().vf_mrope_apply(xc.data_ptr(), out.data_ptr(), _DT[x.dtype], B, H, S, hd, cos.data_ptr(), sin.data_ptr(),
                                 cos.shape[-1], cos.shape[0], pid.data_ptr(), st, sh, sw, _p(norm_weight),
                                 float(norm_eps), _stream()),
            "vf_mrope_apply",
        )
    return out


def mrope_apply_heads_(x2d, col0, head_stride, B, H, S, cos, sin, position_ids, mrope_section, norm_weight=None,
                       norm_eps=1e-6, hd=256):
    """In place on a token-major 2-D tensor [B*S, ld]: head h occupies columns col0 + h*head_stride .. +hd.
    (q/k zero-centred RMSNorm +) MRoPE-I without any transpose; see vf_mrope_apply_strided."""
    _require_cuda(x2d, cos, sin, position_ids, norm_weight)
    assert x2d.dim() == 2 and x2d.stride(1) == 1 and x2d.shape[0] == B * S
    ld = x2d.stride(0)
    st = (C.c_int64 * 3)(S * ld, head_stride, ld)
    pid = position_ids.to(torch.int64).contiguous()
    st_, sh_, sw_ = (int(v) for v in mrope_section)
    base = x2d.data_ptr() + col0 * x2d.element_size()
    with _timed("mrope_apply", bytes=2.0 * B * S * H * hd * x2d.element_size() + 8.0 * pid.numel()):
        check(
            lib().vf_mrope_apply_strided(base, base, _DT[x2d.dtype], B, H, S, hd, st, st, cos.data_ptr(), sin.data_ptr(),
                                         cos.shape[-1], cos.shape[0], pid.data_ptr(), st_, sh_, sw_, _p(norm_weight),
                                         float(norm_eps), _stream()),
            "vf_mrope_apply_strided",
        )
    return x2d


def mrope_position_ids(input_ids, image_mask, image_token_id, feeds_cpu, merge):
    """input_ids int64 [b, seq] (cuda); feeds_cpu: CPU int64 [F,3] or None-equivalent empty."""
    _require_cuda(input_ids, image_mask)
    b, seq = input_ids.shape
    ids = input_ids.to(torch.int64).contiguous()
    mask = None if image_mask is None else image_mask.to(torch.uint8).contiguous()
    feeds = feeds_cpu.to(device="cpu", dtype=torch.int64).contiguous()
    out = torch.empty((3, b, seq), dtype=torch.int64, device=input_ids.device)
    with _timed("mrope_position_ids", bytes=32.0 * b * seq):   # read ids (8 B), write 3 axes (24 B)
        check(
            lib().vf_mrope_position_ids(ids.data_ptr(), _p(mask), int(image_token_id), feeds.data_ptr(), feeds.shape[0],
                                        int(merge), b, seq, out.data_ptr(), _stream()),
            "vf_mrope_position_ids",
        )
    return out


def fuse_scan(input_ids, image_mask, image_token_id, inv_cap=0):
    """Returns (row_map int32 [b*seq], n_placeholders int32 [1], inv_map int32 [inv_cap] or None) —
    all on device, no sync. inv_map[j] = flat token row of the j-th placeholder, -1 beyond."""
    _require_cuda(input_ids, image_mask)
    ids = input_ids.to(torch.int64).contiguous()
    n = ids.numel()
    mask = None if image_mask is None else image_mask.to(torch.uint8).contiguous()
    row_map = torch.empty(n, dtype=torch.int32, device=ids.device)
    count = torch.empty(1, dtype=torch.int32, device=ids.device)
    inv = torch.empty(inv_cap, dtype=torch.int32, device=ids.device) if inv_cap > 0 else None
    scratch = torch.empty(n // 1024 + 1024, dtype=torch.int32, device=ids.device)
    with _timed("fuse_scan", bytes=12.0 * n + 4.0 * inv_cap):   # read ids (8 B), write rank (4 B) + inverse map
        check(
            lib().vf_fuse_scan(ids.data_ptr(), _p(mask), int(image_token_id), n, row_map.data_ptr(), count.data_ptr(),
                               _p(inv), inv_cap, scratch.data_ptr(), _stream()),
            "vf_fuse_scan",
        )
    return row_map, count, inv


def embed_gather_scatter(input_ids, table, vision, row_map, out, skip_vision=False, n_vis=None):
    _require_cuda(input_ids, table, vision, row_map, out)
    ids = input_ids.to(torch.int64).contiguous()
    assert table.dtype == torch.bfloat16 and table.is_contiguous() and out.dtype == torch.bfloat16 and out.is_contiguous()
    D = table.shape[1]
    n_vis, vd = (0 if n_vis is None else int(n_vis)), 1
    if vision is not None:
        assert vision.is_contiguous() and vision.shape[-1] == D
        n_vis, vd = vision.numel() // D, _DT[vision.dtype]
    rows_moved = ids.numel() - (n_vis if skip_vision else 0)   # placeholder rows are skipped in skip mode
    with _timed("embed_gather_scatter", bytes=rows_moved * D * 4.0 + ids.numel() * 12.0):
        check(
            lib().vf_embed_gather_scatter(ids.data_ptr(), table.data_ptr(), table.shape[0], D, _p(vision), vd, n_vis,
                                          _p(row_map), out.data_ptr(), ids.numel(), int(skip_vision), _stream()),
            "vf_embed_gather_scatter",
        )
    return out


def to_bf16(x):
    _require_cuda(x)
    if x.dtype == torch.bfloat16:
        return x.contiguous()
    assert x.dtype == torch.float32
    xc = x.contiguous()
    out = torch.empty(xc.shape, dtype=torch.bfloat16, device=x.device)
    with _timed("cast", bytes=6.0 * xc.numel()):
        check(lib().vf_cast_f32_to_bf16(xc.data_ptr(), out.data_ptr(), xc.numel(), _stream()), "vf_cast_f32_to_bf16")
    return out


def to_f32(x):
    _require_cuda(x)
    if x.dtype == torch.float32:
        return x.contiguous()
    assert x.dtype == torch.bfloat16
    xc = x.contiguous()
    out = torch.empty(xc.shape, dtype=torch.float32, device=x.device)
    with _timed("cast", bytes=6.0 * xc.numel()):
        check(lib().vf_cast_bf16_to_f32(xc.data_ptr(), out.data_ptr(), xc.numel(), _stream()), "vf_cast_bf16_to_f32")
    return out


def preprocess_u8(img_u8, mean, std, T=2, dtype=torch.bfloat16):
    """uint8 [B, H, W, 3] (cuda) -> normalised [B, 3, T, H, W] (every temporal slot = the frame)."""
    _require_cuda(img_u8)
    assert img_u8.dtype == torch.uint8 and img_u8.dim() == 4 and img_u8.shape[-1] == 3 and img_u8.is_contiguous()
    B, H, W, _ = img_u8.shape
    out = torch.empty((B, 3, T, H, W), dtype=dtype, device=img_u8.device)
    m3 = (C.c_float * 3)(*[float(v) for v in mean])
    s3 = (C.c_float * 3)(*[float(v) for v in std])
    with _timed("preprocess_u8", bytes=float(B * H * W * 3 * (1 + T * out.element_size()))):
        check(lib().vf_preprocess_u8(img_u8.data_ptr(), B, H, W, T, m3, s3, out.data_ptr(), _DT[dtype], _stream()),
              "vf_preprocess_u8")
    return out


def gelu(x, tanh_form=False):
    """Stand-alone GELU (erf form by default, tanh form for nn.GELU(approximate="tanh")); fp32 or bf16, any shape."""
    _require_cuda(x)
    if x.dtype not in _DT:
        raise VFuseError(f"vf_gelu handles fp32 and bf16 tensors, got {x.dtype}")
    xc = x.contiguous()
    out = torch.empty_like(xc)
    if xc.numel() == 0:
        return out
    with _timed("gelu", bytes=2.0 * xc.numel() * xc.element_size()):
        check(lib().vf_gelu(xc.data_ptr(), out.data_ptr(), _DT[xc.dtype], xc.numel(), int(bool(tanh_form)), _stream()), "vf_gelu")
    return out


def rmsnorm_zc(x, one_plus_scale, eps):
    """ZeroCenteredRMSNorm.forward over the last dim; x fp32 or bf16, one_plus_scale fp32 [D]."""
    _require_cuda(x, one_plus_scale)
    if x.dtype not in _DT:
        raise VFuseError(f"vf_rmsnorm_zc handles fp32 and bf16 tensors, got {x.dtype}")
    D = x.shape[-1]
    x2d = x.reshape(-1, D)
    if x2d.stride(1) != 1:
        x2d = x2d.contiguous()
    out = torch.empty((x2d.shape[0], D), dtype=x.dtype, device=x.device)
    assert one_plus_scale.dtype == torch.float32 and one_plus_scale.is_contiguous() and one_plus_scale.numel() == D
    with _timed("rmsnorm_zc", bytes=2.0 * x2d.numel() * x2d.element_size()):
        check(lib().vf_rmsnorm_zc(x2d.data_ptr(), x2d.stride(0), one_plus_scale.data_ptr(), out.data_ptr(), D, _DT[x.dtype],
                                  x2d.shape[0], D, float(eps), _stream()), "vf_rmsnorm_zc")
    return out.view(x.shape)


def embed_pos_concat(input_ids, tok_table, pos_table, fused, row_off):
    """fused fp32 [b, n_total, D]: fused[:, row_off:row_off+seq] = tok_table[input_ids] + pos_table[:seq]."""
    _require_cuda(input_ids, tok_table, pos_table, fused)
    b, seq = input_ids.shape
    ids = input_ids.to(torch.int64).contiguous()
    assert tok_table.dtype == pos_table.dtype and tok_table.dtype in _DT and tok_table.is_contiguous() and pos_table.is_contiguous()
    assert fused.dtype == torch.float32 and fused.is_contiguous() and fused.dim() == 3 and fused.shape[0] == b
    D = tok_table.shape[1]
    assert fused.shape[2] == D and pos_table.shape[1] == D
    with _timed("embed_pos_concat", bytes=float(b * seq * D * (2 * tok_table.element_size() + 4))):
        check(lib().vf_embed_pos_concat(ids.data_ptr(), tok_table.data_ptr(), tok_table.shape[0], pos_table.data_ptr(),
                                        pos_table.shape[0], _DT[tok_table.dtype], fused.data_ptr(), b, seq, D, fused.shape[1],
                                        int(row_off), _stream()), "vf_embed_pos_concat")
    return fused


def im2col_patches(x, P):
    """x [B, C, H, W] fp32/bf16 -> bf16 [B*nh*nw, C*P*P (padded to a multiple of 8)] patch rows, columns (c, py, px)."""
    _require_cuda(x)
    assert x.dim() == 4 and x.is_contiguous() and x.dtype in _DT
    B, Cc, H, W = x.shape
    K = Cc * P * P
    ld = (K + 7) // 8 * 8
    out = torch.zeros((B * (H // P) * (W // P), ld), dtype=torch.bfloat16, device=x.device) if ld != K else \
        torch.empty((B * (H // P) * (W // P), ld), dtype=torch.bfloat16, device=x.device)
    with _timed("im2col", bytes=float(x.numel() * (x.element_size() + 2))):
        check(lib().vf_im2col_patches(x.data_ptr(), _DT[x.dtype], B, Cc, H, W, P, out.data_ptr(), ld, _stream()), "vf_im2col_patches")
    return out[:, :K] if ld != K else out


def fill_rows(src, out, B, rows_per_sample, D, add_row0=None):
    _require_cuda(src, out, add_row0)
    assert src.dtype == torch.float32 and src.is_contiguous() and out.dtype == torch.float32 and out.is_contiguous()
    check(lib().vf_fill_rows_f32(src.data_ptr(), _p(add_row0), out.data_ptr(), B, rows_per_sample, D, _stream()), "vf_fill_rows_f32")
    return out
