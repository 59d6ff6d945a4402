"""Thin torch-tensor wrappers over the C ABI (include/g2vlm_b200.h).

PyTorch is used here for device memory and streams only: every function validates its tensors,
fills the plain-C argument struct with raw device pointers and sizes and calls the kernel library
on torch's current CUDA stream.  Nothing in this module computes with torch ops.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch

from . import _lib

EPI_STORE_BF16, EPI_SWIGLU_BF16, EPI_RESID_F32, EPI_STORE_F32 = 0, 1, 2, 3
GEMM_GELU, GEMM_ROUND_AFTER_SCALE, GEMM_RELU, GEMM_ACCUMULATE, GEMM_ROUND_BF16 = 1, 2, 4, 8, 16
GEMM_QUICK_GELU, GEMM_ROUND_SUM = 32, 64
GEMM_NO_TMA_OUT = 512   # RESID_F32: SM-side read-modify-write instead of the TMA reduce-add
GEMM_FORCE_PAIR, GEMM_FORCE_SINGLE = 128, 256   # choose the CTA-pair / 1-CTA kernel per call (tests, A/B timing)


class G2Error(RuntimeError):
    pass


LAUNCHES = 0  # number of C-ABI kernel launches issued by this process (bench.py reads it)


def launches() -> int:
    """Kernels launched so far by this process through the C ABI: per-op calls from Python (LAUNCHES) plus the
    launches issued inside the native stage drivers (g2vlm_dino_forward / g2vlm_mot_forward_geo / g2vlm_recon_heads)."""
    lib = _lib.load()
    lib.g2vlm_driver_launches.restype = ctypes.c_int64
    return LAUNCHES + int(lib.g2vlm_driver_launches())


def _check(rc: int) -> None:
    global LAUNCHES
    LAUNCHES += 1
    if rc != 0:
        msg = _lib.load().g2vlm_last_error().decode()
        raise G2Error(f"g2vlm_b200 C ABI call failed (code {rc}): {msg}")


def num_sms() -> int:
    """SM count of the current device (persistent kernels launch one CTA per SM)."""
    return torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _req(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise G2Error(f"{name} must be a CUDA tensor (there is no CPU fallback)")
    if t.device.index != torch.cuda.current_device():
        # the launch goes to the CURRENT device's stream: a tensor of another GPU would be a wild pointer there
        raise G2Error(f"{name} lives on cuda:{t.device.index} but the current device is cuda:{torch.cuda.current_device()}"
                      "; call inside `with torch.cuda.device(tensor.device):` (G2VLMFast's public methods do)")
    if t.dtype != dtype:
        raise G2Error(f"{name} must be {dtype}, got {t.dtype}")
    if t.dim() >= 2 and t.stride(-1) != 1:
        raise G2Error(f"{name} must be contiguous in its last dimension")


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("A", ctypes.c_void_p), ("lda", ctypes.c_int64), ("a_rows", ctypes.c_int64),
        ("B", ctypes.c_void_p), ("ldb", ctypes.c_int64),
        ("N", ctypes.c_int32), ("K", ctypes.c_int32), ("n_groups", ctypes.c_int32),
        ("group_row0", ctypes.c_int32 * 2), ("group_rows", ctypes.c_int32 * 2),
        ("epilogue", ctypes.c_int32), ("flags", ctypes.c_uint32),
        ("out", ctypes.c_void_p), ("ldo", ctypes.c_int64),
        ("bias", ctypes.c_void_p), ("scale", ctypes.c_void_p), ("scale_groups", ctypes.c_uint32),
        ("residual", ctypes.c_void_p), ("ldr", ctypes.c_int64),
        ("out_col_group", ctypes.c_int32), ("out_col_stride", ctypes.c_int32),
        ("k_chunk_blocks", ctypes.c_int32),
    ]


def gemm(a: torch.Tensor, w: torch.Tensor, out: torch.Tensor, *, epilogue: int,
         groups: Optional[Sequence[tuple]] = None, bias: Optional[torch.Tensor] = None,
         scale: Optional[torch.Tensor] = None, scale_groups: int = 0, flags: int = 0,
         residual: Optional[torch.Tensor] = None, out_col_group: int = 0, out_col_stride: int = 0,
         k_chunk_blocks: int = 0) -> torch.Tensor:
    """out <- epilogue(a @ w_g.T) per token group; see g2vlm_gemm_bf16 in include/g2vlm_b200.h.
    out_col_group / out_col_stride (STORE_BF16): output column c lands at (c // group) * stride + c % group.

    a: bf16 [rows, K]; w: bf16 [n_groups*N, K] (experts stacked); groups: [(row0, rows), ...]
    (default: one group covering all rows); bias: fp32 [n_groups*N]; scale: fp32 [N].
    """
    _req(a, torch.bfloat16, "a")
    _req(w, torch.bfloat16, "w")
    if groups is None:
        groups = [(0, a.shape[0])]
    ng = len(groups)
    if ng not in (1, 2):
        raise G2Error("gemm: 1 or 2 groups")
    K = a.shape[1]
    if w.shape[1] != K or w.shape[0] % ng:
        raise G2Error(f"gemm: weight shape {tuple(w.shape)} does not match K={K}, groups={ng}")
    N = w.shape[0] // ng
    args = GemmArgs()
    args.A, args.lda, args.a_rows = a.data_ptr(), a.stride(0), a.shape[0]
    args.B, args.ldb = w.data_ptr(), w.stride(0)
    args.N, args.K, args.n_groups = N, K, ng
    for g, (r0, rows) in enumerate(groups):
        args.group_row0[g], args.group_rows[g] = int(r0), int(rows)
    args.epilogue, args.flags = epilogue, flags
    want = torch.bfloat16 if epilogue in (EPI_STORE_BF16, EPI_SWIGLU_BF16) else torch.float32
    _req(out, want, "out")
    n_out = N // 2 if epilogue == EPI_SWIGLU_BF16 else N
    if out_col_group:
        n_out = (N // out_col_group - 1) * out_col_stride + out_col_group
        args.out_col_group, args.out_col_stride = out_col_group, out_col_stride
    if out.shape[0] < max(r0 + rows for r0, rows in groups) or out.shape[1] < n_out:
        raise G2Error(f"gemm: out shape {tuple(out.shape)} too small")
    args.out, args.ldo = out.data_ptr(), out.stride(0)
    if bias is not None:
        _req(bias, torch.float32, "bias")
        if bias.numel() != ng * N:
            raise G2Error("gemm: bias must have n_groups*N elements")
        args.bias = bias.data_ptr()
    if scale is not None:
        _req(scale, torch.float32, "scale")
        if scale.numel() != N:
            raise G2Error("gemm: scale must have N elements")
        args.scale, args.scale_groups = scale.data_ptr(), scale_groups
    if residual is not None:
        _req(residual, torch.float32, "residual")
        args.residual, args.ldr = residual.data_ptr(), residual.stride(0)
    args.k_chunk_blocks = int(k_chunk_blocks)
    _check(_lib.load().g2vlm_gemm_bf16(ctypes.byref(args), _stream()))
    return out


class AttnArgs(ctypes.Structure):
    _fields_ = [
        ("q", ctypes.c_void_p), ("ldq", ctypes.c_int64), ("q_rows", ctypes.c_int64),
        ("k", ctypes.c_void_p), ("ldk", ctypes.c_int64),
        ("v", ctypes.c_void_p), ("ldv", ctypes.c_int64), ("kv_rows", ctypes.c_int64),
        ("out", ctypes.c_void_p), ("ldo", ctypes.c_int64),
        ("num_q_heads", ctypes.c_int32), ("num_kv_heads", ctypes.c_int32),
        ("head_dim", ctypes.c_int32), ("causal", ctypes.c_int32),
        ("softmax_scale", ctypes.c_float), ("n_items", ctypes.c_int32),
        ("work_items", ctypes.c_void_p),
        ("out_head_cols", ctypes.c_int32),
        ("lse_out", ctypes.c_void_p), ("max_ctas", ctypes.c_int32),
    ]


ATTN_ROWS_PER_ITEM = 256


def attention_work_table(cu_seqlens_q: Sequence[int], cu_seqlens_k: Sequence[int]) -> torch.Tensor:
    """Host-side: segment table (flash-attn cu_seqlens) -> int32 [n_items, 8] work items (CPU).

    One item per <= 256 query rows of a segment; see g2vlm_attention in include/g2vlm_b200.h.
    """
    items = []
    for i in range(len(cu_seqlens_q) - 1):
        qb, qe = int(cu_seqlens_q[i]), int(cu_seqlens_q[i + 1])
        kb, ke = int(cu_seqlens_k[i]), int(cu_seqlens_k[i + 1])
        for t0 in range(qb, qe, ATTN_ROWS_PER_ITEM):
            items.append([t0, qb, qe, kb, ke, 0, 0, 0])
    return torch.tensor(items, dtype=torch.int32).reshape(-1, 8)


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, work: torch.Tensor,
              *, num_q_heads: int, num_kv_heads: int, head_dim: int, scale: float,
              causal: bool = False, out_head_cols: int = 0, lse: Optional[torch.Tensor] = None,
              max_ctas: int = 0) -> torch.Tensor:
    """out[rows covered by `work`] = softmax(scale * q k^T) v, per segment; q/k/v/out are 2-D bf16
    views [rows, heads*head_dim] (they may be column slices of one fused QKV buffer).  out_head_cols: write only
    that many columns per head, heads packed at that stride (96-wide heads computed in 128-wide slots).
    lse: optional fp32 [q_rows, heads] receiving the log-sum-exp of every covered row (see attention_merge);
    max_ctas: bound on the persistent grid (leaves SMs to a concurrent communication kernel)."""
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        _req(t, torch.bfloat16, n)
    _req(work, torch.int32, "work")
    if work.dim() != 2 or work.shape[1] != 8 or not work.is_contiguous():
        raise G2Error("attention: work table must be a contiguous int32 [n, 8] tensor")
    if k.shape[0] != v.shape[0]:
        raise G2Error("attention: k and v must have the same number of rows")
    args = AttnArgs()
    args.q, args.ldq, args.q_rows = q.data_ptr(), q.stride(0), q.shape[0]
    args.k, args.ldk = k.data_ptr(), k.stride(0)
    args.v, args.ldv, args.kv_rows = v.data_ptr(), v.stride(0), k.shape[0]
    args.out, args.ldo = out.data_ptr(), out.stride(0)
    args.num_q_heads, args.num_kv_heads, args.head_dim = num_q_heads, num_kv_heads, head_dim
    args.causal, args.softmax_scale = int(causal), float(scale)
    args.n_items, args.work_items = work.shape[0], work.data_ptr()
    args.out_head_cols = int(out_head_cols)
    if lse is not None:
        _req(lse, torch.float32, "lse")
        if not lse.is_contiguous() or lse.shape[0] < q.shape[0] or lse.shape[1] != num_q_heads:
            raise G2Error("attention: lse must be a contiguous fp32 [q_rows, num_q_heads] tensor")
        args.lse_out = lse.data_ptr()
    args.max_ctas = int(max_ctas)
    _check(_lib.load().g2vlm_attention(ctypes.byref(args), _stream()))
    return out


def attention_merge(o_a: torch.Tensor, lse_a: torch.Tensor, o_b: torch.Tensor, lse_b: torch.Tensor, out: torch.Tensor,
                    heads: int, head_cols: int, rows: Optional[int] = None) -> torch.Tensor:
    """Combine two attention partials over disjoint key sets (see g2vlm_attention_merge); out may alias o_a."""
    for t, n in ((o_a, "o_a"), (o_b, "o_b"), (out, "out")):
        _req(t, torch.bfloat16, n)
    for t, n in ((lse_a, "lse_a"), (lse_b, "lse_b")):
        _req(t, torch.float32, n)
        if not t.is_contiguous() or t.shape[1] != heads:
            raise G2Error(f"attention_merge: {n} must be contiguous [rows, heads]")
    rows = out.shape[0] if rows is None else rows
    if min(o_a.shape[0], o_b.shape[0], out.shape[0], lse_a.shape[0], lse_b.shape[0]) < rows:
        raise G2Error("attention_merge: a tensor has fewer rows than requested")
    if min(o_a.shape[1], o_b.shape[1], out.shape[1]) < heads * head_cols:
        raise G2Error("attention_merge: a tensor has fewer than heads*head_cols columns")
    _call("g2vlm_attention_merge", _vp(o_a.data_ptr()), _i64(o_a.stride(0)), _vp(lse_a.data_ptr()), _vp(o_b.data_ptr()),
          _i64(o_b.stride(0)), _vp(lse_b.data_ptr()), _vp(out.data_ptr()), _i64(out.stride(0)), _i64(rows), _i32(heads),
          _i32(head_cols))
    return out


# ---------------------------------------------------------------------------------------------------
# memory-bound kernels
# ---------------------------------------------------------------------------------------------------
_i64, _i32, _f32, _vp = ctypes.c_int64, ctypes.c_int32, ctypes.c_float, ctypes.c_void_p


def _call(name: str, *args) -> None:
    _check(getattr(_lib.load(), name)(*args, _stream()))


def gather_rows(src: torch.Tensor, dst: torch.Tensor, idx: Optional[torch.Tensor], n_rows: int,
                scatter: bool = False, row_elems: Optional[int] = None) -> torch.Tensor:
    """dst[i] = src[idx[i]] (or dst[idx[i]] = src[i] if scatter); idx None = plain row copy."""
    if src.dtype != dst.dtype or not src.is_cuda or src.stride(-1) != 1 or dst.stride(-1) != 1:
        raise G2Error("gather_rows: src/dst must be CUDA tensors of the same dtype, contiguous rows")
    if idx is not None:
        _req(idx, torch.int64, "idx")
    width = (row_elems if row_elems is not None else min(src.shape[1], dst.shape[1])) * src.element_size()
    _call("g2vlm_gather_rows", _vp(src.data_ptr()), _i64(src.stride(0) * src.element_size()),
          _vp(dst.data_ptr()), _i64(dst.stride(0) * dst.element_size()), _ptr(idx), _i64(n_rows),
          _i64(width), _i32(int(scatter)))
    return dst


def rmsnorm_routed(x, out, w_a, w_b, n_first: int, eps: float, rows: Optional[int] = None, round_normed: bool = False):
    _req(x, torch.float32, "x")
    rows = x.shape[0] if rows is None else rows
    _call("g2vlm_rmsnorm_routed", _vp(x.data_ptr()), _i64(x.stride(0)), _vp(out.data_ptr()),
          _i64(out.stride(0)), _i32(int(out.dtype == torch.bfloat16) | (2 if round_normed else 0)), _vp(w_a.data_ptr()),
          _vp(w_b.data_ptr()), _i64(rows), _i64(n_first), _i32(x.shape[1]), _f32(eps))
    return out


def layernorm(x, out, w, b, eps: float, rows: Optional[int] = None, seg_in: int = 0, seg_skip: int = 0):
    _req(x, torch.float32, "x")
    rows = x.shape[0] if rows is None else rows
    _call("g2vlm_layernorm", _vp(x.data_ptr()), _i64(x.stride(0)), _vp(out.data_ptr()), _i64(out.stride(0)),
          _i32(int(out.dtype == torch.bfloat16)), _vp(w.data_ptr()), _vp(b.data_ptr()), _i64(rows),
          _i32(x.shape[1]), _f32(eps), _i32(seg_in), _i32(seg_skip))
    return out


def mrope_table(position_ids, inv_freq, cos, sin, sections):
    _req(position_ids, torch.int64, "position_ids")
    rows, half = position_ids.shape[1], inv_freq.numel()
    _call("g2vlm_mrope_table", _vp(position_ids.data_ptr()), _i64(position_ids.stride(0)),
          _vp(inv_freq.data_ptr()), _vp(cos.data_ptr()), _vp(sin.data_ptr()), _i64(rows), _i32(half),
          _i32(sections[0]), _i32(sections[1]))


def qknorm_mrope(qkv, rows, n_first, n_q, n_kv, head_dim, qw_a, kw_a, qw_b, kw_b, cos, sin, eps,
                 round_normed=False):
    _req(qkv, torch.bfloat16, "qkv")
    _call("g2vlm_qknorm_mrope", _vp(qkv.data_ptr()), _i64(qkv.stride(0)), _i64(rows), _i64(n_first),
          _i32(n_q), _i32(n_kv), _i32(head_dim), _vp(qw_a.data_ptr()), _vp(kw_a.data_ptr()),
          _vp(qw_b.data_ptr()), _vp(kw_b.data_ptr()), _vp(cos.data_ptr()), _vp(sin.data_ptr()), _f32(eps),
          _i32(int(round_normed)))


def im2col_patches(images, out, patch: int, mean=None, std=None):
    """mean / std: 3-tuples -> normalise (x - mean[c]) / std[c] on the fly (raw [0,1] images in)."""
    _req(images, torch.float32, "images")
    _req(out, torch.bfloat16, "out")
    n, _, H, W = images.shape
    if not images.is_contiguous() or not out.is_contiguous():
        raise G2Error("im2col: contiguous tensors required")
    m3 = (ctypes.c_float * 3)(*mean) if mean is not None else None
    s3 = (ctypes.c_float * 3)(*std) if std is not None else None
    _call("g2vlm_im2col_patches", _vp(images.data_ptr()), _vp(out.data_ptr()), _i32(n), _i32(H), _i32(W),
          _i32(patch), _i32(out.shape[1]), m3, s3)
    return out


def dino_embed(patch_emb, cls, reg, pos, out, n: int, P: int, n_reg: int):
    _req(patch_emb, torch.bfloat16, "patch_emb")
    _req(out, torch.float32, "out")
    _call("g2vlm_dino_embed", _vp(patch_emb.data_ptr()), _i64(patch_emb.stride(0)), _vp(cls.data_ptr()),
          _vp(reg.data_ptr()), _vp(pos.data_ptr()), _vp(out.data_ptr()), _i32(n), _i32(P), _i32(n_reg),
          _i32(out.shape[1]))
    return out


def rope2d(buf, rows, n_heads_total, head_stride, head_dim, tokens_per_view, grid_w, cos, sin, bf16_ops=True):
    _req(buf, torch.bfloat16, "buf")
    _call("g2vlm_rope2d", _vp(buf.data_ptr()), _i64(buf.stride(0)), _i64(rows), _i32(n_heads_total),
          _i32(head_stride), _i32(head_dim), _i32(tokens_per_view), _i32(grid_w), _vp(cos.data_ptr()),
          _vp(sin.data_ptr()), _i32(int(bf16_ops)))


def points_epilogue(feat, poses, out0, out1, n, H, W, patch, mode):
    _req(feat, torch.float32, "feat")
    _call("g2vlm_points_epilogue", _vp(feat.data_ptr()), _i64(feat.stride(0)), _ptr(poses), _vp(out0.data_ptr()),
          _ptr(out1), _i32(n), _i32(H), _i32(W), _i32(patch), _i32(mode))


def mean_pool(x, out, n_views, tokens):
    _req(x, torch.float32, "x")
    _call("g2vlm_mean_pool", _vp(x.data_ptr()), _i64(x.stride(0)), _vp(out.data_ptr()), _i32(n_views),
          _i32(tokens), _i32(x.shape[1]))
    return out


def split3(x, out, rows: Optional[int] = None):
    _req(x, torch.float32, "x")
    _req(out, torch.bfloat16, "out")
    rows = x.shape[0] if rows is None else rows
    _call("g2vlm_split3_f32", _vp(x.data_ptr()), _i64(x.stride(0)), _vp(out.data_ptr()), _i64(out.stride(0)),
          _i64(rows), _i32(x.shape[1]))
    return out


def cast_bf16(x, out, rows: Optional[int] = None):
    _req(x, torch.float32, "x")
    _req(out, torch.bfloat16, "out")
    rows = x.shape[0] if rows is None else rows
    _call("g2vlm_cast_f32_to_bf16", _vp(x.data_ptr()), _i64(x.stride(0)), _vp(out.data_ptr()),
          _i64(out.stride(0)), _i64(rows), _i32(x.shape[1]))
    return out


def camera_pose(feat, w_t, b_t, w_r, b_r, poses):
    _req(feat, torch.float32, "feat")
    _call("g2vlm_camera_pose", _vp(feat.data_ptr()), _i64(feat.stride(0)), _vp(w_t.data_ptr()),
          _vp(b_t.data_ptr()), _vp(w_r.data_ptr()), _vp(b_r.data_ptr()), _vp(poses.data_ptr()),
          _i32(feat.shape[0]), _i32(feat.shape[1]))
    return poses


def ply_pack(points: torch.Tensor, images: torch.Tensor, filter_nonfinite: bool = True):
    """points fp32 [N,H,W,3], images fp32 [N,3,H,W] (device) -> (uint8 [n_valid*27] packed PLY vertex
    records on the device, n_valid).  Order-preserving; points with a NaN/Inf coordinate are dropped unless
    filter_nonfinite is False (the reference's filter_nan=False)."""
    _req(points, torch.float32, "points")
    _req(images, torch.float32, "images")
    n, H, W, _ = points.shape
    if not points.is_contiguous() or not images.is_contiguous() or images.shape != (n, 3, H, W):
        raise G2Error("ply_pack: contiguous points [N,H,W,3] and images [N,3,H,W] required")
    total = n * H * W
    out = torch.empty(max(total, 1) * 27, dtype=torch.uint8, device=points.device)
    counts = torch.empty((total + 1023) // 1024 + 1, dtype=torch.int32, device=points.device)
    n_valid = torch.zeros(1, dtype=torch.int64, device=points.device)
    _call("g2vlm_ply_pack", _vp(points.data_ptr()), _vp(images.data_ptr()), _i32(n), _i32(H), _i32(W),
          _vp(out.data_ptr()), _vp(counts.data_ptr()), _vp(n_valid.data_ptr()), _i32(int(filter_nonfinite)))
    k = int(n_valid.item())
    return out[: k * 27], k


def argmax_bf16(logits: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    _req(logits, torch.bfloat16, "logits")
    _req(out, torch.int64, "out")
    _call("g2vlm_argmax_bf16", _vp(logits.data_ptr()), _i64(logits.stride(0)), _i64(logits.shape[0]),
          _i32(logits.shape[1]), _vp(out.data_ptr()))
    return out


def attention_decode(q, k, v, out, workspace, *, num_q_heads, num_kv_heads, head_dim, scale, kv_len_dev=None,
                     kv_len_extra: int = 0):
    """One query token against the cached keys (see g2vlm_attention_decode).  Without kv_len_dev the key count
    is k.shape[0]; with it (device int32) the count is *kv_len_dev + kv_len_extra and k.shape[0] is the bound."""
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        _req(t, torch.bfloat16, n)
    _req(workspace, torch.float32, "workspace")
    if kv_len_dev is not None:
        _req(kv_len_dev, torch.int32, "kv_len_dev")
    _call("g2vlm_attention_decode", _vp(q.data_ptr()), _vp(k.data_ptr()), _i64(k.stride(0)), _vp(v.data_ptr()),
          _i64(v.stride(0)), _i64(k.shape[0]), _ptr(kv_len_dev), _i32(kv_len_extra), _vp(out.data_ptr()),
          _i32(num_q_heads), _i32(num_kv_heads), _i32(head_dim), _f32(scale), _vp(workspace.data_ptr()),
          _i64(workspace.numel()))
    return out


def kv_append(src, dst, rows: int, len_dev=None, static_row: int = 0):
    """dst[(*len_dev | static_row) + i] = src[i] for i < rows (2-D views, same dtype)."""
    if src.dtype != dst.dtype or src.stride(-1) != 1 or dst.stride(-1) != 1:
        raise G2Error("kv_append: same dtype, contiguous rows")
    if len_dev is not None:
        _req(len_dev, torch.int32, "len_dev")
    _call("g2vlm_kv_append", _vp(src.data_ptr()), _i64(src.stride(0) * src.element_size()), _vp(dst.data_ptr()),
          _i64(dst.stride(0) * dst.element_size()), _ptr(len_dev), _i64(static_row), _i64(rows),
          _i64(min(src.shape[1], dst.shape[1]) * src.element_size()))


def attention_decode_workspace_floats(kv_len: int, num_q_heads: int) -> int:
    return ((kv_len + 159) // 160) * num_q_heads * 130


class UndLayerWeights(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("wqkv", "bqkv", "wo", "wgu", "wdown", "input_norm", "post_norm",
                                               "q_norm", "k_norm")]


class DecodeStepArgs(ctypes.Structure):
    _fields_ = [
        ("num_layers", ctypes.c_int32), ("hidden", ctypes.c_int32), ("intermediate", ctypes.c_int32),
        ("n_q_heads", ctypes.c_int32), ("n_kv_heads", ctypes.c_int32), ("head_dim", ctypes.c_int32),
        ("vocab", ctypes.c_int32), ("rms_eps", ctypes.c_float), ("mrope_s0", ctypes.c_int32), ("mrope_s1", ctypes.c_int32),
        ("layers", ctypes.POINTER(UndLayerWeights)), ("kv", ctypes.POINTER(ctypes.c_void_p)),
        ("kv_capacity", ctypes.c_int64), ("kv_bound", ctypes.c_int64),
        ("embed", ctypes.c_void_p), ("final_norm", ctypes.c_void_p), ("lm_head", ctypes.c_void_p),
        ("inv_freq", ctypes.c_void_p), ("cur_token", ctypes.c_void_p), ("position", ctypes.c_void_p),
        ("cache_len", ctypes.c_void_p),
        ("x", ctypes.c_void_p), ("h", ctypes.c_void_p), ("qkv", ctypes.c_void_p), ("attn", ctypes.c_void_p),
        ("act", ctypes.c_void_p), ("y", ctypes.c_void_p), ("cos_sin", ctypes.c_void_p),
        ("attn_ws", ctypes.c_void_p), ("attn_ws_floats", ctypes.c_int64), ("logits", ctypes.c_void_p),
        ("fused_ws", ctypes.c_void_p), ("fused_ws_bytes", ctypes.c_int64), ("keep_token", ctypes.c_int32),
    ]


def und_decode_step(args: DecodeStepArgs) -> None:
    """One greedy decode step issued natively (see g2vlm_und_decode_step): one persistent kernel when args.fused_ws
    is set, ~280 launches otherwise."""
    _check(_lib.load().g2vlm_und_decode_step(ctypes.byref(args), _stream()))


class UndPrefillArgs(ctypes.Structure):
    _fields_ = [
        ("num_layers", ctypes.c_int32), ("hidden", ctypes.c_int32), ("intermediate", ctypes.c_int32),
        ("n_q_heads", ctypes.c_int32), ("n_kv_heads", ctypes.c_int32), ("head_dim", ctypes.c_int32),
        ("rms_eps", ctypes.c_float), ("mrope_s0", ctypes.c_int32), ("mrope_s1", ctypes.c_int32),
        ("layers", ctypes.POINTER(UndLayerWeights)), ("kv", ctypes.POINTER(ctypes.c_void_p)),
        ("kv_capacity", ctypes.c_int64), ("cache_len", ctypes.c_int64), ("rows", ctypes.c_int32), ("causal", ctypes.c_int32),
        ("final_norm", ctypes.c_void_p), ("inv_freq", ctypes.c_void_p), ("position_ids", ctypes.c_void_p),
        ("work", ctypes.c_void_p), ("n_items", ctypes.c_int32),
        ("x", ctypes.c_void_p), ("y", ctypes.c_void_p), ("h", ctypes.c_void_p), ("qkv", ctypes.c_void_p),
        ("attn", ctypes.c_void_p), ("act", ctypes.c_void_p), ("cos", ctypes.c_void_p), ("sin", ctypes.c_void_p),
    ]


def und_prefill(args: UndPrefillArgs) -> None:
    """Text prefill / ViT step of the und expert (T >= 2 rows) issued natively (see g2vlm_und_prefill)."""
    _check(_lib.load().g2vlm_und_prefill(ctypes.byref(args), _stream()))


def und_decode_workspace_bytes(n_q_heads: int, n_kv_heads: int) -> int:
    """Bytes of the zero-initialised workspace of the one-kernel decode step (current device)."""
    fn = _lib.load().g2vlm_und_decode_workspace_bytes
    fn.restype = ctypes.c_int64
    return int(fn(_i32(n_q_heads), _i32(n_kv_heads)))


def rope_vision(buf, rows, n_heads_total, head_stride, head_dim, cos, sin):
    _req(buf, torch.bfloat16, "buf")
    _req(cos, torch.float32, "cos")
    _req(sin, torch.float32, "sin")
    _call("g2vlm_rope_vision", _vp(buf.data_ptr()), _i64(buf.stride(0)), _i64(rows), _i32(n_heads_total),
          _i32(head_stride), _i32(head_dim), _vp(cos.data_ptr()), _vp(sin.data_ptr()))


def resize_lanczos_u8(src: torch.Tensor, htab, vtab, out_h: int, out_w: int, out_u8: Optional[torch.Tensor] = None,
                      out_f32: Optional[torch.Tensor] = None):
    """Pillow-exact LANCZOS resize of one uint8 [H, W, 3] image on the device (see g2vlm_resize_lanczos_u8).
    htab / vtab: (bounds, coef) int32 device tensors from host_prep.lanczos_tables, or None when that axis keeps
    its size.  out_u8 uint8 [out_h, out_w, 3] and/or out_f32 fp32 [3, out_h, out_w]."""
    if src.dtype != torch.uint8 or not src.is_cuda or src.dim() != 3 or src.shape[2] != 3 or src.stride(2) != 1 \
            or src.stride(1) != 3:
        raise G2Error("resize_lanczos_u8: src must be a CUDA uint8 [H, W, 3] tensor with packed pixels")
    H, W = int(src.shape[0]), int(src.shape[1])
    for t, shape, dt, n in ((out_u8, (out_h, out_w, 3), torch.uint8, "out_u8"), (out_f32, (3, out_h, out_w), torch.float32, "out_f32")):
        if t is not None and (tuple(t.shape) != shape or t.dtype != dt or not t.is_cuda or not t.is_contiguous()):
            raise G2Error(f"resize_lanczos_u8: {n} must be a contiguous CUDA {dt} tensor of shape {shape}")
    tmp = torch.empty(H, out_w, 3, dtype=torch.uint8, device=src.device) if htab is not None else None
    hb, hc = htab if htab is not None else (None, None)
    vb, vc = vtab if vtab is not None else (None, None)
    for t in (hb, hc, vb, vc):
        if t is not None and (t.dtype != torch.int32 or not t.is_cuda or not t.is_contiguous()):
            raise G2Error("resize_lanczos_u8: tables must be contiguous CUDA int32 tensors")
    _call("g2vlm_resize_lanczos_u8", _vp(src.data_ptr()), _i32(H), _i32(W), _i64(src.stride(0)), _ptr(hb), _ptr(hc),
          _i32(hc.shape[1] if hc is not None else 0), _ptr(vb), _ptr(vc), _i32(vc.shape[1] if vc is not None else 0),
          _ptr(tmp), _i32(out_h), _i32(out_w), _ptr(out_u8), _ptr(out_f32))
    return out_f32 if out_f32 is not None else out_u8


# ---------------------------------------------------------------------------------------------------
# fp32 mode (see include/g2vlm_b200.h): fp32 operands, no bf16 rounding point
# ---------------------------------------------------------------------------------------------------
def attention_f32(q, k, v, out, work, *, num_q_heads: int, num_kv_heads: int, head_dim: int, scale: float,
                  causal: bool = False) -> torch.Tensor:
    """fp32 counterpart of `attention` (same segment / work-table semantics), head_dim in {16,32,64,96,128}."""
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        _req(t, torch.float32, n)
    _req(work, torch.int32, "work")
    if work.dim() != 2 or work.shape[1] != 8 or not work.is_contiguous():
        raise G2Error("attention_f32: work table must be a contiguous int32 [n, 8] tensor")
    args = AttnArgs()
    args.q, args.ldq, args.q_rows = q.data_ptr(), q.stride(0), q.shape[0]
    args.k, args.ldk = k.data_ptr(), k.stride(0)
    args.v, args.ldv, args.kv_rows = v.data_ptr(), v.stride(0), k.shape[0]
    args.out, args.ldo = out.data_ptr(), out.stride(0)
    args.num_q_heads, args.num_kv_heads, args.head_dim = num_q_heads, num_kv_heads, head_dim
    args.causal, args.softmax_scale = int(causal), float(scale)
    args.n_items, args.work_items = work.shape[0], work.data_ptr()
    _check(_lib.load().g2vlm_attention_f32(ctypes.byref(args), _stream()))
    return out


def im2col_patches_f32(images, out, patch: int, mean=None, std=None):
    _req(images, torch.float32, "images")
    _req(out, torch.float32, "out")
    n, _, H, W = images.shape
    if not images.is_contiguous() or not out.is_contiguous():
        raise G2Error("im2col_f32: contiguous tensors required")
    m3 = (ctypes.c_float * 3)(*mean) if mean is not None else None
    s3 = (ctypes.c_float * 3)(*std) if std is not None else None
    _call("g2vlm_im2col_patches_f32", _vp(images.data_ptr()), _vp(out.data_ptr()), _i32(n), _i32(H), _i32(W),
          _i32(patch), _i32(out.shape[1]), m3, s3)
    return out


def dino_embed_f32(patch_emb, cls, reg, pos, out, n: int, P: int, n_reg: int):
    _req(patch_emb, torch.float32, "patch_emb")
    _req(out, torch.float32, "out")
    _call("g2vlm_dino_embed_f32", _vp(patch_emb.data_ptr()), _i64(patch_emb.stride(0)), _vp(cls.data_ptr()),
          _vp(reg.data_ptr()), _vp(pos.data_ptr()), _vp(out.data_ptr()), _i32(n), _i32(P), _i32(n_reg),
          _i32(out.shape[1]))
    return out


def qknorm_mrope_f32(qkv, rows, n_first, n_q, n_kv, head_dim, qw_a, kw_a, qw_b, kw_b, cos, sin, eps):
    _req(qkv, torch.float32, "qkv")
    _call("g2vlm_qknorm_mrope_f32", _vp(qkv.data_ptr()), _i64(qkv.stride(0)), _i64(rows), _i64(n_first),
          _i32(n_q), _i32(n_kv), _i32(head_dim), _vp(qw_a.data_ptr()), _vp(kw_a.data_ptr()),
          _vp(qw_b.data_ptr()), _vp(kw_b.data_ptr()), _vp(cos.data_ptr()), _vp(sin.data_ptr()), _f32(eps))


def rope2d_f32(buf, rows, n_heads_total, head_stride, head_dim, tokens_per_view, grid_w, cos, sin):
    _req(buf, torch.float32, "buf")
    _call("g2vlm_rope2d_f32", _vp(buf.data_ptr()), _i64(buf.stride(0)), _i64(rows), _i32(n_heads_total),
          _i32(head_stride), _i32(head_dim), _i32(tokens_per_view), _i32(grid_w), _vp(cos.data_ptr()),
          _vp(sin.data_ptr()))


def swiglu_f32(gate_up, out, rows: Optional[int] = None):
    """out[r, c] = silu(gate_up[r, c]) * gate_up[r, I + c], I = out.shape[1]."""
    _req(gate_up, torch.float32, "gate_up")
    _req(out, torch.float32, "out")
    rows = out.shape[0] if rows is None else rows
    if gate_up.shape[1] != 2 * out.shape[1]:
        raise G2Error("swiglu_f32: gate_up must have 2 * out.shape[1] columns")
    _call("g2vlm_swiglu_f32", _vp(gate_up.data_ptr()), _i64(gate_up.stride(0)), _vp(out.data_ptr()), _i64(out.stride(0)),
          _i64(rows), _i32(out.shape[1]))
    return out


def split6(x, out, rows: Optional[int] = None):
    """fp32 [rows, k] -> bf16 [rows, 6k] = [h|h|m|h|l|m] (see g2vlm_split6_f32)."""
    _req(x, torch.float32, "x")
    _req(out, torch.bfloat16, "out")
    rows = x.shape[0] if rows is None else rows
    _call("g2vlm_split6_f32", _vp(x.data_ptr()), _i64(x.stride(0)), _vp(out.data_ptr()), _i64(out.stride(0)),
          _i64(rows), _i32(x.shape[1]))
    return out


# ---------------------------------------------------------------------------------------------------
# stack-level entry points (opaque context + one call per stage; see include/g2vlm_b200.h)
# ---------------------------------------------------------------------------------------------------
class Dims(ctypes.Structure):
    _fields_ = [
        ("hidden", ctypes.c_int32), ("layers", ctypes.c_int32), ("heads", ctypes.c_int32), ("kv_heads", ctypes.c_int32),
        ("intermediate", ctypes.c_int32), ("rms_eps", ctypes.c_float), ("mrope_s0", ctypes.c_int32), ("mrope_s1", ctypes.c_int32),
        ("dino_hidden", ctypes.c_int32), ("dino_layers", ctypes.c_int32), ("dino_heads", ctypes.c_int32),
        ("dino_mlp_ratio", ctypes.c_int32), ("dino_patch", ctypes.c_int32), ("dino_registers", ctypes.c_int32),
        ("dino_ln_eps", ctypes.c_float),
        ("dec_depth", ctypes.c_int32), ("dec_heads", ctypes.c_int32), ("dec_mlp_ratio", ctypes.c_int32),
        ("point_dim", ctypes.c_int32), ("camera_dim", ctypes.c_int32), ("train_conf", ctypes.c_int32),
    ]


class NativeContext:
    """Owner of one `g2vlm_ctx` (destroyed with the object); thin ctypes plumbing only."""

    def __init__(self, cfg):
        lib = _lib.load()
        lib.g2vlm_workspace_bytes.restype = ctypes.c_int64
        d = Dims(hidden=cfg.hidden_size, layers=cfg.num_layers, heads=cfg.num_heads, kv_heads=cfg.num_kv_heads,
                 intermediate=cfg.intermediate_size, rms_eps=cfg.rms_norm_eps, mrope_s0=cfg.mrope_section[0],
                 mrope_s1=cfg.mrope_section[1], dino_hidden=cfg.dino_hidden, dino_layers=cfg.dino_layers,
                 dino_heads=cfg.dino_heads, dino_mlp_ratio=cfg.dino_mlp_ratio, dino_patch=cfg.dino_patch,
                 dino_registers=cfg.dino_registers, dino_ln_eps=cfg.dino_ln_eps, dec_depth=cfg.dec_depth,
                 dec_heads=cfg.dec_heads, dec_mlp_ratio=cfg.dec_mlp_ratio, point_dim=cfg.point_dim,
                 camera_dim=cfg.camera_dim, train_conf=int(cfg.train_conf_pi3))
        self._h = ctypes.c_void_p()
        self._lib = lib
        self._keep = []          # registered tensors must outlive the context
        rc = lib.g2vlm_ctx_create(ctypes.byref(d), ctypes.byref(self._h))
        if rc != 0:
            raise G2Error(f"g2vlm_ctx_create failed (code {rc}): {lib.g2vlm_last_error().decode()}")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.g2vlm_ctx_destroy(h)

    def call(self, name: str, *args) -> None:
        rc = getattr(self._lib, name)(self._h, *args)
        if rc != 0:
            raise G2Error(f"{name} failed (code {rc}): {self._lib.g2vlm_last_error().decode()}")

    def load(self, name: str, t: torch.Tensor) -> None:
        if t.dtype not in (torch.float32, torch.bfloat16) or not t.is_cuda or not t.is_contiguous():
            raise G2Error(f"weight {name}: contiguous CUDA fp32 / bf16 tensor required")
        rows = t.shape[0] if t.dim() > 1 else 1
        self._keep.append(t)
        self.call("g2vlm_load_weights", name.encode(), _vp(t.data_ptr()), _i32(int(t.dtype == torch.bfloat16)), _i64(rows),
                  _i64(t.numel() // max(rows, 1)))

    def region(self, ws: torch.Tensor, name: str, dtype, cols: int) -> torch.Tensor:
        """View of a named intermediate inside the planned workspace (see g2vlm_workspace_region)."""
        off, nb = ctypes.c_int64(), ctypes.c_int64()
        self.call("g2vlm_workspace_region", name.encode(), ctypes.byref(off), ctypes.byref(nb))
        flat = ws[off.value: off.value + nb.value].view(dtype)
        return flat.view(-1, cols)

    def workspace_bytes(self, n_views: int, H: int, W: int, n_prompt: int) -> int:
        n = int(self._lib.g2vlm_workspace_bytes(self._h, _i32(n_views), _i32(H), _i32(W), _i32(n_prompt)))
        if n < 0:
            raise G2Error(f"g2vlm_workspace_bytes: {self._lib.g2vlm_last_error().decode()}")
        return n
