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


class G2Error(RuntimeError):
    pass


def _check(rc: int) -> None:
    if rc != 0:
        msg = _lib.load().g2vlm_last_error().decode()
        raise G2Error(f"g2vlm_b200 C ABI call failed (code {rc}): {msg}")


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _req(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise G2Error(f"{name} must be a CUDA tensor (there is no CPU fallback)")
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
    ]


def gemm(a: torch.Tensor, w: torch.Tensor, out: torch.Tensor, *, epilogue: int,
         groups: Optional[Sequence[tuple]] = None, bias: Optional[torch.Tensor] = None,
         scale: Optional[torch.Tensor] = None, scale_groups: int = 0, flags: int = 0,
         residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out <- epilogue(a @ w_g.T) per token group; see g2vlm_gemm_bf16 in include/g2vlm_b200.h.

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
              causal: bool = False) -> torch.Tensor:
    """out[rows covered by `work`] = softmax(scale * q k^T) v, per segment; q/k/v/out are 2-D bf16
    views [rows, heads*head_dim] (they may be column slices of one fused QKV buffer)."""
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
    _check(_lib.load().g2vlm_attention(ctypes.byref(args), _stream()))
    return out
