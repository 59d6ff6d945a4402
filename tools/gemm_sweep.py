#!/usr/bin/env python
"""Times the grouped GEMM at every shape the recon path uses (config 2) — TFLOP/s per call site."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from g2vlm_b200 import ops

g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g, device="cuda")
SHAPES = [  # name, M, K, N, epilogue, flags, bias
    ("mot.qkv", 21936, 1536, 2048, ops.EPI_STORE_BF16, 0, True),
    ("mot.o", 21936, 1536, 1536, ops.EPI_RESID_F32, ops.GEMM_ROUND_AFTER_SCALE, False),
    ("mot.gateup", 21936, 1536, 17920, ops.EPI_SWIGLU_BF16, 0, False),
    ("mot.down", 21936, 8960, 1536, ops.EPI_RESID_F32, ops.GEMM_ROUND_AFTER_SCALE, False),
    ("dino.qkv", 21984, 1024, 3072, ops.EPI_STORE_BF16, 0, True),
    ("dino.dense", 21984, 1024, 1024, ops.EPI_RESID_F32, 0, True),
    ("dino.fc1", 21984, 1024, 4096, ops.EPI_STORE_BF16, ops.GEMM_GELU, True),
    ("dino.fc2", 21984, 4096, 1024, ops.EPI_RESID_F32, 0, True),
    ("pi3.qkv", 21904, 1536, 6144, ops.EPI_STORE_BF16, 0, True),
    ("pi3.proj", 21904, 2048, 1536, ops.EPI_RESID_F32, 0, True),
    ("pi3.fc1", 21904, 1536, 6144, ops.EPI_STORE_BF16, ops.GEMM_GELU, True),
    ("pi3.fc2", 21904, 6144, 1536, ops.EPI_RESID_F32, 0, True),
    ("pi3.out", 21904, 1536, 1024, ops.EPI_STORE_BF16, 0, True),
    ("head.pts", 21904, 1024, 588, ops.EPI_STORE_F32, 0, True),
    ("cam.res", 21904, 1536, 512, ops.EPI_STORE_F32, ops.GEMM_RELU, True),
]
for name, M, K, N, epi, flags, has_bias in SHAPES:
    a = (rnd(M, K) * 0.5).to(torch.bfloat16)
    w = (rnd(N, K) * 0.05).to(torch.bfloat16)
    bias = rnd(N) if has_bias else None
    scale = torch.ones(N, device="cuda") if epi == ops.EPI_RESID_F32 else None
    if epi in (ops.EPI_STORE_BF16,):
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    elif epi == ops.EPI_SWIGLU_BF16:
        out = torch.empty(M, N // 2, device="cuda", dtype=torch.bfloat16)
    else:
        out = torch.zeros(M, N, device="cuda")
    fn = lambda: ops.gemm(a, w, out, epilogue=epi, bias=bias, scale=scale, scale_groups=1 if scale is not None else 0, flags=flags)
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:12s} M={M} K={K:5d} N={N:5d}  {ms*1000:8.1f} us  {2*M*K*N/ms/1e9:7.1f} TFLOP/s")
