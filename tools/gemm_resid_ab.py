"""A/B of the RESID_F32 epilogue: TMA reduce-add (default) vs SM-side read-modify-write (GEMM_NO_TMA_OUT) on the
residual GEMM shapes of config 2.  Burst timing (50 back-to-back launches after warm-up)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from g2vlm_b200 import ops

SHAPES = [("dino.dense", 21984, 1024, 1024, None), ("dino.fc2", 21984, 1024, 4096, None),
          ("mot.o_proj", 21943, 1536, 1536, [(0, 21904), (21904, 39)]), ("mot.down", 21943, 1536, 8960, [(0, 21904), (21904, 39)]),
          ("pi3.proj", 21904, 1536, 1536, None), ("pi3.fc2", 21904, 1536, 6144, None), ("cam.res", 21904, 512, 1536, None)]
g = torch.Generator().manual_seed(0)
for name, M, N, K, groups in SHAPES:
    ng = 1 if groups is None else 2
    a = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
    w = (torch.randn(ng * N, K, generator=g) * 0.05).to(torch.bfloat16).cuda()
    bias = torch.randn(ng * N, generator=g).cuda()
    gamma = (torch.rand(N, generator=g) + 0.5).cuda()
    x = torch.zeros(M, N, device="cuda")
    res = {}
    for tag, fl in (("tma", 0), ("rmw", ops.GEMM_NO_TMA_OUT), ("tma", 0), ("rmw", ops.GEMM_NO_TMA_OUT)):
        kw = dict(epilogue=ops.EPI_RESID_F32, groups=groups, bias=bias, scale=gamma, scale_groups=1, flags=fl | ops.GEMM_ROUND_AFTER_SCALE)
        for _ in range(5):
            ops.gemm(a, w, x, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            ops.gemm(a, w, x, **kw)
        e1.record(); torch.cuda.synchronize()
        res.setdefault(tag, []).append(e0.elapsed_time(e1) / 50 * 1e3)
    t, r = min(res["tma"]), min(res["rmw"])
    print(f"{name:11s} M={M} N={N} K={K}:  tma reduce-add {t:7.1f} us ({2 * M * N * K / t / 1e6:6.0f} TFLOP/s)   "
          f"rmw {r:7.1f} us ({2 * M * N * K / r / 1e6:6.0f} TFLOP/s)   {100 * (r - t) / r:+.1f} %")
