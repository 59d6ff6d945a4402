#!/usr/bin/env python
"""HBM-bound kernels of the recon step at their config-2 shapes: time (CUDA events, L2 flushed by the working set
of the loop) and achieved algorithmic GB/s against the measured copy bandwidth in MEASURED_PEAKS.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from g2vlm_b200 import ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6553.9
T, H, D = 16 * 1371, 1536, 1024
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g, device="cuda")
NB = 6   # rotate over NB buffer sets (> L2) so every launch streams from HBM


def timed(fn, iters=30):
    for i in range(NB):
        fn(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % NB)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def report(name, us, nbytes):
    print(f"{name:34s} {us:7.1f} us  {nbytes / us / 1e3:7.0f} GB/s  {nbytes / us / 1e3 / PEAK:5.2f} of copy peak")


x = [rnd(T, H) for _ in range(NB)]
ob = [torch.empty(T, H, device="cuda", dtype=torch.bfloat16) for _ in range(NB)]
wa, wb = torch.rand(H, device="cuda") + 0.5, torch.rand(H, device="cuda") + 0.5
report("rmsnorm_routed 21936x1536 f32->bf16", timed(lambda i: ops.rmsnorm_routed(x[i], ob[i], wa, wb, T - 39, 1e-6)), T * H * 6)
bb = torch.randn(H, device="cuda")
report("layernorm 21936x1536 f32->bf16", timed(lambda i: ops.layernorm(x[i], ob[i], wa, bb, 1e-6)), T * H * 6)
xd = [rnd(16 * 1374, D) for _ in range(NB)]
od = [torch.empty(16 * 1374, D, device="cuda", dtype=torch.bfloat16) for _ in range(NB)]
report("layernorm 21984x1024 f32->bf16", timed(lambda i: ops.layernorm(xd[i], od[i], wa[:D].contiguous(), bb[:D].contiguous(), 1e-6)),
       16 * 1374 * D * 6)
qkv = [rnd(T, 2048).to(torch.bfloat16) for _ in range(NB)]
cos, sin = rnd(T, 64), rnd(T, 64)
ws = [torch.rand(128, device="cuda") + 0.5 for _ in range(4)]
report("qknorm_mrope 21936 x (12+2) heads", timed(lambda i: ops.qknorm_mrope(qkv[i], T, T - 39, 12, 2, 128, *ws, cos, sin, 1e-6,
                                                                           round_normed=False)),
       T * 14 * 128 * 2 * 2 + T * 64 * 4 * 2)
