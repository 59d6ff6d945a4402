#!/usr/bin/env python
"""Sustained (power-capped) throughput of the grouped tcgen05 GEMM against torch.matmul (cuBLAS) on the MoT shapes:
each candidate runs back to back for ~2 s, the second half is timed.  cuBLAS here is a yardstick only (it is not on
the product path): it tells how much a 2-CTA / lower-energy kernel could still gain under the power cap."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from g2vlm_b200 import ops

g = torch.Generator(device="cuda").manual_seed(0)
T = 16 * 1371


def sustained(fn, flops, seconds=2.0):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds / 2:     # heat up
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds / 2:
        for _ in range(10):
            fn()
        n += 10
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    return ms, flops / ms / 1e9


for name, K, N in (("gate/up (as plain GEMM)", 1536, 17920), ("down", 8960, 1536), ("qkv", 1536, 2048)):
    a = (torch.randn(T, K, generator=g, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g, device="cuda") * 0.05).to(torch.bfloat16)
    out = torch.empty(T, N, device="cuda", dtype=torch.bfloat16)
    flops = 2 * T * K * N
    ms1, tf1 = sustained(lambda: ops.gemm(a, w, out, epilogue=ops.EPI_STORE_BF16), flops)
    ms2, tf2 = sustained(lambda: torch.matmul(a, w.t(), out=out), flops)
    print(f"{name:26s} ours {ms1 * 1e3:7.1f} us {tf1:6.0f} TFLOP/s | cuBLAS {ms2 * 1e3:7.1f} us {tf2:6.0f} TFLOP/s")
