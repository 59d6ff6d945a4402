#!/usr/bin/env python
"""Per-CTA fixed overhead of the attention kernel on short segments: time the Pi3-decoder shape (16 views x 1369
rows, 16 heads, head_dim 128 padded) and the DINO shape (1374 rows, 16 heads, head_dim 64) with 1x, 2x, 3x the keys
per segment; t = a + b * key_blocks separates the per-CTA prologue/epilogue (a) from the main loop (b)."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from g2vlm_b200 import ops

g = torch.Generator(device="cuda").manual_seed(0)
for name, N, P, heads, d in (("decoder d128", 16, 1369, 16, 128), ("dino d64", 16, 1374, 16, 64)):
    T = N * P
    q = torch.randn(T, heads * d, generator=g, device="cuda").to(torch.bfloat16)
    out = torch.empty_like(q)
    for mult in (1, 2, 3):
        kv = torch.randn(T * mult, 2 * heads * d, generator=g, device="cuda").to(torch.bfloat16)
        cu_q = [v * P for v in range(N + 1)]
        cu_k = [v * P * mult for v in range(N + 1)]
        work = ops.attention_work_table(cu_q, cu_k).cuda()
        fn = lambda: ops.attention(q, kv[:, : heads * d], kv[:, heads * d:], out, work, num_q_heads=heads,
                                   num_kv_heads=heads, head_dim=d, scale=1 / math.sqrt(d))
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        blocks = math.ceil(P * mult / 128)
        flops = 4 * N * P * P * mult * heads * d
        print(f"{name}: keys x{mult} ({blocks} blocks/CTA, {work.shape[0] * heads} CTAs): {ms * 1e3:.1f} us  "
              f"{flops / ms / 1e9:.0f} TFLOP/s")
