#!/usr/bin/env python
"""Stand-alone launcher of one hot kernel at its config-2 shape (for ncu --set full captures and
quick timing).  usage: python tools/run_kernel.py {attention|gateup|down|qkv} [iters]"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from g2vlm_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "attention"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
T, K0, H, I = 16 * 1371, 7, 1536, 8960
n_geo = 16 * 1369
groups = [(0, n_geo), (n_geo, T - n_geo)]
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g, device="cuda")

if which == "attention":
    qkv = rnd(T + K0, 2048).to(torch.bfloat16)
    out = torch.empty(T, 1536, device="cuda", dtype=torch.bfloat16)
    work = ops.attention_work_table([0, T], [0, T + K0]).cuda()
    fn = lambda: ops.attention(qkv[:T, :1536], qkv[:, 1536:1792], qkv[:, 1792:], out, work, num_q_heads=12,
                               num_kv_heads=2, head_dim=128, scale=1 / math.sqrt(128))
    flops = 4 * T * (T + K0) * 12 * 128
elif which == "gateup":
    a = (rnd(T, H) * 0.5).to(torch.bfloat16); w = (rnd(2 * 2 * I, H) * 0.05).to(torch.bfloat16)
    out = torch.empty(T, I, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.gemm(a, w, out, epilogue=ops.EPI_SWIGLU_BF16, groups=groups)
    flops = 2 * T * H * 2 * I
elif which == "down":
    a = (rnd(T, I) * 0.5).to(torch.bfloat16); w = (rnd(2 * H, I) * 0.05).to(torch.bfloat16)
    out = torch.zeros(T, H, device="cuda")
    sc = torch.ones(H, device="cuda")
    fn = lambda: ops.gemm(a, w, out, epilogue=ops.EPI_RESID_F32, groups=groups, scale=sc, scale_groups=1,
                          flags=ops.GEMM_ROUND_AFTER_SCALE)
    flops = 2 * T * H * I
else:
    a = (rnd(T, H) * 0.5).to(torch.bfloat16); w = (rnd(2 * 2048, H) * 0.05).to(torch.bfloat16)
    b = rnd(2 * 2048)
    out = torch.empty(T, 2048, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.gemm(a, w, out, epilogue=ops.EPI_STORE_BF16, groups=groups, bias=b)
    flops = 2 * T * H * 2048

for _ in range(3):
    fn()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"{which}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s")
