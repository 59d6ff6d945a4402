#!/usr/bin/env python
"""Stand-alone launcher of one hot kernel at its config-2 shape (for ncu --set full captures and
quick timing).  usage: python tools/run_kernel.py {attention|attention64|gateup|down|qkv|oproj|dino_dense|dino_fc1} [iters]"""
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
elif which == "attention64":     # DINO encoder attention: 16 views x 1374 rows, segments of 1369 (quirk Q1), 16 heads x 64
    N, P, S = 16, 1369, 1374
    qkv = rnd(N * S, 3 * 1024).to(torch.bfloat16)
    out = torch.zeros(N * S, 1024, device="cuda", dtype=torch.bfloat16)
    cu = [i * P for i in range(N + 1)]
    work = ops.attention_work_table(cu, cu).cuda()
    fn = lambda: ops.attention(qkv[:, :1024], qkv[:, 1024:2048], qkv[:, 2048:], out, work, num_q_heads=16,
                               num_kv_heads=16, head_dim=64, scale=1 / math.sqrt(64))
    flops = 4 * N * P * P * 16 * 64
elif which == "oproj":           # routed o_proj + LayerScale + residual (TMA reduce-add epilogue)
    a = (rnd(T, H) * 0.5).to(torch.bfloat16); w = (rnd(2 * H, H) * 0.05).to(torch.bfloat16)
    out = torch.zeros(T, H, device="cuda")
    sc = torch.ones(H, device="cuda")
    fn = lambda: ops.gemm(a, w, out, epilogue=ops.EPI_RESID_F32, groups=groups, scale=sc, scale_groups=1,
                          flags=ops.GEMM_ROUND_AFTER_SCALE)
    flops = 2 * T * H * H
elif which == "dino_dense":      # DINO attention.output.dense + lambda1 + residual (K = N = 1024)
    M = 16 * 1374
    a = (rnd(M, 1024) * 0.5).to(torch.bfloat16); w = (rnd(1024, 1024) * 0.05).to(torch.bfloat16)
    b, sc = rnd(1024), torch.ones(1024, device="cuda")
    out = torch.zeros(M, 1024, device="cuda")
    fn = lambda: ops.gemm(a, w, out, epilogue=ops.EPI_RESID_F32, bias=b, scale=sc, scale_groups=1)
    flops = 2 * M * 1024 * 1024
elif which == "dino_fc1":        # DINO mlp.fc1 + exact-erf GELU
    M = 16 * 1374
    a = (rnd(M, 1024) * 0.5).to(torch.bfloat16); w = (rnd(4096, 1024) * 0.05).to(torch.bfloat16)
    b = rnd(4096)
    out = torch.empty(M, 4096, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.gemm(a, w, out, epilogue=ops.EPI_STORE_BF16, bias=b, flags=ops.GEMM_GELU)
    flops = 2 * M * 1024 * 4096
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
