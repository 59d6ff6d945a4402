#!/usr/bin/env python
"""Back-to-back timing (no launch gaps) of the decode-step kernels at full-model shapes."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from g2vlm_b200 import ops

def timeit(fn, n=200):
    for _ in range(5): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000

g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g, device="cuda")
L, nq, nkv, hd = 22000, 12, 2, 128
# 28 distinct cache buffers so the kernel streams from HBM like the real step
bufs = [rnd(L + 64, 512).to(torch.bfloat16) for _ in range(28)]
q = rnd(nq * hd).to(torch.bfloat16); out = torch.empty(nq * hd, device="cuda", dtype=torch.bfloat16)
ws = torch.empty(ops.attention_decode_workspace_floats(L + 64, nq), device="cuda")
i = [0]
def attn():
    b = bufs[i[0] % 28]; i[0] += 1
    ops.attention_decode(q, b[:L, :256], b[:L, 256:], out, ws, num_q_heads=nq, num_kv_heads=nkv, head_dim=hd, scale=1/math.sqrt(hd))
us = timeit(attn)
print(f"attention_decode L={L}: {us:.1f} us  ({L*1024/us/1e6:.2f} TB/s of K+V)")
for name, N, K, epi in (("qkv", 2048, 1536, ops.EPI_STORE_BF16), ("o", 1536, 1536, ops.EPI_RESID_F32),
                        ("gateup", 17920, 1536, ops.EPI_SWIGLU_BF16), ("down", 1536, 8960, ops.EPI_RESID_F32),
                        ("lm_head", 151936, 1536, ops.EPI_STORE_BF16)):
    ws_ = [(rnd(N, K) * 0.05).to(torch.bfloat16) for _ in range(8 if N < 100000 else 2)]
    x = rnd(1, K).to(torch.bfloat16)
    n_out = N // 2 if epi == ops.EPI_SWIGLU_BF16 else N
    o = torch.zeros(1, n_out, device="cuda", dtype=torch.bfloat16 if epi in (ops.EPI_STORE_BF16, ops.EPI_SWIGLU_BF16) else torch.float32)
    j = [0]
    def gv():
        w = ws_[j[0] % len(ws_)]; j[0] += 1
        ops.gemm(x, w, o, epilogue=epi)
    us = timeit(gv, 100)
    print(f"gemv {name:8s} N={N:6d} K={K:5d}: {us:7.1f} us  ({N*K*2/us/1e6:.2f} TB/s of weights)")
