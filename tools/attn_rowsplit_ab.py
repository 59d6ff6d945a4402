"""A/B of the two-threads-per-row attention kernel (attention_rowsplit_kernel, G2VLM_ATTN_ROWSPLIT) against the
one-thread-per-row kernel on the three attention shapes of config 2 (16 views of 518 px):
  mot   28 launches per step: T = 21 936 rows, 12 Q heads / 2 KV heads x 128, one segment of T + 7 keys
  pi3   20 launches per step: 16 views x 1369 rows, 16 heads of 96 computed in 128-wide slots, per-view segments
  dino  24 launches per step: 16 views x 1374 rows, 16 heads x 64, segments of 1369 (quirk Q1)
usage: python tools/attn_rowsplit_ab.py [iters]"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from g2vlm_b200 import ops

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
g = torch.Generator().manual_seed(0)


def shape(name):
    if name == "mot":
        T, K0 = 16 * 1371, 7
        qkv = torch.randn(T + K0, 2048, generator=g).to(torch.bfloat16).cuda()
        out = torch.zeros(T, 1536, device="cuda", dtype=torch.bfloat16)
        work = ops.attention_work_table([0, T], [0, T + K0]).cuda()
        kw = dict(num_q_heads=12, num_kv_heads=2, head_dim=128, scale=1 / math.sqrt(128))
        return (qkv[:T, :1536], qkv[:, 1536:1792], qkv[:, 1792:], out, work), kw, 4 * T * (T + K0) * 12 * 128, 1
    if name == "pi3":
        N, P, H = 16, 1369, 16
        qkv = torch.randn(N * P, 3 * H * 128, generator=g).to(torch.bfloat16).cuda()
        out = torch.zeros(N * P, H * 96, device="cuda", dtype=torch.bfloat16)
        cu = [i * P for i in range(N + 1)]
        work = ops.attention_work_table(cu, cu).cuda()
        kw = dict(num_q_heads=H, num_kv_heads=H, head_dim=128, scale=1 / math.sqrt(96), out_head_cols=96)
        return (qkv[:, :H * 128], qkv[:, H * 128:2 * H * 128], qkv[:, 2 * H * 128:], out, work), kw, 4 * N * P * P * H * 96, 1
    N, P, S, H = 16, 1369, 1374, 16
    qkv = torch.randn(N * S, 3 * H * 64, generator=g).to(torch.bfloat16).cuda()
    out = torch.zeros(N * S, H * 64, device="cuda", dtype=torch.bfloat16)
    cu = [i * P for i in range(N + 1)]
    work = ops.attention_work_table(cu, cu).cuda()
    kw = dict(num_q_heads=H, num_kv_heads=H, head_dim=64, scale=1 / math.sqrt(64))
    return (qkv[:, :H * 64], qkv[:, H * 64:2 * H * 64], qkv[:, 2 * H * 64:], out, work), kw, 4 * N * P * P * H * 64, 2


for name in (sys.argv[2:] or ["mot", "pi3", "dino"]):
    args, kw, flops, bit = shape(name)
    ref = None
    modes = [0, bit, 0, bit]
    if os.environ.get("AB_MODES"):
        modes = [int(m) for m in os.environ["AB_MODES"].split(",")]
    for mode in modes:
        os.environ["G2VLM_ATTN_ROWSPLIT"] = str(mode)
        args[3].zero_()
        for _ in range(3):
            ops.attention(*args, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ops.attention(*args, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        if ref is None:
            ref = args[3].float().clone()
        err = ((args[3].float() - ref).abs().max() / ref.abs().max()).item()
        print(f"{name:5s} rowsplit={mode}: {ms * 1e3:8.1f} us  {flops / ms / 1e9:7.1f} TFLOP/s  max rel diff vs mode 0: {err:.2e}", flush=True)
