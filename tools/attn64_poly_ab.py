"""A/B of the FMA-pipe exp2 offload in attention_tcgen05_kernel<64> on the DINO shape of config 2
(16 views x (1369 + 5) rows, segments of 1369 per quirk Q1, 16 heads x 64): G2VLM_ATTN_POLY = 0..4 of every 8 exp2 pairs."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from g2vlm_b200 import ops

N, P, S, H, D = 16, 1369, 1374, 16, 64
rows = N * S
g = torch.Generator().manual_seed(0)
qkv = (torch.randn(rows, 3 * H * D, generator=g)).to(torch.bfloat16).cuda()
cu = [i * P for i in range(N + 1)]
work = ops.attention_work_table(cu, cu).cuda()
out = torch.zeros(rows, H * D, device="cuda", dtype=torch.bfloat16)
q, k, v = qkv[:, :H * D], qkv[:, H * D:2 * H * D], qkv[:, 2 * H * D:]
flops = 4 * N * P * P * H * D
ref = None
for poly in (0, 2, 0, 2):   # the library keeps the 0/8 and 2/8 instantiations (1, 3, 4 of 8 were measured in r02)
    os.environ["G2VLM_ATTN_POLY"] = str(poly)
    for _ in range(5):
        ops.attention(q, k, v, out, work, num_q_heads=H, num_kv_heads=H, head_dim=D, scale=1 / math.sqrt(D))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        ops.attention(q, k, v, out, work, num_q_heads=H, num_kv_heads=H, head_dim=D, scale=1 / math.sqrt(D))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    if ref is None:
        ref = out.float().clone()
    err = ((out.float() - ref).abs().max() / ref.abs().max()).item()
    print(f"poly {poly}/8: {ms * 1e3:7.1f} us  {flops / ms / 1e9:7.1f} TFLOP/s  max rel diff vs poly 0: {err:.2e}")
