#!/usr/bin/env python
"""A few one-kernel decode steps of the full model behind a 22k-row cache — the command the ncu capture of
und_decode_fused_kernel profiles (profiles/r02_ncu_decode_fused.txt).  usage: python tools/run_decode_step.py [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from g2vlm_b200 import schema
from g2vlm_b200.model import G2VLMFast, KVCache

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = schema.FULL
model = G2VLMFast(cfg, schema.init_synthetic(cfg, seed=0, device="cuda"))
torch.cuda.empty_cache()
cache = KVCache(cfg.num_layers, cfg.num_kv_heads, cfg.head_dim, model.device)
L = 22144
cache.reserve(L + steps + 8)
for b in cache.buf:
    b.normal_()
cache.len = L
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
model.generate_text(cache, None, None, torch.tensor([7]), torch.full((3, 1), L), 2, end_token_id=None, fused_step=True)
e0.record()
model.generate_text(cache, None, None, torch.tensor([7]), torch.full((3, 1), L + 2), steps, end_token_id=None, fused_step=True)
e1.record()
torch.cuda.synchronize()
print(f"{steps} one-kernel decode steps behind {L} cached rows: {e0.elapsed_time(e1) / steps:.3f} ms/token")
