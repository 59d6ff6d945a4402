#!/usr/bin/env python
"""Time the 7-token prompt prefill of the recon path (und expert, 28 layers) on its own.
usage: python tools/text_prefill_bench.py [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from g2vlm_b200 import ops, schema
from g2vlm_b200.model import G2VLMFast, NaiveCache

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20


class Tok:
    def encode(self, prompt):
        return [11, 12, 13, 14, 15]


TOKENS = dict(bos_token_id=1, eos_token_id=2, start_of_image=3, end_of_image=4)
cfg = schema.FULL
model = G2VLMFast(cfg, schema.init_synthetic(cfg, seed=0, embed_rows=32, device="cuda"))
gi, _, _ = model.prepare_prompts_addbos([0], [0], ["x"], Tok(), TOKENS)
gi = {k: (v if k.endswith('lens') else v.cuda()) for k, v in gi.items()}
fn = lambda: model.forward_cache_update_text(NaiveCache(cfg.num_layers), **gi)
for _ in range(3):
    fn()
torch.cuda.synchronize()
n0 = ops.LAUNCHES
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    fn()
e1.record()
torch.cuda.synchronize()
print(f"text prefill ({gi['packed_text_ids'].numel()} tokens): {e0.elapsed_time(e1) / iters:.3f} ms per call, "
      f"{(ops.LAUNCHES - n0) // iters} launches; und-expert weights 2.62 GB -> "
      f"{2.62 / (e0.elapsed_time(e1) / iters):.2f} TB/s")
import time
t0 = time.perf_counter()
for _ in range(iters):
    fn()
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"host enqueue time per call: {(t1 - t0) / iters * 1e3:.3f} ms")
# GPU time without the host in the way: keep the GPU busy with filler work while the host enqueues the calls
big = torch.randn(16384, 16384, device="cuda", dtype=torch.bfloat16)
torch.cuda.synchronize()
for _ in range(40):          # ~0.25 s of filler
    big @ big
e0.record()
for _ in range(3):           # 768 launches: fits the driver's launch queue, so the host never blocks on the GPU
    fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"GPU time per call (host enqueue hidden behind filler work): {ms:.3f} ms -> {2.62 / ms:.2f} TB/s of und-expert weights")
