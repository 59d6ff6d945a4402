#!/usr/bin/env python
"""Row f1 measurement: decode tokens/s of generate_text on the full model after a 16-view 518px scene is in
the KV cache (L = 21 943 + prompt tokens), and the prefill times, for the ~280-launch step of round 1 and the one-kernel step.
usage: python tools/decode_bench.py [steps]
       G2VLM_DECODE_OPT_SWEEP=1,0,1,0 python tools/decode_bench.py [steps]   same-process A/B of G2VLM_DECODE_OPT values
       (bit 0 of G2VLM_DECODE_OPT: grid barrier without explicit fences, the default)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from g2vlm_b200 import schema
from g2vlm_b200.model import G2VLMFast, NaiveCache

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = schema.FULL
sd = schema.init_synthetic(cfg, seed=0, device="cuda")  # full 151936-row embedding and lm_head
model = G2VLMFast(cfg, sd)
del sd
torch.cuda.empty_cache()
ids = dict(bos_token_id=1, eos_token_id=2, start_of_image=3, end_of_image=4)
views = schema.synthetic_views(16, 518, 518, seed=1)


def text_inputs(n, kvlen, rope):
    return dict(text_token_lens=torch.tensor([n], dtype=torch.int), packed_text_ids=torch.arange(10, 10 + n),
                packed_text_position_ids=(rope + torch.arange(n)).expand(3, -1), packed_text_indexes=kvlen + torch.arange(n),
                packed_key_value_indexes=torch.arange(kvlen), key_values_lens=torch.tensor([kvlen], dtype=torch.int)), kvlen + n, rope + n


def run(graph=True, fused=True):
    past = NaiveCache(cfg.num_layers)
    gi, kvlen, rope = text_inputs(20, 0, 0)
    past = model.forward_cache_update_text(past, **gi)
    gi, nl, nr = model.prepare_dino_images_pi3([kvlen], [rope], views, None, ids)
    past, _ = model.forward_cache_update_dino(past, **gi)
    gi, kvlen, rope = text_inputs(60, nl[0], nr[0])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    past = model.forward_cache_update_text(past, **gi)
    torch.cuda.synchronize(); t_prefill = time.perf_counter() - t0
    t0 = time.perf_counter()
    out = model.generate_text(past, None, None, torch.tensor([7]), torch.full((3, 1), rope), steps, end_token_id=None,
                              use_cuda_graph=graph, fused_step=fused)
    torch.cuda.synchronize(); t_dec = time.perf_counter() - t0
    return t_prefill, t_dec, past.seq_lens, out


run()
if os.environ.get("G2VLM_DECODE_OPT_SWEEP"):
    for opt in os.environ["G2VLM_DECODE_OPT_SWEEP"].split(","):
        os.environ["G2VLM_DECODE_OPT"] = opt
        run(graph=False, fused=True)
        tp, td, L, out = run(graph=False, fused=True)
        print(f"G2VLM_DECODE_OPT={opt}: {td / steps * 1e3:.3f} ms/token")
    sys.exit(0)
for fused in (False, True):
    name = "one persistent kernel per step" if fused else "~280 launches per step"
    tp, td, L, out = run(graph=False, fused=fused)
    print(f"[{name}] eager: {td / steps * 1e3:.3f} ms/token")
    tp, td, L, out = run(fused=fused)
    gb = 2.62 + 0.47 + L * 28 * 1024 / 1e9
    print(f"[{name}] CUDA graph: cache length {L}; 60-token question prefill {tp * 1e3:.2f} ms; decode {steps} tokens in "
          f"{td * 1e3:.1f} ms = {td / steps * 1e3:.3f} ms/token = {steps / td:.1f} tokens/s (batch 1, greedy, lm_head "
          f"{cfg.vocab_size}-way); {gb:.2f} GB per token -> {gb / (td / steps) / 1e3:.2f} TB/s")
