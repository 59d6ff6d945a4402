#!/usr/bin/env python
"""BASELINE configs[4] measurement: spatial-reasoning chat, interleaved geometry + semantic tokens, prefill +
decode.  1 image: 518x518 geo step (DINO + MoT geo expert) + 756x756 ViT step (2916 patches -> 729 tokens, und
expert) + ~60-token question + 100 greedy decode steps; full G2VLM-2B-MoT + Qwen2-VL ViT, random-init weights.
usage: python tools/chat_bench.py [decode_steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from g2vlm_b200 import schema
from g2vlm_b200.model import G2VLMFast

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
cfg = schema.FULL_CHAT
sd = schema.init_synthetic(cfg, seed=0, device="cuda")
model = G2VLMFast(cfg, sd)
del sd
torch.cuda.empty_cache()
ids = dict(bos_token_id=151644, eos_token_id=151645, start_of_image=151652, end_of_image=151653)


class Tok:
    def encode(self, p, add_special_tokens=True):
        n = 20 if "system" in p else (3 if "your text" in p else 60)
        return list(range(1000, 1000 + n))

    def decode(self, t):
        return " ".join(str(int(i)) for i in t)


def vit_transform(images):
    g = torch.Generator().manual_seed(0)
    return torch.randn(54 * 54, 3 * 2 * 14 * 14, generator=g), torch.tensor([[1, 54, 54]])


view = schema.synthetic_views(1, 518, 518, seed=1)


def sync_t():
    torch.cuda.synchronize()
    return time.perf_counter()


def run():
    from g2vlm_b200.model import KVCache
    t = {}
    past = KVCache(cfg.num_layers, cfg.num_kv_heads, cfg.head_dim, model.device)
    t0 = sync_t()
    gi, nl, nr = model.prepare_prompts_pure_text([0], [0], ["system"], Tok(), ids)
    past = model.forward_cache_update_text(past, **gi); t["system_prompt"] = sync_t() - t0
    t0 = sync_t()
    gi, nl, nr = model.prepare_dino_images_pi3(nl, nr, view, None, ids)
    past, _ = model.forward_cache_update_dino(past, **gi); t["geo_step_1view"] = sync_t() - t0
    t0 = sync_t()
    gi, nl, nr = model.prepare_vit_images(nl, nr, [None], vit_transform, ids)
    past = model.forward_cache_update_vit(past, **gi); t["vit_step_729tok"] = sync_t() - t0
    t0 = sync_t()
    gi, nl, nr = model.prepare_prompts_pure_text(nl, nr, ["question"], Tok(), ids)
    past = model.forward_cache_update_text(past, **gi); t["question_60tok"] = sync_t() - t0
    st = model.prepare_start_tokens(nl, nr, Tok(), ids)
    t0 = sync_t()
    out = model.generate_text(past_key_values=past, max_length=steps, end_token_id=None, **st)
    t["decode"] = sync_t() - t0
    return t, past.seq_lens


run()
t, L = run()
pre = sum(v for k, v in t.items() if k != "decode")
print("chat (configs[4]) on one B200: " + ", ".join(f"{k} {v * 1e3:.2f} ms" for k, v in t.items() if k != "decode"))
print(f"prefill total {pre * 1e3:.2f} ms for {L - steps} cached tokens; decode {steps} tokens: {t['decode'] / steps * 1e3:.3f} ms/token "
      f"= {steps / t['decode']:.1f} tokens/s (batch 1, greedy)")
