#!/usr/bin/env python
"""How much of a recon step is GPU idle time between our kernels?  Wraps every C-ABI launch of one
device-resident step in CUDA events and compares sum(kernel durations) with the step time."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from g2vlm_b200 import ops, schema
from g2vlm_b200.model import G2VLMFast, NaiveCache

cfg = schema.FULL
sd = schema.init_synthetic(cfg, seed=0, embed_rows=32, device="cuda")
model = G2VLMFast(cfg, sd)
del sd
views = schema.synthetic_views(16, 518, 518, seed=1)


class Tok:
    def encode(self, p):
        return [11, 12, 13, 14, 15, 16]


ids = dict(bos_token_id=1, eos_token_id=2, start_of_image=3, end_of_image=4)
gi_t, nl, nr = model.prepare_prompts_addbos([0], [0], ["x"], Tok(), ids)
gi_t = {k: (v if k.endswith('lens') else v.cuda()) for k, v in gi_t.items()}
gi, _, _ = model.prepare_dino_images_pi3(nl, nr, views, None, ids)
gi = {k: (v if k.endswith('lens') else v.cuda()) for k, v in gi.items()}


def step():
    past, last = model.forward_cache_update_dino(NaiveCache(cfg.num_layers), update_past_key_values=False,
                                                 prompt=gi_t, **gi)   # fused prompt prefill, as recon() does
    return model.reconstruct(past_key_values=past, selected_hidden_states=last, **gi)


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
plain = e0.elapsed_time(e1)

events = []
lib = ops._lib.load()
names = [n for n in ops._lib.declared_symbols() if n not in ("g2vlm_abi_version", "g2vlm_last_error")]
orig = {}


class Wrapped:
    def __init__(self, name, fn):
        self.name, self.fn = name, fn

    def __call__(self, *a):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r = self.fn(*a)
        e.record()
        events.append((self.name, s, e))
        return r


class LibProxy:
    def __getattr__(self, n):
        f = getattr(lib, n)
        return Wrapped(n, f) if n in names else f


ops._lib.load = lambda: LibProxy()
e0.record(); step(); e1.record(); torch.cuda.synchronize()
inst = e0.elapsed_time(e1)
agg = collections.defaultdict(float)
for n, s, e in events:
    agg[n] += s.elapsed_time(e)
tot = sum(agg.values())
print(f"step (plain) {plain:.2f} ms; instrumented {inst:.2f} ms; sum of {len(events)} kernel intervals {tot:.2f} ms; "
      f"idle/gap estimate {plain - tot:.2f} ms ({100 * (plain - tot) / plain:.1f}%)")
for n, t in sorted(agg.items(), key=lambda kv: -kv[1]):
    print(f"  {t:8.2f} ms  {n}")
