#!/usr/bin/env python
"""List every implicit host<->device synchronisation of one steady-state recon() call and of the device-resident
step (torch.cuda.set_sync_debug_mode('warn')).  A sync stops the host from enqueueing the next scene while this
one computes.  usage: python tools/find_syncs.py [--full]"""
import os
import sys
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from g2vlm_b200 import schema
from g2vlm_b200.model import G2VLMFast
from g2vlm_b200.serving import ReconServer


class Tok:
    def encode(self, prompt):
        return [11, 12, 13, 14, 15]


TOKENS = dict(bos_token_id=1, eos_token_id=2, start_of_image=3, end_of_image=4)
full = "--full" in sys.argv
cfg = schema.FULL if full else schema.TINY
model = G2VLMFast(cfg, schema.init_synthetic(cfg, seed=0, embed_rows=32, device="cuda"))
views = schema.synthetic_views(4, 518 if full else 70, 518 if full else 98, seed=1).pin_memory()
server = ReconServer(model, Tok(), TOKENS)
for _ in range(3):
    server.submit(views)
server.drain()
torch.cuda.set_sync_debug_mode("warn")
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter("always")
    server.submit(views)
    server.submit(views)
torch.cuda.set_sync_debug_mode("default")
server.drain()
seen = {}
for x in w:
    key = f"{os.path.relpath(x.filename)}:{x.lineno}"
    seen[key] = seen.get(key, 0) + 1
print(f"{len(w)} synchronising calls in two ReconServer.submit() calls")
for k, v in sorted(seen.items(), key=lambda kv: -kv[1]):
    print(f"  {v:4d} x {k}")
