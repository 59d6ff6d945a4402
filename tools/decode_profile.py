#!/usr/bin/env python
"""Per-kernel time of ONE eager decode step on the full model with a 22k-token cache."""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from g2vlm_b200 import schema, ops
from g2vlm_b200.model import G2VLMFast, KVCache
cfg = schema.FULL
sd = schema.init_synthetic(cfg, seed=0, embed_rows=4096, device="cuda")
model = G2VLMFast(cfg, sd); del sd
cache = KVCache(cfg.num_layers, cfg.num_kv_heads, cfg.head_dim, model.device)
L = 22000
cache.reserve(L + 64)
for b in cache.buf:
    b.normal_()
cache.len = L
lib = ops._lib.load()
names = [n for n in ops._lib.declared_symbols() if n not in ("g2vlm_abi_version", "g2vlm_last_error")]
events = []
class Wrapped:
    def __init__(self, name, fn): self.name, self.fn = name, fn
    def __call__(self, *a):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r = self.fn(*a); e.record(); events.append((self.name, a, s, e)); return r
class LibProxy:
    def __getattr__(self, n):
        f = getattr(lib, n)
        return Wrapped(n, f) if n in names else f
model.generate_text(cache, None, None, torch.tensor([7]), torch.full((3, 1), 700), 4, use_cuda_graph=False)
ops._lib.load = lambda: LibProxy()
cache.len = L
model.generate_text(cache, None, None, torch.tensor([7]), torch.full((3, 1), 700), 1, use_cuda_graph=False)
torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for n, a, s, e in events:
    key = n
    if n == "g2vlm_gemm_bf16":
        ga = a[0]._obj
        key = f"gemv N={ga.N} K={ga.K} epi={ga.epilogue}"
    agg[key][0] += 1; agg[key][1] += s.elapsed_time(e)
tot = sum(t for _, t in agg.values())
print(f"one decode step: {tot:.3f} ms in {len(events)} launches")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {t*1000:8.1f} us total  n={c:3d}  {t/c*1000:7.1f} us each  {n}")
