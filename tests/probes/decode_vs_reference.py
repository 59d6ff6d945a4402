#!/usr/bin/env python
"""Row f.1 next to the reference itself: greedy `generate_text` of the UNMODIFIED reference classes (oracle/_ref, real
flash-attn, CUDA autocast, its per-step NaiveCache re-allocation) and of the CUDA path on the same B200, the same
weights and the same KV cache.   python tests/probes/decode_vs_reference.py [cache_rows] [tokens]
Test infrastructure (uses oracle/)."""
import contextlib
import io
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from g2vlm_b200 import schema  # noqa: E402
from g2vlm_b200.model import G2VLMFast  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 22144
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 32
torch.cuda.set_device(0)
cfg = schema.FULL
sd = schema.init_synthetic(cfg, seed=0, device="cuda")
ref = rh.build_reference_model(rh.dims_from_cfg(cfg), visual_und=False, device="cuda")
msg = ref.load_state_dict(sd, strict=False)
print(f"reference on the GPU: missing {len(msg.missing_keys)} unexpected {len(msg.unexpected_keys)} keys", flush=True)
from modeling.g2vlm.qwen2vl import NaiveCache as RefCache  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(1)
kv = [(torch.randn(L, cfg.num_kv_heads, cfg.head_dim, generator=g, device="cuda").to(torch.bfloat16),
       torch.randn(L, cfg.num_kv_heads, cfg.head_dim, generator=g, device="cuda").to(torch.bfloat16)) for _ in range(cfg.num_layers)]


def ref_cache():
    c = RefCache(cfg.num_layers)
    for i, (k, v) in enumerate(kv):
        c.key_cache[i], c.value_cache[i] = k.clone(), v.clone()
    return c


def run_ref(n):
    c = ref_cache()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad(), torch.amp.autocast("cuda", dtype=torch.bfloat16), contextlib.redirect_stdout(io.StringIO()):
        ids = ref.generate_text(past_key_values=c, packed_key_value_indexes=torch.arange(L, device="cuda"),
                                key_values_lens=torch.tensor([L], dtype=torch.int, device="cuda"),
                                packed_start_tokens=torch.tensor([7], device="cuda"),
                                packed_query_position_ids=torch.full((3, 1), L, device="cuda"), max_length=n)
    torch.cuda.synchronize()
    return ids[:, 0].tolist(), (time.perf_counter() - t0) / n * 1e3


run_ref(2)
tok_ref, ms_ref = run_ref(steps)
print(f"reference (PyTorch 2.11 + flash-attn, bf16 autocast): {ms_ref:.2f} ms/token = {1e3 / ms_ref:.1f} tokens/s behind {L} cached rows", flush=True)
del ref
torch.cuda.empty_cache()
ours = G2VLMFast(cfg, sd)


def run_ours(n, fused):
    c = ref_cache()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ids = ours.generate_text(c, None, None, torch.tensor([7]), torch.full((3, 1), L), n, end_token_id=None,
                             use_cuda_graph=not fused, fused_step=fused)
    torch.cuda.synchronize()
    return ids[:, 0].tolist(), (time.perf_counter() - t0) / n * 1e3


for fused in (False, True):
    run_ours(4, fused)
    tok, ms = run_ours(steps, fused)
    same = sum(a == b for a, b in zip(tok, tok_ref))
    print(f"g2vlm_b200 ({'one persistent kernel per step' if fused else '~280 launches per step, CUDA graph'}): {ms:.3f} ms/token = "
          f"{1e3 / ms:.1f} tokens/s (includes adopting the reference's cache) = x{ms_ref / ms:.1f}; first {same} of {steps} tokens equal "
          f"the reference's (random weights: a near tie ends the common prefix)", flush=True)
