#!/usr/bin/env python
"""Where a one-kernel decode step spends its time: SM cycles of CTA 0 and of the last CTA per phase and per grid barrier
(und_decode_fused_kernel<TIMING>), summed over the layers of ONE step behind a 16-view scene in the KV cache.
usage: python tools/decode_phase_trace.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from g2vlm_b200 import schema
from g2vlm_b200.model import G2VLMFast, NaiveCache

modes = [0]
cfg = schema.FULL
model = G2VLMFast(cfg, schema.init_synthetic(cfg, seed=0, device="cuda"))
torch.cuda.empty_cache()
ids = dict(bos_token_id=1, eos_token_id=2, start_of_image=3, end_of_image=4)
views = schema.synthetic_views(16, 518, 518, seed=1)
gi, nl, nr = model.prepare_dino_images_pi3([0], [0], views, None, ids)
NAMES = ["qkv", "B", "attn", "B", "merge", "B", "o_proj", "B", "gate/up", "B", "down", "B", "lm_head", "B"]
for mode in modes:
    os.environ["G2VLM_DECODE_TRACE"] = "1"
    past, _ = model.forward_cache_update_dino(NaiveCache(cfg.num_layers), **gi)
    model.generate_text(past, None, None, torch.tensor([7]), torch.full((3, 1), nr[0]), 6, end_token_id=None,
                        use_cuda_graph=False, fused_step=True)
    torch.cuda.synchronize()
    ws = [t for k, t in model.buf._b.items() if k[0] == "dec.fused_ws"][0]
    t = ws[64:64 + 30 * 8].view(torch.int64).cpu().view(2, 15).double()
    print(f"cache length {past.seq_lens}; one step, phases summed over {cfg.num_layers} layers")
    gu = ws[64 + 30 * 8:64 + 35 * 8].view(torch.int64).cpu().tolist()
    n = max(gu[2], 1)
    print(f"  gate/up of CTA 0, warp 0: {gu[2]} stages; mean cycles per stage: wait {gu[0] / n:.0f}, multiply + release "
          f"{gu[1] / n:.0f}; cycles per layer: RMSNorm {gu[3] / cfg.num_layers:.0f}, all stages + sync {gu[4] / cfg.num_layers:.0f}")
    at = ws[64 + 35 * 8:64 + 41 * 8].view(torch.int64).cpu().tolist()
    print("  attention of CTA 0, cycles per layer: " + "  ".join(f"{n} {v / cfg.num_layers:.0f}" for n, v in zip(
        ("q/k-norm + rope", "wait for K|V", "scores", "softmax", "P.V", "reduce + store"), at)))
    for c, who in enumerate(("CTA 0", "last CTA")):
        cyc, ns = t[c, :14], t[c, 14].item()
        mhz = cyc.sum().item() / ns * 1e3
        print(f"  {who}: {ns / 1e3:.0f} us at {mhz:.0f} MHz; us per phase: " +
              "  ".join(f"{n} {v / mhz:.0f}" for n, v in zip(NAMES, cyc.tolist())))
