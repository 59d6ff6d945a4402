#!/usr/bin/env python
"""One-off check (too slow for the test suite): FULL-depth, FULL-width G2VLM-2B-MoT (28 MoT + 24 DINO layers,
5 blocks per decoder) on N views of 518x518 — CUDA path vs the CPU oracle (bf16 mode), same weights.
usage: python tests/probes/full_depth_parity.py [n_views] [layerscale|-] [H W]   (N=1, 518x518: ~6 TFLOP on the CPU;
       8 294 518 = BASELINE configs[0], the reference's CPU-runnable shape)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from g2vlm_b200 import schema
from g2vlm_b200.model import G2VLMFast
from oracle import restate

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ls_value = float(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] != "-" else None   # e.g. 0.01 = the reference's LayerScale init
cfg = schema.FULL


class Tok:
    def encode(self, p):
        return [11, 12, 13, 14, 15, 16]


ids = dict(bos_token_id=1, eos_token_id=2, start_of_image=3, end_of_image=4)
t0 = time.time()
sd = schema.init_synthetic(cfg, seed=0, embed_rows=32, device="cuda")
if ls_value is not None:  # "released regime": LayerScale gammas at the reference's init value (qwen2vl.py:765-766)
    for k in sd:
        if k.endswith("ls1.gamma") or k.endswith("ls2.gamma") or k.endswith(".lambda1"):
            sd[k].fill_(ls_value)
model = G2VLMFast(cfg, sd)
sd_cpu = {k: v.cpu() for k, v in sd.items()}
del sd
torch.cuda.empty_cache()
HH, WW = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (518, 518)
v = (schema.synthetic_views(n, HH, WW, seed=1) * 255).round() / 255
c_out, c_ref = {}, {}
out = model.recon(Tok(), dict(ids), None, v, collect=c_out)
torch.cuda.synchronize()
print(f"setup + gpu {time.time() - t0:.1f}s", flush=True)
t0 = time.time()
torch.set_num_threads(os.cpu_count())
ref = restate.recon(sd_cpu, cfg, v, mode="bf16", collect=c_ref)
print(f"oracle {time.time() - t0:.1f}s", flush=True)
t0 = time.time()
c32 = {}
ref32 = restate.recon(sd_cpu, cfg, v, mode="fp32", collect=c32)
print(f"oracle fp32 {time.time() - t0:.1f}s", flush=True)
rel = lambda a, b: ((a.float().cpu() - b.float()).abs().max() / b.float().abs().max()).item()
res = {"n_views": n, "image_hw": [HH, WW], "layerscale": ls_value}
# distance of each bf16 implementation from the fp32 ground truth (the reference never runs fp32)
for k in ("local_points", "points", "global_points", "camera_poses"):
    res[f"cuda_vs_fp32.{k}"] = rel(out[k], ref32[k])
    res[f"oracle_bf16_vs_fp32.{k}"] = rel(ref[k], ref32[k])
res["cuda_vs_fp32.last_hidden"] = rel(c_out["last_hidden"], c32["last_hidden"])
res["oracle_bf16_vs_fp32.last_hidden"] = rel(c_ref["last_hidden"], c32["last_hidden"])
# log-depth (the head output before exp): the exp amplifies an absolute error dz into a relative error dz
lz = lambda p: p["local_points"][..., 2].float().cpu().log()
res["logz_abs_err.cuda_vs_oracle_bf16"] = (lz(out) - lz(ref)).abs().max().item()
res["logz_abs_err.oracle_bf16_vs_fp32"] = (lz(ref) - lz(ref32)).abs().max().item()
res["logz_range"] = [lz(ref).min().item(), lz(ref).max().item()]
res["dino_tokens"] = rel(c_out["dino_tokens"].view_as(c_ref["dino_tokens"]), c_ref["dino_tokens"])
for i in (0, 6, 13, 20, 27):
    res[f"mot{i}"] = rel(c_out["mot_layers"][i], c_ref["mot_layers"][i])
res["last_hidden"] = rel(c_out["last_hidden"], c_ref["last_hidden"])
for k in ("point_hidden", "camera_hidden", "global_hidden"):
    res[k] = rel(c_out[k], c_ref[k])
for k in ("local_points", "points", "global_points", "camera_poses"):
    res[k] = rel(out[k], ref[k])
print(json.dumps(res, indent=1))
