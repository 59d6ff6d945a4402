"""B200 probe: the UNMODIFIED reference classes (oracle/_ref, real flash-attn, CUDA autocast) next to the CUDA path.

    python tests/probes/gpu_reference.py [--views 16 --height 518 --width 518] [--ls 0.01|synthetic] [--depth full|N]

Prints, for one scene: max-rel error of every output between {ours, reference-GPU, restatement bf16 (GPU tensors),
restatement fp32 (GPU tensors)} and the reference's views/s (stock flash-attn call, CUDA events).  Test
infrastructure (uses oracle/): the findings are turned into tests/test_reference_gpu.py and bench.py's gpu_reference."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from dataclasses import replace

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from PIL import Image  # noqa: E402

from g2vlm_b200 import schema  # noqa: E402
from g2vlm_b200.model import G2VLMFast  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402
from oracle import restate  # noqa: E402

KEYS = ("points", "local_points", "global_points", "camera_poses")


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max()).item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=16)
    ap.add_argument("--height", type=int, default=518)
    ap.add_argument("--width", type=int, default=518)
    ap.add_argument("--ls", default="0.01")
    ap.add_argument("--depth", default="full")
    ap.add_argument("--no-restate", action="store_true")
    ap.add_argument("--time-steps", type=int, default=3)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    res = dict(args=vars(args))

    # 1. does the flash-attn wheel run on this GPU, and does it leave uncovered rows unwritten (quirk Q1)?
    from flash_attn import flash_attn_varlen_func
    q = torch.randn(300, 4, 64, device="cuda", dtype=torch.bfloat16)
    poison = torch.full_like(q, 777.0)
    del poison                                     # the caching allocator hands the same block to flash-attn's `out`
    cu = torch.tensor([0, 128, 256], dtype=torch.int32, device="cuda")
    o = flash_attn_varlen_func(q, q, q, cu, cu, 128, 128, causal=False)
    res["flash_attn_ok"] = bool(torch.isfinite(o[:256].float()).all())
    res["q1_uncovered_rows_unwritten"] = bool((o[256:] == 777.0).all())
    print("flash-attn ok:", res["flash_attn_ok"], "| rows >= cu_seqlens[-1] left unwritten:", res["q1_uncovered_rows_unwritten"], flush=True)

    cfg = schema.FULL
    if args.depth != "full":
        d = int(args.depth)
        cfg = replace(cfg, num_layers=d, dino_layers=d, dec_depth=min(d, cfg.dec_depth))
    sd = schema.init_synthetic(cfg, seed=0, embed_rows=32, device="cuda")
    if args.ls != "synthetic":
        for k in sd:
            if k.endswith("ls1.gamma") or k.endswith("ls2.gamma") or k.endswith(".lambda1"):
                sd[k].fill_(float(args.ls))
    u8 = (schema.synthetic_views(args.views, args.height, args.width, seed=1) * 255).round().to(torch.uint8)
    pil = [Image.fromarray(u8[i].permute(1, 2, 0).numpy()) for i in range(u8.shape[0])]
    views = u8.float() / 255.0

    t0 = time.time()
    ref = rh.build_reference_model(rh.dims_from_cfg(cfg, vocab_size=32), visual_und=False, device="cuda")
    msg = ref.load_state_dict(sd, strict=False)
    res["ref_missing_keys"] = list(msg.missing_keys)[:8]
    res["ref_unexpected_keys"] = list(msg.unexpected_keys)[:8]
    print(f"reference model on GPU in {time.time() - t0:.1f}s; missing {len(msg.missing_keys)} unexpected {len(msg.unexpected_keys)}",
          res["ref_missing_keys"], res["ref_unexpected_keys"], flush=True)
    pred_ref = rh.run_reference_recon(ref, pil)
    torch.cuda.synchronize()

    ours = G2VLMFast(cfg, sd)
    tok = rh.StubTokenizer()
    pred = ours.recon(tok, dict(rh.NEW_TOKEN_IDS), None, views)
    torch.cuda.synchronize()
    outs = {"ours": pred, "ref_gpu": pred_ref}
    if not args.no_restate:
        with torch.device("cuda"):
            for mode in ("bf16", "fp32"):
                t0 = time.time()
                outs["restate_" + mode] = restate.recon(sd, cfg, views.cuda(), mode=mode)
                torch.cuda.synchronize()
                print(f"restate {mode} on GPU tensors: {time.time() - t0:.1f}s", flush=True)
    names = list(outs)
    table = {}
    for i, a in enumerate(names):
        for b in names[i + 1:]:
            table[f"{a} vs {b}"] = {k: rel(outs[a][k], outs[b][k]) for k in KEYS}
            print(f"{a:>13} vs {b:<13}", " ".join(f"{k}={v:.2e}" for k, v in table[f'{a} vs {b}'].items()), flush=True)
    res["max_rel"] = table

    # 2. speed of the stock reference (unwrapped flash-attn) vs ours, same inputs, CUDA events
    import modeling.g2vlm.dinov2_model as _dm
    _dm.flash_attn_varlen_func = flash_attn_varlen_func

    def timed(fn, n):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    import contextlib, io
    def ref_step():
        with contextlib.redirect_stdout(io.StringIO()):
            rh.run_reference_recon(ref, pil)
    ms_ref = timed(ref_step, args.time_steps)
    ms_ours = timed(lambda: ours.recon(tok, dict(rh.NEW_TOKEN_IDS), None, views), args.time_steps)
    res.update(ref_ms=ms_ref, ours_ms=ms_ours, ref_views_per_s=args.views / ms_ref * 1e3,
               ours_views_per_s=args.views / ms_ours * 1e3, mem_gb=torch.cuda.max_memory_allocated() / 1e9)
    print(f"reference (PyTorch + FA2) {ms_ref:.1f} ms/scene = {res['ref_views_per_s']:.2f} views/s | ours {ms_ours:.1f} ms = "
          f"{res['ours_views_per_s']:.2f} views/s | x{ms_ref / ms_ours:.1f} | peak mem {res['mem_gb']:.1f} GB", flush=True)
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
