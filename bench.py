#!/usr/bin/env python
"""Benchmark of the recon hot path (BASELINE.json: recon views/sec at 518 px; MoT layer ms; % roofline).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on host cores

A step = one `recon` of one synthetic scene (BASELINE.json configs[1]: G2VLM-2B-MoT, 16 views of
518x518, bf16) per GPU; with N > 1 every rank reconstructs its own scene (scenes are independent
units -> weak scaling, no data-path collective).  Rank 0 prints ONE JSON line.

  value     views/s with the prepared inputs already resident in HBM (device-timed, max over ranks)
  e2e       views/s through the public API from HOST image tensors: scenes streamed through
            `g2vlm_b200.serving.ReconServer` (= `G2VLMFast.recon(...)` per scene with the upload of scene i+1 and
            the download of scene i-1 on side streams): host index construction + H2D copies + the full forward
            + D2H copy of every output inside the timed region; `single_call_ms` is one blocking recon() call
  roofline  the dominant kernel (MoT shared attention, tcgen05): algorithmic FLOPs per launch
            4*T*(T+K0)*heads*head_dim / its mean launch duration measured with CUDA events on the
            launching stream inside the timed region, against MEASURED_PEAKS.json (sustained bf16)
  cpu_baseline  the reference's OWN classes (oracle/_ref = the unmodified tree staged by oracle/stage_ref.py, run
            under oracle/ref_harness.py) on the host cores: full-width, full-depth model, a bounded sample of the
            scene's views (rank 0, N = 1 only).  `--impl reference` times ONE OR MORE FULL steps of the same
            16-view scene with those classes (capped by a time budget, never by shrinking the model).
  gpu_reference the same classes on the same B200 with their own flash-attn calls (PyTorch + FA2): views/s on the
            same weights and views — the number to beat — and max-rel error of our outputs against it
  view_sharded  (N > 1 only) ONE long scene split by view over the N GPUs (BASELINE configs[3]): 64- and 256-view
            scenes, strong-scaling efficiency against the single-GPU time of the same scene, exposed K/V exchange
            time per layer, and `sp_parity_max_rel` = sharded result vs this rank's own single-GPU result

Synthetic data: seeded random-init weights of the full architecture (no checkpoint / network) and
blurred-noise views (g2vlm_b200.schema).  Inputs per step (5 GB of bf16 weights + activations) are far
larger than the 126 MB L2, so no explicit L2 flush is needed between iterations.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "recon views/sec (518px)"
UNIT = "views/s"


class StubTokenizer:
    def encode(self, prompt):
        return [11, 12, 13, 14, 15, 16]


TOKENS = dict(bos_token_id=1, eos_token_id=2, start_of_image=3, end_of_image=4)


_REAL_STDOUT = None


def isolate_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL's version banner, the reference's progress
    prints): from here on file descriptor 1 goes to stderr and only emit() reaches the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(tflops=p.get("bf16_tflops_sustained", 1407.1), hbm=p.get("hbm_gbs", 6553.9), src="measured")
    return dict(tflops=1400.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for n, v in zip(names, r[5:9]):
                if "Active" in v and "Not" not in v:
                    reasons.add(n)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def algorithmic_flops(cfg, n_views, P, K0=7):
    """SURVEY.md §8(d) formulas (multiply-add = 2)."""
    T = n_views * (P + 2)
    H, I = cfg.hidden_size, cfg.intermediate_size
    nq, nkv, hd = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim
    mot_lin = 2 * T * (H * (nq + 2 * nkv) * hd + nq * hd * H + 3 * H * I)
    mot_att = 4 * T * (T + K0) * nq * hd
    D, S = cfg.dino_hidden, P + 1 + cfg.dino_registers
    dino = cfg.dino_layers * (n_views * S * 2 * (4 * D * D + 2 * D * D * cfg.dino_mlp_ratio) + 4 * D * n_views * P * P) \
        + 2 * 588 * D * n_views * P + 2 * D * H * n_views * P
    blk, att = 2 * 12 * H * H, 4 * P * H
    tok = n_views * P
    heads = tok * (5 * (blk + att) + 2 * H * cfg.point_dim + 2 * cfg.point_dim * 588) \
        + tok * (5 * (blk + att) + 2 * H * cfg.camera_dim + 12 * cfg.camera_dim ** 2) \
        + tok * (5 * (blk + 2 * att + 4 * H * H) + 2 * H * cfg.point_dim + 2 * cfg.point_dim * 588) + 5 * P * 4 * H * H
    return dict(total=cfg.num_layers * (mot_lin + mot_att) + dino + heads, mot_layer=mot_lin + mot_att,
                mot_attention_launch=mot_att)


# --------------------------------------------------------------------------------------------------
# reference legs (test infrastructure: the only places where bench.py touches oracle/)
# --------------------------------------------------------------------------------------------------
DATA = "synthetic (seeded random-init weights: N(0,0.02^2) Linear/bias, U(0.5,1.5) norm weights, LayerScale {ls}; blurred-noise views rounded to 8 bit)"


def bench_weights(cfg, device, layerscale):
    """The benchmark's weights; LayerScale (ls1/ls2 gamma, DINO lambda1) either at the reference's
    init value 0.01 — the regime a trained checkpoint lives in, default — or 'synthetic' ~U(0.5,1.5)."""
    from g2vlm_b200 import schema
    sd = schema.init_synthetic(cfg, seed=0, embed_rows=32, device=device)
    if layerscale != "synthetic":
        for k in sd:
            if k.endswith("ls1.gamma") or k.endswith("ls2.gamma") or k.endswith(".lambda1"):
                sd[k].fill_(float(layerscale))
    return sd


def bench_views_u8(n_views, size, seed, width=None):
    from g2vlm_b200 import schema
    return (schema.synthetic_views(n_views, size, width or size, seed=seed) * 255).round().to(torch.uint8)


def to_pil(u8):
    from PIL import Image
    return [Image.fromarray(u8[i].permute(1, 2, 0).numpy()) for i in range(u8.shape[0])]


def build_reference(cfg, device, layerscale):
    """The reference's own G2VLM (oracle/_ref or /root/reference) with the benchmark's weights loaded."""
    from oracle import ref_harness as rh
    gen_dev = "cuda" if torch.cuda.is_available() else "cpu"   # same generator stream as our arm whenever a GPU exists
    sd = bench_weights(cfg, gen_dev, layerscale)
    ref = quiet(rh.build_reference_model, rh.dims_from_cfg(cfg, vocab_size=32), visual_und=False, device=device,
                zero_fill_uncovered=False, skip_init=True)     # stock flash-attn call for timing
    msg = ref.load_state_dict(sd, strict=False)
    params = dict(ref.named_parameters())
    with torch.no_grad():
        for k in msg.missing_keys:                     # nothing recon reads; left uninitialised by skip_init
            if k in params:
                params[k].zero_()
    del sd
    return rh, ref


def quiet(fn, *a, **k):
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):       # the reference prints progress lines: stdout carries ONE JSON line
        return fn(*a, **k)


def cpu_reference_steps(cfg, n_views, size, layerscale, threads, budget_s, max_steps, warm_views=1):
    """Times the reference's own classes on the host cores (bf16 CPU autocast -> oneDNN/AMX; SDPA stand-in for
    flash-attn): one untimed `warm_views`-view step (pages in oneDNN primitives), then full `n_views`-view steps
    until `budget_s` is spent (at least one).  Returns (seconds per step, steps run, build seconds)."""
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    rh, ref = build_reference(cfg, "cpu", layerscale)
    build_s = time.perf_counter() - t0
    if warm_views:
        quiet(rh.run_reference_recon, ref, to_pil(bench_views_u8(warm_views, size, seed=7)))
    pil = to_pil(bench_views_u8(n_views, size, seed=1))
    times = []
    t_start = time.perf_counter()
    while len(times) < max_steps and (not times or time.perf_counter() - t_start + times[-1] < budget_s):
        t0 = time.perf_counter()
        quiet(rh.run_reference_recon, ref, pil)
        times.append(time.perf_counter() - t0)
    return sum(times) / len(times), len(times), build_s


def cpu_port_sample(n_views: int, size: int, threads: int, layer_frac: int = 4):
    """Fallback when no reference tree is reachable: the oracle port on a bounded sample (full-width model at
    1/layer_frac depth), scaled to full depth.  Returns (views/s, description)."""
    from dataclasses import replace

    from g2vlm_b200 import schema
    from oracle import restate

    torch.set_num_threads(threads)
    full = schema.FULL
    cfg = replace(full, num_layers=max(1, full.num_layers // layer_frac),
                  dino_layers=max(1, full.dino_layers // layer_frac), dec_depth=max(1, full.dec_depth // layer_frac))
    sd = schema.init_synthetic(cfg, seed=0, embed_rows=32)
    views = schema.synthetic_views(n_views, size, size, seed=1)
    t0 = time.perf_counter()
    restate.recon(sd, cfg, views, mode="bf16")
    per_step = time.perf_counter() - t0
    scale = full.num_layers / cfg.num_layers
    desc = (f"oracle PORT (oracle/restate.py; no reference tree on this box), full-width model at 1/{layer_frac} depth, "
            f"{n_views} view(s) of {size}px, {threads} threads; {per_step:.2f} s per sample step, scaled x{scale:.0f}")
    return n_views / (per_step * scale), desc


def workload_string(n_views, size):
    return (f"G2VLM-2B-MoT recon bf16, {n_views} views {size}px, single B200 (BASELINE configs[1]); one scene per GPU"
            if (n_views, size) == (16, 518) else f"G2VLM-2B-MoT recon bf16, {n_views} views {size}px; one scene per GPU")


def run_reference_arm(args):
    """`--impl reference`: the UNMODIFIED reference classes on the box's host cores, on OUR arm's config (full
    G2VLM-2B-MoT, `--views` views of `--size` px, bf16 autocast), every step one full scene."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from g2vlm_b200 import schema
    from oracle import ref_harness as rh
    threads = os.cpu_count() or 1
    P = (args.size // 14) ** 2
    cfg = schema.TINY if args.tiny else schema.FULL
    if rh.available():
        per_step, n_steps, build_s = cpu_reference_steps(cfg, args.views, args.size, args.layerscale, threads,
                                                         args.ref_budget_s, max(1, args.steps))
        v = args.views / per_step
        kind = "reference"
        sample = (f"reference classes (oracle/_ref, unmodified; CPU bf16 autocast, SDPA stand-in for flash-attn), full-width "
                  f"full-depth model, FULL scene of {args.views} views {args.size}px per step, {threads} threads; "
                  f"{n_steps} timed step(s) of {per_step:.1f} s after one untimed 1-view step (time budget {args.ref_budget_s:.0f} s, "
                  f"requested steps {args.steps}); model build {build_s:.0f} s untimed")
    else:
        v, sample = cpu_port_sample(1, args.size, threads)
        per_step, n_steps, kind = args.views / v, 1, "port"
    line = dict(metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                steps_run=n_steps, ms_per_step=per_step * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="bf16", data=DATA.format(ls=args.layerscale), impl="reference",
                config=dict(workload=workload_string(args.views, args.size) if not args.tiny else "TINY DEBUG MODEL (invalid as a benchmark)",
                            views_per_scene=args.views, image_size=args.size, tokens=args.views * (P + 2),
                            scenes_per_step=1, parallelism="host cores (reference classes, CPU)"),
                cpu_baseline=dict(value=v, unit=UNIT, cores=threads, kind=kind, sample=sample),
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    emit(line)


def gpu_reference_leg(cfg, layerscale, u8, ours_pred, steps=3):
    """The reference's own classes on THIS GPU (PyTorch + flash-attn 2, CUDA autocast), same weights and views:
    views/s (CUDA events, 1 warm-up) and max-rel error of our outputs against it."""
    try:
        from oracle import ref_harness as rh
        if not rh.gpu_available():
            return dict(unavailable="no reference tree (oracle/_ref) or no flash_attn wheel on this box")
        rh, ref = build_reference(cfg, "cuda", layerscale)
        pil = to_pil(u8)
        # parity sample: the DINO rows flash-attn never writes (quirk Q1: `out = empty_like(q)`, rows >= cu_seqlens[-1])
        # are zero-filled, which is the definition every implementation here uses; the TIMED steps below run the
        # stock call (whatever the allocator left in those rows then flows into the result: 0.15 max-rel observed)
        import modeling.g2vlm.dinov2_model as _dm
        stock = _dm.flash_attn_varlen_func
        _dm.flash_attn_varlen_func = rh.flash_attn_zero_fill(stock)
        pred = quiet(rh.run_reference_recon, ref, pil)        # warm-up + parity sample
        _dm.flash_attn_varlen_func = stock
        torch.cuda.synchronize()
        max_rel = {}
        for k in ("points", "local_points", "global_points", "camera_poses"):
            a, b = ours_pred[k].float(), pred[k].float()
            max_rel[k] = ((a - b).abs().max() / b.abs().max()).item()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            quiet(rh.run_reference_recon, ref, pil)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        import flash_attn
        out = dict(value=u8.shape[0] / ms * 1e3, unit=UNIT, ms_per_step=ms, steps=steps,
                   impl=f"reference classes (oracle/_ref, unmodified) on this GPU: PyTorch {torch.__version__} + flash_attn "
                        f"{flash_attn.__version__}, bf16 autocast, input = the same 8-bit views as PIL images",
                   max_rel_ours_vs_reference=max_rel,
                   note="timed steps = stock flash-attn call; parity sample = the same with the DINO rows flash-attn never "
                        "writes (quirk Q1) zero-filled")
        del ref, pred
        torch.cuda.empty_cache()
        return out
    except Exception as e:  # a baseline must never take the measurement down with it
        return dict(unavailable=f"{type(e).__name__}: {e}")


# --------------------------------------------------------------------------------------------------
# view-sharded long scene (BASELINE configs[3]): sequence parallel over the ranks
# --------------------------------------------------------------------------------------------------
def decode_leg(cfg, layerscale, views_host, tokens=64):
    """Row f.1 next to the headline: greedy decode tokens/s behind this scene in the KV cache (20-token system prompt, the
    16-view geo step with cache update, a 60-token question), full 151 936-row embedding / lm_head, batch 1.  Device-timed
    (CUDA events around the token loop); `multi_launch` = the ~280-launch step of round 1 on the same cache, CUDA graph."""
    from g2vlm_b200 import schema
    from g2vlm_b200.model import G2VLMFast, NaiveCache
    sd = schema.init_synthetic(cfg, seed=0, device="cuda")
    if layerscale != "synthetic":
        for k in sd:
            if k.endswith("ls1.gamma") or k.endswith("ls2.gamma") or k.endswith(".lambda1"):
                sd[k].fill_(float(layerscale))
    model = G2VLMFast(cfg, sd)
    del sd
    torch.cuda.empty_cache()

    def text(n, kvlen, rope):
        return dict(text_token_lens=torch.tensor([n], dtype=torch.int), packed_text_ids=torch.arange(10, 10 + n),
                    packed_text_position_ids=(rope + torch.arange(n)).expand(3, -1), packed_text_indexes=kvlen + torch.arange(n),
                    packed_key_value_indexes=torch.arange(kvlen), key_values_lens=torch.tensor([kvlen], dtype=torch.int)), kvlen + n, rope + n

    def run(n, fused):
        gi, kvlen, rope = text(20, 0, 0)
        past = model.forward_cache_update_text(NaiveCache(cfg.num_layers), **gi)
        gi, nl, nr = model.prepare_dino_images_pi3([kvlen], [rope], views_host, None, TOKENS)
        past, _ = model.forward_cache_update_dino(past, **gi)
        gi, kvlen, rope = text(60, nl[0], nr[0])
        past = model.forward_cache_update_text(past, **gi)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        model.generate_text(past, None, None, torch.tensor([7]), torch.full((3, 1), rope), n, end_token_id=None,
                            use_cuda_graph=not fused, fused_step=fused)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, kvlen

    run(8, True)
    ms, rows = run(tokens, True)
    run(8, False)
    ms_multi, _ = run(max(16, tokens // 2), False)
    H, I, nq, nkv, hd = cfg.hidden_size, cfg.intermediate_size, cfg.num_heads, cfg.num_kv_heads, cfg.head_dim
    w_layer = 2 * ((nq + 2 * nkv) * hd * H + H * nq * hd + 3 * H * I)           # bf16 und-expert weights of a layer
    nbytes = cfg.num_layers * (w_layer + 2 * 2 * nkv * hd * rows) + 2 * cfg.vocab_size * H
    pk = peaks()
    traffic = None
    try:  # dram bytes of one step from the committed ncu capture (same cache length within 0.6 %)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "kernel_traffic.json")))["und_decode_fused_kernel@L=22144"]["traffic_bytes"]
    except Exception:
        pass
    return dict(metric="greedy decode tokens/sec (batch 1)", value=1e3 / ms, unit="tokens/s", ms_per_token=ms, cached_rows=rows,
                tokens_timed=tokens, bytes_per_token=nbytes,
                roofline=dict(bound="hbm", achieved=nbytes / ms / 1e6, peak=pk["hbm"], unit="GB/s", frac=nbytes / ms / 1e6 / pk["hbm"],
                              traffic=traffic, kernel="und_decode_fused_kernel (one persistent cooperative launch per token)"),
                multi_launch_ms_per_token=ms_multi,
                note="`next` row f.1 (chat path), not the headline metric; algorithmic bytes = und-expert weights + lm_head + "
                     "the K|V rows of the cache, each read once per token")


def view_sharded_scene(model, dist, rank, world, n_views, size, steps, warmup, compare_single):
    """ONE scene of n_views views split by view over the ranks (sequence parallel).  Returns a dict with the max-over-
    ranks time per scene, the exposed K/V-exchange time per layer and, if `compare_single`, the single-GPU time of
    the same scene on this rank plus sp_parity_max_rel (sharded result vs this rank's own single-GPU result)."""
    tok = StubTokenizer()
    views = bench_views_u8(n_views, size, seed=11).float() / 255.0

    def allmax(x):
        t = torch.tensor([float(x)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, n):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return allmax(e0.elapsed_time(e1) / n), out

    res = dict(views=n_views, tokens=n_views * ((size // 14) ** 2 + 2), mode=model.sp_mode, sm_margin=model.sp_sm_margin)
    for _ in range(warmup):
        model.recon_view_sharded(tok, dict(TOKENS), views)
    model.sp_events = []
    ms, pred = timed(lambda: model.recon_view_sharded(tok, dict(TOKENS), views), steps)
    ev, model.sp_events = model.sp_events, None
    exposed = [a.elapsed_time(b) for a, b in ev if b is not None]
    res.update(ms_per_scene=ms, views_per_s=n_views / ms * 1e3,
               exposed_exchange_ms_per_layer=allmax(sum(exposed) / max(1, len(exposed))),
               mot_layer_ms=None)
    if compare_single:
        for _ in range(warmup):
            model.recon(tok, dict(TOKENS), None, views)
        ms1, full = timed(lambda: model.recon(tok, dict(TOKENS), None, views), steps)
        v0, v1 = pred["view_range"]
        worst = 0.0
        for k in ("points", "local_points", "global_points", "camera_poses"):
            a, b = pred[k].float(), full[k][:, v0:v1].float()
            worst = max(worst, ((a - b).abs().max() / full[k].float().abs().max()).item())
        worst = max(worst, ((pred["camera_poses_all"] - full["camera_poses"]).abs().max()
                            / full["camera_poses"].abs().max()).item())
        res.update(single_gpu_ms_per_scene=ms1, strong_scaling_efficiency=ms1 / (ms * world),
                   sp_parity_max_rel=allmax(worst))
    return res


def view_sharded_block(model, cfg, dist, rank, world, args):
    """`view_sharded` block of the N > 1 bench line (BASELINE configs[3])."""
    out = dict(note="ONE scene split by view over the N GPUs: DINO by attention segment + neighbour exchange, per MoT "
                    "layer the K|V rows are exchanged over NVLink (mode peer: symmetric memory + copy-engine pulls) behind "
                    "the attention over the local keys, then remote keys + log-sum-exp merge; context broadcast; times are "
                    "max over ranks, CUDA events",
               scenes=[])
    try:
        P = (args.size // 14) ** 2
        for n_views, cmp_single in ((64, True), (256, False)):
            r = view_sharded_scene(model, dist, rank, world, n_views, args.size, steps=1, warmup=1, compare_single=cmp_single)
            fl = algorithmic_flops(cfg, n_views, P)
            r["achieved_tflops_per_gpu"] = fl["total"] / 1e12 / (r["ms_per_scene"] / 1e3) / world
            r["mot_layer_algorithmic_ms_at_peak_per_gpu"] = fl["mot_layer"] / 1e12 / peaks()["tflops"] / world * 1e3
            out["scenes"].append(r)
            if n_views == 64:
                out["sp_parity_max_rel"] = r["sp_parity_max_rel"]
        # A/B of the exchange on the 64-view scene: v2 (NCCL point-to-point under the local attention), v1 (all-gather)
        keep = model.sp_mode
        for tag, mode in (("nccl_p2p_v2_64views", "overlap"), ("allgather_v1_64views", "allgather")):
            model.sp_mode = mode
            r = view_sharded_scene(model, dist, rank, world, 64, args.size, steps=1, warmup=1, compare_single=False)
            out[tag] = dict(ms_per_scene=r["ms_per_scene"], exposed_exchange_ms_per_layer=r["exposed_exchange_ms_per_layer"])
        model.sp_mode = keep
    except Exception as e:   # must not take the scene-DP line down with it
        out["error"] = f"{type(e).__name__}: {e}"
    return out


def run_view_sharded(args, cfg, dist, rank, world, local_rank):
    """`--workload views`: the standalone long-scene run (one JSON line, strong scaling)."""
    from g2vlm_b200 import ops
    from g2vlm_b200.model import G2VLMFast

    n_views, size = args.views, args.size
    P = (size // 14) ** 2
    model = G2VLMFast(cfg, bench_weights(cfg, "cuda", args.layerscale))
    model.sp_mode, model.sp_sm_margin = args.sp_mode, args.sp_sm_margin
    torch.cuda.empty_cache()
    views_host = bench_views_u8(n_views, size, seed=1).float() / 255.0
    tok = StubTokenizer()

    def step():
        if dist is None:
            return model.recon(tok, dict(TOKENS), None, views_host)
        return model.recon_view_sharded(tok, dict(TOKENS), views_host)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(1, args.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ops.launches()
    model.sp_events = []
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    exposed = [a.elapsed_time(b) for a, b in model.sp_events if b is not None]
    model.sp_events = None
    exposed_ms = sum(exposed) / max(1, len(exposed))
    if dist is not None:
        t = torch.tensor([ms, exposed_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, exposed_ms = float(t[0].item()), float(t[1].item())
    clocks = sampler.stop()
    barrier()
    if rank == 0:
        fl = algorithmic_flops(cfg, n_views, P)
        pk = peaks()
        per = ms / args.steps
        emit(dict(
            metric=METRIC, value=n_views / (per / 1e3), unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=per, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="bf16",
            data=DATA.format(ls=args.layerscale),
            config=dict(workload=f"G2VLM-2B-MoT long-sequence recon, ONE scene of {n_views} views {size}px, view-sharded "
                                 f"over {world} GPU(s): DINO by segment + neighbour exchange, per-layer K/V exchange "
                                 f"({args.sp_mode}, NCCL/NVLink), context broadcast", views_per_scene=n_views, image_size=size,
                        tokens=n_views * (P + 2), parallelism=f"view-sp{world}", sp_mode=args.sp_mode, sp_sm_margin=args.sp_sm_margin),
            clocks=clocks, gpu_launches=ops.launches() - l0, exposed_exchange_ms_per_layer=exposed_ms,
            e2e=dict(value=n_views / (per / 1e3), unit=UNIT, h2d_bytes_per_step=views_host.numel() * 4 // world,
                     d2h_bytes_per_step=0, note="timed through recon_view_sharded from host views; outputs stay on the GPUs"),
            whole_step=dict(algorithmic_tflop=fl["total"] / 1e12,
                            achieved_tflops_per_gpu=fl["total"] / 1e12 / (per / 1e3) / world,
                            frac_of_peak=fl["total"] / 1e12 / (per / 1e3) / world / pk["tflops"])))
    if dist is not None:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------
# fp32 mode (north_star "fp32 mode <= 1e-4"; BASELINE configs[0] "8 views fp32")
# --------------------------------------------------------------------------------------------------
def run_fp32(args, cfg, local_rank):
    """`--mode fp32`: the fp32-mode path (g2vlm_b200/model_fp32.py) on BASELINE configs[0]'s geometry by default
    (8 views of 294x518 = examples/dl3dv frames after load_and_resize14).  One JSON line, dtype "fp32"; `value` and
    `e2e` as in the bf16 line; `max_rel_vs_fp32_oracle` = outputs against oracle/restate.py mode="fp32" evaluated on GPU
    tensors in this run (the checker; not timed)."""
    from g2vlm_b200 import ops
    from g2vlm_b200.model import G2VLMFast, NaiveCache
    n_views, Hh, Ww = args.views, args.height, args.width
    P = (Hh // 14) * (Ww // 14)
    sd = bench_weights(cfg, "cuda", args.layerscale)
    model = G2VLMFast(cfg, sd, mode="fp32")
    views_u8 = bench_views_u8(n_views, Hh, seed=1, width=Ww)
    views_host = (views_u8.float() / 255.0).pin_memory()
    tok = StubTokenizer()
    gi_text, newlens, new_rope = model.prepare_prompts_addbos([0], [0], ["x"], tok, TOKENS)
    HOST_META = ("text_token_lens", "key_values_lens", "packed_seqlens", "dino_token_seqlens")
    gi, _, _ = model.prepare_dino_images_pi3(newlens, new_rope, views_host, None, TOKENS)
    gi = {k: (v if k in HOST_META else v.cuda()) for k, v in gi.items()}

    def step_resident():
        past, last = model.forward_cache_update_dino(NaiveCache(cfg.num_layers), prompt=gi_text, **gi)
        return model.reconstruct(past_key_values=past, selected_hidden_states=last, **gi)

    out_host = {}

    def step_e2e():
        pred = model.recon(tok, dict(TOKENS), None, views_host)
        for k in ("points", "local_points", "global_points", "camera_poses"):
            if k not in out_host:
                out_host[k] = torch.empty(pred[k].shape, dtype=pred[k].dtype, pin_memory=True)
            out_host[k].copy_(pred[k], non_blocking=True)
        return pred

    def timed(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ops.launches()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, (ops.launches() - l0)

    for _ in range(max(3, args.warmup)):
        step_resident()
    sampler = ClockSampler(local_rank)
    ms, launches = timed(step_resident, args.steps)
    clocks = sampler.stop()
    step_e2e()
    ms_e2e, _ = timed(step_e2e, max(1, min(args.steps, 3)))
    pred = model.recon(tok, dict(TOKENS), None, views_host)
    try:
        from oracle import restate
        with torch.device("cuda"):
            ref = restate.recon(sd, cfg, views_host.cuda(), mode="fp32")
        max_rel = {k: ((pred[k].double() - ref[k].double()).abs().max() / ref[k].double().abs().max()).item()
                   for k in ("points", "local_points", "global_points", "camera_poses")}
    except Exception as e:
        max_rel = dict(error=f"{type(e).__name__}: {e}")
    fl = algorithmic_flops(cfg, n_views, P)
    d2h = sum(v.numel() * v.element_size() for v in out_host.values())
    emit(dict(
        metric=METRIC, value=n_views / ms * 1e3, unit=UNIT, n_gpus=1, steps=args.steps, warmup=max(3, args.warmup),
        ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="fp32",
        data=DATA.format(ls=args.layerscale),
        config=dict(workload=f"G2VLM-2B-MoT recon fp32 mode, {n_views} views {Hh}x{Ww} (BASELINE configs[0] geometry: "
                             f"examples/dl3dv frames after load_and_resize14), single B200", views_per_scene=n_views,
                    image_height=Hh, image_width=Ww, tokens=n_views * (P + 2), parallelism="scene-dp1",
                    arithmetic="every Linear = three-piece split-bf16 GEMM on tcgen05 (6 products, fp32 RN chunk sums); "
                               "fp32 FFMA attention; fp32 elementwise", l2="inputs larger than L2; no flush"),
        clocks=clocks, gpu_launches=launches,
        e2e=dict(value=n_views / ms_e2e * 1e3, unit=UNIT, h2d_bytes_per_step=views_host.numel() * 4, d2h_bytes_per_step=d2h,
                 ms_per_step=ms_e2e, note="one blocking recon() from pinned host views + download of all outputs per step"),
        max_rel_vs_fp32_oracle=max_rel, tolerance=1e-4,
        whole_step=dict(algorithmic_tflop=fl["total"] / 1e12, achieved_tflops=fl["total"] / 1e12 / (ms / 1e3))))


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="g2vlm_b200", choices=["g2vlm_b200", "reference"])
    ap.add_argument("--views", type=int, default=16)
    ap.add_argument("--size", type=int, default=518)
    ap.add_argument("--cpu-views", type=int, default=2,
                    help="views of the cpu_baseline sample (reference classes, full depth) inside our arm's run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-view-sharded", action="store_true", help="N > 1: skip the view_sharded block")
    ap.add_argument("--no-decode", action="store_true", help="N = 1: skip the `decode` block (row f.1: tokens/s behind this scene)")
    ap.add_argument("--layerscale", default="0.01",
                    help="LayerScale value of the synthetic weights: 0.01 = the reference's init (g2vlm/qwen2vl.py:765-766), "
                         "the regime of a trained checkpoint; 'synthetic' = U(0.5,1.5)")
    ap.add_argument("--ref-budget-s", type=float, default=120.0,
                    help="--impl reference: stop adding full-scene steps once this much time is spent (>= 1 step)")
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"],
                    help="fp32: the fp32-mode path on BASELINE configs[0]'s geometry (--views 8 --height 294 --width 518 unless given)")
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--sp-mode", default="peer", choices=["peer", "overlap", "allgather"])
    ap.add_argument("--sp-sm-margin", type=int, default=-1,
                    help="SMs the local-key attention leaves to the NCCL exchange kernel (= NCCL_MAX_NCHANNELS); -1: min(32, 4*N)")
    ap.add_argument("--no-native", action="store_true",
                    help="per-op calls from Python (~630 per step) instead of the three native stage calls (bit-identical results)")
    ap.add_argument("--no-fuse-prompt", action="store_true",
                    help="run the 7-token prompt prefill as a separate und pass (reference order) instead of fused into the geo step")
    ap.add_argument("--profile", action="store_true",
                    help="profiling run under ncu: skip the e2e arm and the CPU baseline, allow warmup < 3 (not a bench value)")
    ap.add_argument("--tiny", action="store_true", help="tiny model dims (smoke / debugging only; INVALID as a benchmark)")
    ap.add_argument("--workload", default="scenes", choices=["scenes", "views"],
                    help="scenes: one 16-view scene per GPU (default, BASELINE configs[1]/[2]); views: ONE scene of "
                         "--views views split by view over the GPUs (BASELINE configs[3], sequence parallel)")
    args = ap.parse_args()
    isolate_stdout()

    if args.impl == "reference":
        run_reference_arm(args)
        return

    from g2vlm_b200 import ops, schema
    from g2vlm_b200.model import G2VLMFast, NaiveCache
    from g2vlm_b200.serving import RESULT_KEYS, ReconServer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        # the view-sharded K/V exchange runs UNDER the local-key attention (a persistent kernel that leaves
        # --sp-sm-margin SMs free): a few NCCL channels move the 45-180 MB per layer far faster than the attention
        # it hides behind, and every channel is an SM taken from that attention
        if args.sp_sm_margin < 0:
            args.sp_sm_margin = min(32, 4 * world)   # more peers -> more P2P channels; the local-key attention they
                                                     # take SMs from is only 1/world of the layer's attention
        os.environ.setdefault("NCCL_MAX_NCHANNELS", str(max(1, args.sp_sm_margin)))
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.warmup < 3 and not args.profile:
        args.warmup = 3  # timing rule: at least 3 warm-up steps

    cfg = schema.TINY if args.tiny else schema.FULL
    if args.mode == "fp32":
        if "--views" not in sys.argv:
            args.views = 8
        args.height, args.width = args.height or 294, args.width or 518
        if rank == 0:
            run_fp32(args, cfg, local_rank)
        if dist is not None:
            dist.destroy_process_group()
        return
    if args.workload == "views":
        run_view_sharded(args, cfg, dist, rank, world, local_rank)
        return
    n_views, size = args.views, args.size
    P = (size // 14) ** 2
    model = G2VLMFast(cfg, bench_weights(cfg, "cuda", args.layerscale))
    model.fuse_prompt = not args.no_fuse_prompt
    model.native = not args.no_native
    model.sp_mode, model.sp_sm_margin = args.sp_mode, args.sp_sm_margin
    torch.cuda.empty_cache()
    views_u8 = bench_views_u8(n_views, size, seed=1 + rank)          # 8-bit views: the reference legs read the same pixels
    views_host = (views_u8.float() / 255.0).pin_memory()
    tok = StubTokenizer()

    # ---- device-resident arm ("value") ---------------------------------------------------------
    gi_text, newlens, new_rope = model.prepare_prompts_addbos([0], [0], ["x"], tok, TOKENS)
    # the *_lens tensors are host-side control metadata (reading them from the device would stall the host every step)
    HOST_META = ("text_token_lens", "key_values_lens", "packed_seqlens", "dino_token_seqlens")
    gi_text = {k: (v if k in HOST_META else v.cuda()) for k, v in gi_text.items()}
    gi, _, _ = model.prepare_dino_images_pi3(newlens, new_rope, views_host, None, TOKENS)
    gi = {k: (v if k in HOST_META else v.cuda()) for k, v in gi.items()}

    T = n_views * (P + 2)
    K_PROMPT = int(gi_text["packed_text_ids"].numel())   # prompt rows ride along in the fused path

    def step_resident():
        # same sequence as G2VLMFast.recon(): the 7-token prompt prefill rides along with the geo step
        # (model.fuse_prompt; --no-fuse-prompt runs it as the reference's separate und pass first)
        past = NaiveCache(cfg.num_layers)
        if not model.fuse_prompt:
            past = model.forward_cache_update_text(past, **gi_text)
        past, last = model.forward_cache_update_dino(past, update_past_key_values=False,
                                                     prompt=gi_text if model.fuse_prompt else None, **gi)
        return model.reconstruct(past_key_values=past, selected_hidden_states=last, **gi)

    out_host = {}

    def step_single_call():
        # one blocking user call: upload, compute and download back to back on one stream
        pred = model.recon(tok, dict(TOKENS), None, views_host)
        for k in RESULT_KEYS:
            if k not in out_host:
                out_host[k] = torch.empty(pred[k].shape, dtype=pred[k].dtype, pin_memory=True)
            out_host[k].copy_(pred[k], non_blocking=True)
        return pred

    # the e2e arm: scenes fed one after another through the serving loop (g2vlm_b200/serving.py): every step uploads
    # its views from pinned host memory and downloads its point maps to pinned host memory; the copies of step i+-1
    # overlap the kernels of step i on side streams.  drain() closes the timed region after the last download.
    server = ReconServer(model, tok, dict(TOKENS))

    def step_e2e():
        server.submit(views_host)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = ops.launches()
        e0.record()
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms, ops.launches() - launches0

    for _ in range(args.warmup):
        step_resident()
    # CUDA events around every MoT shared-attention launch, recorded on the launching stream by the native stage
    # driver (g2vlm_mot_forward_geo's attention_events) or, on the per-op path, by a wrapper around ops.attention
    model.attn_events = []
    orig_attention = ops.attention

    def timed_attention(q, *a, **k):
        if q.shape[0] in (T, T + K_PROMPT) and k.get("num_kv_heads") == cfg.num_kv_heads and not k.get("causal", False):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = orig_attention(q, *a, **k)
            e1.record()
            model.attn_events.append((e0, e1))
            return r
        return orig_attention(q, *a, **k)

    ops.attention = timed_attention
    model.stage_events = []
    sampler = ClockSampler(local_rank)
    ms_total, launches = timed(step_resident, args.steps)
    clocks = sampler.stop()
    ops.attention = orig_attention
    stage_events, model.stage_events = model.stage_events, None
    att_events, model.attn_events = model.attn_events, None

    att_ms = [a.elapsed_time(b) for a, b in att_events]
    att_mean_ms = sum(att_ms) / max(1, len(att_ms))
    # stage times from the boundary events of the timed steps
    stage = {}
    names = [n for n, _ in stage_events]
    per = len(names) // max(1, args.steps)
    for s in range(args.steps):
        ev = dict(stage_events[s * per:(s + 1) * per])
        if {"dino_begin", "dino_end", "mot_end", "heads_end"} <= set(ev):
            stage.setdefault("dino_ms", []).append(ev["dino_begin"].elapsed_time(ev["dino_end"]))
            stage.setdefault("mot_ms", []).append(ev["dino_end"].elapsed_time(ev["mot_end"]))
            stage.setdefault("heads_ms", []).append(ev["mot_end"].elapsed_time(ev["heads_end"]))
    stage = {k: sum(v) / len(v) for k, v in stage.items()}

    if args.profile:
        ms_e2e = ms_single = float("nan")
        args.no_cpu_baseline = True
    else:
        for _ in range(2):
            step_e2e()
        server.drain()
        ms_e2e, _ = timed(step_e2e, args.steps, finish=server.drain)
        step_single_call()
        ms_single, _ = timed(step_single_call, 2)
        ms_single /= 2

    view_sharded = None
    if dist is not None and not args.no_view_sharded and not args.profile and not args.tiny:
        view_sharded = view_sharded_block(model, cfg, dist, rank, world, args)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    pk = peaks()
    fl = algorithmic_flops(cfg, n_views, P)
    traffic = None
    try:  # dram bytes per launch of the dominant kernel from the committed ncu --set full capture
        tr = json.load(open(os.path.join(ROOT, "profiles", "kernel_traffic.json")))
        if n_views == 16 and size == 518 and not args.tiny:
            traffic = tr["attention_tcgen05_kernel<128>@T=21936"]["traffic_bytes"]
    except Exception:
        traffic = None
    ms_per_step = ms_total / args.steps
    value = world * n_views / (ms_per_step / 1e3)
    e2e_value = world * n_views / (ms_e2e / args.steps / 1e3)
    achieved = fl["mot_attention_launch"] / (att_mean_ms / 1e3) / 1e12 if att_mean_ms > 0 else 0.0
    # e2e: the raw views cross PCIe once (normalised on the device) + the index / position tensors
    h2d = views_host.numel() * views_host.element_size() \
        + sum(v.numel() * v.element_size() for k, v in gi.items() if k not in ("packed_dino_images", "original_images")) \
        + sum(v.numel() * v.element_size() for v in gi_text.values())
    d2h = sum(v.numel() * v.element_size() for v in server.host_out[0].values())
    line = dict(
        metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
        data=DATA.format(ls=args.layerscale),
        config=dict(workload=workload_string(n_views, size) if not args.tiny else "TINY DEBUG MODEL (invalid as a benchmark)",
                    views_per_scene=n_views, image_size=size, tokens=T, scenes_per_step=world, parallelism=f"scene-dp{world}",
                    l2="inputs larger than L2 (5 GB of weights streamed per step); no flush"),
        clocks=clocks,
        e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                 ms_per_step=ms_e2e / args.steps, single_call_ms=ms_single,
                 note="scenes streamed through ReconServer: copies of step i+-1 overlap the kernels of step i; "
                      "single_call_ms = one blocking recon() + download with nothing overlapped"),
        gpu_launches=launches,
        roofline=dict(bound="tensor", kernel="attention_tcgen05_kernel<128> (MoT shared attention)", achieved=achieved,
                      peak=pk["tflops"], unit="TFLOP/s", frac=achieved / pk["tflops"], traffic=traffic,
                      peak_source=pk["src"] + " sustained bf16", launch_ms=att_mean_ms, launches_timed=len(att_ms)),
        mot_layer_ms=stage.get("mot_ms", 0.0) / cfg.num_layers,
        stage_ms=stage,
        whole_step=dict(algorithmic_tflop=fl["total"] / 1e12, achieved_tflops=fl["total"] / 1e12 / (ms_per_step / 1e3),
                        frac_of_peak=fl["total"] / 1e12 / (ms_per_step / 1e3) / pk["tflops"]),
    )
    if view_sharded is not None:
        line["view_sharded"] = view_sharded
        if "sp_parity_max_rel" in view_sharded:
            line["sp_parity_max_rel"] = view_sharded["sp_parity_max_rel"]
    if world == 1 and not args.profile and not args.tiny:
        if not args.no_gpu_reference:
            # the number to beat: the reference's own classes on this GPU (PyTorch + flash-attn 2), same weights / views
            ours = model.recon(tok, dict(TOKENS), None, views_host)
            line["gpu_reference"] = gpu_reference_leg(cfg, args.layerscale, views_u8, ours)
            if "value" in line["gpu_reference"]:
                line["gpu_reference"]["speedup_e2e_vs_gpu_reference"] = e2e_value / line["gpu_reference"]["value"]
        if not args.no_decode:
            try:
                del model, server
                torch.cuda.empty_cache()
                line["decode"] = decode_leg(cfg, args.layerscale, views_host)
            except Exception as e:  # a `next`-row block must never take the headline number down with it
                line["decode"] = dict(failed=f"{type(e).__name__}: {e}")
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            try:
                from oracle import ref_harness as rh
                if rh.available():
                    per_step, n_steps, build_s = cpu_reference_steps(cfg, args.cpu_views, size, args.layerscale, cores,
                                                                     budget_s=20.0, max_steps=2, warm_views=1)
                    line["cpu_baseline"] = dict(
                        value=args.cpu_views / per_step, unit=UNIT, cores=cores, kind="reference",
                        sample=f"reference classes (oracle/_ref, unmodified; CPU bf16 autocast), full-width full-depth model, "
                               f"bounded sample: {args.cpu_views} of the scene's {n_views} views ({size}px) per step, {n_steps} timed "
                               f"step(s) of {per_step:.1f} s after one 1-view warm-up; under-counts the quadratic attention of "
                               f"the full scene, i.e. favours the CPU (`--impl reference` times the full {n_views}-view scene)")
                else:
                    v, desc = cpu_port_sample(1, size, cores)
                    line["cpu_baseline"] = dict(value=v, unit=UNIT, cores=cores, kind="port", sample=desc)
            except Exception as e:  # the baseline must never take the GPU number down with it
                line["cpu_baseline"] = dict(value=None, unit=UNIT, cores=cores, kind="reference",
                                            sample=f"failed: {type(e).__name__}: {e}")
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
