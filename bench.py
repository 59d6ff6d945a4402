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
  cpu_baseline  the oracle port (oracle/restate.py = CPU restatement of the reference algorithm) timed
            on the host cores on a bounded sample (rank 0, N = 1 only)

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
# reference arm / cpu baseline: the oracle port on host cores, bounded sample
# --------------------------------------------------------------------------------------------------
def cpu_sample(n_views: int, size: int, steps: int, warmup: int, threads: int, layer_frac: int = 1):
    """Times oracle/restate.recon (CPU restatement of the reference algorithm, bf16 mode) on a bounded
    sample: the FULL-width G2VLM-2B-MoT architecture at 1/`layer_frac` of every stack's depth, `n_views`
    views of `size` px.  Returns (views/s extrapolated to full depth, description)."""
    from dataclasses import replace

    from g2vlm_b200 import schema
    from oracle import restate

    torch.set_num_threads(threads)
    full = schema.FULL
    cfg = replace(full, num_layers=max(1, full.num_layers // layer_frac),
                  dino_layers=max(1, full.dino_layers // layer_frac), dec_depth=max(1, full.dec_depth // layer_frac))
    sd = schema.init_synthetic(cfg, seed=0, embed_rows=32)
    views = schema.synthetic_views(n_views, size, size, seed=1)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        restate.recon(sd, cfg, views, mode="bf16")
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    per_step = sum(times) / len(times)
    # depth scaling: every stack's cost is linear in its depth; the non-layer work (embeddings, heads'
    # final linears, epilogue) is < 1% and is counted at its measured value
    scale = full.num_layers / cfg.num_layers
    est_full = per_step * scale
    desc = (f"oracle port (oracle/restate.py, torch CPU fp32 math with bf16 rounding points), full-width "
            f"G2VLM-2B-MoT at 1/{layer_frac} depth ({cfg.num_layers} MoT + {cfg.dino_layers} DINO layers, "
            f"{cfg.dec_depth} blocks per decoder), {n_views} view(s) of {size}x{size}, {threads} threads; "
            f"{per_step:.2f} s per sample step, scaled x{scale:.0f} to full depth")
    return n_views / est_full, desc, per_step


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    v, desc, per_step = cpu_sample(args.cpu_views, 518, max(1, args.steps), min(args.warmup, 1), threads,
                                   layer_frac=args.cpu_layer_frac)
    P = (args.size // 14) ** 2
    line = dict(metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=per_step * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                data="synthetic (seeded random-init weights, blurred-noise views)", impl="reference",
                # same workload as our arm; every step is a BOUNDED SAMPLE of it (cpu_baseline.sample): the
                # 1-view sample under-counts the quadratic attention of the 16-view scene, i.e. it favours the CPU
                config=dict(workload="G2VLM-2B-MoT recon bf16, 16 views 518px, single B200 (BASELINE configs[1]); one scene per GPU",
                            views_per_scene=args.views, image_size=args.size, tokens=args.views * (P + 2),
                            scenes_per_step=1, parallelism="host cores (reference algorithm, CPU)"),
                cpu_baseline=dict(value=v, unit=UNIT, cores=threads, kind="port", sample=desc),
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# view-sharded long scene (BASELINE configs[3]): sequence parallel over the ranks
# --------------------------------------------------------------------------------------------------
def run_view_sharded(args, cfg, dist, rank, world, local_rank):
    from g2vlm_b200 import ops, schema
    from g2vlm_b200.model import G2VLMFast

    n_views, size = args.views, args.size
    P = (size // 14) ** 2
    sd = schema.init_synthetic(cfg, seed=0, embed_rows=32, device="cuda")
    model = G2VLMFast(cfg, sd)
    model.fuse_prompt = not args.no_fuse_prompt
    del sd
    torch.cuda.empty_cache()
    views_host = schema.synthetic_views(n_views, size, size, seed=1)
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
    l0 = ops.LAUNCHES
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop()
    barrier()
    if rank == 0:
        fl = algorithmic_flops(cfg, n_views, P)
        pk = peaks()
        per = ms / args.steps
        print(json.dumps(dict(
            metric=METRIC, value=n_views / (per / 1e3), unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=per, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="bf16",
            data="synthetic (seeded random-init weights, blurred-noise views)",
            config=dict(workload=f"G2VLM-2B-MoT long-sequence recon, ONE scene of {n_views} views {size}px, view-sharded "
                                 f"over {world} GPU(s): DINO by segment + neighbour exchange, per-layer K/V all-gather "
                                 f"(NCCL/NVLink), context broadcast", views_per_scene=n_views, image_size=size,
                        tokens=n_views * (P + 2), parallelism=f"view-sp{world}"),
            clocks=clocks, gpu_launches=ops.LAUNCHES - l0,
            e2e=dict(value=n_views / (per / 1e3), unit=UNIT, h2d_bytes_per_step=views_host.numel() * 4 // world,
                     d2h_bytes_per_step=0, note="timed through recon_view_sharded from host views; outputs stay on the GPUs"),
            whole_step=dict(algorithmic_tflop=fl["total"] / 1e12,
                            achieved_tflops_per_gpu=fl["total"] / 1e12 / (per / 1e3) / world,
                            frac_of_peak=fl["total"] / 1e12 / (per / 1e3) / world / pk["tflops"]))))
    if dist is not None:
        dist.destroy_process_group()


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
    ap.add_argument("--cpu-views", type=int, default=1)
    ap.add_argument("--cpu-layer-frac", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fuse-prompt", action="store_true",
                    help="run the 7-token prompt prefill as a separate und pass (reference order) instead of fused into the geo step")
    ap.add_argument("--profile", action="store_true",
                    help="profiling run under ncu: skip the e2e arm and the CPU baseline, allow warmup < 3 (not a bench value)")
    ap.add_argument("--tiny", action="store_true", help="tiny model dims (smoke / debugging only; INVALID as a benchmark)")
    ap.add_argument("--workload", default="scenes", choices=["scenes", "views"],
                    help="scenes: one 16-view scene per GPU (default, BASELINE configs[1]/[2]); views: ONE scene of "
                         "--views views split by view over the GPUs (BASELINE configs[3], sequence parallel)")
    args = ap.parse_args()

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
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.warmup < 3 and not args.profile:
        args.warmup = 3  # timing rule: at least 3 warm-up steps

    cfg = schema.TINY if args.tiny else schema.FULL
    if args.workload == "views":
        run_view_sharded(args, cfg, dist, rank, world, local_rank)
        return
    n_views, size = args.views, args.size
    P = (size // 14) ** 2
    sd = schema.init_synthetic(cfg, seed=0, embed_rows=32, device="cuda")
    model = G2VLMFast(cfg, sd)
    model.fuse_prompt = not args.no_fuse_prompt
    del sd
    torch.cuda.empty_cache()
    views_host = schema.synthetic_views(n_views, size, size, seed=1 + rank).pin_memory()
    tok = StubTokenizer()

    # ---- device-resident arm ("value") ---------------------------------------------------------
    gi_text, newlens, new_rope = model.prepare_prompts_addbos([0], [0], ["x"], tok, TOKENS)
    # the *_lens tensors are host-side control metadata (reading them from the device would stall the host every step)
    HOST_META = ("text_token_lens", "key_values_lens", "packed_seqlens", "dino_token_seqlens")
    gi_text = {k: (v if k in HOST_META else v.cuda()) for k, v in gi_text.items()}
    gi, _, _ = model.prepare_dino_images_pi3(newlens, new_rope, views_host, None, TOKENS)
    gi = {k: (v if k in HOST_META else v.cuda()) for k, v in gi.items()}

    att_events = []
    orig_attention = ops.attention
    T = n_views * (P + 2)
    K_PROMPT = int(gi_text["packed_text_ids"].numel())   # prompt rows ride along in the fused path

    def timed_attention(q, *a, **k):
        if q.shape[0] in (T, T + K_PROMPT) and k.get("num_kv_heads") == cfg.num_kv_heads and not k.get("causal", False):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = orig_attention(q, *a, **k)
            e1.record()
            att_events.append((e0, e1))
            return r
        return orig_attention(q, *a, **k)

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
        launches0 = ops.LAUNCHES
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
        return ms, ops.LAUNCHES - launches0

    for _ in range(args.warmup):
        step_resident()
    ops.attention = timed_attention
    model.stage_events = []
    sampler = ClockSampler(local_rank)
    ms_total, launches = timed(step_resident, args.steps)
    clocks = sampler.stop()
    ops.attention = orig_attention
    stage_events, model.stage_events = model.stage_events, None

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
        data="synthetic (seeded random-init weights, blurred-noise views)",
        config=dict(workload="G2VLM-2B-MoT recon bf16, 16 views 518px, single B200 (BASELINE configs[1]); one scene per GPU"
                    if not args.tiny else "TINY DEBUG MODEL (invalid as a benchmark)",
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
    if world == 1 and not args.no_cpu_baseline and not args.tiny:
        try:
            v, desc, _ = cpu_sample(args.cpu_views, 518, 1, 0, os.cpu_count() or 1, layer_frac=args.cpu_layer_frac)
            line["cpu_baseline"] = dict(value=v, unit=UNIT, cores=os.cpu_count() or 1, kind="port", sample=desc)
        except Exception as e:  # the baseline must never take the GPU number down with it
            line["cpu_baseline"] = dict(value=None, unit=UNIT, cores=os.cpu_count() or 1, kind="port",
                                        sample=f"failed: {type(e).__name__}: {e}")
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
