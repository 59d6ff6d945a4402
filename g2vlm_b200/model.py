"""Host-side mirror of the reference model interface for the recon path, driving the CUDA kernels.

`G2VLMFast` keeps the reference's call shapes:
    recon(tokenizer, new_token_ids, dino_image_transform, images, prompt=...) -> dict
    forward_cache_update_text(...) -> NaiveCache
    forward_cache_update_dino(...) -> (NaiveCache, last_hidden_state)
    reconstruct(...) -> dict
(reference: modeling/g2vlm/g2vlm.py:1240-1303, 701-733, 968-1039, 1143-1238) and consumes the
reference `state_dict` key names unchanged (`from_state_dict`).  All arithmetic runs in the kernel
library through `g2vlm_b200.ops`; torch is used for device buffers, streams and host<->device copies
only.  There is no CPU / PyTorch fallback: constructing the model without the built library raises.

Internal row order of the MoT stack ("gather by expert"): the T packed tokens are permuted ONCE per
forward to [geo rows (packed_dino_token_indexes order) | und rows (packed_text_indexes order)] so each
expert's rows are contiguous and every routed Linear is one grouped GEMM; non-causal attention is
permutation invariant, and the result is scattered back to the packed order at the end.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional, Sequence

import torch

from . import _lib, host_prep, ops
from .schema import G2Config, state_dict_schema
from .sharding import view_ranges


class NaiveCache:
    """Same contract as the reference's NaiveCache (modeling/g2vlm/qwen2vl.py:237-251)."""

    def __init__(self, num_layers: int):
        self.key_cache = {k: None for k in range(num_layers)}
        self.value_cache = {k: None for k in range(num_layers)}

    @property
    def num_layers(self):
        return len(self.key_cache)

    @property
    def seq_lens(self):
        return self.key_cache[0].shape[0] if self.key_cache[0] is not None else 0


class _CacheView:
    """`cache.key_cache[layer]` / `cache.value_cache[layer]` of a KVCache: (L, n_kv, head_dim) bf16 views."""

    def __init__(self, cache, half):
        self._c, self._h = cache, half

    def __len__(self):
        return self._c._layers

    def __getitem__(self, layer):
        c = self._c
        if c.len == 0 or c.buf[layer] is None:
            return None
        w = c.kvw // 2
        return c.buf[layer][: c.len, self._h * w:(self._h + 1) * w].unflatten(1, (c.n_kv, c.hd))


class KVCache(NaiveCache):
    """Append-style KV cache with the NaiveCache read contract (row f1): one preallocated bf16 buffer
    [capacity, 2*n_kv*head_dim] (K | V) per layer; a step appends its rows in place instead of the reference's
    per-step re-allocation + index scatter of the whole cache (g2vlm/qwen2vl.py:621-638, O(L) per token)."""

    def __init__(self, num_layers, n_kv, hd, device):
        self._layers, self.n_kv, self.hd, self.kvw = num_layers, n_kv, hd, 2 * n_kv * hd
        self.device = device
        self.buf = [None] * num_layers
        self.len, self.cap = 0, 0
        self.key_cache, self.value_cache = _CacheView(self, 0), _CacheView(self, 1)

    @property
    def num_layers(self):
        return self._layers

    @property
    def seq_lens(self):
        return self.len

    def reserve(self, rows: int) -> None:
        if rows <= self.cap:
            return
        # headroom so the question prefill + decode steps that follow a big step never re-allocate
        cap = (max(rows + 1024, int(self.cap * 1.25)) + 1023) // 1024 * 1024
        for i in range(self._layers):
            nb = torch.empty(cap, self.kvw, dtype=torch.bfloat16, device=self.device)
            if self.buf[i] is not None and self.len:
                ops.gather_rows(self.buf[i], nb, None, self.len)
            self.buf[i] = nb
        self.cap = cap

    @classmethod
    def adopt(cls, past, cfg, device) -> "KVCache":
        """Accept a reference-style NaiveCache (possibly filled by reference code) or a KVCache."""
        if isinstance(past, KVCache):
            return past
        c = cls(cfg.num_layers, cfg.num_kv_heads, cfg.head_dim, device)
        if past is not None and past.key_cache[0] is not None:
            L = past.key_cache[0].shape[0]
            c.reserve(L)
            w = c.kvw // 2
            for i in range(cfg.num_layers):
                c.buf[i][:L, :w] = past.key_cache[i].reshape(L, w).to(device, torch.bfloat16)
                c.buf[i][:L, w:] = past.value_cache[i].reshape(L, w).to(device, torch.bfloat16)
            c.len = L
        return c


def _pad_dim(hd: int) -> int:
    if hd <= 64:
        return 64
    if hd <= 128:
        return 128
    raise ValueError(f"head_dim {hd} > 128 is not supported by the attention kernel")


def _bf16(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(torch.bfloat16).contiguous().to(device)


def _f32(t: torch.Tensor, device) -> torch.Tensor:
    return t.float().contiguous().to(device)


def _bias_bf16(t: torch.Tensor, device) -> torch.Tensor:
    """autocast casts the bias of a Linear to bf16: keep its bf16-rounded value, stored as fp32."""
    return t.to(torch.bfloat16).float().contiguous().to(device)


def _pad_head_rows(w: torch.Tensor, n_heads: int, hd: int, hp: int) -> torch.Tensor:
    """[n_heads*hd, ...] -> [n_heads*hp, ...] (zero rows appended to every head)."""
    if hd == hp:
        return w
    shape = (n_heads, hd) + tuple(w.shape[1:])
    out = torch.zeros((n_heads, hp) + tuple(w.shape[1:]), dtype=w.dtype, device=w.device)
    out[:, :hd] = w.reshape(shape)
    return out.reshape((n_heads * hp,) + tuple(w.shape[1:]))


def _pad_head_cols(w: torch.Tensor, n_heads: int, hd: int, hp: int) -> torch.Tensor:
    """[out, n_heads*hd] -> [out, n_heads*hp]."""
    if hd == hp:
        return w
    out = torch.zeros(w.shape[0], n_heads, hp, dtype=w.dtype, device=w.device)
    out[:, :, :hd] = w.reshape(w.shape[0], n_heads, hd)
    return out.reshape(w.shape[0], n_heads * hp)


def _interleave_gate_up(gate: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    """[I, H] x 2 -> [2I, H] with blocks of 128 gate rows followed by the matching 128 up rows."""
    I, H = gate.shape
    assert I % 128 == 0, "intermediate_size must be a multiple of 128"
    return torch.stack([gate.view(I // 128, 128, H), up.view(I // 128, 128, H)], dim=1).reshape(2 * I, H)


def _split_hi_lo_hi(w: torch.Tensor) -> torch.Tensor:
    """fp32 weight [N, K] -> bf16 [N, 3K] = [hi | lo | hi] (pairs with activations [hi | hi | lo])."""
    hi = w.to(torch.bfloat16)
    lo = (w - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, lo, hi], dim=1)


def _on_device(fn):
    """Public entry points run with the model's GPU as the current device (ops launch on the current device's stream),
    so a process may hold models on several GPUs."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        if torch.cuda.current_device() == self.device.index:
            return fn(self, *a, **k)
        with torch.cuda.device(self.device):
            return fn(self, *a, **k)
    return wrapped


class _Buffers:
    """Shape-keyed cache of device workspaces (allocated once, reused across calls)."""

    def __init__(self, device):
        self.device = device
        self._b: Dict[tuple, torch.Tensor] = {}

    def get(self, name: str, shape, dtype, zero: bool = False) -> torch.Tensor:
        key = (name, tuple(shape), dtype)
        t = self._b.get(key)
        if t is None:
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.device)
            self._b[key] = t
        return t


class G2VLMFast:
    mode = "bf16"

    def __new__(cls, cfg=None, state_dict=None, device="cuda", mode: str = "bf16"):
        if cls is G2VLMFast and mode == "fp32":
            from .model_fp32 import G2VLMFastFP32
            return super().__new__(G2VLMFastFP32)
        if mode not in ("bf16", "fp32"):
            raise ValueError(f"mode must be 'bf16' or 'fp32', got {mode!r}")
        return super().__new__(cls)

    def __init__(self, cfg: G2Config, state_dict: Dict[str, torch.Tensor], device="cuda", mode: str = "bf16"):
        """mode="bf16": the reference's autocast arithmetic (rounding points reproduced, tolerance 2e-2);
        mode="fp32": no bf16 rounding anywhere (g2vlm_b200/model_fp32.py; <= 1e-4 against oracle/restate.py
        mode="fp32"; recon only)."""
        if not torch.cuda.is_available():
            raise RuntimeError("G2VLMFast needs a CUDA device (sm_100a); there is no CPU fallback")
        _lib.load()  # fail loudly if the kernel library is missing
        if cfg.head_dim != 128:
            raise ValueError("LLM head_dim must be 128 (mrope_section is hard-coded to [16,24,24])")
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.buf = _Buffers(self.device)
        self._stage: Dict[str, dict] = {}
        self.device_resize = True   # recon(): LANCZOS resize of path / PIL inputs on the device (host_prep)
        self.fuse_prompt = True   # recon(): run the prompt prefill inside the geo step (language_model_forward_geo)
        missing = [k for k in state_dict_schema(cfg) if k not in state_dict and k != "language_model.lm_head.weight"]
        if missing:
            raise KeyError(f"state_dict is missing {len(missing)} keys, e.g. {missing[:4]}")
        self._pack(state_dict)
        self._rope2d_cache: Dict[tuple, tuple] = {}
        self._pos_cache: Dict[tuple, torch.Tensor] = {}
        self._work_cache: Dict[tuple, torch.Tensor] = {}
        self._dino_attn_cover: Dict[int, int] = {}   # zero-initialised DINO attention buffer -> rows covered last time
        self.stage_events: Optional[list] = None  # bench.py: [(name, cuda event)] at stage boundaries
        # view-sharded K/V exchange: "peer" (v3: symmetric memory + copy-engine pulls behind the local-key attention, no SM
        # taken), "overlap" (v2: NCCL point-to-point exchange behind the local-key attention; sp_sm_margin SMs left to
        # the NCCL kernel) or "allgather" (v1: one blocking all-gather per layer); all but v1 end in an LSE merge
        self.sp_mode = "peer"
        self.sp_peer_margin = 2                   # SMs the local-key attention leaves to the barrier kernel in peer mode
        self.sp_trace: Optional[list] = None      # tools/sp_peer_trace.py: per-layer events of the peer exchange
        self._sp_sym: Dict[tuple, object] = {}
        self.sp_sm_margin = 4    # = NCCL_MAX_NCHANNELS the launcher sets (bench.py); 2-GPU sweep: profiles/r02_sp_sweep_n2.txt
        self.sp_events: Optional[list] = None     # bench.py: CUDA event pairs around every exchange wait
        # stack-level C ABI (g2vlm_dino_forward / g2vlm_mot_forward_geo / g2vlm_recon_heads): the fused recon path is
        # THREE native calls instead of ~630 per-op calls from Python; False keeps the per-op path (bit-identical)
        self.native = self.mode == "bf16"
        self._nctx = None
        self._nws: Dict[tuple, torch.Tensor] = {}
        self._nplan: Optional[tuple] = None
        self.attn_events: Optional[list] = None   # bench.py: (start, end) CUDA events of every MoT attention launch
        self._raw_images = False  # True while recon() feeds un-normalised views (normalised on device)

    _NVTX = {"dino_begin": (0, "g2vlm.dino_encoder"), "dino_end": (1, "g2vlm.mot_stack"), "mot_end": (1, None),
             "heads_begin": (0, "g2vlm.pi3_heads"), "heads_end": (1, None)}

    def _mark(self, name: str) -> None:
        """Stage boundary: NVTX range (nsys / ncu --nvtx see dino_encoder / mot_stack / pi3_heads) and, for bench.py,
        a CUDA event."""
        pop, push = self._NVTX.get(name, (0, None))
        if pop:
            torch.cuda.nvtx.range_pop()
        if push:
            torch.cuda.nvtx.range_push(push)
        if self.stage_events is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.stage_events.append((name, ev))

    def _sp_mark(self, which: int) -> None:
        """Event pair around the point where the compute stream waits for the K/V exchange (exposed time)."""
        if self.sp_events is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            if which == 0:
                self.sp_events.append([ev, None])
            else:
                self.sp_events[-1][1] = ev

    def _sp_symmetric(self, group, t_max: int, kvw: int):
        """(symmetric K|V send buffer bf16 [2, t_max, kvw], its rendezvous handle, side stream) for the view-sharded
        exchange, or None when symmetric memory is unavailable on ANY rank (decided collectively).  Cached."""
        import torch.distributed as dist
        key = (id(group), t_max, kvw)
        if key in self._sp_sym:
            return self._sp_sym[key]
        t, ok = None, 1
        try:
            import torch.distributed._symmetric_memory as symm
            t = symm.empty((2, t_max, kvw), dtype=torch.bfloat16, device=self.device)
        except Exception:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        res = None
        if int(flag.item()):
            hdl = symm.rendezvous(t, group)
            res = (t, hdl, torch.cuda.Stream(device=self.device))
        self._sp_sym[key] = res
        return res

    @classmethod
    def from_state_dict(cls, cfg: G2Config, state_dict, device="cuda", mode: str = "bf16") -> "G2VLMFast":
        return cls(cfg, state_dict, device, mode=mode)

    RECON_STAGES = ("forward_cache_update_text", "forward_cache_update_dino", "reconstruct")

    @classmethod
    def attach(cls, ref_model, device=None, stages=RECON_STAGES) -> "G2VLMFast":
        """INTEGRATION.md §2 as code: build the CUDA path from a live reference `G2VLM` instance (its `config` and
        `state_dict()`, key names unchanged) and replace the recon stages of THAT object — signatures identical to
        g2vlm.py:701-711, 968-983, 1143-1154 — so `ref_model.recon(...)` (g2vlm.py:1240-1303), its host-side
        `prepare_*` code and every caller (`inference_recon.py:36-43`) keep working unchanged.  The reference's
        `NaiveCache` objects are adopted transparently (KVCache.adopt).  Duck-typed: nothing of the reference is
        imported here.  Returns the G2VLMFast (also stored as `ref_model._g2vlm_b200`)."""
        c = ref_model.config
        vit = c.vit_config if getattr(ref_model, "vit_model", None) is not None else None
        cfg = G2Config.from_reference(c.llm_config, c.dino_config, vit,
                                      train_conf_pi3=bool(getattr(c, "train_conf_pi3", False)))
        if device is None:
            device = next(ref_model.parameters()).device
            if device.type != "cuda":
                device = torch.device("cuda", torch.cuda.current_device())
        fast = cls(cfg, ref_model.state_dict(), device)
        for name in stages:
            setattr(ref_model, name, getattr(fast, name))
        ref_model._g2vlm_b200 = fast
        return fast

    # ------------------------------------------------------------------------------------------
    # weight packing (one-time; reference key names in, kernel-friendly layouts out)
    # ------------------------------------------------------------------------------------------
    def _pack(self, sd):
        cfg, dev = self.cfg, self.device
        g = lambda k: sd[k].detach().float()  # packed on whatever device the checkpoint lives on
        lm = "language_model.model."
        self.embed = _f32(g(lm + "embed_tokens.weight"), dev)
        self.layers = []
        for i in range(cfg.num_layers):
            p = f"{lm}layers.{i}."
            a = p + "self_attn."
            L = {}
            qkv, bias = [], []
            for sfx in ("_moe_geo", ""):  # expert order: geo (group 0), und (group 1)
                qkv += [g(a + f"q_proj{sfx}.weight"), g(a + f"k_proj{sfx}.weight"), g(a + f"v_proj{sfx}.weight")]
                bias += [g(a + f"q_proj{sfx}.bias"), g(a + f"k_proj{sfx}.bias"), g(a + f"v_proj{sfx}.bias")]
            L["wqkv"] = _bf16(torch.cat(qkv, 0), dev)
            L["bqkv"] = _bias_bf16(torch.cat(bias, 0), dev)
            L["wo"] = _bf16(torch.cat([g(a + "o_proj_moe_geo.weight"), g(a + "o_proj.weight")], 0), dev)
            L["wgu"] = _bf16(torch.cat([
                _interleave_gate_up(g(p + "mlp_moe_geo.gate_proj.weight"), g(p + "mlp_moe_geo.up_proj.weight")),
                _interleave_gate_up(g(p + "mlp.gate_proj.weight"), g(p + "mlp.up_proj.weight"))], 0), dev)
            L["wdown"] = _bf16(torch.cat([g(p + "mlp_moe_geo.down_proj.weight"), g(p + "mlp.down_proj.weight")], 0), dev)
            for n in ("input_layernorm", "post_attention_layernorm"):
                L[n + "_geo"] = _f32(g(p + n + "_moe_geo.weight"), dev)
                L[n + "_und"] = _f32(g(p + n + ".weight"), dev)
            for n in ("q_norm", "k_norm"):
                L[n + "_geo"] = _f32(g(a + n + "_moe_geo.weight"), dev)
                L[n + "_und"] = _f32(g(a + n + ".weight"), dev)
            L["ls1"] = _f32(g(p + "ls1.gamma"), dev)
            L["ls2"] = _f32(g(p + "ls2.gamma"), dev)
            self.layers.append(L)
        self.lm_head = (_bf16(g("language_model.lm_head.weight"), dev)
                        if "language_model.lm_head.weight" in sd else None)
        self.norm_geo = _f32(g(lm + "norm_moe_geo.weight"), dev)
        self.norm_und = _f32(g(lm + "norm.weight"), dev)
        hd = cfg.head_dim
        inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
        self.inv_freq = inv.contiguous().to(dev)

        # ---- DINO ----
        D, nh, dhd = cfg.dino_hidden, cfg.dino_heads, cfg.dino_head_dim
        self.dino_hp = _pad_dim(dhd)
        d = "dino_model."
        kp = 3 * cfg.dino_patch ** 2
        self.dino_kpad = (kp + 63) // 64 * 64
        wpatch = g(d + "embeddings.patch_embeddings.projection.weight").reshape(D, kp)
        wp = torch.zeros(D, self.dino_kpad, device=wpatch.device)
        wp[:, :kp] = wpatch
        self.dino_wpatch = _bf16(wp, dev)
        self.dino_bpatch = _bias_bf16(g(d + "embeddings.patch_embeddings.projection.bias"), dev)
        self.dino_cls = _f32(g(d + "embeddings.cls_token").reshape(D), dev)
        self.dino_reg = _f32(g(d + "embeddings.register_tokens").reshape(cfg.dino_registers, D), dev)
        self.dino_pos_table = g(d + "embeddings.position_embeddings")  # master table, resampled per grid
        self.dino_layers = []
        for i in range(cfg.dino_layers):
            p = f"{d}encoder.layer.{i}."
            at = p + "attention.attention."
            L = {}
            L["wqkv"] = _bf16(torch.cat([_pad_head_rows(g(at + f"{n}.weight"), nh, dhd, self.dino_hp)
                                         for n in ("query", "key", "value")], 0), dev)
            L["bqkv"] = _bias_bf16(torch.cat([_pad_head_rows(g(at + f"{n}.bias"), nh, dhd, self.dino_hp)
                                              for n in ("query", "key", "value")], 0), dev)
            L["wdense"] = _bf16(_pad_head_cols(g(p + "attention.output.dense.weight"), nh, dhd, self.dino_hp), dev)
            L["bdense"] = _bias_bf16(g(p + "attention.output.dense.bias"), dev)
            L["wfc1"] = _bf16(g(p + "mlp.fc1.weight"), dev); L["bfc1"] = _bias_bf16(g(p + "mlp.fc1.bias"), dev)
            L["wfc2"] = _bf16(g(p + "mlp.fc2.weight"), dev); L["bfc2"] = _bias_bf16(g(p + "mlp.fc2.bias"), dev)
            for n in ("norm1", "norm2"):
                L[n + "w"] = _f32(g(p + n + ".weight"), dev); L[n + "b"] = _f32(g(p + n + ".bias"), dev)
            L["ls1"] = _f32(g(p + "layer_scale1.lambda1"), dev)
            L["ls2"] = _f32(g(p + "layer_scale2.lambda1"), dev)
            self.dino_layers.append(L)
        self.dino_lnw = _f32(g(d + "layernorm.weight"), dev)
        self.dino_lnb = _f32(g(d + "layernorm.bias"), dev)
        self.w_dino2llm = _bf16(g("dino2llm.weight"), dev)
        self.b_dino2llm = _bias_bf16(g("dino2llm.bias"), dev)

        # ---- Pi3 decoders ----
        H, dh, ehd = cfg.hidden_size, cfg.dec_heads, cfg.dec_head_dim
        self.dec_hp = _pad_dim(ehd)
        # 96-wide Pi3 heads live in 128-wide q/k/v slots for the attention kernel.  When the head width is a multiple
        # of 32 the Linear layers around it stay UNPADDED: the qkv GEMM scatters its columns into the slots
        # (out_col_group/out_col_stride) and the attention kernel writes a compact [rows, heads*96] output, so no
        # zero weight rows / columns are multiplied (25 % of the qkv and proj GEMMs otherwise).
        self.dec_compact = ehd != self.dec_hp and ehd % 32 == 0
        hp_w = ehd if self.dec_compact else self.dec_hp   # head width the packed weights use

        def pack_block(p, cross):
            B = {}
            qkv_w = g(p + "attn.qkv.weight").reshape(3, dh * ehd, H)
            qkv_b = g(p + "attn.qkv.bias").reshape(3, dh * ehd)
            B["wqkv"] = _bf16(torch.cat([_pad_head_rows(qkv_w[j], dh, ehd, hp_w) for j in range(3)], 0), dev)
            B["bqkv"] = _bias_bf16(torch.cat([_pad_head_rows(qkv_b[j], dh, ehd, hp_w) for j in range(3)], 0), dev)
            B["wproj"] = _bf16(_pad_head_cols(g(p + "attn.proj.weight"), dh, ehd, hp_w), dev)
            B["bproj"] = _bias_bf16(g(p + "attn.proj.bias"), dev)
            names = ["norm1", "norm2"] + (["norm3", "norm_y"] if cross else [])
            for n in names:
                B[n + "w"] = _f32(g(p + n + ".weight"), dev); B[n + "b"] = _f32(g(p + n + ".bias"), dev)
            if cross:
                c = p + "cross_attn."
                B["wcq"] = _bf16(_pad_head_rows(g(c + "q_proj.weight"), dh, ehd, hp_w), dev)
                B["bcq"] = _bias_bf16(_pad_head_rows(g(c + "q_proj.bias"), dh, ehd, hp_w), dev)
                B["wckv"] = _bf16(torch.cat([_pad_head_rows(g(c + "k_proj.weight"), dh, ehd, hp_w),
                                             _pad_head_rows(g(c + "v_proj.weight"), dh, ehd, hp_w)], 0), dev)
                B["bckv"] = _bias_bf16(torch.cat([_pad_head_rows(g(c + "k_proj.bias"), dh, ehd, hp_w),
                                                  _pad_head_rows(g(c + "v_proj.bias"), dh, ehd, hp_w)], 0), dev)
                B["wcproj"] = _bf16(_pad_head_cols(g(c + "proj.weight"), dh, ehd, hp_w), dev)
                B["bcproj"] = _bias_bf16(g(c + "proj.bias"), dev)
            B["wfc1"] = _bf16(g(p + "mlp.fc1.weight"), dev); B["bfc1"] = _bias_bf16(g(p + "mlp.fc1.bias"), dev)
            B["wfc2"] = _bf16(g(p + "mlp.fc2.weight"), dev); B["bfc2"] = _bias_bf16(g(p + "mlp.fc2.bias"), dev)
            return B

        self.decoders = {}
        dec_names = [("point_decoder", False), ("camera_decoder", False), ("global_points_decoder", True)]
        head_names = ["point_head", "global_point_head"]
        if cfg.train_conf_pi3:  # confidence branch (g2vlm.py:209-226): a copy of the point decoder + 1-channel head
            dec_names.append(("conf_decoder", False))
            head_names.append("conf_head")
        for name, cross in dec_names:
            blocks = [pack_block(f"{name}.blocks.{i}.", cross) for i in range(cfg.dec_depth)]
            self.decoders[name] = dict(blocks=blocks, cross=cross,
                                       wout=_bf16(g(f"{name}.linear_out.weight"), dev),
                                       bout=_bias_bf16(g(f"{name}.linear_out.bias"), dev))
        # fp32 heads (autocast disabled in the reference): split-bf16 weights, exact fp32 biases
        for hname in head_names:
            w = g(hname + ".proj.weight")
            hi = w.to(torch.bfloat16)
            setattr(self, hname + "_whi", _bf16(hi, dev))
            setattr(self, hname + "_wlo", _bf16(w - hi.float(), dev))
            setattr(self, hname + "_b", _f32(g(hname + ".proj.bias"), dev))
        self.cam = {}
        for i in range(2):
            for j in (1, 2, 3):
                k = f"camera_head.res_conv.{i}.res_conv{j}"
                self.cam[f"r{i}{j}w"] = _bf16(_split_hi_lo_hi(g(k + ".weight")), dev)
                self.cam[f"r{i}{j}b"] = _f32(g(k + ".bias"), dev)
        for j in (0, 2):
            self.cam[f"m{j}w"] = _bf16(_split_hi_lo_hi(g(f"camera_head.more_mlps.{j}.weight")), dev)
            self.cam[f"m{j}b"] = _f32(g(f"camera_head.more_mlps.{j}.bias"), dev)
        for n in ("fc_t", "fc_rot"):
            self.cam[n + "w"] = _f32(g(f"camera_head.{n}.weight"), dev)
            self.cam[n + "b"] = _f32(g(f"camera_head.{n}.bias"), dev)

        # ---- Qwen2-VL ViT (chat path, row f2) ----
        self.vit = None
        if cfg.vit_depth > 0 and "vit_model.patch_embed.proj.weight" in sd:
            E, vh = cfg.vit_embed_dim, cfg.vit_heads
            vhd = E // vh
            self.vit_hp = _pad_dim(vhd)
            v = "vit_model."
            V = dict(wpatch=_bf16(g(v + "patch_embed.proj.weight").reshape(E, -1), dev), blocks=[])
            for i in range(cfg.vit_depth):
                p = f"{v}blocks.{i}."
                qw = g(p + "attn.qkv.weight").reshape(3, vh * vhd, E)
                qb = g(p + "attn.qkv.bias").reshape(3, vh * vhd)
                B = dict(wqkv=_bf16(torch.cat([_pad_head_rows(qw[j], vh, vhd, self.vit_hp) for j in range(3)], 0), dev),
                         bqkv=_bias_bf16(torch.cat([_pad_head_rows(qb[j], vh, vhd, self.vit_hp) for j in range(3)], 0), dev),
                         wproj=_bf16(_pad_head_cols(g(p + "attn.proj.weight"), vh, vhd, self.vit_hp), dev),
                         bproj=_bias_bf16(g(p + "attn.proj.bias"), dev),
                         wfc1=_bf16(g(p + "mlp.fc1.weight"), dev), bfc1=_bias_bf16(g(p + "mlp.fc1.bias"), dev),
                         wfc2=_bf16(g(p + "mlp.fc2.weight"), dev), bfc2=_bias_bf16(g(p + "mlp.fc2.bias"), dev))
                for n in ("norm1", "norm2"):
                    B[n + "w"] = _f32(g(p + n + ".weight"), dev); B[n + "b"] = _f32(g(p + n + ".bias"), dev)
                V["blocks"].append(B)
            V.update(lnw=_f32(g(v + "merger.ln_q.weight"), dev), lnb=_f32(g(v + "merger.ln_q.bias"), dev),
                     wm0=_bf16(g(v + "merger.mlp.0.weight"), dev), bm0=_bias_bf16(g(v + "merger.mlp.0.bias"), dev),
                     wm2=_bf16(g(v + "merger.mlp.2.weight"), dev), bm2=_bias_bf16(g(v + "merger.mlp.2.bias"), dev))
            self.vit = V

    # ------------------------------------------------------------------------------------------
    # host -> device staging of the per-call index tensors
    # ------------------------------------------------------------------------------------------
    def _idx(self, name: str, t: torch.Tensor, dtype=torch.long) -> torch.Tensor:
        """Device copy of a host-built index tensor WITHOUT stalling the host: the values go through one of two
        reusable pinned staging buffers (guarded by an event) into a reusable device buffer, so the copy is a true
        async H2D in stream order — a pageable `.to(device)` would block the host until the stream drains, which
        stops it from enqueueing the next scene while this one computes.  Device tensors pass through."""
        if t.is_cuda:
            return t.to(dtype).contiguous()
        slot = self._stage.setdefault(name, dict(i=0, pin=[None, None], ev=[None, None]))
        i = slot["i"]
        slot["i"] = i ^ 1
        pin = slot["pin"][i]
        if pin is None or pin.shape != t.shape or pin.dtype != dtype:
            pin = slot["pin"][i] = torch.empty(t.shape, dtype=dtype, pin_memory=True)
        elif slot["ev"][i] is not None:
            slot["ev"][i].synchronize()       # the upload issued two calls ago has left this buffer (long done)
        pin.copy_(t)
        d = self.buf.get("idx." + name, tuple(t.shape), dtype)
        d.copy_(pin, non_blocking=True)
        if slot["ev"][i] is None:
            slot["ev"][i] = torch.cuda.Event()
        slot["ev"][i].record()
        return d

    # ------------------------------------------------------------------------------------------
    # small cached host-built tables
    # ------------------------------------------------------------------------------------------
    def _work(self, cu_q: Sequence[int], cu_k: Sequence[int], tag: str) -> torch.Tensor:
        key = (tag, tuple(cu_q), tuple(cu_k))
        t = self._work_cache.get(key)
        if t is None:
            t = ops.attention_work_table(cu_q, cu_k).to(self.device)
            self._work_cache[key] = t
        return t

    def _cross_work(self, n_views: int, P: int) -> torch.Tensor:
        key = ("cross", n_views, P)
        t = self._work_cache.get(key)
        if t is None:
            items = [[t0, v * P, (v + 1) * P, 0, P, 0, 0, 0]
                     for v in range(n_views) for t0 in range(v * P, (v + 1) * P, ops.ATTN_ROWS_PER_ITEM)]
            t = torch.tensor(items, dtype=torch.int32).to(self.device)
            self._work_cache[key] = t
        return t

    def _dino_pos(self, gh: int, gw: int, square: bool) -> torch.Tensor:
        """position_embeddings for a gh x gw grid; bicubic-antialias resample of the learned table when
        the grid differs (interpolate_pos_encoding, modeling_dinov2_with_registers.py:93-145).  This
        depends on the weights and the grid only, so it is computed once per grid and cached."""
        key = (gh, gw, square)
        t = self._pos_cache.get(key)
        if t is None:
            cfg = self.cfg
            pos = self.dino_pos_table
            g0 = cfg.dino_grid
            if not (gh * gw == g0 * g0 and square):
                pp = pos[:, 1:].reshape(1, g0, g0, -1).permute(0, 3, 1, 2)
                pp = torch.nn.functional.interpolate(pp, size=(gh, gw), mode="bicubic", align_corners=False,
                                                     antialias=True)
                pos = torch.cat((pos[:, :1], pp.permute(0, 2, 3, 1).reshape(1, gh * gw, -1)), dim=1)
            t = pos[0].contiguous().to(self.device)
            self._pos_cache[key] = t
        return t

    def _rope2d_tables(self, gh: int, gw: int):
        """cos/sin tables of RoPE2D built exactly like the reference's cache (pos_embed.py:118-128):
        angles cast to bf16 BEFORE cos/sin (quirk Q2), cos/sin stored in bf16.  Unique columns only."""
        key = (gh, gw)
        t = self._rope2d_cache.get(key)
        if t is None:
            D = self.cfg.dec_head_dim // 2
            inv_freq = 1.0 / (self.cfg.rope2d_base ** (torch.arange(0, D, 2).float() / D))
            tt = torch.arange(max(gh, gw), dtype=inv_freq.dtype)
            freqs = torch.einsum("i,j->ij", tt, inv_freq).to(torch.bfloat16)
            t = (freqs.cos().float().contiguous().to(self.device), freqs.sin().float().contiguous().to(self.device))
            self._rope2d_cache[key] = t
        return t

    # ------------------------------------------------------------------------------------------
    # stack-level C ABI: opaque context, caller-owned workspace, one call per stage
    # ------------------------------------------------------------------------------------------
    def _native_ctx(self):
        if self._nctx is None:
            c = ops.NativeContext(self.cfg)
            reg = c.load
            reg("embed", self.embed); reg("inv_freq", self.inv_freq)
            reg("norm_geo", self.norm_geo); reg("norm_und", self.norm_und)
            reg("dino2llm.w", self.w_dino2llm); reg("dino2llm.b", self.b_dino2llm)
            for i, L in enumerate(self.layers):
                for k, t in L.items():
                    reg(f"mot.{i}.{k}", t)
            reg("dino.wpatch", self.dino_wpatch); reg("dino.bpatch", self.dino_bpatch)
            reg("dino.cls", self.dino_cls); reg("dino.reg", self.dino_reg)
            reg("dino.lnw", self.dino_lnw); reg("dino.lnb", self.dino_lnb)
            for i, L in enumerate(self.dino_layers):
                for k, t in L.items():
                    reg(f"dino.{i}.{k}", t)
            for name, dec in self.decoders.items():
                for b, B in enumerate(dec["blocks"]):
                    for k, t in B.items():
                        reg(f"dec.{name}.{b}.{k}", t)
                reg(f"dec.{name}.wout", dec["wout"]); reg(f"dec.{name}.bout", dec["bout"])
            for hname in ["point_head", "global_point_head"] + (["conf_head"] if self.cfg.train_conf_pi3 else []):
                for sfx in ("whi", "wlo", "b"):
                    reg(f"head.{hname}.{sfx}", getattr(self, f"{hname}_{sfx}"))
            for k, t in self.cam.items():
                reg(f"cam.{k}", t)
            self._nctx = c
        return self._nctx

    def _native_plan(self, N: int, Hh: int, Ww: int, Kp: int, seqlens) -> torch.Tensor:
        """Workspace of one geometry (cached) with its tables in place (g2vlm_recon_plan, once per geometry switch)."""
        ctx = self._native_ctx()
        seq = tuple(int(n) for n in seqlens)
        key = (N, Hh, Ww, Kp, seq)
        ws = self._nws.get(key[:4])
        if ws is None:
            ws = self._nws[key[:4]] = torch.empty(ctx.workspace_bytes(N, Hh, Ww, Kp), dtype=torch.uint8, device=self.device)
        if self._nplan != key:
            import ctypes
            arr = (ctypes.c_int32 * N)(*seq)
            ctx.call("g2vlm_recon_plan", ops._i32(N), ops._i32(Hh), ops._i32(Ww), ops._i32(Kp), arr, ops._vp(ws.data_ptr()),
                     ops._i64(ws.numel()), ops._stream())
            self._nplan = key
        return ws

    def _native_dino_mot(self, packed_text_ids, packed_text_indexes, packed_dino_token_indexes, dino_token_seqlens,
                         packed_position_ids, packed_dino_images, prompt):
        """Encoder + dino2llm + the MoT stack as TWO native calls (same kernels as the per-op path)."""
        import ctypes
        cfg, dev = self.cfg, self.device
        img = packed_dino_images.to(dev, torch.float32).contiguous()
        N, _, Hh, Ww = img.shape
        p = cfg.dino_patch
        gh, gw = Hh // p, Ww // p
        P = gh * gw
        T = N * (P + 2)
        if int(packed_text_ids.numel()) != 2 * N or int(packed_dino_token_indexes.numel()) != N * P:
            raise ValueError("index tensors do not match the image geometry")
        Kp = int(prompt["packed_text_ids"].numel())
        ws = self._native_plan(N, Hh, Ww, Kp, dino_token_seqlens.tolist())
        ctx, st = self._native_ctx(), ops._stream()
        tokens = ctypes.c_void_p()
        self._mark("dino_begin")
        ctx.call("g2vlm_dino_forward", ops._vp(img.data_ptr()), ops._i32(N), ops._i32(Hh), ops._i32(Ww),
                 ops._i32(int(self._raw_images)), ops._vp(self._dino_pos(gh, gw, Hh == Ww).data_ptr()), ops._vp(ws.data_ptr()),
                 ctypes.byref(tokens), st)
        self._mark("dino_end")
        ids = self._idx("dino.txt_ids", packed_text_ids)
        tidx = self._idx("dino.txt_idx", packed_text_indexes)
        gidx = self._idx("dino.geo_idx", packed_dino_token_indexes)
        pos = self._idx("mot.pos", packed_position_ids.contiguous())
        pids = self._idx("mot.prompt_ids", prompt["packed_text_ids"])
        ppos = self._idx("mot.prompt_pos", prompt["packed_text_position_ids"].contiguous())
        last = torch.empty(T, cfg.hidden_size, dtype=torch.float32, device=dev)
        events = None
        if self.attn_events is not None:
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(2 * cfg.num_layers)]
            for e in evs:
                e.record()                      # creates the underlying cudaEvent_t; re-recorded by the driver
            events = (ctypes.c_void_p * len(evs))(*[e.cuda_event for e in evs])
            self.attn_events += [(evs[2 * i], evs[2 * i + 1]) for i in range(cfg.num_layers)]
        ctx.call("g2vlm_mot_forward_geo", tokens, ops._vp(ids.data_ptr()), ops._vp(tidx.data_ptr()), ops._vp(gidx.data_ptr()),
                 ops._vp(pos.data_ptr()), ops._vp(pids.data_ptr()), ops._vp(ppos.data_ptr()), ops._vp(ws.data_ptr()),
                 ops._vp(last.data_ptr()), events, st)
        self._mark("mot_end")
        return last

    def _native_heads(self, selected_hidden_states, packed_dino_token_indexes, N, Hh, Ww):
        cfg, dev = self.cfg, self.device
        p = cfg.dino_patch
        gh, gw = Hh // p, Ww // p
        P = gh * gw
        if self._nplan is None or self._nplan[:3] != (N, Hh, Ww):
            self._native_plan(N, Hh, Ww, 0, [P] * N)
        ws = self._nws[self._nplan[:4]]
        geo = self._idx("recon.geo_idx", packed_dino_token_indexes)
        cos, sin = self._rope2d_tables(gh, gw)
        out = {k: torch.empty(N, Hh, Ww, 3, dtype=torch.float32, device=dev) for k in ("points", "local_points", "global_points")}
        poses = torch.empty(N, 4, 4, dtype=torch.float32, device=dev)
        conf = torch.empty(N, Hh, Ww, 1, dtype=torch.float32, device=dev) if cfg.train_conf_pi3 else None
        hid = selected_hidden_states
        if hid.dtype != torch.float32 or not hid.is_contiguous() or hid.device != dev:
            hid = hid.to(dev, torch.float32).contiguous()
        self._native_ctx().call("g2vlm_recon_heads", ops._vp(hid.data_ptr()), ops._vp(geo.data_ptr()), ops._vp(cos.data_ptr()),
                                ops._vp(sin.data_ptr()), ops._vp(ws.data_ptr()), ops._vp(out["points"].data_ptr()),
                                ops._vp(out["local_points"].data_ptr()), ops._vp(out["global_points"].data_ptr()),
                                ops._vp(poses.data_ptr()), ops._ptr(conf), ops._stream())
        self._mark("heads_end")
        return out, poses, conf

    # ------------------------------------------------------------------------------------------
    # MoT language model
    # ------------------------------------------------------------------------------------------
    def _mot_layer(self, L, x, T, n_geo, qkv, attn, act, hbuf, cos, sin, work, kv_rows, causal, round_normed,
                   kv_exchange=None, kv_len_dev=None, prompt_rows=0, train=None, attn_override=None):
        # prompt_rows: the LAST prompt_rows of the T rows are the causal text prefill riding along with the geo
        # step (fused recon path); they take the und branch's rounding of the normed q/k (round_normed=True)
        cfg = self.cfg
        H, I = cfg.hidden_size, cfg.intermediate_size
        nq, nkv, hd = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim
        groups = [(0, n_geo), (n_geo, T - n_geo)]
        # train = dict(perm, qkv_packed, attn_packed): the bf16-module numerics of forward_train (extra bf16 rounding in
        # the norms / rotary, bf16 residual stream) and attention in PACKED row order (its masks depend on positions)
        rn = train is not None
        resid_flags = ops.GEMM_ROUND_AFTER_SCALE | (ops.GEMM_ROUND_SUM if rn else 0)
        ops.rmsnorm_routed(x, hbuf, L["input_layernorm_geo"], L["input_layernorm_und"], n_geo, cfg.rms_norm_eps, rows=T,
                           round_normed=rn)
        ops.gemm(hbuf[:T], L["wqkv"], qkv, epilogue=ops.EPI_STORE_BF16, groups=groups, bias=L["bqkv"])
        Tm = T - prompt_rows
        ops.qknorm_mrope(qkv, Tm, n_geo, nq, nkv, hd, L["q_norm_geo"], L["k_norm_geo"], L["q_norm_und"],
                         L["k_norm_und"], cos, sin, cfg.rms_norm_eps, round_normed=2 if rn else round_normed)
        if prompt_rows:
            ops.qknorm_mrope(qkv[Tm:], prompt_rows, 0, nq, nkv, hd, L["q_norm_geo"], L["k_norm_geo"], L["q_norm_und"],
                             L["k_norm_und"], cos[Tm:], sin[Tm:], cfg.rms_norm_eps, round_normed=True)
        if attn_override is not None:
            pass  # view-sharded overlap mode: the callback runs the exchange and the split attention itself
        elif kv_exchange is None:
            k_all, v_all = qkv[:kv_rows, nq * hd:(nq + nkv) * hd], qkv[:kv_rows, (nq + nkv) * hd:]
        else:  # append-style cache / view-sharded all-gather: K/V rows come from the callback
            k_all, v_all = kv_exchange(qkv)
        if attn_override is not None:
            attn_override(qkv, attn)
        elif rn:
            qp, ap = train["qkv_packed"], train["attn_packed"]
            ops.gather_rows(qkv[:T], qp, train["perm"], T, scatter=True)          # internal -> packed row order
            ops.attention(qp[:, : nq * hd], qp[:, nq * hd:(nq + nkv) * hd], qp[:, (nq + nkv) * hd:], ap, work,
                          num_q_heads=nq, num_kv_heads=nkv, head_dim=hd, scale=1.0 / math.sqrt(hd))
            ops.gather_rows(ap, attn, train["perm"], T)                            # packed -> internal
        elif T == 1 and nq // nkv <= 8:
            # decode step: one query row sees every cached key -> flash-decoding split over the keys
            ws = self.buf.get("und.dec_ws", (ops.attention_decode_workspace_floats(k_all.shape[0], nq),), torch.float32)
            ops.attention_decode(qkv[0, : nq * hd], k_all, v_all, attn[0], ws, num_q_heads=nq, num_kv_heads=nkv,
                                 head_dim=hd, scale=1.0 / math.sqrt(hd), kv_len_dev=kv_len_dev, kv_len_extra=1)
        else:
            ops.attention(qkv[:T, : nq * hd], k_all, v_all, attn, work, num_q_heads=nq, num_kv_heads=nkv, head_dim=hd,
                          scale=1.0 / math.sqrt(hd), causal=causal)
        ops.gemm(attn[:T], L["wo"], x, epilogue=ops.EPI_RESID_F32, groups=groups, scale=L["ls1"], scale_groups=1,
                 flags=resid_flags)
        ops.rmsnorm_routed(x, hbuf, L["post_attention_layernorm_geo"], L["post_attention_layernorm_und"], n_geo,
                           cfg.rms_norm_eps, rows=T, round_normed=rn)
        ops.gemm(hbuf[:T], L["wgu"], act, epilogue=ops.EPI_SWIGLU_BF16, groups=groups)
        ops.gemm(act[:T], L["wdown"], x, epilogue=ops.EPI_RESID_F32, groups=groups, scale=L["ls2"], scale_groups=1,
                 flags=resid_flags)

    def _mot_buffers(self, rows_q: int, rows_kv: int):
        cfg = self.cfg
        nq, nkv, hd = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim
        qkv = self.buf.get("mot.qkv", (rows_kv, (nq + 2 * nkv) * hd), torch.bfloat16)
        attn = self.buf.get("mot.attn", (rows_q, nq * hd), torch.bfloat16)
        act = self.buf.get("mot.act", (rows_q, cfg.intermediate_size), torch.bfloat16)
        hbuf = self.buf.get("mot.h", (rows_q, cfg.hidden_size), torch.bfloat16)
        return qkv, attn, act, hbuf

    def _und_forward(self, x, position_ids, cache: KVCache, causal: bool, update: bool = True, len_dev=None,
                     kv_bound: int = 0):
        """Qwen2VLModel.forward_inference(mode='und') on top of an append-style cache, single sample
        (reference g2vlm/qwen2vl.py:1267-1337 with the und branches :570-576, 859-860, 891-893, 1322-1323):
        text prefill (causal), the ViT step (non-causal) and every decode step.  x: fp32 [T,H] on the device in
        packed order (all rows belong to the und expert); returns the `norm`-ed hidden states fp32 [T,H]."""
        cfg, dev = self.cfg, self.device
        nq, nkv, hd, H = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim, cfg.hidden_size
        T, L = x.shape[0], cache.len
        cache.reserve(L + T)
        if len_dev is not None and T != 1:
            raise ValueError("device-resident cache length is only supported for single-token steps")
        pos = self._idx("und.pos", position_ids)
        cos = self.buf.get("und.cos", (T, hd // 2), torch.float32)
        sin = self.buf.get("und.sin", (T, hd // 2), torch.float32)
        use_native = self.native and T >= 2 and len_dev is None and hd == 128
        if not use_native:
            ops.mrope_table(pos, self.inv_freq, cos, sin, cfg.mrope_section)
        qkv, attn, act, hbuf = self._mot_buffers(T, T)
        if T == 1:
            work = None                   # single-token steps use the split-K decode attention
        elif T <= ops.ATTN_ROWS_PER_ITEM:  # one work item whose key range grows with the cache
            work = self._idx("und.work", torch.tensor([[0, 0, T, 0, L + T, 0, 0, 0]], dtype=torch.int32), torch.int32)
        else:
            work = self._work([0, T], [0, L + T], "und")
        xs = self.buf.get("und.x", (T, H), torch.float32)
        ops.gather_rows(x, xs, None, T)
        y = self.buf.get("und.y", (T, H), torch.float32)   # reused workspace (valid until the next und step)
        if use_native:
            # the whole step as ONE C-ABI call (g2vlm_und_prefill): the same per-op entry points in the same order
            import ctypes
            a = ops.UndPrefillArgs()
            a.num_layers, a.hidden, a.intermediate, a.n_q_heads, a.n_kv_heads, a.head_dim = \
                cfg.num_layers, H, cfg.intermediate_size, nq, nkv, hd
            a.rms_eps, a.mrope_s0, a.mrope_s1 = cfg.rms_norm_eps, cfg.mrope_section[0], cfg.mrope_section[1]
            kvp = (ctypes.c_void_p * cfg.num_layers)(*[b.data_ptr() for b in cache.buf])
            a.layers, a.kv, a.kv_capacity, a.cache_len = self._und_layers(), kvp, cache.cap, L
            a.rows, a.causal = T, int(causal)
            a.final_norm, a.inv_freq, a.position_ids = self.norm_und.data_ptr(), self.inv_freq.data_ptr(), pos.data_ptr()
            a.work, a.n_items = work.data_ptr(), work.shape[0]
            a.x, a.y, a.h, a.qkv, a.attn, a.act = (xs.data_ptr(), y.data_ptr(), hbuf.data_ptr(), qkv.data_ptr(), attn.data_ptr(),
                                                   act.data_ptr())
            a.cos, a.sin = cos.data_ptr(), sin.data_ptr()
            ops.und_prefill(a)
            if update:
                cache.len = L + T
            return y
        w = nkv * hd
        for i, Lw in enumerate(self.layers):
            kvbuf = cache.buf[i]

            def kv_append(qkv_, kvbuf=kvbuf):
                if len_dev is None:
                    ops.kv_append(qkv_[:T, nq * hd:], kvbuf, T, static_row=L)   # append this step's K|V rows
                    return kvbuf[: L + T, :w], kvbuf[: L + T, w:]
                # graph-replayable decode step: the row index and the key count are read on the device
                ops.kv_append(qkv_[:1, nq * hd:], kvbuf, 1, len_dev=len_dev)
                return kvbuf[:kv_bound, :w], kvbuf[:kv_bound, w:]
            # all rows belong to the und expert: group 0 (geo) is empty
            self._mot_layer(Lw, xs, T, 0, qkv, attn, act, hbuf, cos, sin, work, L + T, causal, True,
                            kv_exchange=kv_append, kv_len_dev=len_dev)
        ops.rmsnorm_routed(xs, y, self.norm_geo, self.norm_und, 0, cfg.rms_norm_eps, rows=T)
        if update and len_dev is None:
            cache.len = L + T
        return y

    @_on_device
    @torch.no_grad()
    def forward_cache_update_text(self, past_key_values: NaiveCache, packed_text_ids, packed_text_position_ids,
                                  text_token_lens, packed_text_indexes, packed_key_value_indexes, key_values_lens):
        """Text prefill, mode='und', causal, on top of whatever the cache holds (reference g2vlm.py:701-733).
        Returns a KVCache (NaiveCache-compatible)."""
        cfg, dev = self.cfg, self.device
        if len(text_token_lens) != 1:
            raise NotImplementedError("single-sample path only (the reference's inference drivers use batch 1)")
        cache = KVCache.adopt(past_key_values, cfg, dev)
        # (a device-resident key_values_lens is not validated: reading it would stall the host on the stream)
        if not key_values_lens.is_cuda and int(key_values_lens.sum()) != cache.len:
            raise ValueError("key_values_lens does not match the cache length")
        n = int(packed_text_ids.numel())
        x = self.buf.get("txt.x", (n, cfg.hidden_size), torch.float32)   # reused workspace: no allocator call per step
        ops.gather_rows(self.embed, x, self._idx("txt.ids", packed_text_ids), n)
        self._und_forward(x, packed_text_position_ids, cache, causal=True)
        return cache

    def _und_layers(self):
        """HOST array of the und expert's weight pointers per layer (g2vlm_und_layer_weights), built once."""
        if not hasattr(self, "_und_layer_array"):
            cfg = self.cfg
            H, I, nq, nkv, hd = cfg.hidden_size, cfg.intermediate_size, cfg.num_heads, cfg.num_kv_heads, cfg.head_dim
            qkv_w = (nq + 2 * nkv) * hd
            arr = (ops.UndLayerWeights * cfg.num_layers)()
            for i, L in enumerate(self.layers):
                w = arr[i]
                # expert 1 (und) is the second half of every stacked [geo | und] tensor
                w.wqkv = L["wqkv"].data_ptr() + qkv_w * H * 2
                w.bqkv = L["bqkv"].data_ptr() + qkv_w * 4
                w.wo = L["wo"].data_ptr() + H * nq * hd * 2
                w.wgu = L["wgu"].data_ptr() + 2 * I * H * 2
                w.wdown = L["wdown"].data_ptr() + H * I * 2
                w.input_norm = L["input_layernorm_und"].data_ptr()
                w.post_norm = L["post_attention_layernorm_und"].data_ptr()
                w.q_norm = L["q_norm_und"].data_ptr()
                w.k_norm = L["k_norm_und"].data_ptr()
            self._und_layer_array = arr
        return self._und_layer_array

    def _decode_args(self, cache: "KVCache", bound: int, cur, pos, len_dev, fused: bool = True):
        """Plain-C argument block of g2vlm_und_decode_step for this generation (und-expert weight pointers).
        fused: run the step as one persistent kernel (csrc/decode_fused.cu) instead of ~280 launches."""
        import ctypes
        cfg = self.cfg
        H, I, nq, nkv, hd = cfg.hidden_size, cfg.intermediate_size, cfg.num_heads, cfg.num_kv_heads, cfg.head_dim
        qkv_w = (nq + 2 * nkv) * hd
        self._und_layers()
        V = self.lm_head.shape[0]
        kv = (ctypes.c_void_p * cfg.num_layers)(*[b.data_ptr() for b in cache.buf])
        g = self.buf.get
        bufs = dict(x=g("dec.x", (H,), torch.float32), h=g("dec.h", (max(H, I),), torch.bfloat16),
                    qkv=g("dec.qkv", (qkv_w,), torch.bfloat16), attn=g("dec.attn", (nq * hd,), torch.bfloat16),
                    act=g("dec.act", (I,), torch.bfloat16), y=g("dec.y", (H,), torch.float32),
                    cos_sin=g("dec.cs", (hd,), torch.float32),
                    attn_ws=g("dec.ws", (ops.attention_decode_workspace_floats(bound, nq),), torch.float32),
                    logits=g("dec.logits", ((V + 7) // 8 * 8,), torch.bfloat16))
        a = ops.DecodeStepArgs()
        a.num_layers, a.hidden, a.intermediate, a.n_q_heads, a.n_kv_heads, a.head_dim = cfg.num_layers, H, I, nq, nkv, hd
        a.vocab, a.rms_eps, a.mrope_s0, a.mrope_s1 = V, cfg.rms_norm_eps, cfg.mrope_section[0], cfg.mrope_section[1]
        a.layers, a.kv, a.kv_capacity, a.kv_bound = self._und_layer_array, kv, cache.cap, bound
        a.embed, a.final_norm, a.lm_head = self.embed.data_ptr(), self.norm_und.data_ptr(), self.lm_head.data_ptr()
        a.inv_freq, a.cur_token, a.position, a.cache_len = self.inv_freq.data_ptr(), cur.data_ptr(), pos.data_ptr(), len_dev.data_ptr()
        for k, t in bufs.items():
            setattr(a, k, t.data_ptr())
        a.attn_ws_floats = bufs["attn_ws"].numel()
        if fused:
            # zeroed once (it carries the barrier counter from step to step), one generation at a time
            ws = g("dec.fused_ws", (ops.und_decode_workspace_bytes(nq, nkv),), torch.uint8, zero=True)
            a.fused_ws, a.fused_ws_bytes = ws.data_ptr(), ws.numel()
        return a, (kv, bufs)

    @_on_device
    @torch.no_grad()
    def generate_text(self, past_key_values, packed_key_value_indexes, key_values_lens, packed_start_tokens,
                      packed_query_position_ids, max_length: int, do_sample: bool = False, temperature: float = 1.0,
                      end_token_id: Optional[int] = None, return_logits: bool = False, use_cuda_graph: bool = False,
                      fused_step: Optional[bool] = None):
        """Decode loop of the reference (g2vlm.py:1070-1141), batch 1: returns the generated ids [steps, 1]
        INCLUDING the start token, like the reference.  The KV cache is appended in place.  Every step is the native
        one-call step (g2vlm_und_decode_step: one persistent kernel, or ~280 launches with fused_step=False);
        do_sample=True keeps the reference's sampling tail — softmax(logits / temperature) in fp32 and
        torch.multinomial (g2vlm.py:1122-1124) — as torch ops on the step's logits (the step then leaves the token to
        the sampler; it draws from torch's CUDA generator, so a seed reproduces a run; not bit-comparable to a CPU
        reference).  return_logits: the fp32 copy of every step's bf16 logits."""
        cfg, dev = self.cfg, self.device
        if do_sample and not temperature > 0:
            raise ValueError("temperature must be positive")
        if self.lm_head is None:
            raise RuntimeError("generate_text needs language_model.lm_head.weight in the state_dict")
        cache = KVCache.adopt(past_key_values, cfg, dev)
        L0 = cache.len
        cache.reserve(L0 + max_length)           # no re-allocation while the step graph holds buffer pointers
        bound = L0 + max_length
        cur = packed_start_tokens.to(dev, torch.long).reshape(1).clone()
        pos = packed_query_position_ids.to(dev, torch.long).reshape(3, 1).clone()
        len_dev = torch.tensor([L0], dtype=torch.int32, device=dev)
        H, V = cfg.hidden_size, self.lm_head.shape[0]
        tokens = torch.empty(max_length, dtype=torch.long, device=dev)

        # native driver: the whole step is enqueued by ONE C-ABI call — one persistent kernel (fused_step, default) or ~280
        # launches; with do_sample the call leaves the token alone (keep_token) and the reference's sampling tail follows
        if fused_step is None:
            fused_step = os.environ.get("G2VLM_DECODE_FUSED", "1") != "0"
        fused_active = bool(fused_step)
        args, keep = self._decode_args(cache, bound, cur, pos, len_dev, fused=fused_active)
        args.keep_token = int(do_sample)
        logits = keep[1]["logits"].view(1, -1)

        def step():
            # everything a step depends on (token, position, cache length) lives on the device, so the SAME call serves
            # every step (row f1: no per-token host work)
            ops.und_decode_step(args)
            if do_sample:
                probs = torch.softmax((logits[:, :V] / temperature).float(), dim=-1)   # bf16 divide, fp32 softmax
                cur.copy_(torch.multinomial(probs, num_samples=1).squeeze(1))
        graph = None
        n_done, all_logits = 0, []
        check_every = 8
        while n_done < max_length:
            tokens[n_done].copy_(cur[0])
            # (the one-kernel step is a single launch already: a graph around one cooperative kernel only adds replay cost)
            if use_cuda_graph and n_done == 1 and graph is None and not return_logits and not do_sample and not fused_active:
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                # capture WITHOUT executing: the captured launches read token / position / length from device
                # memory, so state is only advanced by replays
                with torch.cuda.graph(graph):
                    step()
            if graph is not None:
                graph.replay()
            else:
                step()
            if return_logits:
                all_logits.append(logits[0, :V].float().clone())
            n_done += 1
            if end_token_id is not None and (n_done % check_every == 0 or n_done == max_length):
                # `cur` after step i is token i+1; the reference stops when the NEW token is eos (g2vlm.py:1137)
                produced = torch.cat([tokens[1:n_done], cur]).tolist()
                if int(end_token_id) in produced:
                    n_done = produced.index(int(end_token_id)) + 1
                    break
        cache.len = L0 + n_done
        out = [tokens[i:i + 1] for i in range(n_done)]
        ids = torch.stack(out, dim=0)
        return (ids, all_logits) if return_logits else ids

    @_on_device
    @torch.no_grad()
    def language_model_forward_train(self, packed_sequence, sample_lens, split_lens, attn_modes, packed_position_ids,
                                     packed_und_token_indexes, packed_geo_token_indexes):
        """FORWARD of Qwen2VLModel.forward_train (reference g2vlm/qwen2vl.py:1200-1266, layers :783-840 / :445-553)
        for a packed batch of samples, with the numerics of the bf16 module the training path requires (bf16 residual
        stream, bf16 arithmetic in RMSNorm / rotary; pass bf16-representable weights, as FSDP mixed precision holds).
        split_lens / attn_modes: one list per sample ('causal' text splits, 'full' image splits — data_utils.py:10-37:
        a token sees every earlier split of its sample and its own split causally / fully); 'noise' splits are not
        supported.  Each split is one attention segment with its own causal flag; the rows keep the expert-permuted
        order for the grouped GEMMs and are put back in packed order around the attention, whose masks depend on
        position.  Returns the routed final norm, fp32 [T, H] (bf16 values).  No backward: parity of the forward only."""
        cfg, dev = self.cfg, self.device
        nq, nkv, hd, H = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim, cfg.hidden_size
        T = int(sum(sample_lens))
        geo_i, und_i = packed_geo_token_indexes.long().cpu(), packed_und_token_indexes.long().cpu()
        n_geo = int(geo_i.numel())
        perm = self._idx("train.perm", torch.cat([geo_i, und_i]))
        if int(perm.numel()) != T or packed_sequence.shape[0] != T:
            raise ValueError("geo + und indexes must cover every packed row exactly once")
        items, off = [], 0
        for n, lens, modes in zip(sample_lens, split_lens, attn_modes):
            if sum(lens) != n:
                raise ValueError("split_lens must add up to sample_lens")
            c = off
            for ln, mode in zip(lens, modes):
                if mode not in ("causal", "full"):
                    raise NotImplementedError(f"attention mode {mode!r} is not supported")
                items += [[t0, c, c + ln, off, c + ln, int(mode == "causal"), 0, 0]
                          for t0 in range(c, c + ln, ops.ATTN_ROWS_PER_ITEM)]
                c += ln
            off += n
        work = self._idx("train.work", torch.tensor(items, dtype=torch.int32).reshape(-1, 8), torch.int32)
        x = self.buf.get("train.x", (T, H), torch.float32)
        ops.gather_rows(packed_sequence.to(dev, torch.float32).contiguous(), x, perm, T)
        cos_p = self.buf.get("train.cos_p", (T, hd // 2), torch.float32)
        sin_p = self.buf.get("train.sin_p", (T, hd // 2), torch.float32)
        ops.mrope_table(self._idx("train.pos", packed_position_ids.contiguous()), self.inv_freq, cos_p, sin_p,
                        cfg.mrope_section)
        cos = self.buf.get("train.cos", (T, hd // 2), torch.float32)
        sin = self.buf.get("train.sin", (T, hd // 2), torch.float32)
        ops.gather_rows(cos_p, cos, perm, T)
        ops.gather_rows(sin_p, sin, perm, T)
        qkv, attn, act, hbuf = self._mot_buffers(T, T)
        train = dict(perm=perm, qkv_packed=self.buf.get("train.qkv_p", tuple(qkv.shape), torch.bfloat16),
                     attn_packed=self.buf.get("train.attn_p", tuple(attn.shape), torch.bfloat16))
        for L in self.layers:
            self._mot_layer(L, x, T, n_geo, qkv, attn, act, hbuf, cos, sin, work, T, False, False, train=train)
        yb = self.buf.get("train.yb", (T, H), torch.bfloat16)
        ops.rmsnorm_routed(x, yb, self.norm_geo, self.norm_und, n_geo, cfg.rms_norm_eps, rows=T, round_normed=True)
        y = torch.empty(T, H, dtype=torch.bfloat16, device=dev)
        ops.gather_rows(yb, y, perm, T, scatter=True)
        return y.float()

    @_on_device
    @torch.no_grad()
    def language_model_forward_geo(self, packed_sequence, packed_position_ids, packed_geo_token_indexes,
                                   packed_text_indexes, past_key_values: NaiveCache,
                                   update_past_key_values: bool = True, collect: Optional[list] = None,
                                   group=None, prompt: Optional[dict] = None, rank_rows: Optional[Sequence[int]] = None):
        """Qwen2VLModel.forward_inference(mode='geo', is_causal=False) incl. the routed final norm
        (reference g2vlm/qwen2vl.py:1267-1337).  packed_sequence: fp32 [T, H] in packed order.

        prompt (fused recon path): dict(packed_text_ids, packed_text_position_ids) of the text prefill the
        reference runs as a separate und pass before this step (g2vlm.py:1262-1272).  Layer i of that prefill only
        needs its own layer i-1, and this step only needs the prefill's layer-i K/V, so the Kp prompt rows ride
        along as extra und rows [T, T+Kp): their K/V land exactly where the merged prefix is expected (key rows
        [T, T+Kp)), their attention is one extra causal work item over those keys, and the und-expert weights are
        streamed once per layer instead of twice.  Per-row arithmetic is unchanged (same rounding points); only
        the fp32 accumulation order of the prompt rows' Linear layers differs (tensor-core tiles instead of the
        skinny-row kernel).  Requires an empty cache and update_past_key_values=False."""
        cfg, dev = self.cfg, self.device
        nq, nkv, hd, H = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim, cfg.hidden_size
        T = packed_sequence.shape[0]
        n_geo = int(packed_geo_token_indexes.numel())
        if packed_geo_token_indexes.is_cuda or packed_text_indexes.is_cuda:
            perm = torch.cat([packed_geo_token_indexes.to(dev, torch.long), packed_text_indexes.to(dev, torch.long)])
        else:  # internal row i <- packed row perm[i]
            perm = self._idx("mot.perm", torch.cat([packed_geo_token_indexes.long(), packed_text_indexes.long()]))
        if int(perm.numel()) != T:
            raise ValueError("geo + text indexes must cover every packed row exactly once")
        past_key_values = KVCache.adopt(past_key_values, cfg, dev)
        K0 = past_key_values.len
        Kp = 0
        if prompt is not None:
            if K0 or update_past_key_values or group is not None:
                raise ValueError("the fused prompt needs an empty cache, no cache update and no view sharding")
            Kp = int(prompt["packed_text_ids"].numel())
        x = self.buf.get("mot.x", (T + Kp, H), torch.float32)
        ops.gather_rows(packed_sequence, x, perm, T)
        cos_p = self.buf.get("mot.cos_p", (T, hd // 2), torch.float32)
        sin_p = self.buf.get("mot.sin_p", (T, hd // 2), torch.float32)
        ops.mrope_table(self._idx("mot.pos", packed_position_ids.contiguous()), self.inv_freq, cos_p, sin_p,
                        cfg.mrope_section)
        cos = self.buf.get("mot.cos", (T + Kp, hd // 2), torch.float32)
        sin = self.buf.get("mot.sin", (T + Kp, hd // 2), torch.float32)
        ops.gather_rows(cos_p, cos, perm, T)
        ops.gather_rows(sin_p, sin, perm, T)
        if Kp:
            ops.gather_rows(self.embed, x[T:], self._idx("mot.prompt_ids", prompt["packed_text_ids"]), Kp)
            ops.mrope_table(self._idx("mot.prompt_pos", prompt["packed_text_position_ids"].contiguous()), self.inv_freq,
                            cos[T:], sin[T:], cfg.mrope_section)
        qkv, attn, act, hbuf = self._mot_buffers(T + Kp, T + K0 + Kp)
        kvw = 2 * nkv * hd
        if update_past_key_values:
            past_key_values.reserve(K0 + T)
        kv_exchange = None
        attn_override = None
        if group is not None:
            # view-sharded sequence parallelism: this rank holds T of the scene's rows (rank_rows: every rank's count).
            import torch.distributed as dist
            world, me = dist.get_world_size(group), dist.get_rank(group)
            rank_rows = [T] * world if rank_rows is None else [int(r) for r in rank_rows]
            if rank_rows[me] != T:
                raise ValueError("rank_rows does not match this rank's packed rows")
            T_all = sum(rank_rows)
            if update_past_key_values:
                raise NotImplementedError("the merged NaiveCache is not materialised in view-sharded mode")
            kv_send = self.buf.get("mot.kv_send", (T, kvw), torch.bfloat16)
            mode = self.sp_mode
            if mode == "allgather" and len(set(rank_rows)) != 1:
                mode = "overlap"                                   # the all-gather needs equal shards
            sym = self._sp_symmetric(group, max(rank_rows), kvw) if mode == "peer" else None
            if mode == "peer" and sym is None:
                mode = "overlap"                                   # no symmetric memory on this box: NCCL point-to-point
            if mode == "allgather":
                # v1: one blocking all-gather of the rank's K|V rows per layer on the compute stream;
                # keys = [rank 0 rows | ... | rank R-1 rows | K0 prefix rows]
                kv_all = self.buf.get("mot.kv_all", (T_all + K0, kvw), torch.bfloat16)
                work = self._work([0, T], [0, T_all + K0], "geo_sp")

                def kv_exchange(qkv_):
                    ops.gather_rows(qkv_[:T, nq * hd:], kv_send, None, T)
                    self._sp_mark(0)
                    dist.all_gather_into_tensor(kv_all[:T_all], kv_send, group=group)
                    self._sp_mark(1)
                    return kv_all[:, : nkv * hd], kv_all[:, nkv * hd:]
            elif mode == "peer":
                # v3: each rank writes its K|V rows into a SYMMETRIC buffer (peer-mapped over NVLink on every rank); after
                # a device-side barrier on a side stream the COPY ENGINES pull the other ranks' rows straight into
                # kv_remote — the attention over the local keys that runs meanwhile gives up 2 SMs for the barrier kernel only
                # (the NCCL exchange of v2 needs 16-32 SMs at 8 ranks and still leaves 0.2 ms per layer exposed).
                # Two send buffers alternate per layer: rank Y overwrites buffer b at layer i only after its remote
                # attention of layer i-1, which waited for Y's pulls of layer i-1, which were queued behind barrier i-1,
                # which every rank reaches only after ITS pulls of layer i-2 (the last readers of buffer b) completed.
                sym_t, sym_h, side = sym
                T_max = sym_t.shape[1]
                others = [j for j in range(world) if j != me]
                offs, o = {}, 0
                for j in others:
                    offs[j] = o
                    o += rank_rows[j]
                T_rem = o
                kv_remote = self.buf.get("mot.kv_remote", (max(T_rem, 1), kvw), torch.bfloat16)
                attn_b = self.buf.get("mot.attn_b", (T, nq * hd), torch.bfloat16)
                lse_a = self.buf.get("mot.lse_a", (T, nq), torch.float32)
                lse_b = self.buf.get("mot.lse_b", (T, nq), torch.float32)
                work = None
                work_local = self._work([0, T], [0, T + K0], "geo")
                work_remote = self._work([0, T], [0, T_rem], "geo_rem")
                peer_bufs = {(j, b): sym_h.get_buffer(j, (rank_rows[j], kvw), torch.bfloat16, b * T_max * kvw)
                             for j in others for b in (0, 1)}
                scale = 1.0 / math.sqrt(hd)
                state = dict(layer=0)

                def attn_override(qkv_, attn_):
                    b = state["layer"] & 1
                    state["layer"] += 1
                    ops.gather_rows(qkv_[:T, nq * hd:], sym_t[b], None, T)         # my rows -> symmetric send buffer b
                    trace = self.sp_trace is not None
                    written = torch.cuda.Event(enable_timing=trace)
                    written.record()
                    done = torch.cuda.Event(enable_timing=trace)
                    with torch.cuda.stream(side):
                        side.wait_event(written)
                        sym_h.barrier(channel=0)                                    # every rank has written layer i
                        if trace:
                            bar = torch.cuda.Event(enable_timing=True)
                            bar.record()
                        for j in others:                                            # copy-engine pulls over NVLink
                            kv_remote[offs[j]:offs[j] + rank_rows[j]].copy_(peer_bufs[(j, b)], non_blocking=True)
                        done.record()
                    # two SMs stay free: the barrier is a one-CTA kernel and the persistent attention would otherwise
                    # hold every SM until it ends (measured: 0.125 ms per layer exposed with the full grid)
                    ops.attention(qkv_[:T, : nq * hd], qkv_[:T + K0, nq * hd:(nq + nkv) * hd],
                                  qkv_[:T + K0, (nq + nkv) * hd:], attn_, work_local, num_q_heads=nq, num_kv_heads=nkv,
                                  head_dim=hd, scale=scale, lse=lse_a, max_ctas=max(1, ops.num_sms() - self.sp_peer_margin))
                    if trace:
                        local_end = torch.cuda.Event(enable_timing=True)
                        local_end.record()
                        self.sp_trace.append(dict(written=written, barrier=bar, done=done, local_end=local_end))
                    self._sp_mark(0)
                    torch.cuda.current_stream().wait_event(done)
                    self._sp_mark(1)
                    if T_rem:
                        ops.attention(qkv_[:T, : nq * hd], kv_remote[:T_rem, : nkv * hd], kv_remote[:T_rem, nkv * hd:],
                                      attn_b, work_remote, num_q_heads=nq, num_kv_heads=nkv, head_dim=hd, scale=scale,
                                      lse=lse_b)
                        ops.attention_merge(attn_, lse_a, attn_b, lse_b, attn_, nq, hd, rows=T)
            else:
                # v2: the remote K|V rows travel (one NCCL group of point-to-point sends/receives per layer, straight
                # into their slot of kv_remote) WHILE the tensor cores run attention over the keys that are already
                # here — this rank's own rows and the replicated prefix; then attention over the remote keys and an
                # exact log-sum-exp merge of the two partials (the combination flash-attn does across key blocks).
                others = [j for j in range(world) if j != me]
                offs, o = {}, 0
                for j in others:
                    offs[j] = o
                    o += rank_rows[j]
                T_rem = o
                kv_remote = self.buf.get("mot.kv_remote", (max(T_rem, 1), kvw), torch.bfloat16)
                attn_b = self.buf.get("mot.attn_b", (T, nq * hd), torch.bfloat16)
                lse_a = self.buf.get("mot.lse_a", (T, nq), torch.float32)
                lse_b = self.buf.get("mot.lse_b", (T, nq), torch.float32)
                work = None
                work_local = self._work([0, T], [0, T + K0], "geo")       # own rows + prefix rows [T, T+K0) of qkv
                work_remote = self._work([0, T], [0, T_rem], "geo_rem")
                peers = {j: dist.get_global_rank(group, j) for j in others}
                local_ctas = max(1, ops.num_sms() - self.sp_sm_margin)
                scale = 1.0 / math.sqrt(hd)

                def attn_override(qkv_, attn_):
                    ops.gather_rows(qkv_[:T, nq * hd:], kv_send, None, T)          # contiguous send buffer
                    reqs = []
                    for j in others:
                        reqs.append(dist.P2POp(dist.isend, kv_send, peers[j], group))
                        reqs.append(dist.P2POp(dist.irecv, kv_remote[offs[j]:offs[j] + rank_rows[j]], peers[j], group))
                    works = dist.batch_isend_irecv(reqs) if reqs else []
                    ops.attention(qkv_[:T, : nq * hd], qkv_[:T + K0, nq * hd:(nq + nkv) * hd],
                                  qkv_[:T + K0, (nq + nkv) * hd:], attn_, work_local, num_q_heads=nq, num_kv_heads=nkv,
                                  head_dim=hd, scale=scale, lse=lse_a, max_ctas=local_ctas)
                    self._sp_mark(0)
                    for w_ in works:
                        w_.wait()          # stream-side wait only: the host keeps enqueueing
                    self._sp_mark(1)
                    if T_rem:
                        ops.attention(qkv_[:T, : nq * hd], kv_remote[:T_rem, : nkv * hd], kv_remote[:T_rem, nkv * hd:],
                                      attn_b, work_remote, num_q_heads=nq, num_kv_heads=nkv, head_dim=hd, scale=scale,
                                      lse=lse_b)
                        ops.attention_merge(attn_, lse_a, attn_b, lse_b, attn_, nq, hd, rows=T)
        elif Kp:
            key = ("geo_fused", T, Kp)
            work = self._work_cache.get(key)
            if work is None:   # the geo rows see every key incl. the prompt's; the prompt rows are causal among themselves
                items = ops.attention_work_table([0, T], [0, T + Kp]).tolist()
                items += [[t0, T, T + Kp, T, T + Kp, 1, 0, 0] for t0 in range(T, T + Kp, ops.ATTN_ROWS_PER_ITEM)]
                work = self._work_cache[key] = torch.tensor(items, dtype=torch.int32).to(dev)
        else:
            work = self._work([0, T], [0, T + K0], "geo")
        for i, L in enumerate(self.layers):
            if K0:
                # cached K|V rows (text prefill) become key rows [T, T+K0) (KV merge, qwen2vl.py:621-638)
                if kv_exchange is None:
                    ops.gather_rows(past_key_values.buf[i], qkv[T:, nq * hd:], None, K0)
                else:
                    ops.gather_rows(past_key_values.buf[i], kv_all[T_all:], None, K0)
            self._mot_layer(L, x, T + Kp, n_geo, qkv, attn, act, hbuf, cos, sin, work, T + K0 + Kp, False, False,
                            kv_exchange=kv_exchange, prompt_rows=Kp, attn_override=attn_override)
            if update_past_key_values:
                # append in the reference's merged order: cached rows, then the PACKED query rows
                ops.gather_rows(qkv[:T, nq * hd:], past_key_values.buf[i][K0:], perm, T, scatter=True)
            if collect is not None:
                y = torch.empty(T, H, dtype=torch.float32, device=dev)
                ops.gather_rows(x, y, perm, T, scatter=True)
                collect.append(y)
        if update_past_key_values:
            past_key_values.len = K0 + T
        y_int = self.buf.get("mot.y", (T, H), torch.float32)
        ops.rmsnorm_routed(x, y_int, self.norm_geo, self.norm_und, n_geo, cfg.rms_norm_eps, rows=T)
        last = torch.empty(T, H, dtype=torch.float32, device=dev)
        ops.gather_rows(y_int, last, perm, T, scatter=True)
        return last, past_key_values

    # ------------------------------------------------------------------------------------------
    # DINO encoder
    # ------------------------------------------------------------------------------------------
    def _dino_embeddings(self, img, tag="dino"):
        """Dinov2WithRegistersEmbeddings.forward: images (n,3,H,W) fp32 on device -> x fp32 [n*S, D].
        If self._raw_images is set the images are raw [0,1] views and the ImageNet normalisation is fused
        into the im2col kernel."""
        cfg = self.cfg
        n, _, Hh, Ww = img.shape
        p = cfg.dino_patch
        gh, gw = Hh // p, Ww // p
        P, D = gh * gw, cfg.dino_hidden
        S = P + 1 + cfg.dino_registers
        patches = self.buf.get(tag + ".patches", (n * P, self.dino_kpad), torch.bfloat16)
        if self._raw_images:
            ops.im2col_patches(img, patches, p, host_prep.RESNET_MEAN, host_prep.RESNET_STD)
        else:
            ops.im2col_patches(img, patches, p)
        emb = self.buf.get(tag + ".emb", (n * P, D), torch.bfloat16)
        ops.gemm(patches, self.dino_wpatch, emb, epilogue=ops.EPI_STORE_BF16, bias=self.dino_bpatch)
        x = self.buf.get(tag + ".x", (n * S, D), torch.float32)
        ops.dino_embed(emb, self.dino_cls, self.dino_reg, self._dino_pos(gh, gw, Hh == Ww), x, n, P, cfg.dino_registers)
        return x

    def _dino_layers(self, x, rows, cu, collect=None):
        """The 24 encoder layers in place on x[:rows] (fp32); attention segments `cu` (row offsets)."""
        cfg = self.cfg
        D, nh, hp = cfg.dino_hidden, cfg.dino_heads, self.dino_hp
        work = self._work(cu, cu, "dino")
        h = self.buf.get("dino.h", (rows, D), torch.bfloat16)
        qkv = self.buf.get("dino.qkv", (rows, 3 * nh * hp), torch.bfloat16)
        # rows covered by no segment are never written by the attention kernel: they stay ZERO
        # (definition of the uninitialised flash-attn rows, quirk Q1)
        attn = self.buf.get("dino.attn", (rows, nh * hp), torch.bfloat16, zero=True)
        if self._dino_attn_cover.get(attn.data_ptr()) not in (None, cu[-1]) :
            attn.zero_()     # same buffer, different coverage than the previous call: stale rows must not leak in
        self._dino_attn_cover[attn.data_ptr()] = cu[-1]
        mid = self.buf.get("dino.mid", (rows, D * cfg.dino_mlp_ratio), torch.bfloat16)
        scale = 1.0 / math.sqrt(cfg.dino_head_dim)
        x = x[:rows]
        for L in self.dino_layers:
            ops.layernorm(x, h, L["norm1w"], L["norm1b"], cfg.dino_ln_eps)
            ops.gemm(h, L["wqkv"], qkv, epilogue=ops.EPI_STORE_BF16, bias=L["bqkv"])
            ops.attention(qkv[:, : nh * hp], qkv[:, nh * hp: 2 * nh * hp], qkv[:, 2 * nh * hp:], attn, work,
                          num_q_heads=nh, num_kv_heads=nh, head_dim=hp, scale=scale)
            ops.gemm(attn, L["wdense"], x, epilogue=ops.EPI_RESID_F32, bias=L["bdense"], scale=L["ls1"], scale_groups=1)
            ops.layernorm(x, h, L["norm2w"], L["norm2b"], cfg.dino_ln_eps)
            ops.gemm(h, L["wfc1"], mid, epilogue=ops.EPI_STORE_BF16, bias=L["bfc1"], flags=ops.GEMM_GELU)
            ops.gemm(mid, L["wfc2"], x, epilogue=ops.EPI_RESID_F32, bias=L["bfc2"], scale=L["ls2"], scale_groups=1)
            if collect is not None:
                collect.append(x.clone())

    @_on_device
    @torch.no_grad()
    def dino_forward(self, packed_pixel_values, dino_token_seqlens, collect: Optional[list] = None):
        """Dinov2WithRegistersModel.forward with the caller's cu_seqlens (reference
        g2vlm/dinov2_model.py:301-356, g2vlm.py:988-992) -> bf16 tokens [N*P, D] (post final LN,
        cls/registers dropped)."""
        cfg, dev = self.cfg, self.device
        img = packed_pixel_values.to(dev, torch.float32).contiguous()
        N, _, Hh, Ww = img.shape
        P = (Hh // cfg.dino_patch) * (Ww // cfg.dino_patch)
        S, D = P + 1 + cfg.dino_registers, cfg.dino_hidden
        rows = N * S
        x = self._dino_embeddings(img)
        cu = [0]
        for n in dino_token_seqlens.tolist():
            cu.append(cu[-1] + int(n))
        if cu[-1] > rows:
            raise ValueError("dino_token_seqlens exceed the number of DINO rows")
        self._dino_layers(x, rows, cu, collect)
        tokens = self.buf.get("dino.tokens", (N * P, D), torch.bfloat16)
        ops.layernorm(x, tokens, self.dino_lnw, self.dino_lnb, cfg.dino_ln_eps, seg_in=S, seg_skip=1 + cfg.dino_registers)
        return tokens

    @_on_device
    @torch.no_grad()
    def dino_forward_sharded(self, packed_pixel_values_all, shard, group):
        """View-sharded DINO (sharding.ViewShard): this rank runs the encoder on the flattened rows of ITS
        attention segments (quirk Q1 makes segments straddle images, so the shard is by segment, every
        other op being row-local), then one neighbour exchange re-associates rows with images.
        packed_pixel_values_all: normalised images of the WHOLE scene (host or device); only the images
        intersecting this rank's rows are embedded. Returns bf16 tokens [n_local*P, D] of the owned views."""
        import torch.distributed as dist
        cfg, dev = self.cfg, self.device
        D, S, P = cfg.dino_hidden, shard.S, shard.P
        ia, ib = shard.dino_images
        img = packed_pixel_values_all[ia:ib].to(dev, torch.float32).contiguous()
        x_img = self._dino_embeddings(img, tag="dino_sp")
        g0, g1 = shard.dino_rows
        rows = g1 - g0
        x = self.buf.get("dino_sp.xrows", (rows, D), torch.float32)
        ops.gather_rows(x_img[g0 - ia * S:], x, None, rows)
        c0, c1 = shard.dino_covered_rows
        cu = [i * P for i in range((c1 - c0) // P + 1)]  # local segment offsets (c0 == g0)
        self._dino_layers(x, rows, cu)
        r0, r1 = shard.recv_from_next
        s0, s1 = shard.send_to_prev
        normed = self.buf.get("dino_sp.normed", (rows + (r1 - r0), D), torch.bfloat16)
        ops.layernorm(x, normed, self.dino_lnw, self.dino_lnb, cfg.dino_ln_eps, rows=rows)
        reqs = []
        if s1 > s0:
            send = normed[s0 - g0:s1 - g0]
            reqs.append(dist.P2POp(dist.isend, send, dist.get_global_rank(group, shard.rank - 1), group))
        if r1 > r0:
            recv = normed[rows:rows + (r1 - r0)]
            reqs.append(dist.P2POp(dist.irecv, recv, dist.get_global_rank(group, shard.rank + 1), group))
        if reqs:
            for w in dist.batch_isend_irecv(reqs):
                w.wait()
        key = ("dino_sp.idx", shard)
        idx = self._work_cache.get(key)
        if idx is None:   # built once per shard geometry, not per call
            idx = self._work_cache[key] = (torch.tensor(shard.token_row_index(), dtype=torch.long) - g0).to(dev)
        tokens = self.buf.get("dino_sp.tokens", (shard.n_local * P, D), torch.bfloat16)
        ops.gather_rows(normed, tokens, idx, shard.n_local * P)
        return tokens

    @_on_device
    @torch.no_grad()
    def forward_cache_update_dino(self, past_key_values: NaiveCache, packed_text_ids, packed_text_indexes,
                                  packed_dino_token_indexes, dino_token_seqlens, packed_position_ids, packed_seqlens,
                                  packed_indexes, packed_key_value_indexes, key_values_lens, packed_dino_images,
                                  original_images, update_past_key_values: bool = True, collect: Optional[dict] = None,
                                  shard=None, group=None, prompt: Optional[dict] = None):
        """Reference: g2vlm.py:968-1039.  Returns (past_key_values, last_hidden_state [T, H] fp32).
        `prompt`: generation input of the text prefill to fuse into this step instead of running it as a separate
        pass first (see language_model_forward_geo); the index tensors are the ones built on top of that prefill.
        With `shard` (sharding.ViewShard) the index tensors describe the WHOLE scene and this rank processes
        the packed rows of its views only (last_hidden_state then has the local rows)."""
        cfg, dev = self.cfg, self.device
        T, H = int(sum(packed_seqlens.tolist())), cfg.hidden_size
        if packed_dino_images.shape[0] < 1:
            raise ValueError("at least one view is required")
        if (self.native and prompt is not None and collect is None and shard is None and group is None
                and not update_past_key_values and KVCache.adopt(past_key_values, cfg, dev).len == 0):
            last = self._native_dino_mot(packed_text_ids, packed_text_indexes, packed_dino_token_indexes, dino_token_seqlens,
                                         packed_position_ids, packed_dino_images, prompt)
            return past_key_values, last
        self._mark("dino_begin")
        if shard is None:
            tokens = self.dino_forward(packed_dino_images, dino_token_seqlens,
                                       collect=None if collect is None else collect.setdefault("dino_layers", []))
        else:
            tokens = self.dino_forward_sharded(packed_dino_images, shard, group)
            # restrict the scene-wide index tensors to the packed rows [p0, p1) of the owned views
            p0, p1 = shard.packed_rows
            T = p1 - p0
            tsel = (packed_text_indexes >= p0) & (packed_text_indexes < p1)
            gsel = (packed_dino_token_indexes >= p0) & (packed_dino_token_indexes < p1)
            packed_text_ids = packed_text_ids[tsel]
            packed_text_indexes = packed_text_indexes[tsel] - p0
            packed_dino_token_indexes = packed_dino_token_indexes[gsel] - p0
            packed_position_ids = packed_position_ids[:, p0:p1]
        self._mark("dino_end")
        n_geo = tokens.shape[0]
        geo_emb = self.buf.get("mot.geo_emb", (n_geo, H), torch.float32)
        ops.gemm(tokens, self.w_dino2llm, geo_emb, epilogue=ops.EPI_STORE_F32, bias=self.b_dino2llm,
                 flags=ops.GEMM_ROUND_BF16)
        packed = self.buf.get("mot.packed", (T, H), torch.float32)
        n_und = int(packed_text_ids.numel())
        txt = self.buf.get("mot.txt", (n_und, H), torch.float32)
        ops.gather_rows(self.embed, txt, self._idx("dino.txt_ids", packed_text_ids), n_und)
        ops.gather_rows(txt, packed, self._idx("dino.txt_idx", packed_text_indexes), n_und, scatter=True)
        ops.gather_rows(geo_emb, packed, self._idx("dino.geo_idx", packed_dino_token_indexes), n_geo, scatter=True)
        if collect is not None:
            collect["dino_tokens"] = tokens.float().clone()
            collect["packed_sequence"] = packed.clone()
        last, past_key_values = self.language_model_forward_geo(
            packed, packed_position_ids, packed_dino_token_indexes, packed_text_indexes, past_key_values,
            update_past_key_values=update_past_key_values,
            collect=None if collect is None else collect.setdefault("mot_layers", []), group=group, prompt=prompt,
            rank_rows=None if shard is None else [(b - a) * (shard.P + 2) for a, b in view_ranges(shard.n_views, shard.world)])
        self._mark("mot_end")
        return past_key_values, last

    # ------------------------------------------------------------------------------------------
    # Pi3 decoders + heads
    # ------------------------------------------------------------------------------------------
    def _decoder(self, name, hidden, N, P, gh, gw, out, out_fp32_round=False, context=None):
        """Pi3TransformerDecoder / Pi3ContextTransformerDecoder (transformer_head.py:48-56, 122-131)."""
        cfg, dec = self.cfg, self.decoders[name]
        H, dh, hp, ehd = cfg.hidden_size, cfg.dec_heads, self.dec_hp, cfg.dec_head_dim
        rows = N * P
        x = self.buf.get("dec.x", (rows, H), torch.float32)
        ops.gather_rows(hidden, x, None, rows)
        h = self.buf.get("dec.h", (rows, H), torch.bfloat16)
        cmp_ = self.dec_compact
        regroup = dict(out_col_group=ehd, out_col_stride=hp) if cmp_ else {}
        ohc = ehd if cmp_ else 0
        # zero-filled once: the pad columns of every head slot are never written afterwards
        qkv = self.buf.get("dec.qkv", (rows, 3 * dh * hp), torch.bfloat16, zero=True)
        attn = self.buf.get("dec.attn", (rows, dh * (ehd if cmp_ else hp)), torch.bfloat16)
        mid = self.buf.get("dec.mid", (rows, H * cfg.dec_mlp_ratio), torch.bfloat16)
        cos, sin = self._rope2d_tables(gh, gw)
        cu = [v * P for v in range(N + 1)]
        work = self._work(cu, cu, "dec")
        scale = 1.0 / math.sqrt(ehd)
        if dec["cross"]:
            yh = self.buf.get("dec.yh", (P, H), torch.bfloat16)
            kvc = self.buf.get("dec.kvc", (P, 2 * dh * hp), torch.bfloat16, zero=True)
            qc = self.buf.get("dec.qc", (rows, dh * hp), torch.bfloat16, zero=True)
            cwork = self._cross_work(N, P)
        for B in dec["blocks"]:
            ops.layernorm(x, h, B["norm1w"], B["norm1b"], 1e-6)
            ops.gemm(h, B["wqkv"], qkv, epilogue=ops.EPI_STORE_BF16, bias=B["bqkv"], **regroup)
            ops.rope2d(qkv, rows, 2 * dh, hp, ehd, P, gw, cos, sin)
            ops.attention(qkv[:, : dh * hp], qkv[:, dh * hp: 2 * dh * hp], qkv[:, 2 * dh * hp:], attn, work,
                          num_q_heads=dh, num_kv_heads=dh, head_dim=hp, scale=scale, out_head_cols=ohc)
            ops.gemm(attn, B["wproj"], x, epilogue=ops.EPI_RESID_F32, bias=B["bproj"])
            if dec["cross"]:
                # context = view 0's tokens for every view (g2vlm.py:1196): K/V projected ONCE per block
                ops.layernorm(context, yh, B["norm_yw"], B["norm_yb"], 1e-6, rows=P)
                ops.gemm(yh, B["wckv"], kvc, epilogue=ops.EPI_STORE_BF16, bias=B["bckv"], **regroup)
                ops.rope2d(kvc, P, dh, hp, ehd, P, gw, cos, sin)
                ops.layernorm(x, h, B["norm2w"], B["norm2b"], 1e-6)
                ops.gemm(h, B["wcq"], qc, epilogue=ops.EPI_STORE_BF16, bias=B["bcq"], **regroup)
                ops.rope2d(qc, rows, dh, hp, ehd, P, gw, cos, sin)
                ops.attention(qc, kvc[:, : dh * hp], kvc[:, dh * hp:], attn, cwork, num_q_heads=dh, num_kv_heads=dh,
                              head_dim=hp, scale=scale, out_head_cols=ohc)
                ops.gemm(attn, B["wcproj"], x, epilogue=ops.EPI_RESID_F32, bias=B["bcproj"])
                ops.layernorm(x, h, B["norm3w"], B["norm3b"], 1e-6)
            else:
                ops.layernorm(x, h, B["norm2w"], B["norm2b"], 1e-6)
            ops.gemm(h, B["wfc1"], mid, epilogue=ops.EPI_STORE_BF16, bias=B["bfc1"], flags=ops.GEMM_GELU)
            ops.gemm(mid, B["wfc2"], x, epilogue=ops.EPI_RESID_F32, bias=B["bfc2"])
        ops.cast_bf16(x, h)
        if out.dtype == torch.bfloat16:
            ops.gemm(h, dec["wout"], out, epilogue=ops.EPI_STORE_BF16, bias=dec["bout"])
        else:  # `.float()` of the bf16 linear_out result
            ops.gemm(h, dec["wout"], out, epilogue=ops.EPI_STORE_F32, bias=dec["bout"], flags=ops.GEMM_ROUND_BF16)
        return out

    def _camera_head(self, camera_hidden, N, P):
        """Pi3CameraHead.forward (camera_head.py:48-72), all fp32 (autocast disabled in the reference): split-bf16
        Linears + mean pool + fc_t / fc_rot / SVD orthogonalisation.  camera_hidden fp32 [N*P, camera_dim]."""
        cfg, dev = self.cfg, self.device
        rows = N * P
        C = cfg.camera_dim
        feat = camera_hidden
        t1 = self.buf.get("cam.t1", (rows, C), torch.float32)
        t2 = self.buf.get("cam.t2", (rows, C), torch.float32)
        f2 = self.buf.get("cam.f2", (rows, C), torch.float32)
        f3 = self.buf.get("cam.f3", (rows, C), torch.float32)
        for i, dst in ((0, f2), (1, f3)):
            self._linear_fp32(feat, self.cam[f"r{i}1w"], self.cam[f"r{i}1b"], t1, relu=True)
            self._linear_fp32(t1, self.cam[f"r{i}2w"], self.cam[f"r{i}2b"], t2, relu=True)
            self._linear_fp32(t2, self.cam[f"r{i}3w"], self.cam[f"r{i}3b"], dst, relu=True, residual=feat)
            feat = dst
        pooled = self.buf.get("cam.pooled", (N, C), torch.float32)
        ops.mean_pool(feat, pooled, N, P)
        m1 = self.buf.get("cam.m1", (N, C), torch.float32)
        m2 = self.buf.get("cam.m2", (N, C), torch.float32)
        self._linear_fp32(pooled, self.cam["m0w"], self.cam["m0b"], m1, relu=True)
        self._linear_fp32(m1, self.cam["m2w"], self.cam["m2b"], m2, relu=True)
        poses = torch.empty(N, 4, 4, dtype=torch.float32, device=dev)
        ops.camera_pose(m2, self.cam["fc_tw"], self.cam["fc_tb"], self.cam["fc_rotw"], self.cam["fc_rotb"], poses)
        return poses

    def _linear_fp32(self, x, w3, b, out, relu=False, residual=None):
        """True-fp32 nn.Linear (autocast disabled in the reference) on the bf16 tensor cores:
        x = hi + lo, w = hi + lo, x.w ~ hi.hi + hi.lo + lo.hi as ONE GEMM over the concatenated K."""
        rows, k = x.shape
        xs = self.buf.get(f"fp32lin.{rows}x{k}", (rows, 3 * k), torch.bfloat16)
        ops.split3(x, xs)
        ops.gemm(xs, w3, out, epilogue=ops.EPI_STORE_F32, bias=b, flags=ops.GEMM_RELU if relu else 0, residual=residual)
        return out

    @_on_device
    @torch.no_grad()
    def reconstruct(self, past_key_values=None, packed_key_value_indexes=None, key_values_lens=None,
                    selected_hidden_states=None, packed_dino_token_indexes=None, packed_dino_images=None,
                    original_images=None, collect: Optional[dict] = None, shard=None, group=None, **kwargs):
        """Reference: g2vlm.py:1143-1238.  With `shard`: selected_hidden_states holds this rank's packed rows
        and the outputs cover the owned views; the context of the global-points decoder (view 0's hidden,
        g2vlm.py:1196) is broadcast from rank 0."""
        cfg, dev = self.cfg, self.device
        N, _, Hh, Ww = packed_dino_images.shape
        p = cfg.dino_patch
        gh, gw = Hh // p, Ww // p
        P, H = gh * gw, cfg.hidden_size
        self._mark("heads_begin")
        if self.native and collect is None and shard is None:
            out, poses, conf = self._native_heads(selected_hidden_states, packed_dino_token_indexes, N, Hh, Ww)
            if original_images is not None and original_images.dim() == 4:
                original_images = original_images.unsqueeze(0)
            return dict(points=out["points"][None], local_points=out["local_points"][None],
                        conf=None if conf is None else conf[None], camera_poses=poses[None],
                        global_points=out["global_points"][None], images=original_images)
        geo = self._idx("recon.geo_idx", packed_dino_token_indexes)
        if shard is not None:
            p0, p1 = shard.packed_rows
            geo = geo[(geo >= p0) & (geo < p1)] - p0
            N = shard.n_local
            if original_images is not None:
                original_images = original_images[shard.v0:shard.v1]
        rows = N * P
        hidden = self.buf.get("rec.hidden", (rows, H), torch.float32)
        ops.gather_rows(selected_hidden_states, hidden, geo, rows)
        context = hidden[:P]
        if shard is not None:
            import torch.distributed as dist
            context = self.buf.get("rec.context", (P, H), torch.float32)
            if shard.rank == 0:
                ops.gather_rows(hidden, context, None, P)
            dist.broadcast(context, src=dist.get_global_rank(group, 0), group=group)

        point_hidden = self.buf.get("rec.point_hidden", (rows, cfg.point_dim), torch.bfloat16)
        self._decoder("point_decoder", hidden, N, P, gh, gw, point_hidden)
        camera_hidden = self.buf.get("rec.camera_hidden", (rows, cfg.camera_dim), torch.float32)
        self._decoder("camera_decoder", hidden, N, P, gh, gw, camera_hidden)
        global_hidden = self.buf.get("rec.global_hidden", (rows, cfg.point_dim), torch.bfloat16)
        self._decoder("global_points_decoder", hidden, N, P, gh, gw, global_hidden, context=context)
        if collect is not None:
            collect.update(point_hidden=point_hidden.float().view(N, P, -1).clone(),
                           camera_hidden=camera_hidden.view(N, P, -1).clone(),
                           global_hidden=global_hidden.float().view(N, P, -1).clone())

        poses = self._camera_head(camera_hidden, N, P)

        # point heads: fp32 Linear on bf16-exact activations = two bf16 GEMMs (W = hi + lo)
        nf = 3 * p * p
        feat_pts = self.buf.get("rec.feat_pts", (rows, nf), torch.float32)
        local_points = torch.empty(N, Hh, Ww, 3, dtype=torch.float32, device=dev)
        points = torch.empty(N, Hh, Ww, 3, dtype=torch.float32, device=dev)
        global_points = torch.empty(N, Hh, Ww, 3, dtype=torch.float32, device=dev)
        ops.gemm(point_hidden, self.point_head_whi, feat_pts, epilogue=ops.EPI_STORE_F32, bias=self.point_head_b)
        ops.gemm(point_hidden, self.point_head_wlo, feat_pts, epilogue=ops.EPI_STORE_F32, flags=ops.GEMM_ACCUMULATE)
        ops.points_epilogue(feat_pts, poses, local_points, points, N, Hh, Ww, p, 1)
        ops.gemm(global_hidden, self.global_point_head_whi, feat_pts, epilogue=ops.EPI_STORE_F32,
                 bias=self.global_point_head_b)
        ops.gemm(global_hidden, self.global_point_head_wlo, feat_pts, epilogue=ops.EPI_STORE_F32,
                 flags=ops.GEMM_ACCUMULATE)
        ops.points_epilogue(feat_pts, None, global_points, None, N, Hh, Ww, p, 0)

        self._mark("heads_end")
        conf = None
        if cfg.train_conf_pi3:
            conf_hidden = self.buf.get("rec.conf_hidden", (rows, cfg.point_dim), torch.bfloat16)
            self._decoder("conf_decoder", hidden, N, P, gh, gw, conf_hidden)
            feat_conf = self.buf.get("rec.feat_conf", (rows, p * p), torch.float32)
            ops.gemm(conf_hidden, self.conf_head_whi, feat_conf, epilogue=ops.EPI_STORE_F32, bias=self.conf_head_b)
            ops.gemm(conf_hidden, self.conf_head_wlo, feat_conf, epilogue=ops.EPI_STORE_F32, flags=ops.GEMM_ACCUMULATE)
            conf = torch.empty(N, Hh, Ww, 1, dtype=torch.float32, device=dev)
            ops.points_epilogue(feat_conf, None, conf, None, N, Hh, Ww, p, 2)
            conf = conf[None]
        if original_images is not None and original_images.dim() == 4:
            original_images = original_images.unsqueeze(0)
        return dict(points=points[None], local_points=local_points[None], conf=conf, camera_poses=poses[None],
                    global_points=global_points[None], images=original_images)

    @_on_device
    @torch.no_grad()
    def recon_view_sharded(self, tokenizer, new_token_ids, images, group=None):
        """ONE long scene split by view over the ranks of `group` (sequence parallelism; no reference
        counterpart).  Every rank passes the same `images` ((N,3,H,W) tensor or list of
        paths); returns the reference's result dict restricted to the views this rank owns
        (`view_range` gives them) with `camera_poses` of ALL views all-gathered."""
        import torch.distributed as dist

        from .sharding import shard_views
        group = group if group is not None else dist.group.WORLD
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        dev = self.device
        past = NaiveCache(self.cfg.num_layers)
        gi, newlens, new_rope = self.prepare_prompts_addbos([0], [0], ["Reconstruct the 3D scene."], tokenizer,
                                                            new_token_ids)
        past = self.forward_cache_update_text(past, **{k: v.to(dev) for k, v in gi.items()})  # replicated prefix
        gi, _, _ = self.prepare_dino_images_pi3(newlens, new_rope, images, None, new_token_ids,
                                                normalize_on_host=False)  # normalised on the device
        N, _, Hh, Ww = gi["packed_dino_images"].shape
        P = (Hh // self.cfg.dino_patch) * (Ww // self.cfg.dino_patch)
        shard = shard_views(N, P, rank, world, 1 + self.cfg.dino_registers)
        self._raw_images = True
        try:
            past, last = self.forward_cache_update_dino(past, update_past_key_values=False, shard=shard, group=group,
                                                        **gi)
        finally:
            self._raw_images = False
        pred = self.reconstruct(selected_hidden_states=last, shard=shard, group=group, **gi)
        ranges = view_ranges(N, world)
        n_max = max(b - a for a, b in ranges)
        mine = torch.zeros(n_max, 4, 4, dtype=torch.float32, device=dev)
        mine[: shard.n_local] = pred["camera_poses"][0]
        poses_all = torch.empty(world, n_max, 4, 4, dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(poses_all, mine, group=group)
        pred["camera_poses_all"] = torch.cat([poses_all[r, : b - a] for r, (a, b) in enumerate(ranges)])[None]
        pred["view_range"] = (shard.v0, shard.v1)
        return pred

    # ------------------------------------------------------------------------------------------
    # Qwen2-VL ViT + chat path (rows f1 / f2)
    # ------------------------------------------------------------------------------------------
    def _vit_rope_tables(self, grid_thw):
        """cos / sin [T, head_dim/2] of rot_pos_emb (modeling_qwen2_vl.py:1024-1048), built on the host exactly
        like the reference (positions in 2x2-merge-major patch order), cached per grid."""
        key = ("vit_rope",) + tuple(int(x) for x in grid_thw.flatten().tolist())
        t = self._rope2d_cache.get(key)
        if t is None:
            cfg = self.cfg
            m, hd = cfg.vit_merge, cfg.vit_embed_dim // cfg.vit_heads
            pos = []
            for tt, h, w in grid_thw.tolist():
                hp = torch.arange(h).unsqueeze(1).expand(-1, w).reshape(h // m, m, w // m, m).permute(0, 2, 1, 3).flatten()
                wp = torch.arange(w).unsqueeze(0).expand(h, -1).reshape(h // m, m, w // m, m).permute(0, 2, 1, 3).flatten()
                pos.append(torch.stack([hp, wp], dim=-1).repeat(tt, 1))
            pos = torch.cat(pos, dim=0)
            inv_freq = 1.0 / (10000.0 ** (torch.arange(0, hd // 2, 2, dtype=torch.float) / (hd // 2)))
            freqs = torch.outer(torch.arange(int(grid_thw[:, 1:].max())).float(), inv_freq)
            rot = freqs[pos].flatten(1)
            t = (rot.cos().contiguous().to(self.device), rot.sin().contiguous().to(self.device))
            self._rope2d_cache[key] = t
        return t

    @_on_device
    @torch.no_grad()
    def vit_forward(self, pixel_values, grid_thw, out=None):
        """Qwen2VisionTransformerPretrainedModel.forward (modeling_qwen2_vl.py:1050-1072): flattened patches
        [T, 3*2*14*14] + grid_thw [n_img, 3] -> merged tokens fp32 [T/4, hidden] (bf16-rounded values)."""
        if self.vit is None:
            raise RuntimeError("this model was built without the ViT (cfg.vit_depth == 0 or no vit_model.* weights)")
        cfg, dev, V = self.cfg, self.device, self.vit
        E, vh, hp = cfg.vit_embed_dim, cfg.vit_heads, self.vit_hp
        hd = E // vh
        grid_thw = grid_thw.cpu().long()
        pix = pixel_values.to(dev, torch.float32)
        T = pix.shape[0]
        pb = self.buf.get("vit.pix", tuple(pix.shape), torch.bfloat16)
        ops.cast_bf16(pix.contiguous(), pb)
        x = self.buf.get("vit.x", (T, E), torch.float32)     # bf16 residual stream held in fp32 (ROUND_SUM)
        ops.gemm(pb, V["wpatch"], x, epilogue=ops.EPI_STORE_F32, flags=ops.GEMM_ROUND_BF16)
        cos, sin = self._vit_rope_tables(grid_thw)
        cu = [0]
        for tt, h, w in grid_thw.tolist():
            for _ in range(tt):
                cu.append(cu[-1] + h * w)
        work = self._work(cu, cu, "vit")
        h = self.buf.get("vit.h", (T, E), torch.bfloat16)
        qkv = self.buf.get("vit.qkv", (T, 3 * vh * hp), torch.bfloat16)
        attn = self.buf.get("vit.attn", (T, vh * hp), torch.bfloat16, zero=True)
        mid = self.buf.get("vit.mid", (T, E * cfg.vit_mlp_ratio), torch.bfloat16)
        for B in V["blocks"]:
            ops.layernorm(x, h, B["norm1w"], B["norm1b"], 1e-6)
            ops.gemm(h, B["wqkv"], qkv, epilogue=ops.EPI_STORE_BF16, bias=B["bqkv"])
            ops.rope_vision(qkv, T, 2 * vh, hp, hd, cos, sin)
            ops.attention(qkv[:, : vh * hp], qkv[:, vh * hp: 2 * vh * hp], qkv[:, 2 * vh * hp:], attn, work,
                          num_q_heads=vh, num_kv_heads=vh, head_dim=hp, scale=1.0 / math.sqrt(hd))
            ops.gemm(attn, B["wproj"], x, epilogue=ops.EPI_RESID_F32, bias=B["bproj"], flags=ops.GEMM_ROUND_SUM)
            ops.layernorm(x, h, B["norm2w"], B["norm2b"], 1e-6)
            ops.gemm(h, B["wfc1"], mid, epilogue=ops.EPI_STORE_BF16, bias=B["bfc1"], flags=ops.GEMM_QUICK_GELU)
            ops.gemm(mid, B["wfc2"], x, epilogue=ops.EPI_RESID_F32, bias=B["bfc2"], flags=ops.GEMM_ROUND_SUM)
        m2 = cfg.vit_merge ** 2
        ops.layernorm(x, h, V["lnw"], V["lnb"], 1e-6)
        hm = h.view(T // m2, m2 * E)                                  # PatchMerger: 2x2 neighbours are consecutive rows
        mm = self.buf.get("vit.mm", (T // m2, m2 * E), torch.bfloat16)
        ops.gemm(hm, V["wm0"], mm, epilogue=ops.EPI_STORE_BF16, bias=V["bm0"], flags=ops.GEMM_GELU)
        if out is None:
            out = torch.empty(T // m2, cfg.hidden_size, dtype=torch.float32, device=dev)
        ops.gemm(mm, V["wm2"], out, epilogue=ops.EPI_STORE_F32, bias=V["bm2"], flags=ops.GEMM_ROUND_BF16)
        return out

    def prepare_vit_images(self, curr_kvlens, curr_rope, images, transforms, new_token_ids):
        """Reference: g2vlm.py:735-808 (+ get_rope_index_image_3D, data/data_utils.py:142-206), one image per call
        as chat_with_recon uses it. `transforms([image])` -> (pixel_values [T, C*2*14*14], grid_thw [1, 3])."""
        if len(images) != 1:
            raise NotImplementedError("one image per call (the reference's chat driver does the same)")
        kvlen, rope = int(curr_kvlens[0]), int(curr_rope[0])
        pix, grid = transforms([images[0]])
        t, h, w = [int(x) for x in grid[0]]
        m = self.cfg.vit_merge
        gh, gw = h // m, w // m
        n = t * gh * gw
        if pix.shape[0] // (m * m) != n:
            raise ValueError("pixel_values rows do not match grid_thw")
        ti = torch.arange(t).view(-1, 1).expand(-1, gh * gw).flatten()
        hi = torch.arange(gh).view(1, -1, 1).expand(t, -1, gw).flatten()
        wi = torch.arange(gw).view(1, 1, -1).expand(t, gh, -1).flatten()
        pid = torch.stack([ti, hi, wi]) + (rope + 1)
        delta = int(pid.max() - pid.min())
        pos = torch.cat([torch.full((3, 1), rope), pid, torch.full((3, 1), rope + delta + 2)], dim=1)
        gi = {
            "packed_text_ids": torch.tensor([int(new_token_ids["start_of_image"]), int(new_token_ids["end_of_image"])]),
            "packed_text_indexes": torch.tensor([0, n + 1]),
            "vit_token_seqlens": torch.tensor([n], dtype=torch.int),
            "packed_image_grid_thw": grid[:1].clone(),
            "packed_vit_images": pix[None],
            "packed_vit_token_indexes": torch.arange(1, n + 1),
            "packed_position_ids": pos,
            "packed_seqlens": torch.tensor([n + 2], dtype=torch.int),
            "packed_indexes": kvlen + torch.arange(n + 2),
            "packed_key_value_indexes": torch.arange(kvlen),
            "key_values_lens": torch.tensor([kvlen], dtype=torch.int),
        }
        return gi, [kvlen + n + 2], [rope + delta + 3]

    @_on_device
    @torch.no_grad()
    def forward_cache_update_vit(self, past_key_values, packed_text_ids, packed_text_indexes, packed_vit_images,
                                 packed_image_grid_thw, packed_vit_token_indexes, vit_token_seqlens, packed_position_ids,
                                 packed_seqlens, packed_indexes, packed_key_value_indexes, key_values_lens,
                                 packed_vit_tokens=None, packed_vit_position_ids=None):
        """Reference: g2vlm.py:810-866 — ViT tokens between <start>/<end> through the und expert, NON-causal, on
        top of the cache.  Returns the (appended) KVCache."""
        cfg, dev = self.cfg, self.device
        cache = KVCache.adopt(past_key_values, cfg, dev)
        n = int(sum(packed_seqlens.tolist()))
        x = torch.zeros(n, cfg.hidden_size, dtype=torch.float32, device=dev)
        nt = int(packed_text_ids.numel())
        txt = torch.empty(nt, cfg.hidden_size, dtype=torch.float32, device=dev)
        ops.gather_rows(self.embed, txt, packed_text_ids.to(dev, torch.long), nt)
        ops.gather_rows(txt, x, packed_text_indexes.to(dev, torch.long), nt, scatter=True)
        pix = packed_vit_images.reshape(-1, packed_vit_images.shape[-1])
        emb = self.vit_forward(pix, packed_image_grid_thw)
        ops.gather_rows(emb, x, packed_vit_token_indexes.to(dev, torch.long), emb.shape[0], scatter=True)
        self._und_forward(x, packed_position_ids, cache, causal=False)
        return cache

    def prepare_prompts_pure_text(self, curr_kvlens, curr_rope, prompts, tokenizer, new_token_ids):
        """Reference: g2vlm.py:628-662 (no bos / eos added)."""
        kvlen, rope = int(curr_kvlens[0]), int(curr_rope[0])
        ids = list(tokenizer.encode(prompts[0]))
        n = len(ids)
        gi = {
            "text_token_lens": torch.tensor([n], dtype=torch.int),
            "packed_text_ids": torch.tensor(ids, dtype=torch.long),
            "packed_text_position_ids": (rope + torch.arange(n)).expand(3, -1),
            "packed_text_indexes": kvlen + torch.arange(n),
            "packed_key_value_indexes": torch.arange(kvlen),
            "key_values_lens": torch.tensor([kvlen], dtype=torch.int),
        }
        return gi, [kvlen + n], [rope + n]

    def prepare_start_tokens(self, curr_kvlens, curr_rope, tokenizer, new_token_ids):
        """Reference: g2vlm.py:1041-1068: the start token is the LAST id of the chat template."""
        template = "<|im_start|>user\\your text<|im_end|>\n<|im_start|>assistant\n"
        ids = tokenizer.encode(template, add_special_tokens=False)
        start = ids[-1] if ids else (getattr(tokenizer, "eos_token_id", None) or 151643)
        kvlen, rope = int(curr_kvlens[0]), int(curr_rope[0])
        return {
            "packed_start_tokens": torch.tensor([start], dtype=torch.long),
            "packed_query_position_ids": torch.tensor([rope], dtype=torch.long).expand(3, -1),
            "key_values_lens": torch.tensor([kvlen], dtype=torch.int),
            "packed_key_value_indexes": torch.arange(kvlen),
        }

    @_on_device
    @torch.no_grad()
    def chat_with_recon(self, tokenizer, new_token_ids, image_transform, dino_image_transform, images, prompt,
                        max_length: int, do_sample: bool = False, temperature: float = 1.0, return_ids: bool = False):
        """Same call shape as the reference's chat_with_recon (g2vlm.py:1305-1410): system prompt (und, causal)
        -> geo step over the views with cache update -> one ViT step per image (und, non-causal) -> question
        (und, causal) -> greedy decode.  `images`: list of PIL images (or an (N,3,H,W) tensor for the geo branch
        together with `image_transform` accepting its items)."""
        dev = self.device
        past = KVCache(self.cfg.num_layers, self.cfg.num_kv_heads, self.cfg.head_dim, dev)
        system_prompt = "<|im_start|>system\nYou are a helpful assistant.<|im_end|>\n<|im_start|>user\n"
        gi, newlens, new_rope = self.prepare_prompts_pure_text([0], [0], [system_prompt], tokenizer, new_token_ids)
        past = self.forward_cache_update_text(past, **gi)
        gi, newlens, new_rope = self.prepare_dino_images_pi3(newlens, new_rope, images if torch.is_tensor(images) else list(images),
                                                             dino_image_transform, new_token_ids)
        past, _ = self.forward_cache_update_dino(past, **gi)
        for image in images:
            gi, newlens, new_rope = self.prepare_vit_images(newlens, new_rope, [image], image_transform, new_token_ids)
            past = self.forward_cache_update_vit(past, **gi)
        gi, newlens, new_rope = self.prepare_prompts_pure_text(newlens, new_rope, [prompt + "<|im_end|>\n<|im_start|>assistant"],
                                                               tokenizer, new_token_ids)
        past = self.forward_cache_update_text(past, **gi)
        st = self.prepare_start_tokens(newlens, new_rope, tokenizer, new_token_ids)
        ids = self.generate_text(past_key_values=past, max_length=max_length, do_sample=do_sample, temperature=temperature,
                                 end_token_id=new_token_ids["eos_token_id"], **st)
        if return_ids:
            return ids
        return tokenizer.decode(ids[1:, 0])   # skip the start token, like the reference

    # ------------------------------------------------------------------------------------------
    # the public entry point
    # ------------------------------------------------------------------------------------------
    prepare_prompts_addbos = staticmethod(host_prep.prepare_prompts_addbos)

    def prepare_dino_images_pi3(self, curr_kvlens, curr_rope, images, transforms, new_token_ids,
                                normalize_on_host: bool = True):
        return host_prep.prepare_dino_images_pi3(curr_kvlens, curr_rope, images, new_token_ids, self.cfg.dino_patch,
                                                 normalize_on_host=normalize_on_host)

    @_on_device
    @torch.no_grad()
    def recon(self, tokenizer, new_token_ids, dino_image_transform, images, prompt="Reconstruct the 3D scene.",
              collect: Optional[dict] = None):
        """Same call shape and result dict as the reference's G2VLM.recon (g2vlm.py:1240-1303).  Like the
        reference, `dino_image_transform` and `prompt` are ignored (the prompt is hard-coded, :1264).
        `images`: list of paths / PIL images, or an (N,3,H,W) tensor in [0,1] (H, W multiples of 14)."""
        dev = self.device
        past = KVCache(self.cfg.num_layers, self.cfg.num_kv_heads, self.cfg.head_dim, dev)
        self._mark("start")
        gi, newlens, new_rope = self.prepare_prompts_addbos([0], [0], ["Reconstruct the 3D scene."], tokenizer,
                                                            new_token_ids)
        prompt = None
        if self.fuse_prompt:
            prompt = gi                                     # the 7 prompt rows ride along with the geo step
        else:
            past = self.forward_cache_update_text(past, **gi)   # index tensors stay on the host: _idx stages them
        if not torch.is_tensor(images) and self.device_resize:
            # paths / PIL images: only the decoded uint8 pixels cross PCIe; Pillow-exact LANCZOS + ToTensor on the GPU
            from .host_prep import load_and_resize14_device
            images = load_and_resize14_device(images, 518, dev)
        # the raw views cross PCIe once; normalisation happens on the device (bit-identical to the host op)
        gi, newlens, new_rope = self.prepare_dino_images_pi3(newlens, new_rope, images, dino_image_transform,
                                                             new_token_ids, normalize_on_host=False)
        if collect is not None:
            collect["generation_input"] = {k: v.clone() for k, v in gi.items() if torch.is_tensor(v)}
        raw = gi["original_images"].to(dev, non_blocking=True)
        gi = dict(gi)
        gi["packed_dino_images"] = gi["original_images"] = raw
        self._raw_images = True
        try:
            past, last = self.forward_cache_update_dino(past, update_past_key_values=False, collect=collect,
                                                        prompt=prompt, **gi)
        finally:
            self._raw_images = False
        if collect is not None:
            collect["last_hidden"] = last.clone()
        return self.reconstruct(past_key_values=past, selected_hidden_states=last, collect=collect, **gi)
