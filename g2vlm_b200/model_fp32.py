"""fp32 mode of the recon path: `G2VLMFast(cfg, state_dict, mode="fp32")`.

`north_star`: "point maps, confidences and camera poses within a stated bf16 tolerance (max rel err <= 2e-2, fp32 mode
<= 1e-4)"; BASELINE configs[0] is "8 views fp32".  The reference itself cannot run in fp32 (hard bf16 casts at
modeling/g2vlm/qwen2vl.py:579, 617-619; dinov2_model.py:50-52), so the ground truth of this mode is the restatement
`oracle/restate.py` with mode="fp32" — the same algorithm with every bf16 rounding point removed.

Arithmetic: no bf16 rounding of any activation.
* every nn.Linear is a split-bf16 GEMM on the tcgen05 tensor cores: x = h + m + l and w = h + m + l EXACTLY (three
  bf16 pieces hold all 24 significand bits), x.w = mm + lh + hl + mh + hm + hh (+ terms <= 2^-24) as ONE bf16 GEMM over
  the concatenated K ([m|l|h|m|h|h] x [m|h|l|h|m|h]^T, smallest products first).  The tensor core adds into its fp32
  accumulator with truncation, so the accumulator leaves TMEM every 256 K elements and the chunks are summed in fp32
  round-to-nearest by the epilogue warps (`k_chunk_blocks`).  Measured on the way here at full width, depth 1
  (8 x 294x518): two pieces / one accumulation 1.04e-4 on local_points (on the tolerance: exp(z) turns the absolute
  error of z into a relative one); three pieces / one accumulation still 3e-5 per MoT layer (the truncation, not the
  split, was the limit);
* attention: `g2vlm_attention_f32` (FP32-pipe dot products, exp2f softmax), no head padding (Pi3 heads are 96 wide);
* norms / rotary / SwiGLU / GELU / LayerScale / residuals: fp32 kernels or fp32 GEMM epilogues (STORE_F32 with the
  GELU / scale / residual options).
Only `recon()` and its two stages are implemented in this mode (the chat, ViT, training-forward and view-sharded paths
are bf16-only and raise).  This is the verification mode, not the benchmarked path: ~6x the tensor work of bf16 plus
FP32-pipe attention.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import host_prep, ops
from .model import G2VLMFast, KVCache, NaiveCache, _f32, _on_device, _split_hi_lo_hi


def _w3(w: torch.Tensor, dev) -> torch.Tensor:
    """fp32 weight [N, K] -> bf16 [N, 6K] = [m | h | l | h | m | h] on the device, w = h + m + l exactly; pairs with
    activations split as [m | l | h | m | h | h] by g2vlm_split6_f32: the six products mm, lh, hl, mh, hm, hh are
    accumulated smallest first."""
    w = w.float()
    h = w.to(torch.bfloat16)
    r1 = w - h.float()
    m = r1.to(torch.bfloat16)
    l = (r1 - m.float()).to(torch.bfloat16)
    return torch.cat([m, h, l, h, m, h], dim=1).contiguous().to(dev)


K_CHUNK_BLOCKS = 4   # the accumulator leaves the tensor core every 4 x 64 K elements (see g2vlm_gemm_args.k_chunk_blocks)


class G2VLMFastFP32(G2VLMFast):
    mode = "fp32"

    # ------------------------------------------------------------------------------------------
    # weights: reference key names in, [hi|lo|hi] split weights + exact fp32 biases / norm weights out
    # ------------------------------------------------------------------------------------------
    def _pack(self, sd):
        cfg, dev = self.cfg, self.device
        self._split_rows, self._tmp_rows = {}, {}
        g = lambda k: sd[k].detach().float()
        lm = "language_model.model."
        self.embed = _f32(g(lm + "embed_tokens.weight"), dev)
        self.lm_head = None
        self.vit = None
        self.layers = []
        for i in range(cfg.num_layers):
            p = f"{lm}layers.{i}."
            a = p + "self_attn."
            L = {}
            qkv, bias = [], []
            for sfx in ("_moe_geo", ""):  # expert order: geo (group 0), und (group 1)
                qkv += [g(a + f"q_proj{sfx}.weight"), g(a + f"k_proj{sfx}.weight"), g(a + f"v_proj{sfx}.weight")]
                bias += [g(a + f"q_proj{sfx}.bias"), g(a + f"k_proj{sfx}.bias"), g(a + f"v_proj{sfx}.bias")]
            L["wqkv"] = _w3(torch.cat(qkv, 0), dev)
            L["bqkv"] = _f32(torch.cat(bias, 0), dev)
            L["wo"] = _w3(torch.cat([g(a + "o_proj_moe_geo.weight"), g(a + "o_proj.weight")], 0), dev)
            # gate rows then up rows per expert (the fp32 SwiGLU is a separate elementwise kernel: no interleave)
            L["wgu"] = _w3(torch.cat([g(p + "mlp_moe_geo.gate_proj.weight"), g(p + "mlp_moe_geo.up_proj.weight"),
                                      g(p + "mlp.gate_proj.weight"), g(p + "mlp.up_proj.weight")], 0), dev)
            L["wdown"] = _w3(torch.cat([g(p + "mlp_moe_geo.down_proj.weight"), g(p + "mlp.down_proj.weight")], 0), dev)
            for n in ("input_layernorm", "post_attention_layernorm"):
                L[n + "_geo"] = _f32(g(p + n + "_moe_geo.weight"), dev)
                L[n + "_und"] = _f32(g(p + n + ".weight"), dev)
            for n in ("q_norm", "k_norm"):
                L[n + "_geo"] = _f32(g(a + n + "_moe_geo.weight"), dev)
                L[n + "_und"] = _f32(g(a + n + ".weight"), dev)
            L["ls1"] = _f32(g(p + "ls1.gamma"), dev)
            L["ls2"] = _f32(g(p + "ls2.gamma"), dev)
            self.layers.append(L)
        self.norm_geo = _f32(g(lm + "norm_moe_geo.weight"), dev)
        self.norm_und = _f32(g(lm + "norm.weight"), dev)
        hd = cfg.head_dim
        inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
        self.inv_freq = inv.contiguous().to(dev)

        # ---- DINO ----
        D = cfg.dino_hidden
        d = "dino_model."
        kp = 3 * cfg.dino_patch ** 2
        self.dino_kpad = (kp + 63) // 64 * 64
        wp = torch.zeros(D, self.dino_kpad, device=sd[d + "embeddings.patch_embeddings.projection.weight"].device)
        wp[:, :kp] = g(d + "embeddings.patch_embeddings.projection.weight").reshape(D, kp)
        self.dino_wpatch = _w3(wp, dev)
        self.dino_bpatch = _f32(g(d + "embeddings.patch_embeddings.projection.bias"), dev)
        self.dino_cls = _f32(g(d + "embeddings.cls_token").reshape(D), dev)
        self.dino_reg = _f32(g(d + "embeddings.register_tokens").reshape(cfg.dino_registers, D), dev)
        self.dino_pos_table = g(d + "embeddings.position_embeddings")
        self.dino_layers = []
        for i in range(cfg.dino_layers):
            p = f"{d}encoder.layer.{i}."
            at = p + "attention.attention."
            L = {}
            L["wqkv"] = _w3(torch.cat([g(at + f"{n}.weight") for n in ("query", "key", "value")], 0), dev)
            L["bqkv"] = _f32(torch.cat([g(at + f"{n}.bias") for n in ("query", "key", "value")], 0), dev)
            L["wdense"] = _w3(g(p + "attention.output.dense.weight"), dev)
            L["bdense"] = _f32(g(p + "attention.output.dense.bias"), dev)
            L["wfc1"] = _w3(g(p + "mlp.fc1.weight"), dev); L["bfc1"] = _f32(g(p + "mlp.fc1.bias"), dev)
            L["wfc2"] = _w3(g(p + "mlp.fc2.weight"), dev); L["bfc2"] = _f32(g(p + "mlp.fc2.bias"), dev)
            for n in ("norm1", "norm2"):
                L[n + "w"] = _f32(g(p + n + ".weight"), dev); L[n + "b"] = _f32(g(p + n + ".bias"), dev)
            L["ls1"] = _f32(g(p + "layer_scale1.lambda1"), dev)
            L["ls2"] = _f32(g(p + "layer_scale2.lambda1"), dev)
            self.dino_layers.append(L)
        self.dino_lnw = _f32(g(d + "layernorm.weight"), dev)
        self.dino_lnb = _f32(g(d + "layernorm.bias"), dev)
        self.w_dino2llm = _w3(g("dino2llm.weight"), dev)
        self.b_dino2llm = _f32(g("dino2llm.bias"), dev)

        # ---- Pi3 decoders (heads stay 96 wide: the fp32 attention kernel needs no padding) ----
        def pack_block(p, cross):
            B = {}
            B["wqkv"] = _w3(g(p + "attn.qkv.weight"), dev); B["bqkv"] = _f32(g(p + "attn.qkv.bias"), dev)
            B["wproj"] = _w3(g(p + "attn.proj.weight"), dev); B["bproj"] = _f32(g(p + "attn.proj.bias"), dev)
            for n in ["norm1", "norm2"] + (["norm3", "norm_y"] if cross else []):
                B[n + "w"] = _f32(g(p + n + ".weight"), dev); B[n + "b"] = _f32(g(p + n + ".bias"), dev)
            if cross:
                c = p + "cross_attn."
                B["wcq"] = _w3(g(c + "q_proj.weight"), dev); B["bcq"] = _f32(g(c + "q_proj.bias"), dev)
                B["wckv"] = _w3(torch.cat([g(c + "k_proj.weight"), g(c + "v_proj.weight")], 0), dev)
                B["bckv"] = _f32(torch.cat([g(c + "k_proj.bias"), g(c + "v_proj.bias")], 0), dev)
                B["wcproj"] = _w3(g(c + "proj.weight"), dev); B["bcproj"] = _f32(g(c + "proj.bias"), dev)
            B["wfc1"] = _w3(g(p + "mlp.fc1.weight"), dev); B["bfc1"] = _f32(g(p + "mlp.fc1.bias"), dev)
            B["wfc2"] = _w3(g(p + "mlp.fc2.weight"), dev); B["bfc2"] = _f32(g(p + "mlp.fc2.bias"), dev)
            return B

        self.decoders = {}
        dec_names = [("point_decoder", False), ("camera_decoder", False), ("global_points_decoder", True)]
        head_names = ["point_head", "global_point_head"]
        if cfg.train_conf_pi3:
            dec_names.append(("conf_decoder", False))
            head_names.append("conf_head")
        for name, cross in dec_names:
            blocks = [pack_block(f"{name}.blocks.{i}.", cross) for i in range(cfg.dec_depth)]
            self.decoders[name] = dict(blocks=blocks, cross=cross, wout=_w3(g(f"{name}.linear_out.weight"), dev),
                                       bout=_f32(g(f"{name}.linear_out.bias"), dev))
        for hname in head_names:
            setattr(self, hname + "_w3", _w3(g(hname + ".proj.weight"), dev))
            setattr(self, hname + "_b", _f32(g(hname + ".proj.bias"), dev))
        self.cam = {}
        for i in range(2):
            for j in (1, 2, 3):
                k = f"camera_head.res_conv.{i}.res_conv{j}"
                self.cam[f"r{i}{j}w"] = _w3(g(k + ".weight"), dev)
                self.cam[f"r{i}{j}b"] = _f32(g(k + ".bias"), dev)
        for j in (0, 2):
            self.cam[f"m{j}w"] = _w3(g(f"camera_head.more_mlps.{j}.weight"), dev)
            self.cam[f"m{j}b"] = _f32(g(f"camera_head.more_mlps.{j}.bias"), dev)
        for n in ("fc_t", "fc_rot"):
            self.cam[n + "w"] = _f32(g(f"camera_head.{n}.weight"), dev)
            self.cam[n + "b"] = _f32(g(f"camera_head.{n}.bias"), dev)

    # ------------------------------------------------------------------------------------------
    # building block: true-fp32 nn.Linear on the bf16 tensor cores
    # ------------------------------------------------------------------------------------------
    def _lin(self, x, w3, bias, out, rows=None, groups=None, flags=0, scale=None, scale_groups=0, residual=None):
        """out = [residual +] [scale *] [gelu|relu] (x @ w.T + bias), fp32 in / out; x [rows, K] fp32."""
        rows = x.shape[0] if rows is None else rows
        k = x.shape[1]
        xs = self.buf.get(f"f32.split.{k}", (max(rows, self._split_rows.get(k, 0)), 6 * k), torch.bfloat16)
        self._split_rows[k] = xs.shape[0]
        ops.split6(x, xs, rows)
        dst = out
        if residual is not None and residual.data_ptr() == out.data_ptr():
            # x = x + f(x): the K chunks accumulate in the output buffer, so the residual operand must live elsewhere
            n = out.shape[1]
            dst = self.buf.get(f"f32.lin_tmp.{n}", (max(rows, self._tmp_rows.get(n, 0)), n), torch.float32)
            self._tmp_rows[n] = dst.shape[0]
        ops.gemm(xs[:rows], w3, dst, epilogue=ops.EPI_STORE_F32, groups=groups, bias=bias, flags=flags, scale=scale,
                 scale_groups=scale_groups, residual=residual, k_chunk_blocks=K_CHUNK_BLOCKS)
        if dst is not out:
            ops.gather_rows(dst, out, None, rows)
        return out

    def _linear_fp32(self, x, w3, b, out, relu=False, residual=None):
        """The camera head (shared driver `_camera_head`) with this mode's three-piece split."""
        return self._lin(x, w3, b, out, flags=ops.GEMM_RELU if relu else 0, residual=residual)

    def _rope2d_tables_f32(self, gh: int, gw: int):
        """RoPE2D tables of EXACT fp32 angles (pos_embed.py:120-128 without the bf16 cast of the bf16 path)."""
        key = ("f32", gh, gw)
        t = self._rope2d_cache.get(key)
        if t is None:
            D = self.cfg.dec_head_dim // 2
            inv_freq = 1.0 / (self.cfg.rope2d_base ** (torch.arange(0, D, 2).float() / D))
            freqs = torch.einsum("i,j->ij", torch.arange(max(gh, gw), dtype=inv_freq.dtype), inv_freq)
            t = (freqs.cos().contiguous().to(self.device), freqs.sin().contiguous().to(self.device))
            self._rope2d_cache[key] = t
        return t

    # ------------------------------------------------------------------------------------------
    # DINO encoder
    # ------------------------------------------------------------------------------------------
    @_on_device
    @torch.no_grad()
    def dino_forward(self, packed_pixel_values, dino_token_seqlens, collect: Optional[list] = None):
        """Dinov2WithRegistersModel.forward (g2vlm/dinov2_model.py:301-356) -> fp32 tokens [N*P, D]."""
        cfg, dev = self.cfg, self.device
        img = packed_pixel_values.to(dev, torch.float32).contiguous()
        N, _, Hh, Ww = img.shape
        p = cfg.dino_patch
        gh, gw = Hh // p, Ww // p
        P, D, nh = gh * gw, cfg.dino_hidden, cfg.dino_heads
        S = P + 1 + cfg.dino_registers
        rows = N * S
        patches = self.buf.get("f32.dino.patches", (N * P, self.dino_kpad), torch.float32)
        if self._raw_images:
            ops.im2col_patches_f32(img, patches, p, host_prep.RESNET_MEAN, host_prep.RESNET_STD)
        else:
            ops.im2col_patches_f32(img, patches, p)
        emb = self.buf.get("f32.dino.emb", (N * P, D), torch.float32)
        self._lin(patches, self.dino_wpatch, self.dino_bpatch, emb)
        x = self.buf.get("f32.dino.x", (rows, D), torch.float32)
        ops.dino_embed_f32(emb, self.dino_cls, self.dino_reg, self._dino_pos(gh, gw, Hh == Ww), x, N, P, cfg.dino_registers)
        cu = [0]
        for n in dino_token_seqlens.tolist():
            cu.append(cu[-1] + int(n))
        if cu[-1] > rows:
            raise ValueError("dino_token_seqlens exceed the number of DINO rows")
        work = self._work(cu, cu, "dino")
        h = self.buf.get("f32.dino.h", (rows, D), torch.float32)
        qkv = self.buf.get("f32.dino.qkv", (rows, 3 * D), torch.float32)
        attn = self.buf.get("f32.dino.attn", (rows, D), torch.float32)
        attn.zero_()   # rows in no segment (quirk Q1) are defined as zero; re-zeroed per call (coverage may change)
        mid = self.buf.get("f32.dino.mid", (rows, D * cfg.dino_mlp_ratio), torch.float32)
        scale = 1.0 / math.sqrt(cfg.dino_head_dim)
        for L in self.dino_layers:
            ops.layernorm(x, h, L["norm1w"], L["norm1b"], cfg.dino_ln_eps)
            self._lin(h, L["wqkv"], L["bqkv"], qkv)
            ops.attention_f32(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], attn, work, num_q_heads=nh, num_kv_heads=nh,
                              head_dim=cfg.dino_head_dim, scale=scale)
            self._lin(attn, L["wdense"], L["bdense"], x, scale=L["ls1"], scale_groups=1, residual=x)
            ops.layernorm(x, h, L["norm2w"], L["norm2b"], cfg.dino_ln_eps)
            self._lin(h, L["wfc1"], L["bfc1"], mid, flags=ops.GEMM_GELU)
            self._lin(mid, L["wfc2"], L["bfc2"], x, scale=L["ls2"], scale_groups=1, residual=x)
            if collect is not None:
                collect.append(x.clone())
        tokens = self.buf.get("f32.dino.tokens", (N * P, D), torch.float32)
        ops.layernorm(x, tokens, self.dino_lnw, self.dino_lnb, cfg.dino_ln_eps, seg_in=S, seg_skip=1 + cfg.dino_registers)
        return tokens

    # ------------------------------------------------------------------------------------------
    # MoT language model: geo step with the prompt rows riding along (see G2VLMFast.language_model_forward_geo)
    # ------------------------------------------------------------------------------------------
    @_on_device
    @torch.no_grad()
    def language_model_forward_geo(self, packed_sequence, packed_position_ids, packed_geo_token_indexes,
                                   packed_text_indexes, past_key_values=None, update_past_key_values: bool = False,
                                   collect: Optional[list] = None, group=None, prompt: Optional[dict] = None,
                                   rank_rows=None):
        cfg, dev = self.cfg, self.device
        nq, nkv, hd, H, I = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim, cfg.hidden_size, cfg.intermediate_size
        if group is not None or update_past_key_values or prompt is None:
            raise NotImplementedError("fp32 mode implements the fused recon geo step only (prompt rows riding along, "
                                      "no cache update, no view sharding)")
        if past_key_values is not None and KVCache.adopt(past_key_values, cfg, dev).len:
            raise ValueError("fp32 mode needs an empty cache (the prompt prefill is part of the geo step)")
        T = packed_sequence.shape[0]
        n_geo = int(packed_geo_token_indexes.numel())
        perm = self._idx("mot.perm", torch.cat([packed_geo_token_indexes.long().cpu(), packed_text_indexes.long().cpu()]))
        if int(perm.numel()) != T:
            raise ValueError("geo + text indexes must cover every packed row exactly once")
        Kp = int(prompt["packed_text_ids"].numel())
        R = T + Kp
        x = self.buf.get("f32.mot.x", (R, H), torch.float32)
        ops.gather_rows(packed_sequence, x, perm, T)
        cos_p = self.buf.get("f32.mot.cos_p", (T, hd // 2), torch.float32)
        sin_p = self.buf.get("f32.mot.sin_p", (T, hd // 2), torch.float32)
        ops.mrope_table(self._idx("mot.pos", packed_position_ids.contiguous()), self.inv_freq, cos_p, sin_p, cfg.mrope_section)
        cos = self.buf.get("f32.mot.cos", (R, hd // 2), torch.float32)
        sin = self.buf.get("f32.mot.sin", (R, hd // 2), torch.float32)
        ops.gather_rows(cos_p, cos, perm, T)
        ops.gather_rows(sin_p, sin, perm, T)
        ops.gather_rows(self.embed, x[T:], self._idx("mot.prompt_ids", prompt["packed_text_ids"]), Kp)
        ops.mrope_table(self._idx("mot.prompt_pos", prompt["packed_text_position_ids"].contiguous()), self.inv_freq,
                        cos[T:], sin[T:], cfg.mrope_section)
        qkvw = (nq + 2 * nkv) * hd
        qkv = self.buf.get("f32.mot.qkv", (R, qkvw), torch.float32)
        attn = self.buf.get("f32.mot.attn", (R, nq * hd), torch.float32)
        gu = self.buf.get("f32.mot.gu", (R, 2 * I), torch.float32)
        act = self.buf.get("f32.mot.act", (R, I), torch.float32)
        hbuf = self.buf.get("f32.mot.h", (R, H), torch.float32)
        key = ("geo_fused", T, Kp)
        work = self._work_cache.get(key)
        if work is None:   # geo rows see every key incl. the prompt's; the prompt rows are causal among themselves
            items = ops.attention_work_table([0, T], [0, R]).tolist()
            items += [[t0, T, R, T, R, 1, 0, 0] for t0 in range(T, R, ops.ATTN_ROWS_PER_ITEM)]
            work = self._work_cache[key] = torch.tensor(items, dtype=torch.int32).to(dev)
        groups = [(0, n_geo), (n_geo, R - n_geo)]
        scale = 1.0 / math.sqrt(hd)
        for L in self.layers:
            ops.rmsnorm_routed(x, hbuf, L["input_layernorm_geo"], L["input_layernorm_und"], n_geo, cfg.rms_norm_eps, rows=R)
            self._lin(hbuf, L["wqkv"], L["bqkv"], qkv, groups=groups)
            ops.qknorm_mrope_f32(qkv, R, n_geo, nq, nkv, hd, L["q_norm_geo"], L["k_norm_geo"], L["q_norm_und"],
                                 L["k_norm_und"], cos, sin, cfg.rms_norm_eps)
            ops.attention_f32(qkv[:, : nq * hd], qkv[:, nq * hd:(nq + nkv) * hd], qkv[:, (nq + nkv) * hd:], attn, work,
                              num_q_heads=nq, num_kv_heads=nkv, head_dim=hd, scale=scale)
            self._lin(attn, L["wo"], None, x, groups=groups, scale=L["ls1"], scale_groups=1, residual=x)
            ops.rmsnorm_routed(x, hbuf, L["post_attention_layernorm_geo"], L["post_attention_layernorm_und"], n_geo,
                               cfg.rms_norm_eps, rows=R)
            self._lin(hbuf, L["wgu"], None, gu, groups=groups)
            ops.swiglu_f32(gu, act, rows=R)
            self._lin(act, L["wdown"], None, x, groups=groups, scale=L["ls2"], scale_groups=1, residual=x)
            if collect is not None:
                y = torch.empty(T, H, dtype=torch.float32, device=dev)
                ops.gather_rows(x, y, perm, T, scatter=True)
                collect.append(y)
        y_int = self.buf.get("f32.mot.y", (T, H), torch.float32)
        ops.rmsnorm_routed(x, y_int, self.norm_geo, self.norm_und, n_geo, cfg.rms_norm_eps, rows=T)
        last = torch.empty(T, H, dtype=torch.float32, device=dev)
        ops.gather_rows(y_int, last, perm, T, scatter=True)
        return last, past_key_values

    @_on_device
    @torch.no_grad()
    def forward_cache_update_dino(self, past_key_values, packed_text_ids, packed_text_indexes, packed_dino_token_indexes,
                                  dino_token_seqlens, packed_position_ids, packed_seqlens, packed_indexes,
                                  packed_key_value_indexes, key_values_lens, packed_dino_images, original_images,
                                  update_past_key_values: bool = False, collect: Optional[dict] = None, shard=None,
                                  group=None, prompt: Optional[dict] = None):
        """Reference: g2vlm.py:968-1039 (fp32 arithmetic).  Returns (past_key_values, last_hidden_state [T, H] fp32)."""
        cfg = self.cfg
        if shard is not None:
            raise NotImplementedError("view sharding is bf16-only")
        T, H = int(sum(packed_seqlens.tolist())), cfg.hidden_size
        self._mark("dino_begin")
        tokens = self.dino_forward(packed_dino_images, dino_token_seqlens,
                                   collect=None if collect is None else collect.setdefault("dino_layers", []))
        self._mark("dino_end")
        n_geo = tokens.shape[0]
        geo_emb = self.buf.get("f32.mot.geo_emb", (n_geo, H), torch.float32)
        self._lin(tokens, self.w_dino2llm, self.b_dino2llm, geo_emb)
        packed = self.buf.get("f32.mot.packed", (T, H), torch.float32)
        n_und = int(packed_text_ids.numel())
        txt = self.buf.get("f32.mot.txt", (n_und, H), torch.float32)
        ops.gather_rows(self.embed, txt, self._idx("dino.txt_ids", packed_text_ids), n_und)
        ops.gather_rows(txt, packed, self._idx("dino.txt_idx", packed_text_indexes), n_und, scatter=True)
        ops.gather_rows(geo_emb, packed, self._idx("dino.geo_idx", packed_dino_token_indexes), n_geo, scatter=True)
        if collect is not None:
            collect["dino_tokens"] = tokens.clone()
            collect["packed_sequence"] = packed.clone()
        last, past_key_values = self.language_model_forward_geo(
            packed, packed_position_ids, packed_dino_token_indexes, packed_text_indexes, past_key_values,
            update_past_key_values=update_past_key_values,
            collect=None if collect is None else collect.setdefault("mot_layers", []), prompt=prompt)
        self._mark("mot_end")
        return past_key_values, last

    # ------------------------------------------------------------------------------------------
    # Pi3 decoders + heads
    # ------------------------------------------------------------------------------------------
    def _decoder(self, name, hidden, N, P, gh, gw, out, out_fp32_round=False, context=None):
        """Pi3TransformerDecoder / Pi3ContextTransformerDecoder (transformer_head.py:48-56, 122-131), fp32."""
        cfg, dec = self.cfg, self.decoders[name]
        H, dh, ehd = cfg.hidden_size, cfg.dec_heads, cfg.dec_head_dim
        rows = N * P
        x = self.buf.get("f32.dec.x", (rows, H), torch.float32)
        ops.gather_rows(hidden, x, None, rows)
        h = self.buf.get("f32.dec.h", (rows, H), torch.float32)
        qkv = self.buf.get("f32.dec.qkv", (rows, 3 * H), torch.float32)
        attn = self.buf.get("f32.dec.attn", (rows, H), torch.float32)
        mid = self.buf.get("f32.dec.mid", (rows, H * cfg.dec_mlp_ratio), torch.float32)
        cos, sin = self._rope2d_tables_f32(gh, gw)
        cu = [v * P for v in range(N + 1)]
        work = self._work(cu, cu, "dec")
        scale = 1.0 / math.sqrt(ehd)
        att = dict(num_q_heads=dh, num_kv_heads=dh, head_dim=ehd, scale=scale)
        if dec["cross"]:
            yh = self.buf.get("f32.dec.yh", (P, H), torch.float32)
            kvc = self.buf.get("f32.dec.kvc", (P, 2 * H), torch.float32)
            qc = self.buf.get("f32.dec.qc", (rows, H), torch.float32)
            cwork = self._cross_work(N, P)
        for B in dec["blocks"]:
            ops.layernorm(x, h, B["norm1w"], B["norm1b"], 1e-6)
            self._lin(h, B["wqkv"], B["bqkv"], qkv)
            ops.rope2d_f32(qkv, rows, 2 * dh, ehd, ehd, P, gw, cos, sin)          # q heads then k heads
            ops.attention_f32(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], attn, work, **att)
            self._lin(attn, B["wproj"], B["bproj"], x, residual=x)
            if dec["cross"]:
                ops.layernorm(context, yh, B["norm_yw"], B["norm_yb"], 1e-6, rows=P)
                self._lin(yh, B["wckv"], B["bckv"], kvc, rows=P)
                ops.rope2d_f32(kvc, P, dh, ehd, ehd, P, gw, cos, sin)              # k heads only
                ops.layernorm(x, h, B["norm2w"], B["norm2b"], 1e-6)
                self._lin(h, B["wcq"], B["bcq"], qc)
                ops.rope2d_f32(qc, rows, dh, ehd, ehd, P, gw, cos, sin)
                ops.attention_f32(qc, kvc[:, :H], kvc[:, H:], attn, cwork, **att)
                self._lin(attn, B["wcproj"], B["bcproj"], x, residual=x)
                ops.layernorm(x, h, B["norm3w"], B["norm3b"], 1e-6)
            else:
                ops.layernorm(x, h, B["norm2w"], B["norm2b"], 1e-6)
            self._lin(h, B["wfc1"], B["bfc1"], mid, flags=ops.GEMM_GELU)
            self._lin(mid, B["wfc2"], B["bfc2"], x, residual=x)
        return self._lin(x, dec["wout"], dec["bout"], out)

    @_on_device
    @torch.no_grad()
    def reconstruct(self, past_key_values=None, packed_key_value_indexes=None, key_values_lens=None,
                    selected_hidden_states=None, packed_dino_token_indexes=None, packed_dino_images=None,
                    original_images=None, collect: Optional[dict] = None, shard=None, group=None, **kwargs):
        """Reference: g2vlm.py:1143-1238 (fp32 arithmetic throughout)."""
        cfg, dev = self.cfg, self.device
        if shard is not None:
            raise NotImplementedError("view sharding is bf16-only")
        N, _, Hh, Ww = packed_dino_images.shape
        p = cfg.dino_patch
        gh, gw = Hh // p, Ww // p
        P, H = gh * gw, cfg.hidden_size
        rows = N * P
        self._mark("heads_begin")
        geo = self._idx("recon.geo_idx", packed_dino_token_indexes)
        hidden = self.buf.get("f32.rec.hidden", (rows, H), torch.float32)
        ops.gather_rows(selected_hidden_states, hidden, geo, rows)
        context = hidden[:P]
        point_hidden = self.buf.get("f32.rec.point_hidden", (rows, cfg.point_dim), torch.float32)
        self._decoder("point_decoder", hidden, N, P, gh, gw, point_hidden)
        camera_hidden = self.buf.get("f32.rec.camera_hidden", (rows, cfg.camera_dim), torch.float32)
        self._decoder("camera_decoder", hidden, N, P, gh, gw, camera_hidden)
        global_hidden = self.buf.get("f32.rec.global_hidden", (rows, cfg.point_dim), torch.float32)
        self._decoder("global_points_decoder", hidden, N, P, gh, gw, global_hidden, context=context)
        if collect is not None:
            collect.update(point_hidden=point_hidden.view(N, P, -1).clone(), camera_hidden=camera_hidden.view(N, P, -1).clone(),
                           global_hidden=global_hidden.view(N, P, -1).clone())
        poses = self._camera_head(camera_hidden, N, P)
        nf = 3 * p * p
        feat_pts = self.buf.get("f32.rec.feat_pts", (rows, nf), torch.float32)
        local_points = torch.empty(N, Hh, Ww, 3, dtype=torch.float32, device=dev)
        points = torch.empty(N, Hh, Ww, 3, dtype=torch.float32, device=dev)
        global_points = torch.empty(N, Hh, Ww, 3, dtype=torch.float32, device=dev)
        self._lin(point_hidden, self.point_head_w3, self.point_head_b, feat_pts)
        ops.points_epilogue(feat_pts, poses, local_points, points, N, Hh, Ww, p, 1)
        self._lin(global_hidden, self.global_point_head_w3, self.global_point_head_b, feat_pts)
        ops.points_epilogue(feat_pts, None, global_points, None, N, Hh, Ww, p, 0)
        self._mark("heads_end")
        conf = None
        if cfg.train_conf_pi3:
            conf_hidden = self.buf.get("f32.rec.conf_hidden", (rows, cfg.point_dim), torch.float32)
            self._decoder("conf_decoder", hidden, N, P, gh, gw, conf_hidden)
            feat_conf = self.buf.get("f32.rec.feat_conf", (rows, p * p), torch.float32)
            self._lin(conf_hidden, self.conf_head_w3, self.conf_head_b, feat_conf)
            conf = torch.empty(N, Hh, Ww, 1, dtype=torch.float32, device=dev)
            ops.points_epilogue(feat_conf, None, conf, None, N, Hh, Ww, p, 2)
            conf = conf[None]
        if original_images is not None and original_images.dim() == 4:
            original_images = original_images.unsqueeze(0)
        return dict(points=points[None], local_points=local_points[None], conf=conf, camera_poses=poses[None],
                    global_points=global_points[None], images=original_images)

    # ------------------------------------------------------------------------------------------
    # paths that exist in bf16 mode only
    # ------------------------------------------------------------------------------------------
    def _bf16_only(self, *a, **k):
        raise NotImplementedError("this path is implemented in bf16 mode only; fp32 mode covers recon()")

    forward_cache_update_text = generate_text = language_model_forward_train = _bf16_only
    dino_forward_sharded = recon_view_sharded = vit_forward = forward_cache_update_vit = chat_with_recon = _bf16_only
