"""Model dimensions and the reference ``state_dict`` key schema of the recon path.

The product loads the reference's checkpoint keys unchanged (names confirmed by
dumping ``G2VLM(...).state_dict()`` of the reference classes — see tests/golden/state_dict_keys.json).
The JSON configs of the released checkpoint are not in the reference repo, so dimensions are inputs
(``G2Config``), with the two configurations used by tests and the benchmark predefined below.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, Tuple

import torch


@dataclass(frozen=True)
class G2Config:
    # Qwen2-VL MoT language model (modeling/g2vlm/qwen2vl.py:50-234)
    hidden_size: int = 1536
    num_layers: int = 28
    num_heads: int = 12
    num_kv_heads: int = 2
    intermediate_size: int = 8960
    vocab_size: int = 151936
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1000000.0
    mrope_section: Tuple[int, int, int] = (16, 24, 24)  # hard-coded, modeling_qwen2_vl.py:562-566
    # DINOv2-with-registers encoder
    dino_hidden: int = 1024
    dino_layers: int = 24
    dino_heads: int = 16
    dino_mlp_ratio: int = 4
    dino_image_size: int = 518
    dino_patch: int = 14
    dino_registers: int = 4
    dino_ln_eps: float = 1e-6
    # Pi3 heads (modeling/g2vlm/g2vlm.py:162-203, transformer_head.py:9-56)
    dec_depth: int = 5
    dec_heads: int = 16
    dec_mlp_ratio: int = 4
    point_dim: int = 1024
    camera_dim: int = 512
    rope2d_base: float = 100.0  # 'rope100', g2vlm.py:152
    train_conf_pi3: bool = False
    # Qwen2-VL ViT of the chat path (row f2; vit_depth = 0: recon-only model, no vit_model.* keys needed)
    vit_depth: int = 0
    vit_embed_dim: int = 1280
    vit_heads: int = 16
    vit_mlp_ratio: int = 4
    vit_patch: int = 14
    vit_merge: int = 2
    vit_temporal: int = 2

    @classmethod
    def from_reference(cls, llm_config, dino_config, vit_config=None, **overrides) -> "G2Config":
        """Dimensions from the reference's own config objects (`Qwen2VLConfig`, `Dinov2WithRegistersConfig`,
        optionally `Qwen2VLVisionConfig`, as built in g2vlm_utils.py:32-41) — duck-typed, nothing is imported."""
        kw = dict(hidden_size=llm_config.hidden_size, num_layers=llm_config.num_hidden_layers,
                  num_heads=llm_config.num_attention_heads, num_kv_heads=llm_config.num_key_value_heads,
                  intermediate_size=llm_config.intermediate_size, vocab_size=llm_config.vocab_size,
                  rms_norm_eps=llm_config.rms_norm_eps, rope_theta=llm_config.rope_theta,
                  dino_hidden=dino_config.hidden_size, dino_layers=dino_config.num_hidden_layers,
                  dino_heads=dino_config.num_attention_heads, dino_mlp_ratio=dino_config.mlp_ratio,
                  dino_image_size=dino_config.image_size, dino_patch=dino_config.patch_size,
                  dino_registers=dino_config.num_register_tokens, dino_ln_eps=dino_config.layer_norm_eps)
        if vit_config is not None:
            kw.update(vit_depth=vit_config.depth, vit_embed_dim=vit_config.embed_dim, vit_heads=vit_config.num_heads,
                      vit_mlp_ratio=int(vit_config.mlp_ratio), vit_patch=vit_config.patch_size,
                      vit_merge=vit_config.spatial_merge_size, vit_temporal=vit_config.temporal_patch_size)
        kw.update(overrides)
        return cls(**kw)

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_heads

    @property
    def dino_head_dim(self) -> int:
        return self.dino_hidden // self.dino_heads

    @property
    def dec_head_dim(self) -> int:
        return self.hidden_size // self.dec_heads

    @property
    def dino_grid(self) -> int:
        return self.dino_image_size // self.dino_patch


FULL = G2Config()
TINY = G2Config(hidden_size=256, num_layers=2, num_heads=2, num_kv_heads=1, intermediate_size=512,
                vocab_size=512, dino_hidden=64, dino_layers=2, dino_heads=2)
# the same models with the Qwen2-VL ViT (chat path)
FULL_CHAT = G2Config(vit_depth=32)
TINY_CHAT = G2Config(hidden_size=256, num_layers=2, num_heads=2, num_kv_heads=1, intermediate_size=512,
                     vocab_size=512, dino_hidden=64, dino_layers=2, dino_heads=2, vit_depth=1, vit_embed_dim=64,
                     vit_heads=2, vit_mlp_ratio=2)


def state_dict_schema(cfg: G2Config) -> "OrderedDict[str, Tuple[int, ...]]":
    """name -> shape of every parameter the recon path reads, in the reference's registration order."""
    H, I, hd = cfg.hidden_size, cfg.intermediate_size, cfg.head_dim
    nq, nkv = cfg.num_heads * hd, cfg.num_kv_heads * hd
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    lm = "language_model.model."
    s[lm + "embed_tokens.weight"] = (cfg.vocab_size, H)
    for i in range(cfg.num_layers):
        p = f"{lm}layers.{i}."
        s[p + "ls1.gamma"] = (H,)
        s[p + "ls2.gamma"] = (H,)
        for sfx in ("", "_moe_geo"):
            if sfx == "":
                s[p + "self_attn.q_proj.weight"] = (nq, H); s[p + "self_attn.q_proj.bias"] = (nq,)
                s[p + "self_attn.k_proj.weight"] = (nkv, H); s[p + "self_attn.k_proj.bias"] = (nkv,)
                s[p + "self_attn.v_proj.weight"] = (nkv, H); s[p + "self_attn.v_proj.bias"] = (nkv,)
                s[p + "self_attn.o_proj.weight"] = (H, nq)
                s[p + "self_attn.q_norm.weight"] = (hd,); s[p + "self_attn.k_norm.weight"] = (hd,)
                s[p + "self_attn.q_norm_moe_geo.weight"] = (hd,); s[p + "self_attn.k_norm_moe_geo.weight"] = (hd,)
            else:
                s[p + "self_attn.q_proj_moe_geo.weight"] = (nq, H); s[p + "self_attn.q_proj_moe_geo.bias"] = (nq,)
                s[p + "self_attn.k_proj_moe_geo.weight"] = (nkv, H); s[p + "self_attn.k_proj_moe_geo.bias"] = (nkv,)
                s[p + "self_attn.v_proj_moe_geo.weight"] = (nkv, H); s[p + "self_attn.v_proj_moe_geo.bias"] = (nkv,)
                s[p + "self_attn.o_proj_moe_geo.weight"] = (H, nq)
        for m in ("mlp", "mlp_moe_geo"):
            s[p + f"{m}.gate_proj.weight"] = (I, H)
            s[p + f"{m}.up_proj.weight"] = (I, H)
            s[p + f"{m}.down_proj.weight"] = (H, I)
        for n in ("input_layernorm", "input_layernorm_moe_geo", "post_attention_layernorm",
                  "post_attention_layernorm_moe_geo"):
            s[p + n + ".weight"] = (H,)
    s[lm + "norm.weight"] = (H,)
    s[lm + "norm_moe_geo.weight"] = (H,)
    s["language_model.lm_head.weight"] = (cfg.vocab_size, H)  # only read by generate_text (chat path)

    D, g = cfg.dino_hidden, cfg.dino_grid
    d = "dino_model."
    s[d + "embeddings.cls_token"] = (1, 1, D)
    s[d + "embeddings.register_tokens"] = (1, cfg.dino_registers, D)
    s[d + "embeddings.position_embeddings"] = (1, g * g + 1, D)
    s[d + "embeddings.patch_embeddings.projection.weight"] = (D, 3, cfg.dino_patch, cfg.dino_patch)
    s[d + "embeddings.patch_embeddings.projection.bias"] = (D,)
    for i in range(cfg.dino_layers):
        p = f"{d}encoder.layer.{i}."
        s[p + "norm1.weight"] = (D,); s[p + "norm1.bias"] = (D,)
        for n in ("query", "key", "value"):
            s[p + f"attention.attention.{n}.weight"] = (D, D); s[p + f"attention.attention.{n}.bias"] = (D,)
        s[p + "attention.output.dense.weight"] = (D, D); s[p + "attention.output.dense.bias"] = (D,)
        s[p + "layer_scale1.lambda1"] = (D,)
        s[p + "norm2.weight"] = (D,); s[p + "norm2.bias"] = (D,)
        s[p + "mlp.fc1.weight"] = (D * cfg.dino_mlp_ratio, D); s[p + "mlp.fc1.bias"] = (D * cfg.dino_mlp_ratio,)
        s[p + "mlp.fc2.weight"] = (D, D * cfg.dino_mlp_ratio); s[p + "mlp.fc2.bias"] = (D,)
        s[p + "layer_scale2.lambda1"] = (D,)
    s[d + "layernorm.weight"] = (D,); s[d + "layernorm.bias"] = (D,)
    s["dino2llm.weight"] = (H, D); s["dino2llm.bias"] = (H,)

    F = H * cfg.dec_mlp_ratio

    def block(p, cross):
        s[p + "norm1.weight"] = (H,); s[p + "norm1.bias"] = (H,)
        s[p + "attn.qkv.weight"] = (3 * H, H); s[p + "attn.qkv.bias"] = (3 * H,)
        s[p + "attn.proj.weight"] = (H, H); s[p + "attn.proj.bias"] = (H,)
        s[p + "norm2.weight"] = (H,); s[p + "norm2.bias"] = (H,)
        if cross:
            s[p + "norm_y.weight"] = (H,); s[p + "norm_y.bias"] = (H,)
            for n in ("q_proj", "k_proj", "v_proj", "proj"):
                s[p + f"cross_attn.{n}.weight"] = (H, H); s[p + f"cross_attn.{n}.bias"] = (H,)
            s[p + "norm3.weight"] = (H,); s[p + "norm3.bias"] = (H,)
        s[p + "mlp.fc1.weight"] = (F, H); s[p + "mlp.fc1.bias"] = (F,)
        s[p + "mlp.fc2.weight"] = (H, F); s[p + "mlp.fc2.bias"] = (H,)

    def decoder(name, out_dim, cross=False):
        for i in range(cfg.dec_depth):
            block(f"{name}.blocks.{i}.", cross)
        s[f"{name}.linear_out.weight"] = (out_dim, H); s[f"{name}.linear_out.bias"] = (out_dim,)

    ps2 = cfg.dino_patch ** 2
    decoder("point_decoder", cfg.point_dim)
    s["point_head.proj.weight"] = (3 * ps2, cfg.point_dim); s["point_head.proj.bias"] = (3 * ps2,)
    decoder("camera_decoder", cfg.camera_dim)
    C = cfg.camera_dim
    for i in range(2):
        for j in (1, 2, 3):
            s[f"camera_head.res_conv.{i}.res_conv{j}.weight"] = (C, C)
            s[f"camera_head.res_conv.{i}.res_conv{j}.bias"] = (C,)
    for j in (0, 2):
        s[f"camera_head.more_mlps.{j}.weight"] = (C, C); s[f"camera_head.more_mlps.{j}.bias"] = (C,)
    s["camera_head.fc_t.weight"] = (3, C); s["camera_head.fc_t.bias"] = (3,)
    s["camera_head.fc_rot.weight"] = (9, C); s["camera_head.fc_rot.bias"] = (9,)
    decoder("global_points_decoder", cfg.point_dim, cross=True)
    s["global_point_head.proj.weight"] = (3 * ps2, cfg.point_dim); s["global_point_head.proj.bias"] = (3 * ps2,)
    if cfg.vit_depth > 0:
        E, Fv = cfg.vit_embed_dim, cfg.vit_embed_dim * cfg.vit_mlp_ratio
        v = "vit_model."
        s[v + "patch_embed.proj.weight"] = (E, 3, cfg.vit_temporal, cfg.vit_patch, cfg.vit_patch)
        for i in range(cfg.vit_depth):
            p = f"{v}blocks.{i}."
            s[p + "norm1.weight"] = (E,); s[p + "norm1.bias"] = (E,)
            s[p + "norm2.weight"] = (E,); s[p + "norm2.bias"] = (E,)
            s[p + "attn.qkv.weight"] = (3 * E, E); s[p + "attn.qkv.bias"] = (3 * E,)
            s[p + "attn.proj.weight"] = (E, E); s[p + "attn.proj.bias"] = (E,)
            s[p + "mlp.fc1.weight"] = (Fv, E); s[p + "mlp.fc1.bias"] = (Fv,)
            s[p + "mlp.fc2.weight"] = (E, Fv); s[p + "mlp.fc2.bias"] = (E,)
        Em = E * cfg.vit_merge ** 2
        s[v + "merger.ln_q.weight"] = (E,); s[v + "merger.ln_q.bias"] = (E,)
        s[v + "merger.mlp.0.weight"] = (Em, Em); s[v + "merger.mlp.0.bias"] = (Em,)
        s[v + "merger.mlp.2.weight"] = (H, Em); s[v + "merger.mlp.2.bias"] = (H,)
    if cfg.train_conf_pi3:
        decoder("conf_decoder", cfg.point_dim)
        s["conf_head.proj.weight"] = (ps2, cfg.point_dim); s["conf_head.proj.bias"] = (ps2,)
    return s


def init_synthetic(cfg: G2Config, seed: int = 0, embed_rows: int | None = None,
                   device: str = "cpu") -> Dict[str, torch.Tensor]:
    """Seeded synthetic weights: Linear/Conv/Embedding/bias/token tensors ~ N(0, 0.02^2),
    norm weights and LayerScale/lambda ~ U(0.5, 1.5); geo and und experts independent.  Each tensor has
    its own generator keyed by (seed, index) so the result does not depend on generation order.
    ``embed_rows`` truncates the embedding table (the benchmark only needs the few ids recon uses).
    ``device="cuda"`` generates on the GPU (fast for the 3.4 B-parameter full model; a different but
    equally seeded stream than the CPU generator — copy the result to the host when the oracle needs
    the same weights)."""
    sd: Dict[str, torch.Tensor] = {}
    for idx, (name, shape) in enumerate(state_dict_schema(cfg).items()):
        if (name.endswith("embed_tokens.weight") or name.endswith("lm_head.weight")) and embed_rows is not None:
            shape = (embed_rows, shape[1])
        g = torch.Generator(device=device).manual_seed(seed * 1000003 + idx)
        leaf = name.rsplit(".", 2)
        is_norm_w = name.endswith(".weight") and ("norm" in leaf[-2] or leaf[-2] in ("layernorm", "ln_q"))
        if is_norm_w or name.endswith(".gamma") or name.endswith(".lambda1"):
            t = torch.rand(shape, generator=g, device=device) + 0.5
        else:
            t = torch.randn(shape, generator=g, device=device) * 0.02
        sd[name] = t
    return sd


def synthetic_views(n_views: int, height: int, width: int, seed: int = 1) -> torch.Tensor:
    """(N,3,H,W) in [0,1]: uniform noise smoothed by a 7x7 box blur."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n_views, 3, height, width, generator=g)
    k = torch.ones(3, 1, 7, 7) / 49.0
    x = torch.nn.functional.conv2d(torch.nn.functional.pad(x, (3, 3, 3, 3), mode="replicate"), k, groups=3)
    return x.clamp_(0, 1)
