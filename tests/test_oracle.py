"""CPU tests of the oracle: the restatement (oracle/restate.py) against (1) golden fixtures produced
by the unmodified reference (oracle/make_golden.py -> tests/golden) and (2) the live reference when
/root/reference is present (build container).  Also pins the state_dict key schema."""
import json
import os

import pytest
import torch

from g2vlm_b200 import schema
from oracle import ref_harness, restate

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL_BF16 = 2e-2  # BASELINE.json: max rel err <= 2e-2 in bf16 mode


def _views(case):
    v = schema.synthetic_views(case["n"], case["h"], case["w"], seed=case["seed"])
    return (v * 255).round() / 255.0  # the reference reads 8-bit images (PIL -> ToTensor)


def _maxrel(a, b, scale=None):
    scale = b.abs().max() if scale is None else scale
    return ((a - b).abs().max() / scale).item()


@pytest.fixture(scope="module")
def tiny_sd():
    return schema.init_synthetic(schema.TINY, seed=0)


@pytest.mark.parametrize("name", ["a", "b"])
def test_indices_match_reference_exactly(name):
    g = torch.load(os.path.join(GOLDEN, f"recon_tiny_{name}.pt"))
    gi, newlens, new_rope = restate.prepare_prompts_addbos([11, 12, 13, 14, 15, 16], 1)
    for k in ("packed_text_ids", "packed_text_position_ids", "packed_text_indexes", "text_token_lens"):
        assert torch.equal(gi[k], g["text." + k]), k
    gd, _, _ = restate.prepare_dino_images(_views(g["case"]), newlens[0], new_rope[0], 3, 4)
    for k in ("packed_text_ids", "packed_text_indexes", "dino_token_seqlens", "packed_dino_token_indexes",
              "packed_position_ids", "packed_seqlens", "packed_indexes", "packed_key_value_indexes",
              "key_values_lens"):
        assert gd[k].dtype == g["dino." + k].dtype, k
        assert torch.equal(gd[k], g["dino." + k]), k


@pytest.mark.parametrize("name", ["a", "b"])
def test_restatement_matches_reference_golden(name, tiny_sd):
    g = torch.load(os.path.join(GOLDEN, f"recon_tiny_{name}.pt"))
    sh, sw, st = g["stride"]
    collect = {}
    out = restate.recon(tiny_sd, schema.TINY, _views(g["case"]), mode="bf16", collect=collect)
    assert _maxrel(collect["last_hidden"][::st], g["last_hidden"], g["last_hidden.absmax"]) < TOL_BF16
    for k in ("points", "local_points", "global_points"):
        assert _maxrel(out[k][:, :, ::sh, ::sw], g[k], g[k + ".absmax"]) < TOL_BF16, k
    assert _maxrel(out["camera_poses"], g["camera_poses"]) < TOL_BF16
    assert out["conf"] is None


def test_schema_matches_reference_state_dict_keys():
    keys = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))
    for cfg, name in ((schema.TINY, "tiny"), (schema.FULL, "full")):
        ours = {k: list(v) for k, v in schema.state_dict_schema(cfg).items()}
        ref = {k: v for k, v in keys[name].items()
               if not k.startswith("vit_model.") and k not in ("dino_model.embeddings.mask_token",)}
        assert ours == ref


@pytest.mark.skipif(not ref_harness.available(), reason="/root/reference not present (GPU box)")
def test_restatement_matches_live_reference(tiny_sd):
    """Fresh inputs (not the golden ones) through the unmodified reference classes."""
    from oracle.make_golden import to_pil, views_u8
    model = ref_harness.build_reference_model(ref_harness.TINY, visual_und=False)
    model.load_state_dict(tiny_sd, strict=False)
    u8 = views_u8(2, 98, 518, seed=11)
    ref = ref_harness.run_reference_recon(model, to_pil(u8))
    out = restate.recon(tiny_sd, schema.TINY, u8.float() / 255.0, mode="bf16")
    for k in ("points", "local_points", "global_points", "camera_poses"):
        assert _maxrel(out[k], ref[k].float()) < TOL_BF16, k
    assert torch.equal(ref["images"][0], u8.float() / 255.0)


def test_fp32_mode_close_to_bf16_mode(tiny_sd):
    v = _views(dict(n=2, h=42, w=518, seed=5))
    a = restate.recon(tiny_sd, schema.TINY, v, mode="fp32")
    b = restate.recon(tiny_sd, schema.TINY, v, mode="bf16")
    for k in ("points", "global_points", "camera_poses"):
        assert _maxrel(b[k], a[k]) < 3e-2, k


def _oracle_chat(sd, cfg, case):
    """Text-only slice of chat_with_recon through the restatement; returns (tokens, last_hidden, cache)."""
    v = _views(case)
    W = restate._W(sd)
    emb = W("language_model.model.embed_tokens.weight")
    _, cache = restate.lm_forward_und(sd, cfg, emb[torch.tensor([31, 32, 33, 34])], torch.arange(4).expand(3, -1),
                                      None, True, "bf16")
    gi, nl, nr = restate.prepare_dino_images(v, 4, 4, 3, 4)
    geo, und = gi["packed_dino_token_indexes"], gi["packed_text_indexes"]
    T = int(gi["packed_seqlens"][0])
    x = torch.zeros(T, cfg.hidden_size)
    x[und] = emb[gi["packed_text_ids"]]
    d = restate.dino_forward(sd, cfg, gi["packed_dino_images"], gi["dino_token_seqlens"], "bf16")
    x[geo] = restate.linear(d.reshape(-1, d.shape[-1]), W("dino2llm.weight"), W("dino2llm.bias"), "bf16")
    last, cache = restate.lm_forward_geo(sd, cfg, x, gi["packed_position_ids"], geo, und, cache, "bf16")
    p0 = nr[0]
    _, cache = restate.lm_forward_und(sd, cfg, emb[torch.tensor([41, 42, 43, 44, 45])],
                                      (p0 + torch.arange(5)).expand(3, -1), cache, True, "bf16")
    toks, logits, cache2 = restate.generate_text_greedy(sd, cfg, cache, 23, p0 + 5, case["max_length"], 2)
    return toks, logits, last, cache


def test_chat_restatement_reproduces_reference_tokens(tiny_sd):
    """Row f1: cached und prefill + geo step with cache update + greedy decode == the reference's
    generate_text output (token ids are integers: exact)."""
    g = torch.load(os.path.join(GOLDEN, "chat_tiny.pt"))
    toks, logits, last, cache = _oracle_chat(tiny_sd, schema.TINY, g["case"])
    assert toks == g["tokens"].tolist()
    assert cache[0][0].shape[0] == g["cache_len_before_decode"] == g["newlens"][0]
    assert int(g["start_position"][0]) == 4 + (28 // 14 and 2) * (max(2, 37) + 2) + 5
    assert _maxrel(cache[1][0][::7], g["key_cache_layer1"]) < TOL_BF16
    assert _maxrel(last[::13], g["last_hidden"]) < TOL_BF16


def _oracle_chat_vit(sd, cfg, case):
    """chat_with_recon through the restatement (system prompt -> geo -> ViT per image -> question -> decode)."""
    from oracle.vit_stub import StubVitTransform
    from oracle.make_golden import to_pil, views_u8
    u8 = views_u8(case["n"], case["h"], case["w"], case["seed"])
    pil = to_pil(u8)
    v = u8.float() / 255.0
    W = restate._W(sd)
    emb = W("language_model.model.embed_tokens.weight")
    _, cache = restate.lm_forward_und(sd, cfg, emb[torch.tensor([31, 32, 33, 34])], torch.arange(4).expand(3, -1), None, True, "bf16")
    gi, nl, nr = restate.prepare_dino_images(v, 4, 4, 3, 4)
    geo, und = gi["packed_dino_token_indexes"], gi["packed_text_indexes"]
    x = torch.zeros(int(gi["packed_seqlens"][0]), cfg.hidden_size)
    x[und] = emb[gi["packed_text_ids"]]
    d = restate.dino_forward(sd, cfg, gi["packed_dino_images"], gi["dino_token_seqlens"], "bf16")
    x[geo] = restate.linear(d.reshape(-1, d.shape[-1]), W("dino2llm.weight"), W("dino2llm.bias"), "bf16")
    _, cache = restate.lm_forward_geo(sd, cfg, x, gi["packed_position_ids"], geo, und, cache, "bf16")
    kvlen, rope = nl[0], nr[0]
    tf = StubVitTransform(case["vit_h"], case["vit_w"])
    for im in pil:
        pix, grid = tf([im])
        cache, kvlen, rope = restate.vit_step(sd, cfg, pix, grid, cache, kvlen, rope, 3, 4, "bf16")
    _, cache = restate.lm_forward_und(sd, cfg, emb[torch.tensor([41, 42, 43, 44, 45])], (rope + torch.arange(5)).expand(3, -1),
                                      cache, True, "bf16")
    toks, logits, _ = restate.generate_text_greedy(sd, cfg, cache, 23, rope + 5, case["max_length"], 2)
    return toks, logits, cache


def test_chat_with_vit_restatement_reproduces_reference():
    """Rows f1 + f2: the whole chat_with_recon flow incl. the Qwen2-VL ViT step."""
    g = torch.load(os.path.join(GOLDEN, "chat_vit_tiny.pt"))
    sd = schema.init_synthetic(schema.TINY_CHAT, seed=0)
    toks, _, cache = _oracle_chat_vit(sd, schema.TINY_CHAT, g["case"])
    assert toks == g["tokens"].tolist()
    assert cache[0][0].shape[0] == g["cache_len_before_decode"]
    assert _maxrel(cache[1][0][::7], g["key_cache_layer1"]) < TOL_BF16


def test_schema_with_vit_matches_reference_keys():
    keys = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))
    ours = {k: list(v) for k, v in schema.state_dict_schema(schema.TINY_CHAT).items()}
    ref = {k: v for k, v in keys["tiny_chat"].items() if k != "dino_model.embeddings.mask_token"}
    assert ours == ref


def test_training_forward_restatement_matches_reference(tiny_sd):
    """Row f.4 groundwork: restate.lm_forward_train (MoT stack in training layout: two packed samples, causal text
    splits and full image splits, bf16 module as under FSDP mixed precision) against the output of the unmodified
    Qwen2VLModel.forward_train (tests/golden/train_tiny.pt): within one bf16 ulp of the largest activation."""
    g = torch.load(os.path.join(GOLDEN, "train_tiny.pt"))
    case = g["case"]
    x, pos, geo, und = ref_harness.train_case_inputs(schema.TINY.hidden_size, case)
    sd = {k: v.to(torch.bfloat16).float() for k, v in tiny_sd.items()}     # the reference module is cast to bf16
    samples = case["samples"]
    y = restate.lm_forward_train(sd, schema.TINY, x.float(), pos, geo, und, [sum(s) for s, _ in samples],
                                 [s for s, _ in samples], [m for _, m in samples])
    ref = g["y"].float()
    assert ((y - ref).abs().max() / ref.abs().max()).item() < 8e-3
    # mask semantics (data/data_utils.py:205-239)
    m = restate.train_attention_mask([2, 3, 2], ["causal", "full", "noise"])
    assert m[:2, :2].tolist() == [[True, False], [True, True]]
    assert bool(m[2:5, :5].all()) and not bool(m[2:5, 5:].any())
    assert not bool(m[:5, 5:].any()) and bool(m[5:, 5:].all()) and bool(m[5:, :5].any()) is True


def test_chunked_attention_equals_one_piece():
    """attention_segments processes the query rows in chunks (the 16-view scene's score matrix is 23 GB in one piece):
    exact, including the bottom-right aligned causal mask and GQA."""
    import math
    from oracle import restate
    g = torch.Generator().manual_seed(0)
    q = torch.randn(37, 4, 16, generator=g)
    k = torch.randn(50, 2, 16, generator=g)
    v = torch.randn(50, 2, 16, generator=g)
    for causal in (False, True):
        whole = restate.attention_segments(q, k, v, [0, 20, 37], [0, 30, 50], 1 / math.sqrt(16), causal, "fp32")
        keep = restate.ATTN_CHUNK_ELEMS
        restate.ATTN_CHUNK_ELEMS = 4 * 30 * 3          # 3 query rows per chunk
        try:
            parts = restate.attention_segments(q, k, v, [0, 20, 37], [0, 30, 50], 1 / math.sqrt(16), causal, "fp32")
        finally:
            restate.ATTN_CHUNK_ELEMS = keep
        assert torch.allclose(whole, parts, atol=1e-6)
        # against torch SDPA with an explicit mask (segment 0)
        qs, ks, vs = q[:20].transpose(0, 1), k[:30].transpose(0, 1).repeat_interleave(2, 0), v[:30].transpose(0, 1).repeat_interleave(2, 0)
        mask = torch.ones(20, 30, dtype=torch.bool).tril(diagonal=10) if causal else None
        ref = torch.nn.functional.scaled_dot_product_attention(qs[None], ks[None], vs[None], attn_mask=mask)[0].transpose(0, 1)
        assert torch.allclose(whole[:20], ref, atol=1e-5)


def test_reference_staging_recipe_and_config_bridge():
    """oracle/stage_ref.py copies the reference byte for byte (sha256 manifest) and G2Config.from_reference reads the
    reference's own config objects."""
    from oracle import ref_harness as rh
    from oracle import stage_ref
    if not os.path.isdir(os.path.join(stage_ref.SRC, "modeling", "g2vlm")):
        pytest.skip("no /root/reference in this environment")
    dst = stage_ref.stage()
    assert dst and stage_ref.verify() > 50
    import hashlib
    rel = "modeling/g2vlm/g2vlm.py"
    assert hashlib.sha256(open(os.path.join(stage_ref.SRC, rel), "rb").read()).hexdigest() == \
        hashlib.sha256(open(os.path.join(dst, rel), "rb").read()).hexdigest()
    model = rh.build_reference_model(rh.TINY, visual_und=False, skip_init=True)
    cfg = schema.G2Config.from_reference(model.config.llm_config, model.config.dino_config)
    assert cfg == schema.TINY
    assert rh.dims_from_cfg(schema.TINY)["llm"]["hidden_size"] == 256
