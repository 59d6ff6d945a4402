"""TEST INFRASTRUCTURE — generates tests/golden/* by running the UNMODIFIED reference under the CPU
harness (oracle/ref_harness.py).  Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

Fixtures (all produced by the reference's own classes, tiny random-init model, seeded inputs):
  state_dict_keys.json      name -> shape of G2VLM(...).state_dict() for the tiny and the full dims
  train_tiny.pt             Qwen2VLModel.forward_train output (bf16 module, two packed samples, nested masks)
  recon_tiny_{a,b}.pt       routing/permutation/position index tensors from the reference's
                            prepare_prompts_addbos / prepare_dino_images_pi3 (exact-match targets),
                            camera poses, strided samples of the point maps and of last_hidden.
Case a: 3 views 70x518 (non-square -> bicubic pos-embed resample; quirk-Q1 uncovered rows);
case b: 2 views 518x518 (native 37x37 grid).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from g2vlm_b200 import schema  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = {"a": dict(n=3, h=70, w=518, seed=1), "b": dict(n=2, h=518, w=518, seed=2)}
STRIDE_H, STRIDE_W, STRIDE_T = 5, 7, 13


def views_u8(n, h, w, seed):
    v = schema.synthetic_views(n, h, w, seed=seed)
    return (v * 255).round().to(torch.uint8)


def to_pil(u8):
    return [Image.fromarray(u8[i].permute(1, 2, 0).numpy()) for i in range(u8.shape[0])]


def run_case(model, case):
    u8 = views_u8(case["n"], case["h"], case["w"], case["seed"])
    pil = to_pil(u8)
    tok = rh.StubTokenizer()
    ids = dict(rh.NEW_TOKEN_IDS)
    gi_text, newlens, new_rope = model.prepare_prompts_addbos([0], [0], ["Reconstruct the 3D scene."], tok, ids)
    gi_dino, _, _ = model.prepare_dino_images_pi3(newlens, new_rope, pil, None, ids)
    captured = {}
    orig = model.reconstruct

    def spy(*a, **k):
        captured["last_hidden"] = k["selected_hidden_states"].detach().clone()
        return orig(*a, **k)

    model.reconstruct = spy
    try:
        pred = rh.run_reference_recon(model, pil)
    finally:
        model.reconstruct = orig
    out = dict(case=case, stride=(STRIDE_H, STRIDE_W, STRIDE_T))
    for k in ("packed_text_ids", "packed_text_position_ids", "packed_text_indexes", "text_token_lens"):
        out["text." + k] = gi_text[k].clone()
    for k in ("packed_text_ids", "packed_text_indexes", "dino_token_seqlens", "packed_dino_token_indexes",
              "packed_position_ids", "packed_seqlens", "packed_indexes", "packed_key_value_indexes",
              "key_values_lens"):
        out["dino." + k] = gi_dino[k].clone()
    out["camera_poses"] = pred["camera_poses"].float().clone()
    for k in ("points", "local_points", "global_points"):
        out[k] = pred[k].float()[:, :, ::STRIDE_H, ::STRIDE_W].clone()
        out[k + ".absmax"] = pred[k].float().abs().max()
    out["last_hidden"] = captured["last_hidden"].float()[::STRIDE_T].clone()
    out["last_hidden.absmax"] = captured["last_hidden"].float().abs().max()
    out["images_u8_checksum"] = int(u8.to(torch.int64).sum())
    return out


class ChatTokenizer:
    """Deterministic stand-in for the Qwen2 tokenizer on the chat path (tiny vocab)."""

    def encode(self, prompt, add_special_tokens=True):
        if "your text" in prompt:          # the template of prepare_start_tokens (g2vlm.py:1046)
            return [21, 22, 23]
        if "system" in prompt:
            return [31, 32, 33, 34]
        return [41, 42, 43, 44, 45]


CHAT_CASE = dict(n=2, h=28, w=518, seed=3, max_length=8)


def run_chat_case(model):
    """Text-only slice of chat_with_recon (g2vlm.py:1305-1410): system-prompt prefill (und, causal) -> geo
    step on the views with cache update -> question prefill on top of the cache -> greedy generate_text."""
    from modeling.g2vlm.qwen2vl import NaiveCache
    c = CHAT_CASE
    tok, ids = ChatTokenizer(), dict(rh.NEW_TOKEN_IDS)
    pil = to_pil(views_u8(c["n"], c["h"], c["w"], c["seed"]))
    with torch.no_grad(), torch.amp.autocast("cuda", dtype=torch.bfloat16):
        past = NaiveCache(model.config.llm_config.num_hidden_layers)
        gi, nl, nr = model.prepare_prompts_pure_text([0], [0], ["system prompt"], tok, ids)
        past = model.forward_cache_update_text(past, **gi)
        gi, nl, nr = model.prepare_dino_images_pi3(nl, nr, pil, None, ids)
        past, last = model.forward_cache_update_dino(past, **gi)
        gi, nl, nr = model.prepare_prompts_pure_text(nl, nr, ["question"], tok, ids)
        past = model.forward_cache_update_text(past, **gi)
        st = model.prepare_start_tokens(nl, nr, tok, ids)
        cache_len = past.key_cache[0].shape[0]
        k_probe = past.key_cache[1].float().clone()
        out = model.generate_text(past_key_values=past, max_length=c["max_length"], end_token_id=ids["eos_token_id"], **st)
    return dict(case=c, tokens=out[:, 0].clone(), start_token=st["packed_start_tokens"].clone(),
                start_position=st["packed_query_position_ids"][:, 0].clone(), cache_len_before_decode=cache_len,
                newlens=nl, key_cache_layer1=k_probe[::7].clone(), last_hidden=last.float()[::STRIDE_T].clone())


class ChatTokenizerFull(ChatTokenizer):
    def decode(self, ids):
        return " ".join(str(int(i)) for i in ids)


CHAT_VIT_CASE = dict(n=2, h=28, w=518, seed=4, max_length=6, vit_h=56, vit_w=84)


def run_chat_vit_case():
    """The reference's chat_with_recon end to end (g2vlm.py:1305-1410) on the tiny model WITH the Qwen2-VL ViT:
    system prompt -> geo step -> one ViT step per image -> question -> greedy decode."""
    from oracle.vit_stub import StubVitTransform
    c = CHAT_VIT_CASE
    model = rh.build_reference_model(rh.TINY, visual_und=True)
    sd = schema.init_synthetic(schema.TINY_CHAT, seed=0)
    msg = model.load_state_dict(sd, strict=False)
    assert not msg.unexpected_keys and msg.missing_keys == ["dino_model.embeddings.mask_token"], msg
    pil = to_pil(views_u8(c["n"], c["h"], c["w"], c["seed"]))
    captured = {}
    orig = model.generate_text

    def spy(*a, **k):
        captured["cache_len"] = k["past_key_values"].key_cache[0].shape[0]
        captured["key_cache_layer1"] = k["past_key_values"].key_cache[1].float()[::7].clone()
        out = orig(*a, **k)
        captured["ids"] = out.clone()
        return out

    model.generate_text = spy
    with torch.no_grad():
        text = model.chat_with_recon(ChatTokenizerFull(), dict(rh.NEW_TOKEN_IDS), StubVitTransform(c["vit_h"], c["vit_w"]),
                                     None, pil, "question", c["max_length"])
    return dict(case=c, text=text, tokens=captured["ids"][:, 0].clone(), cache_len_before_decode=captured["cache_len"],
                key_cache_layer1=captured["key_cache_layer1"])


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.manual_seed(0)
    model = rh.build_reference_model(rh.TINY, visual_und=False)
    keys = {"tiny": {k: list(v.shape) for k, v in model.state_dict().items()}}
    sd = schema.init_synthetic(schema.TINY, seed=0)
    msg = model.load_state_dict(sd, strict=False)
    assert not msg.unexpected_keys, msg.unexpected_keys
    for name, case in CASES.items():
        out = run_case(model, case)
        torch.save(out, os.path.join(GOLDEN, f"recon_tiny_{name}.pt"))
        print("wrote", name, {k: tuple(v.shape) for k, v in out.items() if torch.is_tensor(v) and v.dim() > 0})
    chat = run_chat_case(model)
    torch.save(chat, os.path.join(GOLDEN, "chat_tiny.pt"))
    print("wrote chat", chat["tokens"].tolist(), chat["cache_len_before_decode"])
    chatv = run_chat_vit_case()
    torch.save(chatv, os.path.join(GOLDEN, "chat_vit_tiny.pt"))
    print("wrote chat+vit", chatv["tokens"].tolist(), chatv["cache_len_before_decode"], repr(chatv["text"]))
    # training forward of the MoT stack (row f.4 groundwork): LAST, because it converts the language model to bf16
    x, pos, geo, und = rh.train_case_inputs(schema.TINY.hidden_size)
    y = rh.run_reference_lm_forward_train(model, x, pos, geo, und)
    torch.save(dict(case=rh.TRAIN_CASE, y=y.clone()), os.path.join(GOLDEN, "train_tiny.pt"))
    print("wrote train_tiny", tuple(y.shape), float(y.float().abs().max()))
    keys["tiny_chat"] = {k: list(v.shape) for k, v in rh.build_reference_model(rh.TINY, visual_und=True).state_dict().items()}
    # full-size key schema: build on the meta device (no 18 GB allocation)
    with torch.device("meta"):
        full = rh.build_reference_model(rh.FULL, visual_und=False)
    keys["full"] = {k: list(v.shape) for k, v in full.state_dict().items()}
    with open(os.path.join(GOLDEN, "state_dict_keys.json"), "w") as f:
        json.dump(keys, f, indent=0, sort_keys=True)
    print("wrote state_dict_keys.json", len(keys["tiny"]), len(keys["full"]))


if __name__ == "__main__":
    main()
