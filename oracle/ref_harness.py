"""TEST INFRASTRUCTURE — harness that runs the UNMODIFIED reference classes: on the host cores (SDPA stand-in for
flash-attn), or on a GPU with the reference's own flash-attn calls (`build_reference_model(device="cuda")`).

Only `tests/`, `oracle/make_golden.py`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module; the product package `g2vlm_b200` never does.
`/root/reference` exists only in the build container; `oracle/stage_ref.py` (run by `__graft_entry__.build()`)
places a byte-identical, git-ignored copy under `oracle/_ref/`, which travels to the GPU box with the snapshot.
Uses: (a) pin the restatement in `oracle/restate.py` against the real reference classes, (b) generate the golden
fixtures under `tests/golden/` (script: `oracle/make_golden.py`), (c) on the B200: bf16 parity of the CUDA path
against the real classes with real flash-attn and the PyTorch+FA2 speed baseline (`bench.py` `gpu_reference`),
(d) `bench.py --impl reference`: the real classes timed on the host cores.

The reference cannot run unmodified on CPU (hard bf16 casts + flash-attn + `.cuda()` + HF-hub
downloads + transformers-4.49 symbols); the shims below are in-process only and
edit nothing under /root/reference:
  1. stub `easydict`; inject three transformers-4.49 symbols removed in 5.x; register
     ROPE_INIT_FUNCTIONS['default'];
  2. redirect `torch.amp.autocast('cuda')` to 'cpu' so the nested enable/disable regions
     (g2vlm.py:1190-1226, camera_head.py:59) behave as on a GPU;
  3. replace `flash_attn_varlen_func` (g2vlm/qwen2vl.py:643, dinov2_model.py:49) by a per-segment
     SDPA with GQA, bottom-right causal mask and ZEROS for rows outside every segment (flash-attn
     leaves them uninitialised — quirk Q1);
  4. build configs in code (the HF JSON files are not in the repo) and a stub tokenizer.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root() -> str:
    """/root/reference in the build container; on the GPU box the byte-identical copy staged under oracle/_ref by
    `oracle/stage_ref.py` (git-ignored, travels with the gpurun snapshot)."""
    cands = [os.environ.get("G2VLM_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "modeling", "g2vlm")):
            return c
    return cands[1]


REFERENCE_ROOT = _find_root()
# autocast('cuda') is redirected to 'cpu' only while the reference runs on the host (set by build_reference_model)
_REDIRECT_AUTOCAST = True


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "modeling", "g2vlm"))


def gpu_available() -> bool:
    """The reference can run on the GPU with its own flash-attn calls (needs the flash_attn wheel + a CUDA device)."""
    if not (available() and torch.cuda.is_available()):
        return False
    try:
        import flash_attn  # noqa: F401
        return True
    except Exception:
        return False


# ---- shim 3: flash_attn_varlen_func stand-in ----------------------------------------------------
def varlen_sdpa(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q=None, max_seqlen_k=None,
                causal=False, **kw):
    """q (Tq,Hq,D), k/v (Tk,Hk,D) -> (Tq,Hq,D); rows outside every segment are zero."""
    out = torch.zeros_like(q)
    hq, hk = q.shape[1], k.shape[1]
    cq = [int(x) for x in cu_seqlens_q]
    ck = [int(x) for x in cu_seqlens_k]
    for i in range(len(cq) - 1):
        qs = q[cq[i]:cq[i + 1]].transpose(0, 1)  # (Hq, Lq, D)
        ks = k[ck[i]:ck[i + 1]].transpose(0, 1)
        vs = v[ck[i]:ck[i + 1]].transpose(0, 1)
        if hq != hk:
            ks = ks.repeat_interleave(hq // hk, dim=0)
            vs = vs.repeat_interleave(hq // hk, dim=0)
        lq, lk = qs.shape[1], ks.shape[1]
        mask = None
        if causal:
            mask = torch.ones(lq, lk, dtype=torch.bool).tril(diagonal=lk - lq)
        o = torch.nn.functional.scaled_dot_product_attention(qs[None], ks[None], vs[None], attn_mask=mask)[0]
        out[cq[i]:cq[i + 1]] = o.transpose(0, 1)
    return out


_installed = False


def install_shims() -> None:
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # the tree is read-only
    for p in (REFERENCE_ROOT, os.path.join(REFERENCE_ROOT, "modeling")):
        if p not in sys.path:
            sys.path.insert(0, p)

    # 1. missing third-party modules / symbols
    if "easydict" not in sys.modules:
        m = types.ModuleType("easydict")
        m.EasyDict = dict
        sys.modules["easydict"] = m
    try:  # the real wheel imports fine without a GPU; its functions are replaced after import
        import flash_attn  # noqa: F401
    except Exception:
        from importlib.machinery import ModuleSpec
        fa = types.ModuleType("flash_attn")
        fa.__spec__ = ModuleSpec("flash_attn", None)
        fa.flash_attn_varlen_func = varlen_sdpa
        fa.flash_attn_func = None
        sys.modules["flash_attn"] = fa

    class _Dummy:  # placeholder classes only referenced in isinstance / annotations
        pass

    def _none2(*a, **k):
        return None, None

    for mod, name, val in [
        ("transformers.cache_utils", "SlidingWindowCache", _Dummy),
        ("transformers.pytorch_utils", "find_pruneable_heads_and_indices", _none2),
        ("transformers.utils.backbone_utils", "get_aligned_output_features_output_indices", _none2),
    ]:
        mm = importlib.import_module(mod)
        if not hasattr(mm, name):
            setattr(mm, name, val)
    from transformers.modeling_rope_utils import ROPE_INIT_FUNCTIONS

    def _default_rope(cfg, device=None, **kw):
        d = cfg.hidden_size // cfg.num_attention_heads
        inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, d, 2, dtype=torch.int64).float() / d))
        return inv, 1.0

    ROPE_INIT_FUNCTIONS.setdefault("default", _default_rope)
    ROPE_INIT_FUNCTIONS.setdefault("mrope", _default_rope)

    # 2. autocast('cuda') -> 'cpu'
    _orig_autocast = torch.amp.autocast_mode.autocast

    class _ac(_orig_autocast):
        def __init__(self, device_type="cuda", *a, **k):
            super().__init__("cpu" if (device_type == "cuda" and _REDIRECT_AUTOCAST) else device_type, *a, **k)

    torch.amp.autocast = _ac
    torch.amp.autocast_mode.autocast = _ac
    torch.autocast = _ac
    torch.cuda.amp.autocast = lambda *a, **k: _ac("cuda", *a, **k)
    _installed = True


class StubTokenizer:
    """`tokenizer.encode(prompt)` -> fixed 6 ids, so K0 = 1 (bos) + 6 = 7."""

    def encode(self, prompt):
        return [11, 12, 13, 14, 15, 16]


NEW_TOKEN_IDS = dict(bos_token_id=1, eos_token_id=2, start_of_image=3, end_of_image=4)

TINY = dict(
    llm=dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=1,
             intermediate_size=512, vocab_size=512, rms_norm_eps=1e-6, rope_theta=1000000.0,
             max_position_embeddings=32768),
    dino=dict(hidden_size=64, num_hidden_layers=2, num_attention_heads=2, mlp_ratio=4,
              image_size=518, patch_size=14, num_register_tokens=4, layer_norm_eps=1e-6),
    vit=dict(depth=1, embed_dim=64, hidden_size=256, num_heads=2, mlp_ratio=2, patch_size=14),
)

FULL = dict(
    llm=dict(hidden_size=1536, num_hidden_layers=28, num_attention_heads=12, num_key_value_heads=2,
             intermediate_size=8960, vocab_size=151936, rms_norm_eps=1e-6, rope_theta=1000000.0,
             max_position_embeddings=32768),
    dino=dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, mlp_ratio=4,
              image_size=518, patch_size=14, num_register_tokens=4, layer_norm_eps=1e-6),
    vit=dict(depth=1, embed_dim=64, hidden_size=1536, num_heads=2, mlp_ratio=2, patch_size=14),
)


def flash_attn_zero_fill(real):
    """GPU runs use the reference's own flash-attn calls.  For PARITY runs the DINO call is wrapped so that query
    rows beyond cu_seqlens[-1] — which flash-attn never writes (`out = torch.empty_like(q)`; quirk Q1,
    g2vlm.py:988-990 vs dinov2_model.py:338-339) — hold zeros instead of whatever the caching allocator handed out.
    That is the only deviation from the stock call; speed runs use the unwrapped function."""
    def wrapped(q, k, v, cu_seqlens_q, cu_seqlens_k, *a, **kw):
        out = real(q, k, v, cu_seqlens_q, cu_seqlens_k, *a, **kw)
        n = int(cu_seqlens_q[-1])
        if n < q.shape[0]:
            out[n:] = 0
        return out
    return wrapped


import contextlib


@contextlib.contextmanager
def _skip_random_init():
    """Parameters are about to be overwritten by load_state_dict: skip the random initialisers (74 s of single-threaded
    RNG for the 4.5 B-parameter model on the host).  Buffers (rotary inv_freq, ...) are still computed normally."""
    import transformers
    saved = []
    for cls in (torch.nn.Linear, torch.nn.Embedding, torch.nn.Conv2d, torch.nn.Conv3d):
        saved.append((cls, "reset_parameters", cls.reset_parameters))
        cls.reset_parameters = lambda self: None
    pm = transformers.PreTrainedModel
    for name in ("init_weights", "_init_weights", "initialize_weights"):
        if hasattr(pm, name):
            saved.append((pm, name, getattr(pm, name)))
            setattr(pm, name, lambda self, *a, **k: None)
    try:
        yield
    finally:
        for cls, name, fn in saved:
            setattr(cls, name, fn)


def build_reference_model(dims=TINY, visual_und=True, device="cpu", zero_fill_uncovered=True, skip_init=False):
    """Constructs the reference G2VLM from its own classes (mirrors g2vlm_utils.py:31-55).
    device="cpu": flash-attn replaced by the SDPA stand-in, autocast('cuda') redirected to 'cpu'.
    device="cuda": the reference's own flash-attn calls and CUDA autocast regions run unmodified (module parameters
    are created directly on the GPU); `zero_fill_uncovered` wraps the DINO flash-attn call (see flash_attn_zero_fill).
    skip_init: leave the parameters uninitialised (the caller loads a state_dict right away)."""
    global _REDIRECT_AUTOCAST
    install_shims()
    _REDIRECT_AUTOCAST = device == "cpu"
    from modeling.dinov2_with_registers.configuration_dinov2_with_registers import Dinov2WithRegistersConfig
    from modeling.g2vlm.dinov2_model import Dinov2WithRegistersModel
    from modeling.g2vlm.g2vlm import G2VLM, G2VLMConfig
    from modeling.g2vlm.qwen2vl import Qwen2VLConfig, Qwen2VLForCausalLM
    from modeling.qwen2vl.configuration_qwen2_vl import Qwen2VLVisionConfig
    from modeling.qwen2vl.modeling_qwen2_vl import Qwen2VisionTransformerPretrainedModel
    import modeling.g2vlm.dinov2_model as _dm
    import modeling.g2vlm.qwen2vl as _qm

    import modeling.qwen2vl.modeling_qwen2_vl as _vm
    if device == "cpu":
        _qm.flash_attn_varlen_func = varlen_sdpa
        _dm.flash_attn_varlen_func = varlen_sdpa
        _vm.flash_attn_varlen_func = varlen_sdpa   # VisionFlashAttention2 (ViT of the chat path)
    else:
        from flash_attn import flash_attn_varlen_func as _real
        _qm.flash_attn_varlen_func = _real
        _vm.flash_attn_varlen_func = _real
        _dm.flash_attn_varlen_func = flash_attn_zero_fill(_real) if zero_fill_uncovered else _real

    llm = Qwen2VLConfig(pad_token_id=None, rope_scaling={"type": "mrope", "mrope_section": [16, 24, 24]},
                        qk_norm=True, layer_module="Qwen2VLMoTDecoderLayer", tie_word_embeddings=False,
                        **dims["llm"])
    dino = Dinov2WithRegistersConfig(**dims["dino"])
    vit = Qwen2VLVisionConfig(**dims["vit"])
    cfg = G2VLMConfig(visual_und=visual_und, visual_recon=True, llm_config=llm, vit_config=vit,
                      dino_config=dino, vit_max_num_patch_per_side=36)
    with torch.device(device), (_skip_random_init() if skip_init else contextlib.nullcontext()):
        lm = Qwen2VLForCausalLM(llm)
        vm = Qwen2VisionTransformerPretrainedModel(vit) if visual_und else None
        dm = Dinov2WithRegistersModel(dino)
        model = G2VLM(lm, vm, dm, cfg).eval()
    return model


def dims_from_cfg(cfg, vocab_size=None):
    """ref_harness dims dict for a g2vlm_b200.schema.G2Config (so tests build the reference at the product's dims)."""
    return dict(
        llm=dict(hidden_size=cfg.hidden_size, num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.num_heads,
                 num_key_value_heads=cfg.num_kv_heads, intermediate_size=cfg.intermediate_size,
                 vocab_size=vocab_size or cfg.vocab_size, rms_norm_eps=cfg.rms_norm_eps, rope_theta=cfg.rope_theta,
                 max_position_embeddings=32768),
        dino=dict(hidden_size=cfg.dino_hidden, num_hidden_layers=cfg.dino_layers, num_attention_heads=cfg.dino_heads,
                  mlp_ratio=cfg.dino_mlp_ratio, image_size=cfg.dino_grid * cfg.dino_patch, patch_size=cfg.dino_patch,
                  num_register_tokens=cfg.dino_registers, layer_norm_eps=cfg.dino_ln_eps),
        vit=dict(depth=1, embed_dim=64, hidden_size=cfg.hidden_size, num_heads=2, mlp_ratio=2, patch_size=14),
    )


def run_reference_recon(model, images):
    """images: list of PIL images (or paths). Returns the reference's prediction dict (CPU)."""
    install_shims()
    with torch.no_grad():
        return model.recon(StubTokenizer(), dict(NEW_TOKEN_IDS), None, images)


TRAIN_CASE = dict(samples=[([5, 20, 4], ["causal", "full", "causal"]), ([12, 7], ["full", "causal"])], seed=0)


def train_case_inputs(hidden_size: int, case=TRAIN_CASE):
    """Packed two-sample training batch for the MoT stack: 'full' splits are image blocks (first/last token und =
    <soi>/<eoi>, the rest geo), 'causal' splits are text (und).  Returns x (bf16), position ids, geo / und indexes."""
    g = torch.Generator().manual_seed(case["seed"])
    samples = case["samples"]
    T = sum(sum(s) for s, _ in samples)
    x = (torch.randn(T, hidden_size, generator=g) * 0.5).to(torch.bfloat16)
    geo, und, off = [], [], 0
    for lens, modes in samples:
        for L, mode in zip(lens, modes):
            if mode == "full":
                und += [off, off + L - 1]
                geo += list(range(off + 1, off + L - 1))
            else:
                und += list(range(off, off + L))
            off += L
    pos = torch.arange(T)[None].expand(3, -1).contiguous()
    return x, pos, torch.tensor(geo), torch.tensor(sorted(und))


def run_reference_lm_forward_train(model, x, pos, geo, und, case=TRAIN_CASE):
    """Qwen2VLModel.forward_train of the UNMODIFIED reference (qwen2vl.py:1200-1266) on CPU with the nested dense
    masks of data/data_utils.py:205-239.  The training path needs a bf16 module (it index-puts bf16 Linear outputs
    into `new_zeros` of the packed sequence) — that is what FSDP mixed precision gives it — and its
    `sdpa_kernel(EFFICIENT_ATTENTION)` context has no CPU backend, so that one context manager is replaced by a null
    context (the SDPA call itself is the reference's).  NOTE: converts model.language_model to bf16 in place."""
    import contextlib

    install_shims()
    import modeling.g2vlm.qwen2vl as q
    from data.data_utils import prepare_attention_mask_per_sample
    q.sdpa_kernel = lambda *a, **k: contextlib.nullcontext()
    lm = model.language_model.model
    lm.train()
    lm.to(torch.bfloat16)
    masks = [prepare_attention_mask_per_sample(s, m) for s, m in case["samples"]]
    with torch.no_grad():
        out = lm.forward_train(packed_sequence=x, sample_lens=[sum(s) for s, _ in case["samples"]],
                               attention_mask=masks, packed_position_ids=pos, packed_und_token_indexes=und,
                               packed_geo_token_indexes=geo)
    lm.eval()
    return out.packed_query_sequence
