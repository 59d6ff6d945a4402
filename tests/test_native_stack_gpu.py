"""Stack-level C ABI: g2vlm_ctx_* / g2vlm_load_weights / g2vlm_workspace_bytes / g2vlm_recon_plan /
g2vlm_dino_forward / g2vlm_mot_forward_geo / g2vlm_recon_heads.  The native stage drivers issue the same kernels in the
same order as the per-op path driven from Python, so their results must be BIT-IDENTICAL to it (and the per-op path is
what every oracle / reference parity test checks stage by stage)."""
from dataclasses import replace

import pytest
import torch

from g2vlm_b200 import schema

pytestmark = pytest.mark.gpu
KEYS = ("points", "local_points", "global_points", "camera_poses")


class Tok:
    def encode(self, prompt):
        return [11, 12, 13, 14, 15, 16]


IDS = dict(bos_token_id=1, eos_token_id=2, start_of_image=3, end_of_image=4)


def _views(n, h, w, seed):
    return (schema.synthetic_views(n, h, w, seed=seed) * 255).round() / 255.0


def _both(model, v):
    model.native = True
    a = model.recon(Tok(), dict(IDS), None, v)
    a = {k: a[k].clone() for k in KEYS + ("conf",) if a.get(k) is not None}
    model.native = False
    b = model.recon(Tok(), dict(IDS), None, v)
    model.native = True
    torch.cuda.synchronize()
    return a, b


@pytest.fixture(scope="module")
def tiny():
    from g2vlm_b200.model import G2VLMFast
    return G2VLMFast(schema.TINY, schema.init_synthetic(schema.TINY, seed=0))


def test_native_stages_are_bit_identical_to_the_per_op_path(tiny):
    from g2vlm_b200 import ops
    geoms = [(3, 70, 518, 1), (2, 518, 518, 2), (5, 14, 518, 3), (3, 70, 518, 4)]   # switches geometry and comes back
    for n, h, w, seed in geoms:
        l0 = ops.launches()
        a, b = _both(tiny, _views(n, h, w, seed))
        assert ops.launches() - l0 > 100
        for k in KEYS:
            assert torch.equal(a[k], b[k]), (k, (n, h, w))
    assert len(tiny._nws) == 3                     # one cached workspace per geometry


def test_native_conf_branch_and_full_width_depth1():
    from g2vlm_b200.model import G2VLMFast
    cfg = replace(schema.TINY, train_conf_pi3=True)
    m = G2VLMFast(cfg, schema.init_synthetic(cfg, seed=2))
    a, b = _both(m, _views(2, 56, 518, 8))
    assert a["conf"].shape == (1, 2, 56, 518, 1) and torch.equal(a["conf"], b["conf"])
    assert all(torch.equal(a[k], b[k]) for k in KEYS)
    del m
    cfg = replace(schema.FULL, num_layers=1, dino_layers=1, dec_depth=1)       # full-size kernels incl. the CTA-pair GEMMs
    m = G2VLMFast(cfg, schema.init_synthetic(cfg, seed=3, embed_rows=32, device="cuda"))
    a, b = _both(m, _views(8, 294, 518, 4))
    assert all(torch.equal(a[k], b[k]) for k in KEYS)
    del m
    torch.cuda.empty_cache()


def test_workspace_and_plan_contract(tiny):
    from g2vlm_b200 import ops
    ctx = tiny._native_ctx()
    small, big = ctx.workspace_bytes(2, 70, 518, 7), ctx.workspace_bytes(16, 518, 518, 7)
    assert 0 < small < big
    with pytest.raises(ops.G2Error):
        ctx.workspace_bytes(2, 71, 518, 7)          # H not a multiple of the patch size
    import ctypes
    ws = torch.empty(small - 256, dtype=torch.uint8, device="cuda")
    with pytest.raises(ops.G2Error, match="workspace smaller"):
        ctx.call("g2vlm_recon_plan", ops._i32(2), ops._i32(70), ops._i32(518), ops._i32(7), (ctypes.c_int32 * 2)(185, 185),
                 ops._vp(ws.data_ptr()), ops._i64(ws.numel()), ops._stream())
    tiny._nplan = None                              # the failed plan left the context's plan untouched or invalid: re-plan
    # a context with no weights names the slot it misses
    bare = ops.NativeContext(schema.TINY)
    ws = torch.empty(small, dtype=torch.uint8, device="cuda")
    bare.call("g2vlm_recon_plan", ops._i32(2), ops._i32(70), ops._i32(518), ops._i32(7), (ctypes.c_int32 * 2)(185, 185),
              ops._vp(ws.data_ptr()), ops._i64(ws.numel()), ops._stream())
    img = torch.zeros(2, 3, 70, 518, device="cuda")
    pos = torch.zeros(186, 64, device="cuda")
    with pytest.raises(ops.G2Error, match="weight not loaded: dino.wpatch"):
        bare.call("g2vlm_dino_forward", ops._vp(img.data_ptr()), ops._i32(2), ops._i32(70), ops._i32(518), ops._i32(0),
                  ops._vp(pos.data_ptr()), ops._vp(ws.data_ptr()), None, ops._stream())
    torch.cuda.synchronize()


def test_stage_calls_are_cuda_graph_capturable(tiny):
    """After the plan the three stage calls only launch kernels: capture them once, replay on new pixels."""
    from g2vlm_b200.model import NaiveCache
    m = tiny
    m.native = True
    tok = Tok()
    v0, v1 = _views(3, 70, 518, 21), _views(3, 70, 518, 22)
    gi_text, newlens, new_rope = m.prepare_prompts_addbos([0], [0], ["x"], tok, IDS)
    meta = ("text_token_lens", "key_values_lens", "packed_seqlens", "dino_token_seqlens")
    gi_text = {k: (v if k in meta else v.cuda()) for k, v in gi_text.items()}
    gi, _, _ = m.prepare_dino_images_pi3(newlens, new_rope, v0, None, IDS)
    gi = {k: (v if k in meta else v.cuda()) for k, v in gi.items()}
    gi1, _, _ = m.prepare_dino_images_pi3(newlens, new_rope, v1, None, IDS)

    def step():
        past, last = m.forward_cache_update_dino(NaiveCache(m.cfg.num_layers), prompt=gi_text, update_past_key_values=False, **gi)
        return m.reconstruct(past_key_values=past, selected_hidden_states=last, **gi)

    eager0 = {k: step()[k].clone() for k in KEYS}          # also plans the geometry
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = step()
    g.replay()
    torch.cuda.synchronize()
    assert all(torch.equal(out[k], eager0[k]) for k in KEYS)
    gi["packed_dino_images"].copy_(gi1["packed_dino_images"])   # new pixels in the captured input buffer
    g.replay()
    torch.cuda.synchronize()
    replayed = {k: out[k].clone() for k in KEYS}
    eager1 = step()
    assert all(torch.equal(replayed[k], eager1[k]) for k in KEYS)
    assert not torch.equal(replayed["points"], eager0["points"])


def test_native_und_prefill_is_bit_identical_to_the_per_op_path():
    """g2vlm_und_prefill (text prefill / ViT step of the und expert as ONE C call) issues the per-op entry points of
    G2VLMFast._und_forward in the same order: caches and hidden states are bit-identical, causal and non-causal, on an
    empty cache and on top of one, short (<= 8 rows: GEMV path) and long (tensor-core tiles) steps."""
    from g2vlm_b200.model import G2VLMFast, KVCache, NaiveCache
    cfg = schema.TINY
    model = G2VLMFast(cfg, schema.init_synthetic(cfg, seed=4, device="cuda"))

    def run(native):
        model.native = native
        outs, cache, kvlen = [], NaiveCache(cfg.num_layers), 0
        for n, causal in ((5, True), (300, True), (9, False), (2, True)):
            ids = (torch.arange(n) * 13 + 7) % cfg.vocab_size
            x = model.embed[ids.cuda()].contiguous()
            cache = KVCache.adopt(cache, cfg, model.device)
            y = model._und_forward(x, (kvlen + torch.arange(n)).expand(3, -1), cache, causal=causal)
            kvlen += n
            outs.append(y.clone())
        torch.cuda.synchronize()
        return outs, [b[:kvlen].clone() for b in cache.buf], cache.seq_lens

    a, ka, la = run(False)
    b, kb, lb = run(True)
    assert la == lb == 316
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    for x, y in zip(ka, kb):
        assert torch.equal(x, y)
