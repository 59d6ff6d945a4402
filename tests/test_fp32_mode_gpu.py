"""fp32 mode (`G2VLMFast(..., mode="fp32")`, north_star "fp32 mode <= 1e-4"; BASELINE configs[0] "8 views fp32").

Ground truth: `oracle/restate.py` mode="fp32" — the reference's algorithm with every bf16 rounding point removed (the
reference itself cannot run in fp32: hard casts at modeling/g2vlm/qwen2vl.py:579, 617-619).  Tolerance written here:
max|a-b| / max|ref| <= 1e-4 on every output and intermediate stage."""
import math

import pytest
import torch

from g2vlm_b200 import schema

pytestmark = pytest.mark.gpu
TOL = 1e-4


class Tok:
    def encode(self, prompt):
        return [11, 12, 13, 14, 15, 16]


IDS = dict(bos_token_id=1, eos_token_id=2, start_of_image=3, end_of_image=4)


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


def _views(n, h, w, seed):
    return (schema.synthetic_views(n, h, w, seed=seed) * 255).round() / 255.0


# ---------------------------------------------------------------------------------------------------------
# kernels
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,hq,hk,cu_q,cu_k,causal,q_rows", [
    (128, 12, 2, [0, 300], [0, 1000], False, None),             # MoT layout, GQA 6:1, ragged tails
    (128, 2, 1, [0, 7], [0, 7], True, None),                    # the prompt rows: 7 causal tokens
    (128, 4, 2, [0, 200], [0, 333], True, None),                # bottom-right aligned causal, lk > lq
    (64, 4, 4, [0, 777, 1554], [0, 777, 1554], False, 1564),    # DINO: two segments, 10 uncovered tail rows (Q1)
    (96, 16, 16, [0, 37, 74], [0, 37, 74], False, None),        # Pi3 heads, 96 wide, short segments
    (32, 2, 2, [0, 190], [0, 190], False, None),
    (16, 16, 16, [0, 260, 520], [0, 260, 520], False, None),    # > 256 rows per segment: two work items
])
def test_attention_f32_matches_fp64(d, hq, hk, cu_q, cu_k, causal, q_rows):
    from g2vlm_b200 import ops
    g = torch.Generator().manual_seed(d + hq)
    q_rows = q_rows or cu_q[-1]
    kv_rows = max(cu_k[-1], q_rows)
    q = torch.randn(q_rows, hq * d, generator=g).cuda()
    k = torch.randn(kv_rows, hk * d, generator=g).cuda()
    v = torch.randn(kv_rows, hk * d, generator=g).cuda()
    out = torch.full((q_rows, hq * d), 7.0, device="cuda")
    scale = 1.0 / math.sqrt(d)
    ops.attention_f32(q, k, v, out, ops.attention_work_table(cu_q, cu_k).cuda(), num_q_heads=hq, num_kv_heads=hk,
                      head_dim=d, scale=scale, causal=causal)
    torch.cuda.synchronize()
    ref = torch.full((q_rows, hq * d), 7.0, dtype=torch.float64)
    for i in range(len(cu_q) - 1):
        qs = q[cu_q[i]:cu_q[i + 1]].double().cpu().view(-1, hq, d).transpose(0, 1)
        ks = k[cu_k[i]:cu_k[i + 1]].double().cpu().view(-1, hk, d).transpose(0, 1).repeat_interleave(hq // hk, 0)
        vs = v[cu_k[i]:cu_k[i + 1]].double().cpu().view(-1, hk, d).transpose(0, 1).repeat_interleave(hq // hk, 0)
        s = (qs @ ks.transpose(1, 2)) * scale
        if causal:
            lq, lk = s.shape[1], s.shape[2]
            s = s.masked_fill(~torch.ones(lq, lk, dtype=torch.bool).tril(diagonal=lk - lq), float("-inf"))
        ref[cu_q[i]:cu_q[i + 1]] = (torch.softmax(s, -1) @ vs).transpose(0, 1).reshape(-1, hq * d)
    assert (out.double().cpu() - ref).abs().max().item() < 2e-5
    if q_rows > cu_q[-1]:
        assert bool((out[cu_q[-1]:] == 7.0).all()), "rows in no segment must not be written"


def test_split_bf16_gemm_with_fp32_epilogue_options():
    """out = x + gamma * gelu(a @ w.T + b) in fp32 through [hi|hi|lo] x [hi|lo|hi]: 2^-16-level agreement with fp64."""
    from g2vlm_b200 import ops
    from g2vlm_b200.model import _split_hi_lo_hi
    g = torch.Generator().manual_seed(3)
    rows, K, N = 700, 1024, 768
    a = torch.randn(rows, K, generator=g).cuda()
    w = (torch.randn(2 * N, K, generator=g) * 0.05).cuda()
    b = torch.randn(2 * N, generator=g).cuda()
    gamma = (torch.rand(N, generator=g) + 0.5).cuda()
    x0 = torch.randn(rows, N, generator=g).cuda()
    groups = [(0, 600), (600, 100)]
    a3 = torch.empty(rows, 3 * K, device="cuda", dtype=torch.bfloat16)
    ops.split3(a, a3)
    x = x0.clone()
    ops.gemm(a3, _split_hi_lo_hi(w), x, epilogue=ops.EPI_STORE_F32, groups=groups, bias=b, flags=ops.GEMM_GELU, scale=gamma,
             scale_groups=1, residual=x)
    ref = torch.empty(rows, N, dtype=torch.float64)
    for e, (r0, n) in enumerate(groups):
        y = torch.nn.functional.gelu(a[r0:r0 + n].double().cpu() @ w[e * N:(e + 1) * N].double().cpu().T + b[e * N:(e + 1) * N].double().cpu())
        ref[r0:r0 + n] = x0[r0:r0 + n].double().cpu() + (y * gamma.double().cpu() if e == 0 else y)
    assert _rel(x, ref) < 2e-5
    # three-piece split (what fp32 mode uses): x = h + m + l exactly, six products -> fp32-level accuracy
    from g2vlm_b200.model_fp32 import _w3
    a6 = torch.empty(rows, 6 * K, device="cuda", dtype=torch.bfloat16)
    ops.split6(a, a6)
    pieces = a6.float().view(rows, 6, K)
    assert torch.equal(pieces[:, 2] + pieces[:, 0] + pieces[:, 1], a)            # h + m + l == x, bit for bit
    errs = {}
    for chunk in (0, 4):   # 0: one accumulation in the tensor core (truncating adds); 4: chunks summed in fp32 RN
        x = torch.full_like(x0, float("nan"))
        ops.gemm(a6, _w3(w, "cuda"), x, epilogue=ops.EPI_STORE_F32, groups=groups, bias=b, flags=ops.GEMM_GELU, scale=gamma,
                 scale_groups=1, residual=x0, k_chunk_blocks=chunk)
        errs[chunk] = _rel(x, ref)
    with pytest.raises(Exception):   # the chunks accumulate in `out`: an aliased residual is rejected
        ops.gemm(a6, _w3(w, "cuda"), x, epilogue=ops.EPI_STORE_F32, groups=groups, residual=x, k_chunk_blocks=4)
    print("\n  three-piece split vs fp64:", errs)
    assert errs[4] < 1e-6 and errs[4] <= errs[0]


def test_fp32_elementwise_kernels():
    from g2vlm_b200 import ops
    g = torch.Generator().manual_seed(5)
    # SwiGLU
    gu = torch.randn(300, 2 * 512, generator=g).cuda()
    out = torch.empty(300, 512, device="cuda")
    ops.swiglu_f32(gu, out)
    assert _rel(out, torch.nn.functional.silu(gu[:, :512].double()) * gu[:, 512:].double()) < 1e-6
    # q/k norm + M-RoPE against the restatement's formulas
    from oracle import restate
    rows, nq, nkv, hd = 50, 4, 2, 128
    qkv = torch.randn(rows, (nq + 2 * nkv) * hd, generator=g)
    wq = [torch.rand(hd, generator=g) + 0.5 for _ in range(2)]
    wk = [torch.rand(hd, generator=g) + 0.5 for _ in range(2)]
    pos = torch.randint(0, 900, (3, rows), generator=g)
    cos, sin = restate.mrope_cos_sin(pos, hd, 1e6, (16, 24, 24))
    n_first = 30
    want = qkv.clone()
    for (c0, nh, ws) in ((0, nq, wq), (nq * hd, nkv, wk)):
        t = qkv[:, c0:c0 + nh * hd].view(rows, nh, hd)
        n = torch.empty_like(t)
        n[:n_first] = restate.rmsnorm(t[:n_first], ws[0], 1e-6)
        n[n_first:] = restate.rmsnorm(t[n_first:], ws[1], 1e-6)
        want[:, c0:c0 + nh * hd] = (n * cos[:, None] + restate.rotate_half(n) * sin[:, None]).reshape(rows, -1)
    buf = qkv.cuda()
    ops.qknorm_mrope_f32(buf, rows, n_first, nq, nkv, hd, wq[0].cuda(), wk[0].cuda(), wq[1].cuda(), wk[1].cuda(),
                         cos[:, :64].contiguous().cuda(), sin[:, :64].contiguous().cuda(), 1e-6)
    assert _rel(buf, want) < 2e-6
    # RoPE2D against the restatement (fp32 mode = exact angles)
    N, gh, gw, heads, d = 2, 3, 5, 4, 96
    P = gh * gw
    t = torch.randn(N * P, 3 * heads * d, generator=g)
    yy, xx = torch.meshgrid(torch.arange(gh), torch.arange(gw), indexing="ij")
    posi = torch.stack([yy.flatten(), xx.flatten()], -1)[None].expand(N, -1, -1)
    want = t.clone()
    for part in range(2):
        tt = t[:, part * heads * d:(part + 1) * heads * d].view(N, P, heads, d).transpose(1, 2)
        want[:, part * heads * d:(part + 1) * heads * d] = restate.rope2d(tt, posi, 100.0, "fp32").transpose(1, 2).reshape(N * P, -1)
    Dh = d // 2
    inv_freq = 1.0 / (100.0 ** (torch.arange(0, Dh, 2).float() / Dh))
    fr = torch.einsum("i,j->ij", torch.arange(max(gh, gw)).float(), inv_freq)
    buf = t.cuda()
    ops.rope2d_f32(buf, N * P, 2 * heads, d, d, P, gw, fr.cos().contiguous().cuda(), fr.sin().contiguous().cuda())
    assert _rel(buf, want) < 2e-6


# ---------------------------------------------------------------------------------------------------------
# the whole path
# ---------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def tiny32():
    from g2vlm_b200.model import G2VLMFast
    sd = schema.init_synthetic(schema.TINY, seed=0)
    m = G2VLMFast(schema.TINY, sd, mode="fp32")
    assert m.mode == "fp32" and type(m).__name__ == "G2VLMFastFP32"
    return sd, m


@pytest.mark.parametrize("case", [dict(n=3, h=70, w=518, seed=1), dict(n=2, h=518, w=518, seed=2), dict(n=1, h=42, w=518, seed=3)])
def test_fp32_recon_matches_fp32_oracle_stagewise(tiny32, case):
    from oracle import restate
    sd, model = tiny32
    v = _views(**case)
    c_ref, c_out = {}, {}
    ref = restate.recon(sd, schema.TINY, v, mode="fp32", collect=c_ref)
    out = model.recon(Tok(), dict(IDS), None, v, collect=c_out)
    torch.cuda.synchronize()
    errs = {}
    for i, (a, b) in enumerate(zip(c_out["dino_layers"], c_ref["dino_layers"])):
        errs[f"dino{i}"] = _rel(a, b)
    errs["dino_tokens"] = _rel(c_out["dino_tokens"].view_as(c_ref["dino_tokens"]), c_ref["dino_tokens"])
    errs["packed_sequence"] = _rel(c_out["packed_sequence"], c_ref["packed_sequence"])
    for i, (a, b) in enumerate(zip(c_out["mot_layers"], c_ref["mot_layers"])):
        errs[f"mot{i}"] = _rel(a, b)
    errs["last_hidden"] = _rel(c_out["last_hidden"], c_ref["last_hidden"])
    for k in ("point_hidden", "camera_hidden", "global_hidden"):
        errs[k] = _rel(c_out[k], c_ref[k])
    for k in ("local_points", "points", "global_points", "camera_poses"):
        errs[k] = _rel(out[k], ref[k])
    print("\n" + "\n".join(f"  {k:18s} {e:.3e}" for k, e in errs.items()))
    assert all(e < TOL for e in errs.values()), errs
    assert torch.equal(out["images"].cpu()[0], v)


def test_fp32_mode_is_closer_to_fp32_truth_than_bf16_mode_by_orders_of_magnitude(tiny32):
    from g2vlm_b200.model import G2VLMFast
    from oracle import restate
    sd, m32 = tiny32
    m16 = G2VLMFast(schema.TINY, sd)
    v = _views(2, 140, 518, 9)
    truth = restate.recon(sd, schema.TINY, v, mode="fp32")
    e32 = _rel(m32.recon(Tok(), dict(IDS), None, v)["points"], truth["points"])
    e16 = _rel(m16.recon(Tok(), dict(IDS), None, v)["points"], truth["points"])
    assert e32 < TOL and e16 > 20 * e32, (e32, e16)


def test_fp32_mode_rejects_bf16_only_paths(tiny32):
    _, model = tiny32
    with pytest.raises(NotImplementedError):
        model.generate_text(None, None, None, None, None, 4)
    model.fuse_prompt = False
    try:
        with pytest.raises(NotImplementedError):
            model.recon(Tok(), dict(IDS), None, _views(1, 42, 518, 1))
    finally:
        model.fuse_prompt = True


@pytest.mark.parametrize("depth", [1, None])
def test_fp32_full_width_configs0_shape(depth):
    """BASELINE configs[0] geometry (8 views of 294x518: non-square grid, bicubic pos-embed resample, T = 6232) at FULL
    width, depth 1 and full depth (28 + 24 + 5), LayerScale at the reference's init 0.01.  The restatement runs on GPU
    tensors (same code; 33.6 TFLOP of fp32 matmuls at full depth)."""
    from dataclasses import replace

    from g2vlm_b200.model import G2VLMFast
    from oracle import restate
    cfg = schema.FULL if depth is None else replace(schema.FULL, num_layers=depth, dino_layers=depth, dec_depth=depth)
    sd = schema.init_synthetic(cfg, seed=0, embed_rows=32, device="cuda")
    for k in sd:
        if k.endswith("ls1.gamma") or k.endswith("ls2.gamma") or k.endswith(".lambda1"):
            sd[k].fill_(0.01)
    model = G2VLMFast(cfg, sd, mode="fp32")
    v = _views(8, 294, 518, 1)
    c_ref, c_out = {}, {}
    out = model.recon(Tok(), dict(IDS), None, v, collect=c_out)
    with torch.device("cuda"):
        ref = restate.recon(sd, cfg, v.cuda(), mode="fp32", collect=c_ref)
    errs = {"dino_tokens": _rel(c_out["dino_tokens"].view_as(c_ref["dino_tokens"]), c_ref["dino_tokens"]),
            "last_hidden": _rel(c_out["last_hidden"], c_ref["last_hidden"])}
    for k in ("point_hidden", "camera_hidden", "global_hidden"):
        errs[k] = _rel(c_out[k], c_ref[k])
    errs.update({k: _rel(out[k], ref[k]) for k in ("local_points", "points", "global_points", "camera_poses")})
    print(f"\n  depth {depth or 'full'}: " + "  ".join(f"{k} {e:.2e}" for k, e in errs.items()))
    assert all(e < TOL for e in errs.values()), errs
    del model
    torch.cuda.empty_cache()
