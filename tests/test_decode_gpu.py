"""Decode-shaped kernels (row f1): the GEMV path of g2vlm_gemm_bf16 (<= 8 rows) and the split-K decode
attention, against fp32 torch references / the tensor-core kernels."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _relerr(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("rows", [1, 2, 3, 7, 8, 9, 16, 20, 33, 60, 64])
def test_gemv_epilogues_match_reference(rows):
    from g2vlm_b200 import ops
    g = torch.Generator().manual_seed(rows)
    K, N = 1536, 640
    x = (torch.randn(rows, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
    w = (torch.randn(2 * N, K, generator=g) * 0.05).to(torch.bfloat16).cuda()      # two stacked experts
    bias = torch.randn(2 * N, generator=g).cuda()
    groups = [(0, 0), (0, rows)]                                                      # all rows -> expert 1 (und)
    ref = x.float() @ w[N:].float().T
    out = torch.empty(rows, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(x, w, out, epilogue=ops.EPI_STORE_BF16, groups=groups, bias=bias)
    assert _relerr(out, ref + bias[N:]) < 1e-2
    ops.gemm(x, w, out, epilogue=ops.EPI_STORE_BF16, groups=groups, bias=bias, flags=ops.GEMM_GELU)
    assert _relerr(out, torch.nn.functional.gelu((ref + bias[N:]).to(torch.bfloat16).float())) < 1e-2
    gamma = (torch.rand(N, generator=g) + 0.5).cuda()
    r0 = torch.randn(rows, N, generator=g).cuda()     # (seeded: an unseeded device tensor made this check flaky at its edge)
    r = r0.clone()
    ops.gemm(x, w, r, epilogue=ops.EPI_RESID_F32, groups=groups, scale=gamma, scale_groups=2,
             flags=ops.GEMM_ROUND_AFTER_SCALE)
    # one bf16 ulp of the largest update (a rounding flip of bf16(acc) between the two accumulation orders) over max |r|
    assert _relerr(r, r0 + (ref.to(torch.bfloat16).float() * gamma).to(torch.bfloat16).float()) < 8e-3
    f = torch.empty(rows, N, device="cuda")
    ops.gemm(x, w, f, epilogue=ops.EPI_STORE_F32, groups=groups, bias=bias, flags=ops.GEMM_RELU, residual=r0)
    assert _relerr(f, torch.relu(ref + bias[N:]) + r0) < 1e-4
    if rows > 8:   # 9..64 rows: the warp-MMA kernel against the tcgen05 tile kernel on the same inputs
        out2 = torch.empty_like(out)
        ops.gemm(x, w, out, epilogue=ops.EPI_STORE_BF16, groups=groups, bias=bias)
        ops.gemm(x, w, out2, epilogue=ops.EPI_STORE_BF16, groups=groups, bias=bias, flags=ops.GEMM_FORCE_SINGLE)
        assert _relerr(out, out2) < 1e-2 and (out != out2).float().mean() < 0.05
    # SwiGLU: interleaved gate/up blocks of 128 rows
    I = 512
    wg = (torch.randn(I, K, generator=g) * 0.08).to(torch.bfloat16).cuda()
    wu = (torch.randn(I, K, generator=g) * 0.08).to(torch.bfloat16).cuda()
    wi = torch.stack([wg.view(I // 128, 128, K), wu.view(I // 128, 128, K)], 1).reshape(2 * I, K).contiguous()
    o = torch.empty(rows, I, device="cuda", dtype=torch.bfloat16)
    ops.gemm(x, wi, o, epilogue=ops.EPI_SWIGLU_BF16)
    gt, up = (x.float() @ wg.float().T).to(torch.bfloat16), (x.float() @ wu.float().T).to(torch.bfloat16)
    assert _relerr(o, torch.nn.functional.silu(gt.float()).to(torch.bfloat16).float() * up.float()) < 1.5e-2


@pytest.mark.parametrize("rows,K,N", [(5, 1536, 648), (7, 1552, 640), (4, 8960, 1536), (7, 4096, 48), (20, 8960, 1536),
                                      (60, 1536, 2048), (64, 8960, 1536), (33, 1568, 656)])
def test_gemv_skinny_shapes(rows, K, N):
    """2..8 rows: the mma.sync kernel (N % 16 == 0, K % 32 == 0; 8 or 16 warps split K, ragged K split) and the
    scalar fallback for the other shapes."""
    from g2vlm_b200 import ops
    g = torch.Generator().manual_seed(K + N + rows)
    x = (torch.randn(rows, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
    w = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).cuda()
    out = torch.full((rows + 1, N), 7.0, device="cuda")
    ops.gemm(x, w, out[:rows], epilogue=ops.EPI_STORE_F32)
    assert _relerr(out[:rows], x.float() @ w.float().T) < 1e-4
    assert bool((out[rows] == 7.0).all())


def test_gemv_agrees_with_tensor_core_path():
    """9 rows -> tcgen05 tiles, 8 rows -> GEMV: the first 8 rows must agree up to accumulation order."""
    from g2vlm_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(9, 8960, generator=g) * 0.3).to(torch.bfloat16).cuda()
    w = (torch.randn(1536, 8960, generator=g) * 0.03).to(torch.bfloat16).cuda()
    a = torch.empty(9, 1536, device="cuda", dtype=torch.bfloat16)
    b = torch.empty(8, 1536, device="cuda", dtype=torch.bfloat16)
    ops.gemm(x, w, a, epilogue=ops.EPI_STORE_BF16)
    ops.gemm(x[:8], w, b, epilogue=ops.EPI_STORE_BF16)
    assert _relerr(b, a[:8]) < 1e-2
    assert (a[:8] != b).float().mean() < 0.05


@pytest.mark.parametrize("L", [1, 100, 128, 1000, 21951, 80000])
def test_attention_decode(L):
    from g2vlm_b200 import ops
    g = torch.Generator().manual_seed(L)
    nq, nkv, hd = 12, 2, 128
    q = torch.randn(nq * hd, generator=g).to(torch.bfloat16).cuda()
    kv = torch.randn(L + 5, 2 * nkv * hd, generator=g).to(torch.bfloat16).cuda()      # K | V, rows past L unused
    k, v = kv[:L, : nkv * hd], kv[:L, nkv * hd:]
    out = torch.empty(nq * hd, device="cuda", dtype=torch.bfloat16)
    ws = torch.empty(ops.attention_decode_workspace_floats(L, nq), device="cuda")
    scale = 1 / math.sqrt(hd)
    ops.attention_decode(q, k, v, out, ws, num_q_heads=nq, num_kv_heads=nkv, head_dim=hd, scale=scale)
    qs = q.float().view(nq, 1, hd)
    ks = k.float().view(L, nkv, hd).transpose(0, 1).repeat_interleave(nq // nkv, 0)
    vs = v.float().view(L, nkv, hd).transpose(0, 1).repeat_interleave(nq // nkv, 0)
    ref = (torch.softmax(qs @ ks.transpose(1, 2) * scale, -1) @ vs).reshape(-1)
    assert (out.float() - ref).abs().max().item() < 1e-2
    # and against the tensor-core kernel on the same inputs
    out2 = torch.empty(1, nq * hd, device="cuda", dtype=torch.bfloat16)
    work = ops.attention_work_table([0, 1], [0, L]).cuda()
    ops.attention(q.view(1, -1), k, v, out2, work, num_q_heads=nq, num_kv_heads=nkv, head_dim=hd, scale=scale)
    assert (out.float() - out2[0].float()).abs().max().item() < 1e-2


def test_attention_decode_device_resident_length_and_kv_append():
    """The graph-replayable form: the key count and the append row are read from device memory."""
    from g2vlm_b200 import ops
    g = torch.Generator().manual_seed(7)
    nq, nkv, hd, cap = 12, 2, 128, 3000
    kvbuf = torch.randn(cap, 2 * nkv * hd, generator=g).to(torch.bfloat16).cuda()
    q = torch.randn(nq * hd, generator=g).to(torch.bfloat16).cuda()
    new = torch.randn(1, 2 * nkv * hd, generator=g).to(torch.bfloat16).cuda()
    len_dev = torch.tensor([1234], dtype=torch.int32, device="cuda")
    ws = torch.empty(ops.attention_decode_workspace_floats(cap, nq), device="cuda")
    out = torch.empty(nq * hd, device="cuda", dtype=torch.bfloat16)
    scale = 1 / math.sqrt(hd)
    for step in range(3):
        before = kvbuf.clone()
        ops.kv_append(new, kvbuf, 1, len_dev=len_dev)
        L = 1234 + step
        assert torch.equal(kvbuf[L], new[0]) and torch.equal(kvbuf[:L], before[:L]) and torch.equal(kvbuf[L + 1:], before[L + 1:])
        ops.attention_decode(q, kvbuf[:cap, : nkv * hd], kvbuf[:cap, nkv * hd:], out, ws, num_q_heads=nq,
                             num_kv_heads=nkv, head_dim=hd, scale=scale, kv_len_dev=len_dev, kv_len_extra=1)
        k, v = kvbuf[:L + 1, : nkv * hd], kvbuf[:L + 1, nkv * hd:]
        ks = k.float().view(L + 1, nkv, hd).transpose(0, 1).repeat_interleave(nq // nkv, 0)
        vs = v.float().view(L + 1, nkv, hd).transpose(0, 1).repeat_interleave(nq // nkv, 0)
        ref = (torch.softmax(q.float().view(nq, 1, hd) @ ks.transpose(1, 2) * scale, -1) @ vs).reshape(-1)
        assert (out.float() - ref).abs().max().item() < 1e-2
        len_dev += 1


@pytest.mark.parametrize("rows,vocab", [(1, 151936), (3, 512), (2, 40000)])
def test_argmax_lowest_index_on_ties(rows, vocab):
    from g2vlm_b200 import ops
    g = torch.Generator().manual_seed(vocab)
    logits = torch.randn(rows, vocab, generator=g).to(torch.bfloat16)
    logits[0, vocab // 3] = 100.0
    logits[0, vocab // 3 + 777 if vocab > 2000 else vocab // 3 + 7] = 100.0      # tie: the lower index must win
    out = torch.full((rows,), -1, dtype=torch.int64, device="cuda")
    for _ in range(2):                                                            # twice: the ticket counter resets
        ops.argmax_bf16(logits.cuda(), out)
        assert out[0].item() == vocab // 3
        assert torch.equal(out.cpu()[1:], logits[1:].float().argmax(-1))


_FUSED_CFGS = {
    "tiny": dict(cfg=dict(hidden_size=256, num_layers=2, num_heads=2, num_kv_heads=1, intermediate_size=512, vocab_size=512,
                          dino_hidden=64, dino_layers=1, dino_heads=2)),
    # the full model's und-expert geometry (12 q / 2 kv heads, 1536 / 8960), 3 layers, a vocabulary with a ragged tail
    "full-width": dict(cfg=dict(num_layers=3, vocab_size=20011, dino_hidden=64, dino_layers=1, dino_heads=2, dec_depth=1)),
}


@pytest.mark.parametrize("name,L0,steps", [("tiny", 5, 6), ("tiny", 700, 4), ("full-width", 37, 4), ("full-width", 3100, 5),
                                           ("full-width", 30011, 3)])   # 406 keys per CTA: three 160-key stages
def test_fused_decode_step_matches_multi_launch(name, L0, steps):
    """Row f1: the one-kernel decode step (csrc/decode_fused.cu; persistent cooperative kernel, grid barriers) against
    the ~280-launch driver it replaces: same tokens, same logits up to the accumulation order of the attention
    partials, same appended K|V rows, device-resident token / position / cache length advanced alike; eager and through
    a captured CUDA graph."""
    from g2vlm_b200 import schema
    from g2vlm_b200.model import G2VLMFast, NaiveCache
    cfg = schema.G2Config(**_FUSED_CFGS[name]["cfg"])
    model = G2VLMFast(cfg, schema.init_synthetic(cfg, seed=3, device="cuda"))
    V = cfg.vocab_size

    def prefill():
        ids = (torch.arange(L0) * 7 + 11) % V
        return model.forward_cache_update_text(
            NaiveCache(cfg.num_layers), text_token_lens=torch.tensor([L0], dtype=torch.int), packed_text_ids=ids,
            packed_text_position_ids=torch.arange(L0).expand(3, -1), packed_text_indexes=torch.arange(L0),
            packed_key_value_indexes=torch.arange(0), key_values_lens=torch.tensor([0], dtype=torch.int))

    def run(fused, graph, n):
        past = prefill()
        ids = model.generate_text(past, None, None, torch.tensor([23]), torch.full((3, 1), L0), n, end_token_id=None,
                                  use_cuda_graph=graph, fused_step=fused)
        torch.cuda.synchronize()
        logits = model.buf.get("dec.logits", ((V + 7) // 8 * 8,), torch.bfloat16)[:V].float().clone()
        rows = [b[L0:L0 + n].float().clone() for b in past.buf]
        return ids[:, 0].tolist(), logits, rows, past.seq_lens

    def run_forced(fused, tokens):
        """One step at a time from given start tokens (teacher forcing): the logits of every step."""
        past = prefill()
        outs = []
        for i, tok in enumerate(tokens):
            model.generate_text(past, None, None, torch.tensor([tok]), torch.full((3, 1), L0 + i), 1, end_token_id=None,
                                fused_step=fused)
            torch.cuda.synchronize()
            outs.append(model.buf.get("dec.logits", ((V + 7) // 8 * 8,), torch.bfloat16)[:V].float().clone())
        return outs, [b[L0:L0 + len(tokens)].float().clone() for b in past.buf], past.seq_lens

    t_ref, _, rows_ref, len_ref = run(False, False, steps)
    assert len_ref == L0 + steps and t_ref[0] == 23
    # the same inputs step by step: logits agree to accumulation-order noise (one bf16 ulp of the largest logit), the
    # argmax agrees unless the multi-launch path itself has a near tie there
    lg_ref, rows_a, _ = run_forced(False, t_ref)
    lg_fus, rows_b, len_b = run_forced(True, t_ref)
    assert len_b == L0 + steps
    for a, b in zip(lg_fus, lg_ref):
        assert _relerr(a, b) < 1.6e-2          # bf16 logits: 3 ulps of the largest one
        top2 = b.topk(2).values
        if (top2[0] - top2[1]) > 2e-2 * b.abs().max():
            assert int(a.argmax()) == int(b.argmax())
    for a, b in zip(rows_b, rows_a):
        assert _relerr(a, b) < 2e-2
    # free-running generation: eager and graph replay of the one-kernel step are the same computation
    t_e, lg_e, rows_e, len_e = run(True, False, steps)
    t_g, lg_g, rows_g, len_g = run(True, True, steps)
    assert t_e == t_g and len_e == len_g == L0 + steps and torch.equal(lg_e, lg_g)
    for a, b in zip(rows_e, rows_g):
        assert torch.equal(a, b)
    if all((l.topk(2).values[0] - l.topk(2).values[1]) > 2e-2 * l.abs().max() for l in lg_ref[:-1]):
        assert t_e == t_ref


def test_fused_decode_step_rejects_small_workspace():
    from g2vlm_b200 import ops
    a = ops.DecodeStepArgs()
    a.num_layers, a.hidden, a.intermediate, a.n_q_heads, a.n_kv_heads, a.head_dim, a.vocab = 1, 256, 512, 2, 1, 128, 512
    ws = torch.zeros(64, dtype=torch.uint8, device="cuda")
    a.fused_ws, a.fused_ws_bytes = ws.data_ptr(), ws.numel()
    dummy = torch.zeros(8, device="cuda")
    layers = (ops.UndLayerWeights * 1)()
    import ctypes
    kv = (ctypes.c_void_p * 1)(dummy.data_ptr())
    a.layers, a.kv, a.kv_capacity, a.kv_bound = layers, kv, 8, 4
    with pytest.raises(RuntimeError, match="workspace"):
        ops.und_decode_step(a)
    assert ops.und_decode_workspace_bytes(12, 2) > 74 * 12 * 128 * 4
