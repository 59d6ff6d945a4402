"""Decode-shaped kernels (row f1): the GEMV path of g2vlm_gemm_bf16 (<= 8 rows) and the split-K decode
attention, against fp32 torch references / the tensor-core kernels."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _relerr(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("rows", [1, 2, 3, 7, 8])
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
    r0 = torch.randn(rows, N, device="cuda")
    r = r0.clone()
    ops.gemm(x, w, r, epilogue=ops.EPI_RESID_F32, groups=groups, scale=gamma, scale_groups=2,
             flags=ops.GEMM_ROUND_AFTER_SCALE)
    assert _relerr(r, r0 + (ref.to(torch.bfloat16).float() * gamma).to(torch.bfloat16).float()) < 5e-3
    f = torch.empty(rows, N, device="cuda")
    ops.gemm(x, w, f, epilogue=ops.EPI_STORE_F32, groups=groups, bias=bias, flags=ops.GEMM_RELU, residual=r0)
    assert _relerr(f, torch.relu(ref + bias[N:]) + r0) < 1e-4
    # SwiGLU: interleaved gate/up blocks of 128 rows
    I = 512
    wg = (torch.randn(I, K, generator=g) * 0.08).to(torch.bfloat16).cuda()
    wu = (torch.randn(I, K, generator=g) * 0.08).to(torch.bfloat16).cuda()
    wi = torch.stack([wg.view(I // 128, 128, K), wu.view(I // 128, 128, K)], 1).reshape(2 * I, K).contiguous()
    o = torch.empty(rows, I, device="cuda", dtype=torch.bfloat16)
    ops.gemm(x, wi, o, epilogue=ops.EPI_SWIGLU_BF16)
    gt, up = (x.float() @ wg.float().T).to(torch.bfloat16), (x.float() @ wu.float().T).to(torch.bfloat16)
    assert _relerr(o, torch.nn.functional.silu(gt.float()).to(torch.bfloat16).float() * up.float()) < 1.5e-2


@pytest.mark.parametrize("rows,K,N", [(5, 1536, 648), (7, 1552, 640), (4, 8960, 1536), (7, 4096, 48)])
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
