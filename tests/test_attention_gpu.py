"""Parity of the tcgen05 attention kernel (through the C ABI) against an fp32 torch reference:
softmax(scale * q k^T) v per segment with GQA, bottom-right causal mask, ragged segments."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_attention(q, k, v, cu_q, cu_k, hq, hk, d, scale, causal):
    out = torch.zeros(q.shape[0], hq * d, device=q.device, dtype=torch.float32)
    qf, kf, vf = q.float(), k.float(), v.float()
    for i in range(len(cu_q) - 1):
        qs = qf[cu_q[i]:cu_q[i + 1]].view(-1, hq, d).transpose(0, 1)
        ks = kf[cu_k[i]:cu_k[i + 1]].view(-1, hk, d).transpose(0, 1).repeat_interleave(hq // hk, 0)
        vs = vf[cu_k[i]:cu_k[i + 1]].view(-1, hk, d).transpose(0, 1).repeat_interleave(hq // hk, 0)
        s = (qs @ ks.transpose(1, 2)) * scale
        if causal:
            lq, lk = s.shape[1], s.shape[2]
            mask = torch.ones(lq, lk, dtype=torch.bool, device=q.device).tril(diagonal=lk - lq)
            s = s.masked_fill(~mask, float("-inf"))
        p = torch.softmax(s, dim=-1)
        o = p @ vs
        out[cu_q[i]:cu_q[i + 1]] = o.transpose(0, 1).reshape(-1, hq * d)
    return out


def _run(cu_q, cu_k, hq, hk, d, causal=False, seed=0, q_rows=None, kv_rows=None, scale=None, std=1.0):
    from g2vlm_b200 import ops
    g = torch.Generator().manual_seed(seed)
    q_rows = q_rows or cu_q[-1]
    kv_rows = kv_rows or cu_k[-1]
    q = (torch.randn(q_rows, hq * d, generator=g) * std).to(torch.bfloat16).cuda()
    k = (torch.randn(kv_rows, hk * d, generator=g) * std).to(torch.bfloat16).cuda()
    v = torch.randn(kv_rows, hk * d, generator=g).to(torch.bfloat16).cuda()
    out = torch.full((q_rows, hq * d), 7.0, device="cuda", dtype=torch.bfloat16)
    work = ops.attention_work_table(cu_q, cu_k).cuda()
    scale = scale or 1.0 / math.sqrt(d)
    ops.attention(q, k, v, out, work, num_q_heads=hq, num_kv_heads=hk, head_dim=d, scale=scale, causal=causal)
    torch.cuda.synchronize()
    ref = _ref_attention(q, k, v, cu_q, cu_k, hq, hk, d, scale, causal)
    covered = torch.zeros(q_rows, dtype=torch.bool, device="cuda")
    for i in range(len(cu_q) - 1):
        covered[cu_q[i]:cu_q[i + 1]] = True
    err = (out[covered].float() - ref[covered]).abs().max().item()
    assert torch.isfinite(out.float()).all()
    assert (out[~covered] == 7.0).all(), "rows outside every segment must not be written"
    return err, ref[covered].abs().max().item()


@pytest.mark.parametrize("lq,lk,hq,hk,d", [
    (128, 128, 1, 1, 128),     # one block, one tile
    (256, 256, 2, 1, 128),     # both tiles, GQA
    (300, 1000, 12, 2, 128),   # ragged q and k tails, MoT head layout
    (1371 * 2 + 7, 1371 * 2 + 7, 12, 2, 128),
    (700, 700, 4, 4, 64),      # DINO head dim
])
def test_single_segment(lq, lk, hq, hk, d):
    err, mag = _run([0, lq], [0, lk], hq, hk, d)
    assert err < 2e-2 * max(mag, 1.0), (err, mag)


def test_sharp_softmax_triggers_rescale():
    # large logits: running max jumps by more than 2^8 between blocks -> lazy rescale path
    err, mag = _run([0, 256], [0, 1024], 2, 1, 128, std=4.0, seed=3)
    assert err < 3e-2 * max(mag, 1.0), (err, mag)


def test_multi_segment_dino_quirk_layout():
    # N=3 "images": segments of P=777 over a sequence of 3*(777+5) rows; last 15 rows uncovered
    P, N = 777, 3
    cu = [i * P for i in range(N + 1)]
    err, mag = _run(cu, cu, 4, 4, 64, q_rows=N * (P + 5), kv_rows=N * (P + 5))
    assert err < 2e-2 * max(mag, 1.0)


def test_cross_attention_to_view0():
    # every view's queries attend to view 0's keys (Pi3 CrossBlockRope context)
    P, N = 500, 3
    cu_q = [i * P for i in range(N + 1)]
    from g2vlm_b200 import ops
    g = torch.Generator().manual_seed(5)
    hq = hk = 2
    d = 128
    q = torch.randn(N * P, hq * d, generator=g).to(torch.bfloat16).cuda()
    k = torch.randn(P, hk * d, generator=g).to(torch.bfloat16).cuda()
    v = torch.randn(P, hk * d, generator=g).to(torch.bfloat16).cuda()
    out = torch.empty(N * P, hq * d, device="cuda", dtype=torch.bfloat16)
    items = []
    for i in range(N):
        for t0 in range(cu_q[i], cu_q[i + 1], 256):
            items.append([t0, cu_q[i], cu_q[i + 1], 0, P, 0, 0, 0])
    work = torch.tensor(items, dtype=torch.int32).cuda()
    scale = 96 ** -0.5
    ops.attention(q, k, v, out, work, num_q_heads=hq, num_kv_heads=hk, head_dim=d, scale=scale)
    ref = torch.cat([_ref_attention(q[cu_q[i]:cu_q[i + 1]], k, v, [0, P], [0, P], hq, hk, d, scale, False) for i in range(N)])
    assert (out.float() - ref).abs().max().item() < 2e-2


@pytest.mark.parametrize("lq,lk", [(7, 7), (300, 300), (100, 356)])
def test_causal(lq, lk):
    err, mag = _run([0, lq], [0, lk], 2, 1, 128, causal=True, seed=2)
    assert err < 2e-2 * max(mag, 1.0)


def test_strided_fused_qkv_buffer_and_timing():
    """Config-2 MoT shape: q/k/v are column slices of one [T+K0, 2048] buffer; 12:2 GQA, d=128."""
    from g2vlm_b200 import ops
    T, K0 = 16 * 1371, 7
    g = torch.Generator().manual_seed(1)
    qkv = torch.randn(T + K0, 2048, generator=g).to(torch.bfloat16).cuda()
    q, k, v = qkv[:, :1536], qkv[:, 1536:1792], qkv[:, 1792:]
    out = torch.empty(T, 1536, device="cuda", dtype=torch.bfloat16)
    work = ops.attention_work_table([0, T], [0, T + K0]).cuda()
    scale = 1 / math.sqrt(128)
    kw = dict(num_q_heads=12, num_kv_heads=2, head_dim=128, scale=scale)
    for _ in range(2):
        ops.attention(q[:T], k, v, out, work, **kw)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(5):
        ops.attention(q[:T], k, v, out, work, **kw)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 5
    print(f"\nattention T={T}: {ms:.3f} ms  {4 * T * (T + K0) * 12 * 128 / ms / 1e9:.1f} TFLOP/s")
    rows = torch.arange(0, T, 997, device="cuda")[:16]
    qs = q[rows].float().view(-1, 12, 128).transpose(0, 1)
    ks = k.float().view(-1, 2, 128).transpose(0, 1).repeat_interleave(6, 0)
    vs = v.float().view(-1, 2, 128).transpose(0, 1).repeat_interleave(6, 0)
    ref = (torch.softmax(qs @ ks.transpose(1, 2) * scale, -1) @ vs).transpose(0, 1).reshape(-1, 1536)
    assert (out[rows].float() - ref).abs().max().item() < 2e-2


def test_per_item_causal_flag_mixes_masks_in_one_launch():
    """Work items carry their own causal flag (column 5): a non-causal segment over all keys and a causal segment
    over the last keys in ONE launch (the fused prompt prefill of the recon path)."""
    from g2vlm_b200 import ops
    hq, hk, d, T, Kp = 12, 2, 128, 300, 7
    g = torch.Generator().manual_seed(5)
    q = torch.randn(T + Kp, hq * d, generator=g).to(torch.bfloat16).cuda()
    k = torch.randn(T + Kp, hk * d, generator=g).to(torch.bfloat16).cuda()
    v = torch.randn(T + Kp, hk * d, generator=g).to(torch.bfloat16).cuda()
    items = ops.attention_work_table([0, T], [0, T + Kp]).tolist() + [[T, T, T + Kp, T, T + Kp, 1, 0, 0]]
    work = torch.tensor(items, dtype=torch.int32).cuda()
    out = torch.zeros(T + Kp, hq * d, device="cuda", dtype=torch.bfloat16)
    scale = 1 / math.sqrt(d)
    ops.attention(q, k, v, out, work, num_q_heads=hq, num_kv_heads=hk, head_dim=d, scale=scale)
    ref_a = _ref_attention(q[:T], k, v, [0, T], [0, T + Kp], hq, hk, d, scale, False)
    ref_b = _ref_attention(q[T:], k[T:], v[T:], [0, Kp], [0, Kp], hq, hk, d, scale, True)
    assert (out[:T].float() - ref_a).abs().max() < 2e-2
    assert (out[T:].float() - ref_b).abs().max() < 2e-2


@pytest.mark.parametrize("hot_key", [10, 100, 128 + 10, 128 + 100, 256 + 70, 383])
def test_lazy_rescale_paths(hot_key):
    """A score far above the running max (> 2^8 in the scaled log2 domain) appearing in the first half, the second
    half, or a later block forces the O rescale through each of its code paths (early check on the first 64 keys,
    late check after the first two P chunks were already computed)."""
    from g2vlm_b200 import ops
    hq, hk, d, T = 2, 1, 128, 384
    g = torch.Generator().manual_seed(hot_key)
    q = (torch.randn(T, hq * d, generator=g) * 0.3).to(torch.bfloat16)
    k = (torch.randn(T, hk * d, generator=g) * 0.3).to(torch.bfloat16)
    v = torch.randn(T, hk * d, generator=g).to(torch.bfloat16)
    k[hot_key] = (q[5, :d].float() * 40).to(torch.bfloat16)          # row 5 (head 0) gets a huge score there
    k[min(hot_key + 1, T - 1)] = (q[200, d:].float() * 25).to(torch.bfloat16)   # and row 200 (head 1) next to it
    q, k, v = q.cuda(), k.cuda(), v.cuda()
    work = ops.attention_work_table([0, T], [0, T]).cuda()
    out = torch.zeros(T, hq * d, device="cuda", dtype=torch.bfloat16)
    scale = 1 / math.sqrt(d)
    ops.attention(q, k, v, out, work, num_q_heads=hq, num_kv_heads=hk, head_dim=d, scale=scale)
    ref = _ref_attention(q, k, v, [0, T], [0, T], hq, hk, d, scale, False)
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max() < 3e-2


def test_compact_head_output():
    """out_head_cols: 96-wide heads computed in zero-padded 128-wide slots leave as [rows, heads*96]."""
    from g2vlm_b200 import ops
    hq, d, dv, T = 4, 128, 96, 300
    g = torch.Generator().manual_seed(9)
    def padded():
        x = torch.zeros(T, hq, d)
        x[:, :, :dv] = torch.randn(T, hq, dv, generator=g)
        return x.reshape(T, hq * d).to(torch.bfloat16).cuda()
    q, k, v = padded(), padded(), padded()
    work = ops.attention_work_table([0, 140, T], [0, 140, T]).cuda()
    scale = 1 / math.sqrt(dv)
    full = torch.zeros(T, hq * d, device="cuda", dtype=torch.bfloat16)
    ops.attention(q, k, v, full, work, num_q_heads=hq, num_kv_heads=hq, head_dim=d, scale=scale)
    compact = torch.full((T + 1, hq * dv), 5.0, device="cuda", dtype=torch.bfloat16)
    ops.attention(q, k, v, compact[:T], work, num_q_heads=hq, num_kv_heads=hq, head_dim=d, scale=scale, out_head_cols=dv)
    assert torch.equal(compact[:T].view(T, hq, dv), full.view(T, hq, d)[:, :, :dv])
    assert bool((compact[T] == 5.0).all())


@pytest.mark.parametrize("d,causal", [(128, False), (64, False), (128, True), (64, True)])
def test_persistent_ctas_walk_many_ragged_units(d, causal):
    """More (item, head) units than SMs, with segment lengths from 1 row to several tiles and key counts from one
    block to many: every CTA chains units of different length through the same TMEM / barrier state (unit boundary
    hand-offs: q_empty, o_free, s_free, running mbarrier parities)."""
    lens_q = [1, 130, 700, 5, 1371, 256, 257, 64, 900, 3, 512, 129]
    lens_k = [300, 130, 700, 40, 1371, 1, 257, 1000, 900, 3, 127, 2048] if not causal else lens_q
    cu_q = [0]
    cu_k = [0]
    for a, b in zip(lens_q, lens_k):
        cu_q.append(cu_q[-1] + a)
        cu_k.append(cu_k[-1] + b)
    err, mag = _run(cu_q, cu_k, 16, 16 if d == 64 else 4, d, causal=causal, seed=d + causal)
    assert err < 2e-2 * max(mag, 1.0), (err, mag)


@pytest.mark.parametrize("d", [64, 128])
def test_causal_units_without_any_visible_key(d):
    """Bottom-right causal mask with fewer keys than queries: the first query rows see no key at all (whole units
    with zero key blocks are skipped by every role and written as zeros, like flash-attn)."""
    from g2vlm_b200 import ops
    lq, lk, hq = 700, 100, 2
    g = torch.Generator().manual_seed(d)
    q = torch.randn(lq, hq * d, generator=g).to(torch.bfloat16).cuda()
    k = torch.randn(lk, hq * d, generator=g).to(torch.bfloat16).cuda()
    v = torch.randn(lk, hq * d, generator=g).to(torch.bfloat16).cuda()
    out = torch.full((lq, hq * d), 7.0, device="cuda", dtype=torch.bfloat16)
    work = ops.attention_work_table([0, lq], [0, lk]).cuda()
    scale = 1 / math.sqrt(d)
    ops.attention(q, k, v, out, work, num_q_heads=hq, num_kv_heads=hq, head_dim=d, scale=scale, causal=True)
    assert bool((out[: lq - lk] == 0).all())
    ref = _ref_attention(q[lq - lk:], k, v, [0, lk], [0, lk], hq, hq, d, scale, True)
    assert (out[lq - lk:].float() - ref).abs().max() < 2e-2


@pytest.mark.parametrize("d,hq,hk,lq,k_split,max_ctas", [
    (128, 12, 2, 700, (391, 1400), 0),       # MoT head layout, ragged split, full grid
    (128, 12, 2, 300, (7, 2048), 100),       # tiny local key set (the 7 prefix rows), capped persistent grid
    (64, 4, 4, 520, (520, 777), 37),
])
def test_lse_output_and_two_way_merge_equal_one_pass(d, hq, hk, lq, k_split, max_ctas):
    """View-sharded v2 path: attention over two disjoint key sets + g2vlm_attention_merge == one pass over all keys;
    the lse output is ln(sum exp(scale q.k)) of the fp32 reference."""
    from g2vlm_b200 import ops
    g = torch.Generator().manual_seed(9)
    ka, kb = k_split
    q = torch.randn(lq, hq * d, generator=g).to(torch.bfloat16).cuda()
    k = torch.randn(ka + kb, hk * d, generator=g).to(torch.bfloat16).cuda()
    v = torch.randn(ka + kb, hk * d, generator=g).to(torch.bfloat16).cuda()
    scale = 1.0 / math.sqrt(d)
    o_a = torch.empty(lq, hq * d, device="cuda", dtype=torch.bfloat16)
    o_b, o_all = torch.empty_like(o_a), torch.empty_like(o_a)
    lse_a = torch.full((lq, hq), float("nan"), device="cuda")
    lse_b, lse_all = lse_a.clone(), lse_a.clone()
    kw = dict(num_q_heads=hq, num_kv_heads=hk, head_dim=d, scale=scale)
    ops.attention(q, k[:ka], v[:ka], o_a, ops.attention_work_table([0, lq], [0, ka]).cuda(), lse=lse_a, max_ctas=max_ctas, **kw)
    ops.attention(q, k[ka:], v[ka:], o_b, ops.attention_work_table([0, lq], [0, kb]).cuda(), lse=lse_b, **kw)
    ops.attention(q, k, v, o_all, ops.attention_work_table([0, lq], [0, ka + kb]).cuda(), lse=lse_all, **kw)
    merged = torch.empty_like(o_a)
    ops.attention_merge(o_a, lse_a, o_b, lse_b, merged, hq, d)
    inplace = o_a.clone()
    ops.attention_merge(inplace, lse_a, o_b, lse_b, inplace, hq, d)      # out may alias o_a
    torch.cuda.synchronize()
    assert torch.equal(merged, inplace)
    ref = _ref_attention(q, k, v, [0, lq], [0, ka + kb], hq, hk, d, scale, False)
    mag = ref.abs().max().item()
    assert (o_all.float() - ref).abs().max().item() < 2e-2 * max(mag, 1.0)
    assert (merged.float() - ref).abs().max().item() < 2e-2 * max(mag, 1.0)
    assert (merged.float() - o_all.float()).abs().max().item() < 1.5e-2 * max(mag, 1.0)   # two extra bf16 roundings
    # log-sum-exp against fp32
    qs = q.float().view(lq, hq, d).transpose(0, 1)
    ks = k.float().view(-1, hk, d).transpose(0, 1).repeat_interleave(hq // hk, 0)
    want = torch.logsumexp((qs @ ks.transpose(1, 2)) * scale, dim=-1).transpose(0, 1)
    assert (lse_all - want).abs().max().item() < 2e-3
    want_a = torch.logsumexp((qs @ ks[:, :ka].transpose(1, 2)) * scale, dim=-1).transpose(0, 1)
    assert (lse_a - want_a).abs().max().item() < 2e-3


def test_merge_ignores_an_empty_partial():
    """A causal row that sees no key in one partial (lse = -inf, zeros) must take the other partial unchanged."""
    from g2vlm_b200 import ops
    hq, d, rows = 2, 128, 64
    o_a = torch.randn(rows, hq * d, device="cuda").to(torch.bfloat16)
    o_b = torch.full_like(o_a, float("nan"))                          # never read with weight > 0
    lse_a = torch.randn(rows, hq, device="cuda")
    lse_b = torch.full((rows, hq), float("-inf"), device="cuda")
    out = torch.empty_like(o_a)
    ops.attention_merge(o_a, lse_a, o_b, lse_b, out, hq, d)
    assert torch.equal(out, o_a)
    lse_a.fill_(float("-inf"))
    ops.attention_merge(o_a, lse_a, o_b, lse_b, out, hq, d)
    assert bool((out == 0).all())
