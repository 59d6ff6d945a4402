"""Parity of the memory-bound kernels (through the C ABI) against the oracle restatement / plain torch."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _err(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def test_gather_scatter_rows_roundtrip_exact():
    from g2vlm_b200 import ops
    g = torch.Generator().manual_seed(0)
    src = torch.randn(1000, 1536, generator=g).cuda()
    perm = torch.randperm(1000, generator=g).cuda()
    dst = torch.empty_like(src)
    ops.gather_rows(src, dst, perm, 1000)
    assert torch.equal(dst, src[perm])
    back = torch.empty_like(src)
    ops.gather_rows(dst, back, perm, 1000, scatter=True)
    assert torch.equal(back, src)                     # scatter o gather = identity (permutation round trip)
    ops.gather_rows(src, dst, None, 0)                # empty input is a no-op
    b = torch.randn(7, 256, generator=g).to(torch.bfloat16).cuda()
    wide = torch.zeros(20, 2048, dtype=torch.bfloat16, device="cuda")
    ops.gather_rows(b, wide[13:, 1536:1792], None, 7)  # strided destination (KV prefix copy)
    assert torch.equal(wide[13:, 1536:1792], b) and wide[:13].abs().sum() == 0 and wide[:, :1536].abs().sum() == 0


@pytest.mark.parametrize("rows,dim,n_first", [(777, 1536, 700), (300, 1024, 100), (33, 256, 0), (5, 1024, 5), (100, 2048, 40)])
def test_rmsnorm_routed(rows, dim, n_first):
    from g2vlm_b200 import ops
    from oracle import restate
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(rows, dim, generator=g) * 3).cuda()
    wa, wb = (torch.rand(dim, generator=g) + 0.5).cuda(), (torch.rand(dim, generator=g) + 0.5).cuda()
    ref = torch.cat([restate.rmsnorm(x[:n_first].cpu(), wa.cpu(), 1e-6), restate.rmsnorm(x[n_first:].cpu(), wb.cpu(), 1e-6)])
    o32 = torch.empty(rows, dim, device="cuda")
    ops.rmsnorm_routed(x, o32, wa, wb, n_first, 1e-6)
    assert _err(o32, ref) < 1e-5
    o16 = torch.empty(rows, dim, device="cuda", dtype=torch.bfloat16)
    ops.rmsnorm_routed(x, o16, wa, wb, n_first, 1e-6)
    assert (o16.float().cpu() - ref.to(torch.bfloat16).float()).abs().max() <= ref.abs().max() * 2 ** -8


@pytest.mark.parametrize("D", [1024, 1536, 512])   # register-cached kernels (1024, 1536) and the generic one
def test_layernorm_and_segment_drop(D):
    from g2vlm_b200 import ops
    g = torch.Generator().manual_seed(2)
    n, S = 3, 190
    x = (torch.randn(n * S, D, generator=g) * 2 + 0.3).cuda()
    w, b = (torch.rand(D, generator=g) + 0.5).cuda(), (torch.randn(D, generator=g) * 0.1).cuda()
    ref = torch.nn.functional.layer_norm(x.cpu(), (D,), w.cpu(), b.cpu(), 1e-6)
    out = torch.empty(n * S, D, device="cuda")
    ops.layernorm(x, out, w, b, 1e-6)
    assert _err(out, ref) < 1e-5
    comp = torch.empty(n * (S - 5), D, device="cuda", dtype=torch.bfloat16)
    ops.layernorm(x, comp, w, b, 1e-6, seg_in=S, seg_skip=5)
    ref_c = ref.view(n, S, D)[:, 5:].reshape(-1, D)
    assert _err(comp, ref_c) < 5e-3


@pytest.mark.parametrize("nq,nkv", [(12, 2), (2, 1)])   # row-per-warp kernel (14 heads) and the per-head kernel
def test_mrope_table_and_qknorm_mrope(nq, nkv):
    from g2vlm_b200 import ops
    from oracle import restate
    g = torch.Generator().manual_seed(3)
    T, hd, n_first = 600, 128, 550
    pos = torch.stack([torch.randint(0, 700, (T,), generator=g) for _ in range(3)])
    inv = (1.0 / (1e6 ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))).cuda()
    cos, sin = torch.empty(T, 64, device="cuda"), torch.empty(T, 64, device="cuda")
    ops.mrope_table(pos.cuda(), inv, cos, sin, (16, 24, 24))
    cr, sr = restate.mrope_cos_sin(pos, hd, 1e6, (16, 24, 24))
    assert (cos.cpu() - cr[:, :64]).abs().max() < 2e-6 and (sin.cpu() - sr[:, :64]).abs().max() < 2e-6
    assert torch.equal(cr[:, :64], cr[:, 64:])        # the table really only has 64 distinct columns
    qkv = torch.randn(T, (nq + 2 * nkv) * hd, generator=g).to(torch.bfloat16)
    ws = [(torch.rand(hd, generator=g) + 0.5) for _ in range(4)]  # q_geo, k_geo, q_und, k_und
    for round_normed in (False, True):
        buf = qkv.clone().cuda()
        ops.qknorm_mrope(buf, T, n_first, nq, nkv, hd, *[w.cuda() for w in ws], cos, sin, 1e-6, round_normed=round_normed)
        x = qkv.float().view(T, nq + 2 * nkv, hd)

        def norm(t, w):
            n = t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + 1e-6)
            if round_normed:
                n = n.to(torch.bfloat16).float()
            return w * n
        q = torch.cat([norm(x[:n_first, :nq], ws[0]), norm(x[n_first:, :nq], ws[2])])
        k = torch.cat([norm(x[:n_first, nq:nq + nkv], ws[1]), norm(x[n_first:, nq:nq + nkv], ws[3])])
        c, s = cr[:, None, :], sr[:, None, :]
        qe = q * c + restate.rotate_half(q) * s
        ke = k * c + restate.rotate_half(k) * s
        ref = torch.cat([qe, ke, x[:, nq + nkv:]], dim=1).reshape(T, -1)
        got = buf.float().cpu()
        assert _err(got, ref) < 1e-2
        assert torch.equal(got[:, (nq + nkv) * hd:], qkv.float()[:, (nq + nkv) * hd:])   # V untouched


def test_im2col_and_dino_embed_match_oracle():
    from g2vlm_b200 import ops, schema
    from oracle import restate
    cfg = schema.TINY
    sd = schema.init_synthetic(cfg, seed=0)
    img = schema.synthetic_views(2, 42, 518, seed=4)
    n, P, D = 2, 3 * 37, cfg.dino_hidden
    patches = torch.empty(n * P, 640, device="cuda", dtype=torch.bfloat16)
    ops.im2col_patches(img.cuda(), patches, 14)
    ref_p = img.reshape(n, 3, 3, 14, 37, 14).permute(0, 2, 4, 1, 3, 5).reshape(n * P, 588)
    assert torch.equal(patches[:, :588].float().cpu(), ref_p.to(torch.bfloat16).float())
    assert patches[:, 588:].abs().sum() == 0
    ref = restate.dino_embeddings(sd, cfg, img, "bf16")
    wt = sd["dino_model.embeddings.patch_embeddings.projection.weight"].reshape(D, -1)
    emb = restate.linear(ref_p, wt, sd["dino_model.embeddings.patch_embeddings.projection.bias"], "bf16")
    pos = sd["dino_model.embeddings.position_embeddings"]
    pp = pos[:, 1:].reshape(1, 37, 37, -1).permute(0, 3, 1, 2)
    pp = torch.nn.functional.interpolate(pp, size=(3, 37), mode="bicubic", align_corners=False, antialias=True)
    pos = torch.cat((pos[:, :1], pp.permute(0, 2, 3, 1).reshape(1, P, -1)), dim=1)[0].contiguous()
    out = torch.empty(n * (P + 5), D, device="cuda")
    ops.dino_embed(emb.to(torch.bfloat16).cuda(), sd["dino_model.embeddings.cls_token"].reshape(D).cuda(),
                   sd["dino_model.embeddings.register_tokens"].reshape(4, D).contiguous().cuda(), pos.cuda(), out, n, P, 4)
    assert _err(out, ref.reshape(-1, D)) < 1e-6


@pytest.mark.parametrize("hd,gh,gw", [(96, 5, 37), (16, 3, 4)])
def test_rope2d_matches_reference_bf16_quirk(hd, gh, gw):
    from g2vlm_b200 import ops
    from oracle import restate
    g = torch.Generator().manual_seed(5)
    B, heads, P = 2, 4, gh * gw
    hp = 128 if hd > 64 else 64
    t = torch.randn(B, heads, P, hd, generator=g).to(torch.bfloat16).float()
    yy, xx = torch.meshgrid(torch.arange(gh), torch.arange(gw), indexing="ij")
    pos = torch.stack([yy.flatten(), xx.flatten()], -1)[None].expand(B, -1, -1)
    ref = restate.rope2d(t, pos, 100.0, "bf16")                      # (B, heads, P, hd)
    buf = torch.zeros(B * P, heads * hp, dtype=torch.bfloat16)
    buf.view(B, P, heads, hp)[..., :hd] = t.permute(0, 2, 1, 3).to(torch.bfloat16)
    D = hd // 2
    inv_freq = 1.0 / (100.0 ** (torch.arange(0, D, 2).float() / D))
    fr = torch.einsum("i,j->ij", torch.arange(max(gh, gw)).float(), inv_freq).to(torch.bfloat16)
    buf = buf.cuda()
    ops.rope2d(buf, B * P, heads, hp, hd, P, gw, fr.cos().float().cuda(), fr.sin().float().cuda())
    got = buf.view(B, P, heads, hp)[..., :hd].permute(0, 2, 1, 3).float().cpu()
    assert (got - ref).abs().max() <= 2 ** -7 * ref.abs().max()      # at most 1 bf16 ulp of the largest value
    assert (got != ref).float().mean() < 0.02
    assert buf.view(B, P, heads, hp)[..., hd:].abs().sum() == 0      # padding columns stay zero


def test_points_epilogue_matches_pixel_shuffle():
    from g2vlm_b200 import ops
    from oracle import restate
    g = torch.Generator().manual_seed(6)
    n, H, W, p = 2, 42, 70, 14
    P = (H // p) * (W // p)
    feat = torch.randn(n * P, 588, generator=g) * 0.3
    poses = torch.eye(4).repeat(n, 1, 1)
    poses[:, :3, :4] = torch.randn(n, 3, 4, generator=g)
    ps = torch.nn.functional.pixel_shuffle(feat.view(n, P, 588).transpose(1, 2).reshape(n, 588, H // p, W // p), p).permute(0, 2, 3, 1)
    xy, z = ps.split([2, 1], -1)
    z = z.exp()
    local = torch.cat([xy * z, z], -1)
    pts = torch.einsum("nij,nhwj->nhwi", poses, torch.cat([local, torch.ones_like(z)], -1))[..., :3]
    o0 = torch.empty(n, H, W, 3, device="cuda"); o1 = torch.empty_like(o0); o2 = torch.empty_like(o0)
    ops.points_epilogue(feat.cuda(), poses.cuda(), o0, o1, n, H, W, p, 1)
    ops.points_epilogue(feat.cuda(), None, o2, None, n, H, W, p, 0)
    assert torch.equal(o2.cpu(), ps)
    assert _err(o0, local) < 1e-6 and _err(o1, pts) < 1e-5


def test_split3_reconstructs_fp32_and_fp32_linear_accuracy():
    from g2vlm_b200 import ops
    from g2vlm_b200.model import _split_hi_lo_hi
    g = torch.Generator().manual_seed(7)
    rows, K, N = 300, 512, 512
    x = torch.randn(rows, K, generator=g); w = torch.randn(N, K, generator=g) * 0.05; b = torch.randn(N, generator=g)
    xs = torch.empty(rows, 3 * K, device="cuda", dtype=torch.bfloat16)
    ops.split3(x.cuda(), xs)
    hi, hi2, lo = xs[:, :K].float().cpu(), xs[:, K:2 * K].float().cpu(), xs[:, 2 * K:].float().cpu()
    assert torch.equal(hi, hi2) and (hi + lo - x).abs().max() < 2 ** -16 * x.abs().max()
    out = torch.empty(rows, N, device="cuda")
    ops.gemm(xs, _split_hi_lo_hi(w).cuda(), out, epilogue=ops.EPI_STORE_F32, bias=b.cuda())
    ref = (x.double() @ w.double().T + b.double()).float()
    assert _err(out, ref) < 2e-5                                     # fp32-class accuracy from bf16 tensor cores


def test_mean_pool_and_cast():
    from g2vlm_b200 import ops
    g = torch.Generator().manual_seed(8)
    x = torch.randn(3 * 185, 512, generator=g).cuda()
    out = torch.empty(3, 512, device="cuda")
    ops.mean_pool(x, out, 3, 185)
    assert _err(out, x.view(3, 185, 512).mean(1)) < 1e-5
    c = torch.empty(3 * 185, 512, device="cuda", dtype=torch.bfloat16)
    ops.cast_bf16(x, c)
    assert torch.equal(c, x.to(torch.bfloat16))


def test_camera_pose_matches_svd_orthogonalize():
    from g2vlm_b200 import ops
    from oracle import restate
    g = torch.Generator().manual_seed(9)
    n, C = 64, 512
    f = torch.randn(n, C, generator=g)
    wt, bt = torch.randn(3, C, generator=g) * 0.05, torch.randn(3, generator=g)
    wr, br = torch.randn(9, C, generator=g) * 0.05, torch.randn(9, generator=g)
    poses = torch.empty(n, 4, 4, device="cuda")
    ops.camera_pose(f.cuda(), wt.cuda(), bt.cuda(), wr.cuda(), br.cuda(), poses)
    R_ref = restate.svd_orthogonalize(f @ wr.T + br)
    t_ref = f @ wt.T + bt
    P = poses.cpu()
    assert (P[:, :3, 3] - t_ref).abs().max() < 1e-4
    assert (P[:, 3] - torch.tensor([0.0, 0, 0, 1])).abs().max() == 0
    # compare rotations by geodesic angle (SVD sign/order ambiguities cancel in R)
    cosang = ((torch.einsum("nij,nij->n", P[:, :3, :3], R_ref) - 1) / 2).clamp(-1, 1)
    assert torch.acos(cosang).max() < 2e-3
    det = torch.det(P[:, :3, :3])
    assert (det - 1).abs().max() < 1e-4                               # proper rotations, also for reflected inputs
    ortho = P[:, :3, :3] @ P[:, :3, :3].transpose(1, 2)
    assert (ortho - torch.eye(3)).abs().max() < 1e-5
