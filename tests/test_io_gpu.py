"""PLY export (row f3): GPU filtering + packing against a numpy restatement of the reference's
save_ply_visualization post-processing (g2vlm_utils.py:119-143)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_ply_pack_matches_numpy_filtering(tmp_path):
    from g2vlm_b200 import io
    g = torch.Generator().manual_seed(0)
    N, H, W = 3, 70, 98                                   # 20580 points: > 20 scan blocks, ragged last block
    pts = torch.randn(N, H, W, 3, generator=g) * 5
    img = (torch.randint(0, 256, (N, 3, H, W), generator=g).float() / 255.0)
    bad = torch.rand(N, H, W, generator=g) < 0.1
    pts[bad] = torch.tensor([float("nan"), 1.0, 2.0])
    pts[0, 0, 0] = torch.tensor([float("inf"), 0.0, 0.0])
    pts[2, 69, 97] = torch.tensor([0.0, float("-inf"), 0.0])
    pred = dict(points=pts[None].cuda(), images=img[None].cuda())
    path = str(tmp_path / "out" / "scene.ply")
    n = io.save_ply_visualization(pred, path)
    # numpy restatement of the reference
    p = pts.numpy().reshape(-1, 3)
    c = img.permute(0, 2, 3, 1).numpy().reshape(-1, 3)
    valid = ~(np.any(np.isnan(p), axis=1) | np.any(np.isinf(p), axis=1))
    p, c = p[valid], c[valid]
    assert n == len(p)
    rec = io.read_ply(path)
    assert len(rec) == n
    got = np.stack([rec["x"], rec["y"], rec["z"]], 1)
    assert np.array_equal(got, p.astype(np.float64))      # order preserved, fp32 -> double exact
    col = np.stack([rec["red"], rec["green"], rec["blue"]], 1)
    expect = np.rint(np.clip(c.astype(np.float64), 0.0, 1.0) * 255.0).astype(np.uint8)   # Open3D ColorToUint8
    assert np.array_equal(col, expect)
    # colours are 8-bit pixel values k / 255 in float32: every one must come back as exactly k (a truncating
    # conversion returns k - 1 for roughly half of them)
    k = torch.arange(256, dtype=torch.float32)
    assert np.array_equal(np.rint((k / 255.0).double().numpy() * 255.0), k.numpy())
    src = (img.permute(0, 2, 3, 1).reshape(-1, 3)[torch.from_numpy(valid)] * 255.0).round().numpy().astype(np.uint8)
    assert np.array_equal(col, src)
    head = open(path, "rb").read(200).decode("ascii", "ignore")
    assert head.startswith("ply\nformat binary_little_endian 1.0") and f"element vertex {n}" in head


def test_ply_filter_nan_false_keeps_every_point(tmp_path):
    """The reference's filter_nan=False (g2vlm_utils.py:126): non-finite vertices are written as they are."""
    from g2vlm_b200 import io
    pts = torch.randn(1, 6, 7, 3)
    pts[0, 2, 3] = torch.tensor([float("nan"), 1.0, float("inf")])
    img = torch.rand(1, 3, 6, 7)
    pred = dict(points=pts[None].cuda(), images=img[None].cuda())
    path = str(tmp_path / "all.ply")
    assert io.save_ply_visualization(pred, path, filter_nan=False) == 42
    rec = io.read_ply(path)
    assert len(rec) == 42 and np.isnan(rec["x"][2 * 7 + 3]) and np.isinf(rec["z"][2 * 7 + 3])
    assert io.save_ply_visualization(pred, path, filter_nan=True) == 41


def test_ply_pack_all_invalid_and_large():
    from g2vlm_b200 import ops
    pts = torch.full((1, 4, 5, 3), float("nan"), device="cuda")
    img = torch.zeros(1, 3, 4, 5, device="cuda")
    rec, n = ops.ply_pack(pts, img)
    assert n == 0 and rec.numel() == 0
    # config-2 sized output: 16 x 518 x 518 points, all valid
    pts = torch.randn(16, 518, 518, 3, device="cuda")
    img = torch.rand(16, 3, 518, 518, device="cuda")
    rec, n = ops.ply_pack(pts, img)
    assert n == 16 * 518 * 518
    a = np.frombuffer(rec[:27 * 1000].cpu().numpy().tobytes(), dtype=np.dtype([("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("r", "u1"), ("g", "u1"), ("b", "u1")]))
    assert np.array_equal(a["x"], pts.view(-1, 3)[:1000, 0].double().cpu().numpy())
    tail = np.frombuffer(rec[-27:].cpu().numpy().tobytes(), dtype=a.dtype)
    assert tail["z"][0] == float(pts.view(-1, 3)[-1, 2])


@pytest.mark.parametrize("H,W", [(540, 960), (300, 518), (200, 300), (333, 777)])
def test_device_lanczos_resize_is_bit_identical_to_pillow(H, W):
    """g2vlm_resize_lanczos_u8 + ToTensor on the GPU == the reference's host path (PIL LANCZOS resize, ToTensor,
    identity antialias-bilinear), bit for bit: down-scaling, up-scaling, width already 518 (no horizontal pass)."""
    from PIL import Image

    from g2vlm_b200 import host_prep
    rng = np.random.default_rng(H + W)
    views = []
    for i in range(3):
        a = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        a[:, : W // 2] = np.linspace(0, 255, H)[:, None, None].astype(np.uint8) + i   # smooth half, noisy half
        views.append(Image.fromarray(a))
    ref = host_prep.load_and_resize14(views, 518)
    got = host_prep.load_and_resize14_device(views, 518, "cuda")
    assert got.shape == ref.shape and got.dtype == torch.float32
    assert torch.equal(got.cpu(), ref)


def test_device_resize_uint8_output_and_errors():
    from g2vlm_b200 import host_prep, ops
    from PIL import Image
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, (90, 130, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(a).resize((56, 42), Image.Resampling.LANCZOS))
    tabs = [tuple(torch.from_numpy(t).cuda() for t in host_prep.lanczos_tables(n_in, n_out)) for n_in, n_out in ((130, 56), (90, 42))]
    out = torch.zeros(42, 56, 3, dtype=torch.uint8, device="cuda")
    ops.resize_lanczos_u8(torch.from_numpy(a).cuda(), tabs[0], tabs[1], 42, 56, out_u8=out)
    assert np.array_equal(out.cpu().numpy(), ref)
    with pytest.raises(Exception):   # the width changes but no horizontal tables
        ops.resize_lanczos_u8(torch.from_numpy(a).cuda(), None, tabs[1], 42, 56, out_u8=out)
