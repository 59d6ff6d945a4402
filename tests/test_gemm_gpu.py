"""Parity of the tcgen05 grouped GEMM (through the C ABI) against an fp32 torch reference of the
same op with the reference's rounding points (nn.Linear under bf16 autocast returns bf16)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(rows, K, N, ng, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = (torch.randn(rows, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
    w = (torch.randn(ng * N, K, generator=g) * 0.05).to(torch.bfloat16).cuda()
    bias = (torch.randn(ng * N, generator=g) * 0.1).float().cuda()
    return a, w, bias


def _ref_linear(a, w, bias, groups, N):
    out = torch.zeros(a.shape[0], N, device=a.device, dtype=torch.float32)
    for g, (r0, rows) in enumerate(groups):
        wg = w[g * N:(g + 1) * N].float()
        y = a[r0:r0 + rows].float() @ wg.T
        if bias is not None:
            y = y + bias[g * N:(g + 1) * N]
        out[r0:r0 + rows] = y
    return out


def _relerr(x, ref):
    return ((x.float() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("rows,K,N,groups", [
    (300, 256, 512, None),                       # single group, ragged M
    (1000, 1536, 2048, [(0, 937), (937, 63)]),   # two experts, second tiny (MoT und rows)
    (128, 64, 256, None),                        # one k-block
    (77, 328, 600, None),                        # N tail, K tail (not multiples of the 256 / 64 tile)
    (2 * 148 * 128 + 5, 512, 768, None),         # > 2 waves of tiles, persistent loop
])
def test_store_bf16(rows, K, N, groups):
    from g2vlm_b200 import ops
    ng = 1 if groups is None else len(groups)
    a, w, bias = _mk(rows, K, N, ng)
    groups_ = groups or [(0, rows)]
    out = torch.full((rows, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, out, epilogue=ops.EPI_STORE_BF16, groups=groups, bias=bias)
    torch.cuda.synchronize()
    ref = _ref_linear(a, w, bias, groups_, N)
    covered = torch.zeros(rows, dtype=torch.bool, device="cuda")
    for r0, n in groups_:
        covered[r0:r0 + n] = True
    assert torch.isfinite(out[covered].float()).all()
    assert _relerr(out[covered], ref[covered]) < 1e-2  # bf16 output rounding (2^-8)
    # exact after rounding the fp32 reference to bf16, up to accumulation-order noise
    diff = (out[covered].float() - ref[covered].to(torch.bfloat16).float()).abs()
    assert (diff > 0).float().mean() < 0.05


def test_empty_group_is_noop():
    from g2vlm_b200 import ops
    a, w, bias = _mk(16, 64, 256, 1)
    out = torch.zeros(16, 256, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, out, epilogue=ops.EPI_STORE_BF16, groups=[(0, 0)], bias=bias)
    torch.cuda.synchronize()
    assert (out == 0).all()


def test_gelu():
    from g2vlm_b200 import ops
    a, w, bias = _mk(500, 256, 1024, 1, seed=1)
    out = torch.empty(500, 1024, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, out, epilogue=ops.EPI_STORE_BF16, bias=bias, flags=ops.GEMM_GELU)
    ref = torch.nn.functional.gelu(_ref_linear(a, w, bias, [(0, 500)], 1024).to(torch.bfloat16).float())
    assert _relerr(out, ref) < 1e-2


def test_swiglu():
    from g2vlm_b200 import ops
    rows, K, I = 700, 256, 512
    g = torch.Generator().manual_seed(2)
    a = (torch.randn(rows, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
    wg = (torch.randn(2, I, K, generator=g) * 0.08).to(torch.bfloat16).cuda()   # per expert
    wu = (torch.randn(2, I, K, generator=g) * 0.08).to(torch.bfloat16).cuda()
    # interleave gate/up in blocks of 128 rows per expert (layout the SwiGLU epilogue expects)
    w = torch.stack([wg.view(2, I // 128, 128, K), wu.view(2, I // 128, 128, K)], dim=2).reshape(2 * 2 * I, K).contiguous()
    groups = [(0, 600), (600, 100)]
    out = torch.empty(rows, I, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, out, epilogue=ops.EPI_SWIGLU_BF16, groups=groups)
    ref = torch.empty(rows, I, device="cuda")
    for e, (r0, n) in enumerate(groups):
        x = a[r0:r0 + n].float()
        gt = (x @ wg[e].float().T).to(torch.bfloat16)
        up = (x @ wu[e].float().T).to(torch.bfloat16)
        ref[r0:r0 + n] = (torch.nn.functional.silu(gt.float()).to(torch.bfloat16).float() * up.float())
    assert _relerr(out, ref) < 1.5e-2


def test_resid_f32_layerscale():
    from g2vlm_b200 import ops
    rows, K, N = 900, 512, 512
    a, w, _ = _mk(rows, K, N, 2, seed=3)
    groups = [(0, 850), (850, 50)]
    gamma = (torch.rand(N) + 0.5).cuda()
    x0 = torch.randn(rows, N, device="cuda")
    x = x0.clone()
    ops.gemm(a, w, x, epilogue=ops.EPI_RESID_F32, groups=groups, scale=gamma, scale_groups=1,
             flags=ops.GEMM_ROUND_AFTER_SCALE)
    y = _ref_linear(a, w, None, groups, N).to(torch.bfloat16).float()
    y[:850] = (y[:850] * gamma).to(torch.bfloat16).float()
    ref = x0 + y
    assert _relerr(x, ref) < 5e-3


def test_store_f32_flags():
    from g2vlm_b200 import ops
    rows, K, N = 333, 256, 512
    a, w, bias = _mk(rows, K, N, 1, seed=4)
    res = torch.randn(rows, N, device="cuda")
    out = torch.empty(rows, N, device="cuda")
    ops.gemm(a, w, out, epilogue=ops.EPI_STORE_F32, bias=bias, flags=ops.GEMM_RELU, residual=res)
    ref = torch.relu(_ref_linear(a, w, bias, [(0, rows)], N)) + res
    assert _relerr(out, ref) < 1e-4
    # accumulate pass (split-bf16 second term)
    out2 = out.clone()
    ops.gemm(a, w, out2, epilogue=ops.EPI_STORE_F32, flags=ops.GEMM_ACCUMULATE)
    ref2 = ref + _ref_linear(a, w, None, [(0, rows)], N)
    assert _relerr(out2, ref2) < 1e-4


def test_store_f32_pixel_shuffle_head_shape():
    """N = 588 = 3*14*14 (Pi3LinearPts3d.proj): N tail inside a 32-column chunk."""
    from g2vlm_b200 import ops
    rows, K, N = 200, 1024, 588
    a, w, bias = _mk(rows, K, N, 1, seed=6)
    out = torch.full((rows, N), float("nan"), device="cuda")
    ops.gemm(a, w, out, epilogue=ops.EPI_STORE_F32, bias=bias)
    assert _relerr(out, _ref_linear(a, w, bias, [(0, rows)], N)) < 1e-4


def test_full_size_mot_qkv_timing():
    """Config-2 sized routed QKV projection: M = 16*1369 geo + 32 und rows, K = 1536, N = 2048."""
    from g2vlm_b200 import ops
    n_geo, n_und, K, N = 16 * 1369, 32, 1536, 2048
    rows = n_geo + n_und
    a, w, bias = _mk(rows, K, N, 2, seed=5)
    groups = [(0, n_geo), (n_geo, n_und)]
    out = torch.empty(rows, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(a, w, out, epilogue=ops.EPI_STORE_BF16, groups=groups, bias=bias)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(10):
        ops.gemm(a, w, out, epilogue=ops.EPI_STORE_BF16, groups=groups, bias=bias)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 10
    tflops = 2 * rows * K * N / ms / 1e9
    print(f"\nqkv gemm {ms:.3f} ms  {tflops:.1f} TFLOP/s")
    idx = torch.randint(0, rows, (512,), device="cuda")
    ref = _ref_linear(a, w, bias, groups, N)
    assert _relerr(out[idx], ref[idx]) < 1e-2


@pytest.mark.parametrize("epi", ["bf16", "f32", "resid"])
def test_guard_bands_are_never_written(epi):
    """compute-sanitizer is closed on this pool, so out-of-bounds writes are hunted with guard bands: rows
    past the last group row and columns past N (inside the leading dimension) must keep their sentinel."""
    from g2vlm_b200 import ops
    rows, K, N, pad_r, pad_c = 333, 192, 600, 70, 40
    a, w, bias = _mk(rows + pad_r, K, N, 1, seed=11)
    dt = torch.bfloat16 if epi == "bf16" else torch.float32
    out = torch.full((rows + pad_r, N + pad_c), 3.0, device="cuda", dtype=dt)
    e = {"bf16": ops.EPI_STORE_BF16, "f32": ops.EPI_STORE_F32, "resid": ops.EPI_RESID_F32}[epi]
    ops.gemm(a, w, out, epilogue=e, groups=[(0, rows)], bias=bias)
    torch.cuda.synchronize()
    assert (out[rows:] == 3.0).all() and (out[:, N:] == 3.0).all()
    ref = _ref_linear(a, w, bias, [(0, rows)], N)[:rows] + (3.0 if epi == "resid" else 0.0)
    assert _relerr(out[:rows, :N], ref) < 1e-2


def test_store_bf16_column_regroup():
    """out_col_group / out_col_stride: heads of 96 output columns land in 128-wide slots; pad columns untouched."""
    from g2vlm_b200 import ops
    g = torch.Generator().manual_seed(11)
    rows, K, heads = 300, 256, 6
    a = (torch.randn(rows, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
    w = (torch.randn(heads * 96, K, generator=g) * 0.1).to(torch.bfloat16).cuda()
    bias = torch.randn(heads * 96, generator=g).cuda()
    out = torch.full((rows, heads * 128), 3.0, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, out, epilogue=ops.EPI_STORE_BF16, bias=bias, out_col_group=96, out_col_stride=128)
    ref = (a.float() @ w.float().T + bias).view(rows, heads, 96)
    got = out.float().view(rows, heads, 128)
    assert ((got[:, :, :96] - ref).abs().max() / ref.abs().max()) < 1e-2
    assert bool((got[:, :, 96:] == 3.0).all())
    with pytest.raises(Exception):
        ops.gemm(a, w, out, epilogue=ops.EPI_STORE_BF16, out_col_group=48, out_col_stride=128)


# ------------------------------------------------------------------------------------------------------
# Both tile schedulers on the SAME shapes: the CTA-pair kernel (256x256, cta_group::2) is
# what o_proj / down / DINO dense+fc2 / Pi3 proj+fc2 run in the benchmark; FORCE_PAIR / FORCE_SINGLE pick it per call.
# ------------------------------------------------------------------------------------------------------
_KERNELS = ["pair", "single"]


def _force(kernel):
    from g2vlm_b200 import ops
    return ops.GEMM_FORCE_PAIR if kernel == "pair" else ops.GEMM_FORCE_SINGLE


@pytest.mark.parametrize("kernel", _KERNELS)
@pytest.mark.parametrize("rows,K,N,groups", [
    (900, 512, 512, [(0, 850), (850, 50)]),          # few tiles, odd number of M tiles per group (pair padding)
    (6500, 1024, 1536, [(0, 6400), (6400, 100)]),    # >= 148 pair tiles at N = 1536 (the automatic pair regime)
    (130, 4096, 1024, None),                         # long K (DINO fc2 shape), second CTA of the pair nearly empty
])
def test_resid_f32_both_kernels(kernel, rows, K, N, groups):
    from g2vlm_b200 import ops
    ng = 1 if groups is None else len(groups)
    a, w, bias = _mk(rows, K, N, ng, seed=21)
    groups_ = groups or [(0, rows)]
    gamma = (torch.rand(N) + 0.5).cuda()
    x0 = torch.randn(rows, N, device="cuda")
    for flags, rounded in ((ops.GEMM_ROUND_AFTER_SCALE, True), (0, False), (ops.GEMM_ROUND_SUM, False)):
        x = x0.clone()
        ops.gemm(a, w, x, epilogue=ops.EPI_RESID_F32, groups=groups, bias=bias, scale=gamma, scale_groups=1,
                 flags=flags | _force(kernel))
        y = _ref_linear(a, w, bias, groups_, N).to(torch.bfloat16).float()
        n0 = groups_[0][1]
        y[:n0] = y[:n0] * gamma                        # scale_groups = 1: LayerScale on group 0 only
        if rounded:
            y[:n0] = y[:n0].to(torch.bfloat16).float()
        ref = x0 + y
        if flags & ops.GEMM_ROUND_SUM:
            ref = ref.to(torch.bfloat16).float()
            assert _relerr(x, ref) < 1e-2
        else:
            assert _relerr(x, ref) < 5e-3, (kernel, flags)
            # the TMA reduce-add (fp32 add in the L2 reduction units) and the SM-side read-modify-write perform the
            # same single fp32 addition per element: bit-identical residual streams
            x2 = x0.clone()
            ops.gemm(a, w, x2, epilogue=ops.EPI_RESID_F32, groups=groups, bias=bias, scale=gamma, scale_groups=1,
                     flags=flags | _force(kernel) | ops.GEMM_NO_TMA_OUT)
            assert torch.equal(x, x2), (kernel, flags)


@pytest.mark.parametrize("kernel", _KERNELS)
def test_store_f32_both_kernels(kernel):
    from g2vlm_b200 import ops
    rows, K, N = 6500, 512, 1536
    a, w, bias = _mk(rows, K, N, 1, seed=22)
    res = torch.randn(rows, N, device="cuda")
    out = torch.full((rows, N), float("nan"), device="cuda")
    ops.gemm(a, w, out, epilogue=ops.EPI_STORE_F32, bias=bias, flags=ops.GEMM_RELU | _force(kernel), residual=res)
    ref = torch.relu(_ref_linear(a, w, bias, [(0, rows)], N)) + res
    assert _relerr(out, ref) < 1e-4
    out2 = out.clone()
    ops.gemm(a, w, out2, epilogue=ops.EPI_STORE_F32, flags=ops.GEMM_ACCUMULATE | _force(kernel))
    assert _relerr(out2, ref + _ref_linear(a, w, None, [(0, rows)], N)) < 1e-4
    out3 = torch.empty(rows, N, device="cuda")
    ops.gemm(a, w, out3, epilogue=ops.EPI_STORE_F32, bias=bias, flags=ops.GEMM_ROUND_BF16 | _force(kernel))
    ref3 = _ref_linear(a, w, bias, [(0, rows)], N).to(torch.bfloat16).float()
    assert ((out3 - ref3).abs() > 0).float().mean() < 0.05


@pytest.mark.parametrize("kernel", _KERNELS)
@pytest.mark.parametrize("act", ["gelu", "quick_gelu"])
def test_gelu_both_kernels(kernel, act):
    from g2vlm_b200 import ops
    rows, K, N = 3000, 1024, 4096
    a, w, bias = _mk(rows, K, N, 1, seed=23)
    out = torch.empty(rows, N, device="cuda", dtype=torch.bfloat16)
    flag = ops.GEMM_GELU if act == "gelu" else ops.GEMM_QUICK_GELU
    ops.gemm(a, w, out, epilogue=ops.EPI_STORE_BF16, bias=bias, flags=flag | _force(kernel))
    y = _ref_linear(a, w, bias, [(0, rows)], N).to(torch.bfloat16).float()
    ref = torch.nn.functional.gelu(y) if act == "gelu" else y * torch.sigmoid(1.702 * y)
    assert _relerr(out, ref) < 1e-2
    # the activation itself, on the kernel's OWN bf16 pre-activation (a 1-ulp flip of the pre-activation moves the
    # negative tail of GELU by several per cent, so the fp32 reference above cannot resolve this): <= 1 bf16 ulp
    pre = torch.empty_like(out)
    ops.gemm(a, w, pre, epilogue=ops.EPI_STORE_BF16, bias=bias, flags=_force(kernel))
    p32 = pre.float()
    exact = (torch.nn.functional.gelu(p32) if act == "gelu" else p32 * torch.sigmoid(1.702 * p32))
    err = (out.float() - exact).abs()
    assert bool((err <= exact.abs() * 2.0 ** -7 + 1e-30).all())


@pytest.mark.parametrize("kernel", _KERNELS)
def test_swiglu_both_kernels(kernel):
    from g2vlm_b200 import ops
    rows, K, I = 2100, 512, 1024
    g = torch.Generator().manual_seed(24)
    a = (torch.randn(rows, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
    wg = (torch.randn(2, I, K, generator=g) * 0.08).to(torch.bfloat16).cuda()
    wu = (torch.randn(2, I, K, generator=g) * 0.08).to(torch.bfloat16).cuda()
    w = torch.stack([wg.view(2, I // 128, 128, K), wu.view(2, I // 128, 128, K)], dim=2).reshape(2 * 2 * I, K).contiguous()
    groups = [(0, 2000), (2000, 100)]
    out = torch.empty(rows, I, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, out, epilogue=ops.EPI_SWIGLU_BF16, groups=groups, flags=_force(kernel))
    ref = torch.empty(rows, I, device="cuda")
    for e, (r0, n) in enumerate(groups):
        x = a[r0:r0 + n].float()
        gt = (x @ wg[e].float().T).to(torch.bfloat16)
        up = (x @ wu[e].float().T).to(torch.bfloat16)
        ref[r0:r0 + n] = (torch.nn.functional.silu(gt.float()).to(torch.bfloat16).float() * up.float())
    assert _relerr(out, ref) < 1.5e-2


def test_pair_and_single_kernels_agree_bitwise_on_store_bf16():
    """Same K order, same epilogue arithmetic: the two schedulers must produce identical bf16 outputs."""
    from g2vlm_b200 import ops
    rows, K, N = 1500, 768, 1024
    a, w, bias = _mk(rows, K, N, 1, seed=25)
    o1 = torch.empty(rows, N, device="cuda", dtype=torch.bfloat16)
    o2 = torch.empty_like(o1)
    ops.gemm(a, w, o1, epilogue=ops.EPI_STORE_BF16, bias=bias, flags=ops.GEMM_FORCE_PAIR)
    ops.gemm(a, w, o2, epilogue=ops.EPI_STORE_BF16, bias=bias, flags=ops.GEMM_FORCE_SINGLE)
    assert torch.equal(o1, o2)
    with pytest.raises(Exception):
        ops.gemm(a, w, o1, epilogue=ops.EPI_STORE_BF16, flags=ops.GEMM_FORCE_PAIR | ops.GEMM_FORCE_SINGLE)
