"""Timing of the GELU epilogue (table over the bf16 input domain) against the same GEMM without activation, on the
fc1 shapes of config 2 (DINO 21984 x 4096 x 1024, Pi3 21904 x 6144 x 1536)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from g2vlm_b200 import ops
g = torch.Generator().manual_seed(0)
for name, M, N, K in (("dino.fc1", 21984, 4096, 1024), ("pi3.fc1", 21904, 6144, 1536)):
    a = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
    w = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).cuda()
    bias = torch.randn(N, generator=g).cuda()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    t = {}
    for tag, fl in (("gelu", ops.GEMM_GELU), ("none", 0), ("gelu", ops.GEMM_GELU), ("none", 0)):
        for _ in range(5):
            ops.gemm(a, w, out, epilogue=ops.EPI_STORE_BF16, bias=bias, flags=fl)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            ops.gemm(a, w, out, epilogue=ops.EPI_STORE_BF16, bias=bias, flags=fl)
        e1.record(); torch.cuda.synchronize()
        t.setdefault(tag, []).append(e0.elapsed_time(e1) / 50 * 1e3)
    ge, no = min(t["gelu"]), min(t["none"])
    print(f"{name}: with GELU {ge:.1f} us ({2 * M * N * K / ge / 1e6:.0f} TFLOP/s)   without {no:.1f} us ({2 * M * N * K / no / 1e6:.0f} TFLOP/s)")
    # exactness: the table reproduces torch's erf GELU of the kernel's own bf16 pre-activation
    pre = torch.empty_like(out)
    ops.gemm(a, w, pre, epilogue=ops.EPI_STORE_BF16, bias=bias)
    ops.gemm(a, w, out, epilogue=ops.EPI_STORE_BF16, bias=bias, flags=ops.GEMM_GELU)
    want = torch.nn.functional.gelu(pre.float()).to(torch.bfloat16)
    diff = (out.view(torch.int16).int() - want.view(torch.int16).int()).abs()
    print(f"   bit-identical to bf16(gelu_erf(pre)): {(diff == 0).float().mean().item():.6f}, max ulp distance {int(diff.max())}")
