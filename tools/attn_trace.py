"""Clock-stamped timeline of attention_tcgen05_kernel<128> on CTA 0 (steady-state key blocks of the first unit, MoT shape of
config 2; --pi3: the 11-block units of the Pi3 decoders, across a unit boundary).  Needs a debug build of the library with -DG2_ATTN_TRACE (the stamps cost ~10 %, product builds carry none):
    python tools/attn_trace.py --build          # compiles g2vlm_b200/libg2vlm_b200_trace.so next to the product library
    G2VLM_B200_LIB=g2vlm_b200/libg2vlm_b200_trace.so python tools/attn_trace.py
Events per (block, tile) — softmax warp (sub-partition 0, lane 0): wait_s (phase begins), s_full (S_t complete), ld_done
(S in registers), max_done (row maximum / rescale decision taken), published (all of P handed over), first_piece (keys
[0,96) handed over);
MMA issuer thread: pv_begin, pv_chunk0 (first P chunk seen), pv_last (last P chunk seen), pv_issued, qk_issued (next block's QK^T)."""
import ctypes, math, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if "--build" in sys.argv:
    from g2vlm_b200 import _lib
    _lib.build()
    obj = "/tmp/attention_trace.o"
    subprocess.check_call(["nvcc", *_lib.NVCC_FLAGS, "-DG2_ATTN_TRACE", "-c", "-o", obj, str(_lib.CSRC / "attention.cu")])
    others = [str(o) for o in sorted(_lib.OBJ_DIR.glob("*.o")) if o.name != "attention.o"]
    out = str(_lib.PKG_DIR / "libg2vlm_b200_trace.so")
    subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, obj, *others])
    print("built", out)
    sys.exit(0)

import torch
from g2vlm_b200 import ops, _lib

shape = "pi3" if "--pi3" in sys.argv else "mot"
g = torch.Generator().manual_seed(0)
if shape == "mot":
    T, K0 = 16 * 1371, 7
    qkv = torch.randn(T + K0, 2048, generator=g).to(torch.bfloat16).cuda()
    out = torch.zeros(T, 1536, device="cuda", dtype=torch.bfloat16)
    work = ops.attention_work_table([0, T], [0, T + K0]).cuda()
    call = lambda: ops.attention(qkv[:T, :1536], qkv[:, 1536:1792], qkv[:, 1792:], out, work, num_q_heads=12, num_kv_heads=2,
                                 head_dim=128, scale=1 / math.sqrt(128))
    first, last = 8, 14
else:   # per-view self-attention of the Pi3 decoders: 16 views x 1369 rows, 16 heads of 96 in 128-wide slots, 11-block units
    N, P, H = 16, 1369, 16
    qkv = torch.randn(N * P, 3 * H * 128, generator=g).to(torch.bfloat16).cuda()
    out = torch.zeros(N * P, H * 96, device="cuda", dtype=torch.bfloat16)
    cu = [i * P for i in range(N + 1)]
    work = ops.attention_work_table(cu, cu).cuda()
    call = lambda: ops.attention(qkv[:, :H * 128], qkv[:, H * 128:2 * H * 128], qkv[:, 2 * H * 128:], out, work, num_q_heads=H,
                                 num_kv_heads=H, head_dim=128, scale=1 / math.sqrt(96), out_head_cols=96)
    first, last = 7, 16   # the unit boundary is between blocks 10 and 11
for _ in range(2):
    call()
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (3 * 32 * 2 * 16))()
lib = _lib.load()
if not hasattr(lib, "g2vlm_debug_attn_trace"):
    sys.exit("this library has no trace points: build with --build and point G2VLM_B200_LIB at the trace library")
assert lib.g2vlm_debug_attn_trace(buf) == 0
tr = torch.tensor(list(buf)).view(3, 32, 2, 16)
t0 = int(tr[0, first, 0, 0])
sm = ["wait_s", "s_full", "ld_done", "max_done", "published", "first_piece"]
iss = ["pv_begin", "pv_chunk0", "pv_last", "pv_issued", "qk_issued"]
print(f"# {shape} shape; clocks relative to block {first} / tile 0 / wait_s (CTA 0, its first units)")
for blk in range(first, last):
    for t in range(2):
        print(f"blk {blk:2d} tile {t} softmax: " + " ".join(f"{n}={int(tr[t, blk, t, i]) - t0:6d}" for i, n in enumerate(sm)))
        print(f"blk {blk:2d} tile {t} issuer : " + " ".join(f"{n}={int(tr[2, blk, t, i]) - t0:6d}" for i, n in enumerate(iss)))
if shape == "mot":
    per = (int(tr[0, 24, 0, 0]) - int(tr[0, 8, 0, 0])) / 16
    print(f"# period per key block (tile 0, blocks 8..24): {per:.0f} clocks for 2 x 1024 clocks of tensor work")
else:
    per = (int(tr[0, 22, 0, 0]) - int(tr[0, 0, 0, 0])) / 2
    print(f"# one 11-block unit (tile 0, start of unit 0 to start of unit 2, halved): {per:.0f} clocks = {per / 11:.0f} per block")
