"""Per-block timeline of the two-threads-per-row attention kernel on CTA 0 (debug build with -DG2_ATTN_TRACE; see
attention.cu G2_TR).  Build:  nvcc ... -DG2_ATTN_TRACE -c attention.cu ; link with the other objects into
g2vlm_b200/libg2vlm_b200_trace.so.  usage: G2VLM_B200_LIB=g2vlm_b200/libg2vlm_b200_trace.so python tools/attn_trace.py [mode]"""
import ctypes, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from g2vlm_b200 import ops, _lib

mode = sys.argv[1] if len(sys.argv) > 1 else "1"
os.environ["G2VLM_ATTN_ROWSPLIT"] = mode
T, K0 = 16 * 1371, 7
g = torch.Generator().manual_seed(0)
qkv = torch.randn(T + K0, 2048, generator=g).to(torch.bfloat16).cuda()
out = torch.zeros(T, 1536, device="cuda", dtype=torch.bfloat16)
work = ops.attention_work_table([0, T], [0, T + K0]).cuda()
for _ in range(2):
    ops.attention(qkv[:T, :1536], qkv[:, 1536:1792], qkv[:, 1792:], out, work, num_q_heads=12, num_kv_heads=2,
                  head_dim=128, scale=1 / math.sqrt(128))
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (3 * 32 * 2 * 8))()
lib = _lib.load()
assert lib.g2vlm_debug_attn_trace(buf) == 0
tr = torch.tensor(list(buf)).view(3, 32, 2, 8)
t0 = int(tr[0, 8, 0, 0])
names = {0: "sm.h0", 1: "sm.h1", 2: "issue"}
ev = {0: ["begin", "m_known", "published", "next_s_loaded", "next_max"], 2: ["pv_begin", "pv_chunk0", "pv_last_chunk", "pv_issued", "qk_issued"]}
print(f"# rowsplit mode {mode}; clocks relative to block 8 tile 0 'wait_s' of softmax half 0 (CTA 0, first unit)")
for blk in range(8, 13):
    for t in range(2):
        for role in (0, 1, 2):
            e = ev[2 if role == 2 else 0]
            vals = " ".join(f"{n}={int(tr[role, blk, t, i]) - t0:6d}" for i, n in enumerate(e))
            print(f"blk {blk:2d} tile {t} {names[role]}: {vals}")
