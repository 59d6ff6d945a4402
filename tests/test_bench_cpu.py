"""bench.py host logic that needs no GPU: the one-JSON-line contract of stdout and the algorithmic work figures the
roofline numbers are built from."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stdout_carries_only_the_json_line():
    """Libraries print to file descriptor 1 too (NCCL's version banner under torchrun, the reference's progress lines):
    after isolate_stdout() only emit() reaches the real stdout."""
    code = ("import os, bench\n"
            "bench.isolate_stdout()\n"
            "print('python-level noise')\n"
            "os.system('echo c-level noise')\n"
            "bench.emit({'metric': 'x', 'value': 1.5})\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert len(lines) == 1 and json.loads(lines[0]) == {"metric": "x", "value": 1.5}
    assert "python-level noise" in r.stderr and "c-level noise" in r.stderr


def test_algorithmic_work_of_config_2():
    """16 views of 518 px: T = 21 936 tokens; the MoT attention launch is 4 T (T + 7) heads head_dim FLOPs (2.957 TFLOP), the
    whole step 180.3 TFLOP — the figures quoted in DESIGN.md and used for roofline.achieved."""
    sys.path.insert(0, ROOT)
    import bench
    from g2vlm_b200 import schema
    cfg = schema.FULL
    P = (518 // 14) ** 2
    fl = bench.algorithmic_flops(cfg, 16, P)
    T = 16 * (P + 2)
    assert T == 21936
    assert fl["mot_attention_launch"] == 4 * T * (T + 7) * cfg.num_heads * cfg.head_dim
    assert abs(fl["mot_attention_launch"] / 1e12 - 2.957) < 2e-3
    assert abs(fl["total"] / 1e12 - 180.3) < 0.1
    # decode: bytes one greedy token has to read behind that scene (und-expert weights + lm_head + the cache's K|V rows)
    H, I, nq, nkv, hd = cfg.hidden_size, cfg.intermediate_size, cfg.num_heads, cfg.num_kv_heads, cfg.head_dim
    w_layer = 2 * ((nq + 2 * nkv) * hd * H + H * nq * hd + 3 * H * I)
    nbytes = cfg.num_layers * (w_layer + 2 * 2 * nkv * hd * 22144) + 2 * cfg.vocab_size * H
    assert abs(nbytes / 1e9 - 3.72) < 0.01
