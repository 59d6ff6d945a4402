#!/usr/bin/env python
"""Input side of recon() for path / PIL inputs: host LANCZOS resize (the reference's load_and_resize14) versus the
device resize (g2vlm_resize_lanczos_u8), 16 decoded 1920x1080 RGB views -> 16 x 3 x 294 x 518 fp32."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from PIL import Image

from g2vlm_b200 import host_prep

rng = np.random.default_rng(0)
views = [Image.fromarray(rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)) for _ in range(16)]
host_prep.load_and_resize14_device(views, 518, "cuda")   # warm-up: tables, allocator
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    ref = host_prep.load_and_resize14(views, 518).cuda()
torch.cuda.synchronize()
t1 = time.perf_counter()
for _ in range(3):
    got = host_prep.load_and_resize14_device(views, 518, "cuda")
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"16 views 1920x1080 -> 294x518: host PIL LANCZOS + upload {1e3 * (t1 - t0) / 3:.1f} ms per scene; "
      f"upload uint8 + device LANCZOS {1e3 * (t2 - t1) / 3:.1f} ms per scene; identical: {torch.equal(ref, got)}")
