"""Output side of the recon path: PLY export without open3d / numpy post-processing.

Mirrors `save_ply_visualization` (reference g2vlm_utils.py:84-149): world points of every view, coloured
by the input images, points with a NaN/Inf coordinate dropped, written as a binary little-endian PLY with
`double x y z, uchar red green blue` vertices (what Open3D's `write_point_cloud` emits).  The reference
resamples `points` to the image size with antialiased bilinear interpolation first (:113-117) — an identity
here because recon already produces H x W maps — and does the filtering / reshaping in numpy on the host;
here validity filtering, compaction and record packing run on the GPU (`g2vlm_ply_pack`) and the host only
writes the header and one contiguous buffer.
"""
from __future__ import annotations

import os
from typing import Dict

import torch

from . import ops

PLY_DTYPE = [("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("red", "u1"), ("green", "u1"), ("blue", "u1")]


def ply_header(n_vertices: int) -> bytes:
    return ("ply\nformat binary_little_endian 1.0\ncomment Created by g2vlm_b200\n"
            f"element vertex {n_vertices}\nproperty double x\nproperty double y\nproperty double z\n"
            "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n").encode("ascii")


def save_ply_visualization(pred_dict: Dict[str, torch.Tensor], save_path: str, filter_nan: bool = True,
                           verbose: bool = False) -> int:
    """Same call shape as the reference function; returns the number of vertices written."""
    points = pred_dict["points"][0].contiguous()            # (N, H, W, 3)
    images = pred_dict["images"][0].to(points.device).contiguous()   # (N, 3, H, W)
    if images.shape[-2:] != points.shape[1:3]:
        raise ValueError("points and images must have the same spatial size (recon guarantees it)")
    records, n = ops.ply_pack(points, images, filter_nonfinite=filter_nan)
    host = records.cpu().numpy()
    d = os.path.dirname(save_path)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(save_path, "wb") as f:
        f.write(ply_header(n))
        f.write(host.tobytes())
    if verbose:
        total = points.shape[0] * points.shape[1] * points.shape[2]
        print(f"wrote {n} / {total} points to {save_path}")
    return n


def read_ply(path: str):
    """Minimal reader for the files written above (tests / debugging): structured numpy array."""
    import numpy as np
    with open(path, "rb") as f:
        n = None
        while True:
            line = f.readline().decode("ascii").strip()
            if line.startswith("element vertex"):
                n = int(line.split()[-1])
            if line == "end_header":
                break
        return np.frombuffer(f.read(), dtype=np.dtype(PLY_DTYPE), count=n)
