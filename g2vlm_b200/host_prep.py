"""Host-side routing / permutation / position-index construction of the recon path.

Mirrors `G2VLM.prepare_prompts_addbos` (modeling/g2vlm/g2vlm.py:561-594) and
`G2VLM.prepare_dino_images_pi3` (:868-966) with `patchify` / `get_rope_index_image_3D_dino`
(data/data_utils.py:40, 78-137) and the image loader `load_and_resize14`
(data/transforms_vggt.py:411-462).  Integer outputs are bit-exact with the reference (tests compare
them with fixtures produced by the reference itself); they are built vectorised instead of with the
reference's per-image Python loops.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple, Union

import torch

RESNET_MEAN = (0.485, 0.456, 0.406)
RESNET_STD = (0.229, 0.224, 0.225)


def load_and_resize14(images: Sequence, new_width: int = 518) -> torch.Tensor:
    """paths / PIL images -> (N,3,H,W) fp32 in [0,1]; every view is resized (LANCZOS) to the target
    size derived from the FIRST image (data/transforms_vggt.py:421-429), then antialias-bilinear
    resampled to a multiple of 14 (:454-462)."""
    from PIL import Image
    import numpy as np

    pil = [Image.open(im) if isinstance(im, str) else im for im in images]
    w0, h0 = pil[0].size
    tw, th = new_width, round(h0 * (new_width / w0) / 14) * 14
    out = []
    for im in pil:
        im = im.resize((tw, th), Image.Resampling.LANCZOS)
        a = np.asarray(im.convert("RGB") if im.mode != "RGB" else im, dtype=np.uint8)
        out.append(torch.from_numpy(a.copy()).permute(2, 0, 1).float().div(255.0))
    x = torch.stack(out, dim=0)
    ph, pw = x.shape[-2] // 14, x.shape[-1] // 14
    return torch.nn.functional.interpolate(x, (ph * 14, pw * 14), mode="bilinear", align_corners=False,
                                           antialias=True)


def lanczos_tables(in_size: int, out_size: int):
    """Pillow's 8-bit resampling tables for one axis (restates precompute_coeffs + normalize_coeffs_8bpc of
    Pillow's Resample.c with the Lanczos-3 filter): bounds int32 [out,2] = (first input index, taps) and coef
    int32 [out, ksize] in 22-bit fixed point.  Checked against PIL bit for bit (tests/test_host_prep_cpu.py)."""
    import math

    import numpy as np
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 3.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        t = (np.arange(xmax, dtype=np.float64) + xmin - center + 0.5) * ss

        def sinc(v):
            out = np.ones_like(v)
            nz = v != 0.0
            vv = v[nz] * math.pi
            out[nz] = np.sin(vv) / vv
            return out

        w = np.where((t >= -3.0) & (t < 3.0), sinc(t) * sinc(t / 3.0), 0.0)
        ww = 0.0
        for v in w:          # sequential double sum, as the C loop
            ww += float(v)
        if ww != 0.0:
            w = w / ww
        kk[xx, :xmax] = w
        bounds[xx] = (xmin, xmax)
    coef = np.where(kk < 0, np.trunc(-0.5 + kk * (1 << 22)), np.trunc(0.5 + kk * (1 << 22))).astype(np.int32)
    return bounds, coef


_TABLE_CACHE: Dict[tuple, tuple] = {}
_STAGING: Dict[object, object] = {}


def load_and_resize14_device(images: Sequence, new_width: int = 518, device="cuda") -> torch.Tensor:
    """Same result as load_and_resize14 (bit-identical, tests/test_io_gpu.py), but only the decoded uint8 pixels
    cross PCIe and the LANCZOS resampling + ToTensor run on the device (g2vlm_resize_lanczos_u8).  The trailing
    antialiased-bilinear resample of the reference is the identity here because the LANCZOS target is already a
    multiple of 14 in both directions whenever new_width is.  Views that are not plain RGB take the host path."""
    from PIL import Image
    import numpy as np

    from . import ops
    pil = [Image.open(im) if isinstance(im, str) else im for im in images]
    w0, h0 = pil[0].size
    tw, th = new_width, round(h0 * (new_width / w0) / 14) * 14
    if tw % 14 or th % 14 or th <= 0 or any(im.mode != "RGB" for im in pil):
        return load_and_resize14(pil, new_width).to(device)
    out = torch.empty(len(pil), 3, th, tw, dtype=torch.float32, device=device)

    def tables(n_in, n_out):
        if n_in == n_out:
            return None
        key = (n_in, n_out, str(device))
        if key not in _TABLE_CACHE:
            b, c = lanczos_tables(n_in, n_out)
            _TABLE_CACHE[key] = (torch.from_numpy(b).to(device), torch.from_numpy(c).to(device))
        return _TABLE_CACHE[key]

    # one pinned + one device staging slot per view, kept across calls (pinning memory costs ~1 ms per image)
    ev = _STAGING.get("event")
    if ev is not None:
        ev.synchronize()             # the previous call's uploads have left the pinned slots
    for i, im in enumerate(pil):
        a = np.asarray(im, dtype=np.uint8)                                    # [H, W, 3] view of the decoded pixels
        key = (i, a.shape, str(device))
        slot = _STAGING.get(key)
        if slot is None:
            slot = _STAGING[key] = (torch.empty(a.shape, dtype=torch.uint8, pin_memory=True),
                                    torch.empty(a.shape, dtype=torch.uint8, device=device))
        pin, src = slot
        pin.numpy()[...] = a
        src.copy_(pin, non_blocking=True)
        ops.resize_lanczos_u8(src, tables(a.shape[1], tw), tables(a.shape[0], th), th, tw, out_f32=out[i])
    if ev is None:
        ev = _STAGING["event"] = torch.cuda.Event()
    ev.record()
    return out


def prepare_prompts_addbos(curr_kvlens: List[int], curr_rope: List[int], prompts: List[str], tokenizer,
                           new_token_ids: Dict[str, int]) -> Tuple[Dict[str, torch.Tensor], List[int], List[int]]:
    """Same signature and outputs as the reference method (g2vlm.py:561-594)."""
    ids_all, pos_all, lens, idx_all, kv_idx = [], [], [], [], []
    curr = 0
    newlens, new_rope = [], []
    for prompt, kvlen, pos in zip(prompts, curr_kvlens, curr_rope):
        kv_idx.extend(range(curr, curr + kvlen))
        curr += kvlen
        ids = [int(new_token_ids["bos_token_id"])] + list(tokenizer.encode(prompt))
        lens.append(len(ids))
        ids_all.extend(ids)
        pos_all.extend(range(pos, pos + len(ids)))
        idx_all.extend(range(curr, curr + len(ids)))
        newlens.append(kvlen + len(ids))
        new_rope.append(pos + len(ids))
        curr += len(ids)
    gi = {
        "text_token_lens": torch.tensor(lens, dtype=torch.int),
        "packed_text_ids": torch.tensor(ids_all, dtype=torch.long),
        "packed_text_position_ids": torch.tensor(pos_all, dtype=torch.long).expand(3, -1),
        "packed_text_indexes": torch.tensor(idx_all, dtype=torch.long),
        "packed_key_value_indexes": torch.tensor(kv_idx, dtype=torch.long),
        "key_values_lens": torch.tensor(curr_kvlens, dtype=torch.int),
    }
    return gi, newlens, new_rope


def prepare_dino_images_pi3(curr_kvlens: List[int], curr_rope: List[int],
                            images: Union[Sequence, torch.Tensor], new_token_ids: Dict[str, int],
                            patch: int = 14, normalize_on_host: bool = True
                            ) -> Tuple[Dict[str, torch.Tensor], List[int], List[int]]:
    """Reference: g2vlm.py:868-966.  `images` may be paths / PIL images (loaded like the reference)
    or an already loaded (N,3,H,W) tensor in [0,1] with H, W multiples of 14 (benchmark / tests).
    normalize_on_host=False defers the ImageNet normalisation (:950) to the device (fused into the
    im2col kernel, bit-identical): `packed_dino_images` is then the SAME tensor as `original_images`, so
    only one copy of the raw views crosses PCIe."""
    if not torch.is_tensor(images):
        images = load_and_resize14(images, 518)
    assert images.dim() == 4 and images.shape[1] == 3
    N, _, H, W = images.shape
    gh, gw = H // patch, W // patch
    P = gh * gw
    K0, r0 = int(curr_kvlens[0]), int(curr_rope[0])
    step = P + 2                        # packed rows per view: <start>, P patches, <end>
    rope_step = max(gh, gw) + 2         # positions consumed per view (1 + max(gh,gw) + 1)

    v = torch.arange(N, dtype=torch.long)
    base = v * step
    text_idx = torch.stack([base, base + P + 1], dim=1).flatten()
    geo_idx = (base[:, None] + 1 + torch.arange(P, dtype=torch.long)[None]).flatten()
    packed_idx = K0 + torch.arange(N * step, dtype=torch.long)

    pstart = r0 + v * rope_step         # position of <start> of each view
    rr = torch.arange(gh, dtype=torch.long).view(-1, 1).expand(-1, gw).flatten()
    cc = torch.arange(gw, dtype=torch.long).view(1, -1).expand(gh, -1).flatten()
    grid = torch.stack([torch.zeros(P, dtype=torch.long), rr, cc])      # (3, P): t, h, w offsets
    pos = torch.empty(3, N, step, dtype=torch.long)
    pos[:, :, 0] = pstart
    pos[:, :, 1:P + 1] = (pstart + 1)[None, :, None] + grid[:, None, :]
    pos[:, :, P + 1] = pstart + 1 + max(gh, gw)

    soi, eoi = int(new_token_ids["start_of_image"]), int(new_token_ids["end_of_image"])
    mean = torch.tensor(RESNET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(RESNET_STD).view(1, 3, 1, 1)
    gi = {
        "packed_dino_images": (images - mean) / std if normalize_on_host else images,
        "original_images": images.clone() if normalize_on_host else images,
        "packed_text_ids": torch.tensor([soi, eoi] * N, dtype=torch.long),
        "packed_text_indexes": text_idx,
        "dino_token_seqlens": torch.full((N,), P, dtype=torch.int),
        "packed_dino_token_indexes": geo_idx,
        "packed_position_ids": pos.reshape(3, N * step),
        "packed_seqlens": torch.tensor([N * step], dtype=torch.int),
        "packed_indexes": packed_idx,
        "packed_key_value_indexes": torch.arange(K0, dtype=torch.long),
        "key_values_lens": torch.tensor(curr_kvlens, dtype=torch.int),
    }
    return gi, [K0 + N * step], [r0 + N * rope_step]
