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
