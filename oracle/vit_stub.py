"""TEST INFRASTRUCTURE — deterministic stand-in for `QwenVL2ImageTransform` (data/transforms.py:151-176), which
needs the HF hub.  Same output contract as Qwen2VLImageProcessor._preprocess
(modeling/qwen2vl/image_processing_qwen2_vl.py:155-290): CLIP-normalised pixels, temporal patch 2 (the still
image is repeated), 14x14 patches emitted in 2x2-merge-major order, `pixel_values [gh*gw, 3*2*14*14]` and
`image_grid_thw [[1, gh, gw]]`.  Used by the golden generator and by the tests (both sides get the same input)."""
import numpy as np
import torch

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


class StubVitTransform:
    def __init__(self, height=56, width=84, patch=14, merge=2, temporal=2):
        self.h, self.w, self.p, self.m, self.t = height, width, patch, merge, temporal

    def __call__(self, images):
        img = images[0]
        if not torch.is_tensor(img):
            img = img.resize((self.w, self.h), 3)
            x = torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()).permute(2, 0, 1).float() / 255.0
        else:
            x = torch.nn.functional.interpolate(img[None].float(), (self.h, self.w), mode="bilinear", align_corners=False)[0]
        x = (x - torch.tensor(CLIP_MEAN).view(3, 1, 1)) / torch.tensor(CLIP_STD).view(3, 1, 1)
        gh, gw, p, m, t = self.h // self.p, self.w // self.p, self.p, self.m, self.t
        patches = x[None].repeat(t, 1, 1, 1)[None]                                   # (1, t, C, H, W)
        patches = patches.reshape(1, t, 3, gh // m, m, p, gw // m, m, p)
        patches = patches.permute(0, 3, 6, 4, 7, 2, 1, 5, 8)
        flat = patches.reshape(gh * gw, 3 * t * p * p).contiguous()
        return flat, torch.tensor([[1, gh, gw]])
