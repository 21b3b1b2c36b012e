"""Image processor with the duck type of ``LanguageBindImageProcessor``
(processing_image.py:33-82): ``preprocess(image_or_path, return_tensors='pt')`` →
``{'pixel_values': f32 [N,3,224,224]}``, ``__call__(images=...)``, ``image_mean``, ``crop_size``.

Transform chain (processing_image.py:15-25): ToTensor (u8 HWC → f32 CHW / 255) → Resize(224,
bicubic, short side, on the tensor; torchvision 0.17 defaults antialias=True) → CenterCrop(224) → Normalize(mean, std).
Implemented with torch only (no torchvision import) so the op order is explicit.

The raw-uint8 fast path (`to_uint8_nhwc`) skips normalisation on the host: 224×224 frames are
shipped as NHWC uint8 and normalised inside the patchify kernel on the GPU (north star).
"""
from __future__ import annotations

from typing import List, Union

import numpy as np
import torch
import torch.nn.functional as F

from .constants import OPENAI_DATASET_MEAN, OPENAI_DATASET_STD


def _load_rgb_u8(image) -> np.ndarray:
    """path | PIL image | ndarray[H,W,3] u8 → ndarray[H,W,3] u8 (processing_image.py:28-31)."""
    if isinstance(image, str):
        from PIL import Image
        image = Image.open(image).convert("RGB")
    if isinstance(image, np.ndarray):
        arr = image
    elif isinstance(image, torch.Tensor):
        arr = image.cpu().numpy()
    else:  # PIL
        arr = np.asarray(image.convert("RGB") if getattr(image, "mode", "RGB") != "RGB" else image)
    if arr.ndim != 3 or arr.shape[2] != 3 or arr.dtype != np.uint8:
        raise ValueError(f"expected HxWx3 uint8 image, got {arr.shape} {arr.dtype}")
    return arr


class TeoImageProcessor:
    def __init__(self, image_size: int = 224):
        self.image_size = int(image_size)
        self.image_mean = OPENAI_DATASET_MEAN
        self.image_std = OPENAI_DATASET_STD
        self.crop_size = {"height": self.image_size, "width": self.image_size}

    # -- reference-equivalent float path ------------------------------------------------------
    def _transform(self, image) -> torch.Tensor:
        s = self.image_size
        x = torch.from_numpy(_load_rgb_u8(image).copy()).permute(2, 0, 1).to(torch.float32) / 255.0
        _, h, w = x.shape
        if min(h, w) != s:                                   # Resize(short side → s), bicubic
            if h <= w:
                nh, nw = s, int(s * w / h)
            else:
                nh, nw = int(s * h / w), s
            x = F.interpolate(x[None], size=(nh, nw), mode="bicubic", align_corners=False,
                              antialias=True)[0]
            h, w = nh, nw
        top, left = int(round((h - s) / 2.0)), int(round((w - s) / 2.0))   # CenterCrop
        x = x[:, top:top + s, left:left + s]
        mean = torch.tensor(self.image_mean, dtype=torch.float32)[:, None, None]
        std = torch.tensor(self.image_std, dtype=torch.float32)[:, None, None]
        return (x - mean) / std

    def __call__(self, images=None, text=None, return_tensors=None, **kwargs):
        if images is None:
            raise ValueError("You have to specify either text or images. Both cannot be none.")
        if text is not None:
            raise ValueError("the CLIP text side is not part of the TEOChat hot path")
        if not isinstance(images, list):
            images = [images]
        return {"pixel_values": torch.stack([self._transform(im) for im in images])}

    def preprocess(self, images, return_tensors="pt"):
        return self.__call__(images=images, return_tensors=return_tensors)

    # -- the same chain on the GPU ----------------------------------------------------------------
    def preprocess_device(self, images, device) -> torch.Tensor:
        """``preprocess`` with the arithmetic on the GPU (`teo_resize_crop_normalize_u8`): each image travels as raw
        uint8 HWC (a quarter of the float bytes, and at its own size) and ToTensor → Resize → CenterCrop → Normalize run
        in two kernels per image.  Returns f32 [N,3,S,S] on ``device`` — the reference's pixel_values to fp32 rounding."""
        import ctypes as C

        from . import lib as L
        if not isinstance(images, list):
            images = [images]
        lib, s = L.load(), self.image_size
        device = torch.device(device)
        stream = torch.cuda.current_stream(device).cuda_stream
        out = torch.empty(len(images), 3, s, s, dtype=torch.float32, device=device)
        mean, std = (C.c_float * 3)(*self.image_mean), (C.c_float * 3)(*self.image_std)
        for n, im in enumerate(images):
            arr = np.array(_load_rgb_u8(im), order="C")            # own, writable copy (PIL hands out read-only views)
            h, w = arr.shape[:2]
            nh, nw = (s, int(s * w / h)) if h <= w else (int(s * h / w), s)
            top, left = int(round((nh - s) / 2.0)), int(round((nw - s) / 2.0))
            src = torch.from_numpy(arr).to(device, non_blocking=True)
            ws = torch.empty(lib.teo_resize_workspace_bytes(h, w, s), dtype=torch.uint8, device=device)
            L.check(lib.teo_resize_crop_normalize_u8(src.data_ptr(), h, w, nh, nw, top, left, s, mean, std, out[n].data_ptr(),
                                                     ws.data_ptr(), ws.numel(), stream), "teo_resize_crop_normalize_u8")
        return out

    # -- raw uint8 fast path --------------------------------------------------------------------
    def to_uint8_nhwc(self, images: Union[List, "np.ndarray"]) -> torch.Tensor:
        """Frames that are already image_size² → u8 [N,H,W,3] (pinned-host friendly)."""
        if not isinstance(images, list):
            images = [images]
        arrs = [_load_rgb_u8(im) for im in images]
        for a in arrs:
            if a.shape[:2] != (self.image_size, self.image_size):
                raise ValueError("raw path needs frames already at image_size; use preprocess()")
        return torch.from_numpy(np.stack(arrs))
