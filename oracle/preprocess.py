"""ORACLE — test infrastructure only (see oracle/__init__.py).

numpy restatement of the reference processor's transform chain for images that are not already
image_size² (videollava/model/multimodal_encoder/languagebind/image/processing_image.py:15-25):

    ToTensor()                       u8 HWC → f32 CHW / 255                                   (:18)
    Resize(224, BICUBIC)             short side → 224 on the TENSOR; torchvision 0.17 (pyproject.toml:15 pins
                                     torch 2.2.1 / torchvision 0.17.1) runs the anti-aliased separable bicubic    (:19)
    CenterCrop(224)                                                                           (:20)
    Normalize(OPENAI mean, std)      (x - mean) / std                                         (:22)

torchvision / ATen are third-party here; the anti-aliased kernel restated below is ATen's
(`upsample_bicubic2d_aa`, UpSampleKernel.cpp: area-pixel scale = in/out, support = 2·max(scale, 1), cubic
convolution with a = -0.5, per-output-pixel weights normalised to sum 1, width pass then height pass, all in fp32).
Pinned in tests/test_preprocess_cpu.py against the installed torch's F.interpolate(..., mode="bicubic",
antialias=True) — the call torchvision's Resize makes — to fp32 rounding.
"""
import numpy as np

f32 = np.float32


def resized_geometry(h: int, w: int, s: int):
    """torchvision Resize(int) + CenterCrop(int): (new_h, new_w, top, left)."""
    if h <= w:
        nh, nw = s, int(s * w / h)
    else:
        nh, nw = int(s * h / w), s
    return nh, nw, int(round((nh - s) / 2.0)), int(round((nw - s) / 2.0))


def _cubic_aa(x):
    a = f32(-0.5)
    x = np.abs(x)
    if x < f32(1):
        return ((a + f32(2)) * x - (a + f32(3))) * x * x + f32(1)
    if x < f32(2):
        return (((x - f32(5)) * x + f32(8)) * x - f32(4)) * a
    return f32(0)


def aa_taps(in_size: int, out_size: int):
    """[(first source index, normalised weights f32[n])] for every output index."""
    scale = f32(in_size) / f32(out_size)
    support = f32(2) * scale if scale >= 1 else f32(2)
    invscale = f32(1) / scale if scale >= 1 else f32(1)
    taps = []
    for i in range(out_size):
        center = scale * (f32(i) + f32(0.5))
        lo = max(int(center - support + f32(0.5)), 0)
        n = min(int(center + support + f32(0.5)), in_size) - lo
        w = np.array([_cubic_aa((f32(j + lo) - center + f32(0.5)) * invscale) for j in range(n)], dtype=f32)
        tot = w.sum(dtype=f32)
        if tot != 0:
            w = w / tot
        taps.append((lo, w))
    return taps


def resize_bicubic_aa(x: np.ndarray, nh: int, nw: int) -> np.ndarray:
    """x f32 [C,H,W] → f32 [C,nh,nw]."""
    c, h, w = x.shape
    mid = np.zeros((c, h, nw), f32)
    for i, (lo, wt) in enumerate(aa_taps(w, nw)):
        mid[:, :, i] = (x[:, :, lo:lo + len(wt)] * wt).sum(-1, dtype=f32)
    out = np.zeros((c, nh, nw), f32)
    for i, (lo, wt) in enumerate(aa_taps(h, nh)):
        out[:, i, :] = (mid[:, lo:lo + len(wt), :] * wt[None, :, None]).sum(1, dtype=f32)
    return out


def preprocess_u8_hwc(img: np.ndarray, s: int, mean, std) -> np.ndarray:
    """u8 [H,W,3] → f32 [3,s,s], the reference's pixel_values for one image."""
    x = img.transpose(2, 0, 1).astype(f32) / f32(255)
    h, w = x.shape[1:]
    nh, nw, top, left = resized_geometry(h, w, s)
    if min(h, w) != s:
        x = resize_bicubic_aa(x, nh, nw)
    x = x[:, top:top + s, left:left + s]
    return (x - np.asarray(mean, f32)[:, None, None]) / np.asarray(std, f32)[:, None, None]
