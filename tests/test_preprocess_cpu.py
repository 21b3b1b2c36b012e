"""The oracle's restatement of the processor chain (oracle/preprocess.py) against the installed torch — the
call torchvision's Resize makes on a tensor (processing_image.py:19) — and against the product's host processor."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import preprocess as OP
from teochat_b200.constants import OPENAI_DATASET_MEAN, OPENAI_DATASET_STD
from teochat_b200.processor import TeoImageProcessor

SIZES = [(300, 400), (400, 300), (1024, 1024), (100, 150), (225, 224), (512, 333), (64, 64), (897, 640), (224, 224), (224, 500)]


@pytest.mark.parametrize("h,w", SIZES[:6])
def test_resize_restatement_matches_torch(h, w):
    rng = np.random.default_rng(h * 7 + w)
    x = rng.integers(0, 256, (3, h, w)).astype(np.float32) / np.float32(255)
    nh, nw, _, _ = OP.resized_geometry(h, w, 224)
    want = F.interpolate(torch.from_numpy(x)[None], size=(nh, nw), mode="bicubic", align_corners=False, antialias=True)[0].numpy()
    got = OP.resize_bicubic_aa(x, nh, nw)
    assert np.abs(got - want).max() <= 4e-6                   # fp32 summation order only


@pytest.mark.parametrize("h,w", SIZES)
def test_chain_matches_host_processor(h, w):
    rng = np.random.default_rng(h * 13 + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    want = TeoImageProcessor(224).preprocess(img)["pixel_values"][0].numpy()
    got = OP.preprocess_u8_hwc(img, 224, OPENAI_DATASET_MEAN, OPENAI_DATASET_STD)
    assert got.shape == want.shape == (3, 224, 224)
    assert np.abs(got - want).max() <= 2e-5                   # 4e-6 / min(std)
    if min(h, w) == 224:                                      # no resampling: crop + normalise only, exact
        assert np.array_equal(got, want)


def test_identity_taps_when_scale_is_one():
    for lo, w in OP.aa_taps(224, 224)[2:-2]:
        assert w.tolist().count(1.0) == 1 and np.count_nonzero(w) == 1
