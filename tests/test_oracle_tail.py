"""CPU: the oracle's restatement of the reference's numpy/OpenCV frame tail (utils/common.py:255-257) against
numpy and cv2 themselves (opencv-python 4.13, the third-party code the reference calls) -- bit-exact."""
import cv2
import numpy as np
import pytest

import oracle

SIZES = [(768, 1024, 921, 691), (768, 1024, 920, 690), (768, 1024, 819, 614), (96, 128, 115, 86), (96, 128, 100, 75),
         (97, 131, 90, 60), (48, 64, 64, 48), (48, 64, 70, 50), (64, 48, 100, 120), (50, 70, 7, 5)]


@pytest.mark.parametrize("H,W,pw,ph", SIZES)
def test_crop_and_resize_match_cv2(H, W, pw, ph):
    rng = np.random.default_rng(H * 1000 + pw)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    ref = cv2.getRectSubPix(image=img, patchSize=(pw, ph), center=(W / 2.0, H / 2.0))
    assert np.array_equal(oracle.getrectsubpix(img, pw, ph, W / 2.0, H / 2.0), ref)
    ref2 = cv2.resize(src=ref, dsize=(W, H), fx=0.0, fy=0.0, interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(oracle.resize_linear(ref, W, H), ref2)


def test_to_uint8_matches_numpy_expression():
    rng = np.random.default_rng(1)
    r = rng.uniform(-0.2, 1.2, (4, 32, 48)).astype(np.float32)
    r[0, 0, :4] = [0.0, 1.0, 254.9999 / 255.0, 0.5]
    ref = (r[0:3].transpose(1, 2, 0) * 255.0).clip(0.0, 255.0).astype(np.uint8)     # utils/common.py:255
    assert np.array_equal(oracle.to_uint8(r), ref)


def test_median5_binary_matches_reference_formulation():
    """spatial_filter(x, 'median-5') of the reference (unfold + median, reflect pad) on a binary mask."""
    import torch
    rng = np.random.default_rng(2)
    m = (rng.random((1, 1, 40, 56)) > 0.4).astype(np.float32)
    t = torch.from_numpy(m)
    x = torch.nn.functional.pad(t, [2, 2, 2, 2], mode='reflect').unfold(2, 5, 1).unfold(3, 5, 1).contiguous()
    ref = x.view(1, 1, 40, 56, 25).median(-1, False)[0].numpy()
    assert np.array_equal(oracle.median5_binary(m), ref)
