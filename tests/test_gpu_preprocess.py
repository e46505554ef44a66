"""GPU parity of the frame preprocessing: `lv_frame_preprocess` must reproduce
ImageProcessor.process_images + .to(bfloat16) BIT FOR BIT (integer resampling, IEEE float32 normalisation) - against
the committed outputs of the reference's own code (tests/golden/ref_preprocess.pt) and against the numpy oracle at the
real 448 x 448 target size on video-like frame sizes."""
import os

import numpy as np
import pytest
import torch

from oracle import preprocess as P

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_preprocess_matches_the_references_own_outputs(lib_built):
    from long_vita_b200.preprocess import preprocess_frames

    g = torch.load(os.path.join(GOLD, "ref_preprocess.pt"))
    for frames, want in zip(g["frames"], g["out"]):
        got = preprocess_frames(frames.cuda(), image_size=g["image_size"])
        assert got.dtype == torch.bfloat16 and got.shape == want.shape
        assert torch.equal(got.cpu(), want.to(torch.bfloat16)), frames.shape


@pytest.mark.parametrize("h,w,n", [(360, 640, 3), (1080, 1920, 2), (500, 333, 1), (448, 448, 2), (100, 100, 1), (720, 448, 1)])
def test_preprocess_matches_the_oracle_at_448(lib_built, h, w, n):
    from long_vita_b200.preprocess import preprocess_frames

    rng = np.random.default_rng(h * 7 + w)
    f = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    want = torch.from_numpy(P.process_frames(list(f))).to(torch.bfloat16)
    got = preprocess_frames(torch.from_numpy(f).cuda())
    assert torch.equal(got.cpu(), want)


def test_preprocess_guards_and_empty_batch(lib_built):
    from long_vita_b200.preprocess import preprocess_frames

    with pytest.raises(ValueError):
        preprocess_frames(torch.zeros(1, 8, 8, 3, dtype=torch.uint8))                    # CPU tensor: no fallback
    with pytest.raises(ValueError):
        preprocess_frames(torch.zeros(1, 8, 8, 3, dtype=torch.float32, device="cuda"))
    out = preprocess_frames(torch.zeros(0, 64, 64, 3, dtype=torch.uint8, device="cuda"))
    assert out.shape == (0, 3, 448, 448)
