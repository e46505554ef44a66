"""GPU parity of the dynamic-patch tiling of still images: `lv_image_tiles_preprocess` behind
`preprocess_image_dynamic` must reproduce ImageProcessor.process_dynamic + .to(bfloat16) BIT FOR BIT - against the
committed outputs of the reference's own code (tests/golden/ref_preprocess_dynamic.pt) and against the numpy oracle at
the real 448-pixel tile on photo-like sizes.  (The kernels' per-element bodies are also run on the host against the same
fixture: tests/test_preprocess_tiles_host.py.)"""
import os

import numpy as np
import pytest
import torch

from oracle import preprocess as P

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_image_tiles_match_the_references_own_outputs(lib_built):
    from long_vita_b200.preprocess import preprocess_image_dynamic

    g = torch.load(os.path.join(GOLD, "ref_preprocess_dynamic.pt"))
    for im, want, gp in zip(g["images"], g["out"], g["grid_pixels"]):
        got, grid = preprocess_image_dynamic(im.cuda(), g["min_patch_grid"], g["max_patch_grid"], g["image_size"])
        assert grid == tuple(gp)
        assert got.dtype == torch.bfloat16 and got.shape == want.shape
        assert torch.equal(got.cpu(), want.to(torch.bfloat16)), tuple(im.shape)


@pytest.mark.parametrize("h,w", [(768, 1024), (1400, 700), (448, 448), (300, 2000), (2160, 3840)])
def test_image_tiles_match_the_oracle_at_448(lib_built, h, w):
    from long_vita_b200.preprocess import preprocess_image_dynamic

    rng = np.random.default_rng(h * 3 + w)
    base = rng.integers(0, 256, (h // 16 + 1, w // 16 + 1, 3)).repeat(16, axis=0).repeat(16, axis=1)[:h, :w]
    im = np.clip(base + rng.integers(-30, 31, base.shape), 0, 255).astype(np.uint8)
    want, grid_want = P.process_dynamic(im, 1, 12, 448)
    got, grid = preprocess_image_dynamic(torch.from_numpy(im).cuda())
    assert grid == grid_want
    assert torch.equal(got.cpu(), torch.from_numpy(want).to(torch.bfloat16))
