"""Frame preprocessing on the GPU: decoded uint8 frames -> the bf16 [N, 3, 448, 448] tensor of the
vision tower, bit-identical to `ImageProcessor.process_images` + `.to(bfloat16)`
(long_vita/data/processor/image_processor.py:183-223; called per video by tools/inference_long_vita.py:248-290).

Host side: the resampling windows and fixed-point weights of Pillow's bicubic filter are computed here in float64
exactly as Pillow's `precompute_coeffs` / `normalize_coeffs_8bpc` do (they depend only on the canvas side and the
target size, so one small table serves every frame of a video) and cached on the device; the pixel work - padding to
a square with the mean colour, both resampling passes, scaling, normalisation, channel-first bf16 output - runs in
`lv_frame_preprocess` (csrc/preprocess.cu).

Still images go through the reference's dynamic-patch tiling instead (`ImageProcessor.process_dynamic`,
image_processor.py:263-285; tools/inference_long_vita.py:643-645): `dynamic_tile_grid` picks the grid of 448-pixel tiles
(host arithmetic on two integers), `preprocess_image_dynamic` resizes the image to that grid without keeping the aspect
ratio, cuts it into the tiles and puts a 448 x 448 thumbnail of the whole image first (`lv_image_tiles_preprocess`).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Sequence, Tuple

import numpy as np
import torch

from . import _lib

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)      # long_vita/constants.py; normalize_type="imagenet"
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)       # (tools/inference_long_vita.py:828-834)
_PRECISION_BITS = 32 - 8 - 2

_tables: Dict[Tuple[int, int, str], Tuple[torch.Tensor, torch.Tensor, torch.Tensor, int]] = {}


def resample_table(in_size: int, out_size: int):
    """Windows and 22-bit fixed-point weights of PIL's BICUBIC resize from `in_size` to `out_size` pixels
    (Resample.c precompute_coeffs + normalize_coeffs_8bpc).  Returns (xmin int32 [out], count int32 [out],
    coeff int32 [out, ksize], ksize).  Vectorised over the output coordinates in float64 with Pillow's operation order
    per element - the window sum runs left to right, one column at a time - so the table is bit-identical to Pillow's
    (tests/test_oracle_pinning.py, tests/test_preprocess_tiles_host.py); a still image needs four such tables, which a
    per-pixel Python loop made slower than the GPU work itself."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    center = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    lo = np.maximum(np.trunc(center - support + 0.5).astype(np.int64), 0)        # C's (int) cast truncates toward zero
    hi = np.minimum(np.trunc(center + support + 0.5).astype(np.int64), in_size)
    cnt = hi - lo
    x = np.arange(ksize, dtype=np.int64)[None, :]
    valid = x < cnt[:, None]
    t = np.abs(((x + lo[:, None]).astype(np.float64) - center[:, None] + 0.5) * ss)
    a = -0.5
    w = np.where(t < 1.0, ((a + 2.0) * t - (a + 3.0)) * t * t + 1, np.where(t < 2.0, (((t - 5) * t + 8) * t - 4) * a, 0.0))
    w = np.where(valid, w, 0.0)
    ww = np.zeros(out_size, np.float64)
    for c in range(ksize):                     # Pillow accumulates the window sum in this order
        ww = np.where(valid[:, c], ww + w[:, c], ww)
    nz = ww != 0.0
    w = np.where(nz[:, None], w / np.where(nz, ww, 1.0)[:, None], w)
    one = float(1 << _PRECISION_BITS)
    fixed = np.where(w < 0, np.trunc(w * one - 0.5), np.trunc(w * one + 0.5))
    rows = np.where(valid, fixed, 0.0).astype(np.int32)
    return lo.astype(np.int32), cnt.astype(np.int32), np.ascontiguousarray(rows), ksize


def _device_table(in_size: int, out_size: int, device):
    key = (in_size, out_size, str(device))
    if key not in _tables:
        xmin, cnt, rows, ksize = resample_table(in_size, out_size)
        _tables[key] = (torch.from_numpy(xmin).to(device), torch.from_numpy(cnt).to(device), torch.from_numpy(rows).to(device), ksize)
        if len(_tables) > 16:
            _tables.pop(next(iter(_tables)))
    return _tables[key]


def preprocess_frames(frames: torch.Tensor, image_size: int = 448, mean: Sequence[float] = IMAGENET_DEFAULT_MEAN,
                      std: Sequence[float] = IMAGENET_DEFAULT_STD) -> torch.Tensor:
    """frames uint8 [N, H, W, 3] on the GPU (RGB, as decord / PIL decode them) -> bf16 [N, 3, image_size, image_size]."""
    if not (frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[-1] == 3):
        raise ValueError("preprocess_frames expects a CUDA uint8 tensor [N, H, W, 3]")
    frames = frames.contiguous()
    n, H, W, _ = frames.shape
    side = max(H, W)
    xmin, cnt, coeff, ksize = _device_table(side, image_size, frames.device)
    lib = _lib.lib()
    out = torch.empty((n, 3, image_size, image_size), dtype=torch.bfloat16, device=frames.device)
    ws = torch.empty(int(lib.lv_frame_preprocess_ws_bytes(n, H, W, image_size)), dtype=torch.uint8, device=frames.device)
    bg = (C.c_int32 * 3)(*[int(x * 255) for x in mean])                    # image_processor.py:205
    m = (C.c_float * 3)(*mean)
    s = (C.c_float * 3)(*std)
    _lib.check(lib.lv_frame_preprocess(frames.data_ptr(), out.data_ptr(), ws.data_ptr(), xmin.data_ptr(), cnt.data_ptr(),
                                       coeff.data_ptr(), ksize, n, H, W, image_size, bg, m, s,
                                       torch.cuda.current_stream().cuda_stream), "lv_frame_preprocess")
    return out


def dynamic_tile_grid(width: int, height: int, min_patch_grid: int = 1, max_patch_grid: int = 12,
                      image_size: int = 448) -> Tuple[int, int]:
    """(columns, rows) of the tile grid `dynamic_preprocess` resizes a width x height image to
    (image_processor.py:404-426 with find_closest_aspect_ratio :387-401): among all grids of min..max tiles the one
    whose aspect ratio is closest to the image's; on an exact tie a later (larger) grid wins only if the image has more
    than half its pixels.  Candidates are visited in the reference's order - the iteration order of a `set` of the
    (i, j) pairs, stably sorted by i * j - because that order decides ties."""
    aspect = width / height
    seen = set()
    for n in range(min_patch_grid, max_patch_grid + 1):
        for i in range(1, n + 1):
            for j in range(1, n + 1):
                if min_patch_grid <= i * j <= max_patch_grid:
                    seen.add((i, j))
    best, best_diff = (1, 1), float("inf")
    for gx, gy in sorted(seen, key=lambda g: g[0] * g[1]):
        diff = abs(aspect - gx / gy)
        if diff < best_diff:
            best, best_diff = (gx, gy), diff
        elif diff == best_diff and width * height > 0.5 * image_size * image_size * gx * gy:
            best = (gx, gy)
    return best


def _tiles_call(lib, image, out, H, W, out_h, out_w, S, tile_base, m, s):
    xt = _device_table(W, out_w, image.device)
    yt = _device_table(H, out_h, image.device)
    ws = torch.empty(int(lib.lv_image_tiles_ws_bytes(H, out_w)), dtype=torch.uint8, device=image.device)
    _lib.check(lib.lv_image_tiles_preprocess(image.data_ptr(), out.data_ptr(), ws.data_ptr(), xt[0].data_ptr(), xt[1].data_ptr(),
                                             xt[2].data_ptr(), xt[3], yt[0].data_ptr(), yt[1].data_ptr(), yt[2].data_ptr(), yt[3],
                                             H, W, out_h, out_w, S, tile_base, m, s, torch.cuda.current_stream().cuda_stream),
               "lv_image_tiles_preprocess")


def preprocess_image_dynamic(image: torch.Tensor, min_patch_grid: int = 1, max_patch_grid: int = 12, image_size: int = 448,
                             mean: Sequence[float] = IMAGENET_DEFAULT_MEAN, std: Sequence[float] = IMAGENET_DEFAULT_STD):
    """image uint8 [H, W, 3] on the GPU -> (bf16 [n_tiles, 3, image_size, image_size], (grid width, grid height) in pixels)
    like `ImageProcessor.process_dynamic` (image_processor.py:263-285): n_tiles = columns * rows, plus the thumbnail of
    the whole image in front when the grid has more than one tile (use_thumbnail=True, :442-447)."""
    if not (image.is_cuda and image.dtype == torch.uint8 and image.dim() == 3 and image.shape[-1] == 3):
        raise ValueError("preprocess_image_dynamic expects a CUDA uint8 tensor [H, W, 3]")
    image = image.contiguous()
    H, W, _ = image.shape
    gx, gy = dynamic_tile_grid(W, H, min_patch_grid, max_patch_grid, image_size)
    S = image_size
    thumb = 1 if gx * gy > 1 else 0
    lib = _lib.lib()
    out = torch.empty((gx * gy + thumb, 3, S, S), dtype=torch.bfloat16, device=image.device)
    m = (C.c_float * 3)(*mean)
    s = (C.c_float * 3)(*std)
    _tiles_call(lib, image, out, H, W, gy * S, gx * S, S, thumb, m, s)
    if thumb:
        _tiles_call(lib, image, out, H, W, S, S, S, 0, m, s)
    return out, (gx * S, gy * S)
