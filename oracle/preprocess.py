"""CPU oracle - frame preprocessing.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates `ImageProcessor.process_images` (long_vita/data/processor/image_processor.py:183-223): every frame is
padded to a square with the mean colour (`expand2square`, :192-203, background = int(mean * 255) per channel :205),
resized to image_size x image_size with PIL's BICUBIC filter (:207-209), scaled by 1/255 and normalised with the
ImageNet mean / std in float32 (:211-216), and returned channel-first (:218-221).

The arithmetic of the resize lives in a third-party dependency that IS installed here but is not part of the reference:
Pillow (unpinned in requirements.txt; 12.2.0 in this image), `ImagingResample` in src/libImaging/Resample.c.  Its 8-bit
path is integer arithmetic, restated here with numpy:
  * per output coordinate a window [xmin, xmin + n) of input pixels and float64 weights of the bicubic kernel
    (a = -0.5) stretched by the down-scale factor (antialiasing), normalised to sum 1  (precompute_coeffs);
  * weights converted to fixed point with PRECISION_BITS = 32 - 8 - 2 = 22, round half away from zero
    (normalize_coeffs_8bpc);
  * horizontal pass over the rows the vertical pass needs, then vertical pass, each accumulating in int32 from
    1 << 21 and clipping (acc >> 22) to [0, 255] - so the intermediate image is uint8 again.
Dynamic-patch tiling of still images (`process_dynamic` :263-285 = `dynamic_preprocess` :404-448 + `process_images`) is
restated in `dynamic_grid` / `process_dynamic` below and pinned the same way.
PINNING: tests/test_oracle_pinning.py runs the reference's own `process_images` (with Pillow) from /root/reference on
seeded synthetic frames and requires bit-identical float32 output; tests/golden/ref_preprocess.pt carries the same
frames and outputs for the GPU box.
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)      # long_vita/constants.py (normalize_type="imagenet")
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)
PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the whole-image box [0, in_size).
    Returns (xmin int32 [out], count int32 [out], coeff int32 [out, ksize])."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    cnt = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        lo = max(lo, 0)
        hi = int(center + support + 0.5)
        hi = min(hi, in_size)
        n = hi - lo
        w = [_bicubic((x + lo - center + 0.5) * ss) for x in range(n)]
        ww = sum(w)          # Pillow accumulates in the same left-to-right order
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(v * (1 << PRECISION_BITS) - 0.5) if v < 0 else int(v * (1 << PRECISION_BITS) + 0.5)
        xmin[xx], cnt[xx] = lo, n
    return xmin, cnt, kk


def _pass(img: np.ndarray, xmin, cnt, kk, axis: int) -> np.ndarray:
    """One resampling pass of a uint8 [H, W, C] image along `axis` (0 = vertical, 1 = horizontal)."""
    src = img.astype(np.int64)
    out_n = xmin.shape[0]
    shape = list(img.shape)
    shape[axis] = out_n
    out = np.empty(shape, np.uint8)
    for o in range(out_n):
        acc = np.full(src.take(0, axis=axis).shape, 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(int(cnt[o])):
            acc += src.take(int(xmin[o]) + x, axis=axis) * int(kk[o, x])
        v = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
        if axis == 0:
            out[o] = v
        else:
            out[:, o] = v
    return out


def resize_bicubic_u8(img: np.ndarray, out_size: int) -> np.ndarray:
    """PIL `Image.resize((out, out), BICUBIC)` of a uint8 [H, W, 3] image: horizontal pass, then vertical pass;
    a pass whose size does not change is skipped (Resample.c need_horizontal / need_vertical)."""
    h, w, _ = img.shape
    cur = img
    if w != out_size:
        xm, cn, kk = resample_coeffs(w, out_size)
        if h != out_size:
            # Pillow resamples only the rows the vertical pass will read; the others never influence the result
            pass
        cur = _pass(cur, xm, cn, kk, axis=1)
    if h != out_size:
        ym, cn, kk = resample_coeffs(h, out_size)
        cur = _pass(cur, ym, cn, kk, axis=0)
    return cur


def resize_bicubic_u8_rect(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """PIL `Image.resize((out_w, out_h))` (default filter of Pillow >= 7: BICUBIC) of a uint8 [H, W, 3] image."""
    h, w, _ = img.shape
    cur = img
    if w != out_w:
        cur = _pass(cur, *resample_coeffs(w, out_w), axis=1)
    if h != out_h:
        cur = _pass(cur, *resample_coeffs(h, out_h), axis=0)
    return cur


def dynamic_grid(width: int, height: int, min_num: int = 1, max_num: int = 12, image_size: int = 448) -> Tuple[int, int]:
    """Tile grid (columns, rows) of `dynamic_preprocess` (image_processor.py:404-426): candidate grids with
    min_num <= columns * rows <= max_num collected in a set, sorted by tile count (:409-416), and the one closest in
    aspect ratio chosen by `find_closest_aspect_ratio` (:387-401; ties go to the later candidate when the image area
    exceeds half the grid's area)."""
    ratio = width / height
    cands = sorted({(i, j) for n in range(min_num, max_num + 1) for i in range(1, n + 1) for j in range(1, n + 1)
                    if min_num <= i * j <= max_num}, key=lambda x: x[0] * x[1])
    best, best_d = (1, 1), float("inf")
    area = width * height
    for c in cands:
        d = abs(ratio - c[0] / c[1])
        if d < best_d:
            best_d, best = d, c
        elif d == best_d and area > 0.5 * image_size * image_size * c[0] * c[1]:
            best = c
    return best


def process_dynamic(img: np.ndarray, min_num: int = 1, max_num: int = 12, image_size: int = 448, mean=IMAGENET_DEFAULT_MEAN,
                    std=IMAGENET_DEFAULT_STD):
    """uint8 [H, W, 3] image -> (float32 [n, 3, S, S], (grid width, grid height)): process_dynamic (:263-285) - resize to
    the grid (:424-429), crop the tiles row-major (:431-440), thumbnail of the whole image first if there is more than
    one tile (:442-447), then process_images on the list (tiles are square and already S x S, so its padding and resize
    leave them unchanged)."""
    h, w, _ = img.shape
    gx, gy = dynamic_grid(w, h, min_num, max_num, image_size)
    S = image_size
    big = resize_bicubic_u8_rect(img, gx * S, gy * S)
    tiles = [big[(i // gx) * S : (i // gx + 1) * S, (i % gx) * S : (i % gx + 1) * S] for i in range(gx * gy)]
    if len(tiles) != 1:
        tiles = [resize_bicubic_u8_rect(img, S, S)] + tiles
    return process_frames(tiles, image_size, mean, std), (gx * S, gy * S)


def expand2square(img: np.ndarray, background: Sequence[int]) -> np.ndarray:
    """image_processor.py:192-203: paste centred on a square canvas of the mean colour."""
    h, w, _ = img.shape
    if h == w:
        return img
    n = max(h, w)
    out = np.empty((n, n, 3), np.uint8)
    out[:] = np.asarray(background, np.uint8)
    if w > h:
        top = (w - h) // 2
        out[top : top + h] = img
    else:
        left = (h - w) // 2
        out[:, left : left + w] = img
    return out


def process_frames(frames: Sequence[np.ndarray], image_size: int = 448, mean=IMAGENET_DEFAULT_MEAN,
                   std=IMAGENET_DEFAULT_STD) -> np.ndarray:
    """uint8 [H, W, 3] frames -> float32 [N, 3, image_size, image_size] (process_images, :183-223)."""
    out = np.ones((len(frames), 3, image_size, image_size), np.float32)
    bg = tuple(int(x * 255) for x in mean)
    m = np.array(mean, np.float32)
    s = np.array(std, np.float32)
    for i, f in enumerate(frames):
        sq = expand2square(np.asarray(f, np.uint8), bg)
        r = resize_bicubic_u8(sq, image_size).astype(np.float32)
        r = r * np.float32(1.0) / np.float32(255.0)
        r = (r - m) / s
        out[i] = r.transpose(2, 0, 1)
    return out
