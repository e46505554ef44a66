"""Frame preprocessing on the GPU: decoded uint8 frames -> the bf16 [N, 3, 448, 448] tensor of the
vision tower, bit-identical to `ImageProcessor.process_images` + `.to(bfloat16)`
(long_vita/data/processor/image_processor.py:183-223; called per video by tools/inference_long_vita.py:248-290).

Host side: the resampling windows and fixed-point weights of Pillow's bicubic filter are computed here in float64
exactly as Pillow's `precompute_coeffs` / `normalize_coeffs_8bpc` do (they depend only on the canvas side and the
target size, so one small table serves every frame of a video) and cached on the device; the pixel work - padding to
a square with the mean colour, both resampling passes, scaling, normalisation, channel-first bf16 output - runs in
`lv_frame_preprocess` (csrc/preprocess.cu).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Sequence, Tuple

import torch

from . import _lib

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)      # long_vita/constants.py; normalize_type="imagenet"
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)       # (tools/inference_long_vita.py:828-834)
_PRECISION_BITS = 32 - 8 - 2

_tables: Dict[Tuple[int, int, str], Tuple[torch.Tensor, torch.Tensor, torch.Tensor, int]] = {}


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_table(in_size: int, out_size: int):
    """Windows and 22-bit fixed-point weights of PIL's BICUBIC resize from `in_size` to `out_size` pixels
    (Resample.c precompute_coeffs + normalize_coeffs_8bpc).  Returns python lists (xmin, count, coeff rows, ksize)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    xmin, cnt, rows = [], [], []
    one = float(1 << _PRECISION_BITS)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = max(int(center - support + 0.5), 0)
        hi = min(int(center + support + 0.5), in_size)
        w = [_bicubic((x + lo - center + 0.5) * ss) for x in range(hi - lo)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        row = [int(v * one - 0.5) if v < 0 else int(v * one + 0.5) for v in w]
        rows.append(row + [0] * (ksize - len(row)))
        xmin.append(lo)
        cnt.append(hi - lo)
    return xmin, cnt, rows, ksize


def _device_table(in_size: int, out_size: int, device):
    key = (in_size, out_size, str(device))
    if key not in _tables:
        xmin, cnt, rows, ksize = resample_table(in_size, out_size)
        _tables[key] = (torch.tensor(xmin, dtype=torch.int32, device=device), torch.tensor(cnt, dtype=torch.int32, device=device),
                        torch.tensor(rows, dtype=torch.int32, device=device).contiguous(), ksize)
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
