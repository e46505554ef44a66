"""Synthetic inputs of the reference's shape (there is no network for datasets).

Prompt layout follows tools/inference_long_vita.py:730-748: every frame contributes
[VID_START] + 256 x [VID_CONTEXT] + [VID_END]; `image_indices[0]` is the batch index (0) and
`image_indices[1]` the 256 sequence positions of the frame's context tokens.  Text ids are
uniform over the ordinary vocabulary.  Frames are N(0,1) bf16 (the post-ImageNet-normalise
distribution)."""
from __future__ import annotations

from typing import Tuple

import torch

from .config import LongVITAConfig

VID_START_ID, VID_CONTEXT_ID, VID_END_ID = 151652, 151654, 151653   # <|begin/context/end_of_video|>-like ids
TEXT_VOCAB = 151643


def build_prompt(cfg: LongVITAConfig, n_frames: int, n_text: int = 16, pad_multiple: int = 256, seed: int = 1234
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (input_ids [1, S] int64, image_indices [2, n_frames, 256] int64) on the CPU.
    S = n_frames * 258 + n_text, padded with text tokens up to a multiple of `pad_multiple`."""
    g = torch.Generator().manual_seed(seed)
    tpi = cfg.visual.tokens_per_image
    vocab = min(TEXT_VOCAB, cfg.vocab_size)
    special = [min(t, cfg.vocab_size - 1) for t in (VID_START_ID, VID_CONTEXT_ID, VID_END_ID)]
    per = tpi + 2
    s_raw = n_frames * per + n_text
    S = (s_raw + pad_multiple - 1) // pad_multiple * pad_multiple
    ids = torch.randint(0, vocab, (S,), generator=g)
    if n_frames:
        frame = torch.tensor([special[0]] + [special[1]] * tpi + [special[2]])
        ids[: n_frames * per] = frame.repeat(n_frames)
    starts = torch.arange(n_frames) * per + 1
    idx_s = starts[:, None] + torch.arange(tpi)[None, :]
    image_indices = torch.stack([torch.zeros_like(idx_s), idx_s])
    return ids.unsqueeze(0), image_indices


def synthetic_frames(cfg: LongVITAConfig, n_frames: int, seed: int = 1234, device="cpu", pin: bool = False) -> torch.Tensor:
    g = torch.Generator(device=device).manual_seed(seed + 17)
    size = cfg.visual.image_size
    x = torch.randn((n_frames, 3, size, size), generator=g, device=device, dtype=torch.float32).to(torch.bfloat16)
    if pin and x.device.type == "cpu":
        x = x.pin_memory()
    return x
