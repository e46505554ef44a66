"""K/V cache for incremental decoding.

The reference's Megatron serving loop re-runs the whole prefill for every generated token
(long_vita_megatron/inference/text_generation/generation.py:127-135: with `use_kv_cache` off the full
`tokens` are fed each step) - O(S^2) per token.  Its HF path delegates the cache to transformers'
`DynamicCache` (modeling_long_vita.py:150-151), which grows by `torch.cat` and so re-copies the whole
cache every step.  Here the cache is one pre-allocated [capacity, hkv, d] bf16 buffer per layer for K and
for V, holding post-RoPE rows; a step appends one row per layer and the decode attention reads each row
once (`ops.attention_decode`).

Duck-types the part of transformers' `Cache` the reference touches: `len(cache)` (`modeling_long_vita.py:90`)
and `get_seq_length()` (:154).
"""
from __future__ import annotations

from typing import List, Tuple

import torch


class KVCache:
    def __init__(self, num_layers: int, capacity: int, num_kv_heads: int, head_dim: int, device, dtype=torch.bfloat16):
        self.capacity = int(capacity)
        self.k: List[torch.Tensor] = [torch.empty((self.capacity, num_kv_heads, head_dim), dtype=dtype, device=device)
                                      for _ in range(num_layers)]
        self.v: List[torch.Tensor] = [torch.empty((self.capacity, num_kv_heads, head_dim), dtype=dtype, device=device)
                                      for _ in range(num_layers)]
        self._len = 0            # tokens whose K/V rows are valid in EVERY layer
        self._pending = 0        # rows appended by the forward pass in flight (committed at its end)

    def __len__(self) -> int:
        return self._len

    def get_seq_length(self, layer_idx: int = 0) -> int:
        return self._len

    def append(self, layer_idx: int, k: torch.Tensor, v: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, int]:
        """Write k / v [t, hkv, d] behind the committed rows of `layer_idx`; returns (K, V, new_length)
        where K / V are the cache buffers (valid rows [0, new_length))."""
        t = k.shape[0]
        end = self._len + t
        if end > self.capacity:
            raise RuntimeError(f"KV cache overflow: {end} tokens > capacity {self.capacity}")
        self.k[layer_idx][self._len : end].copy_(k)
        self.v[layer_idx][self._len : end].copy_(v)
        self._pending = t
        return self.k[layer_idx], self.v[layer_idx], end

    def commit(self) -> None:
        """Called once per forward pass, after the last layer."""
        self._len += self._pending
        self._pending = 0

    def bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.k + self.v)
