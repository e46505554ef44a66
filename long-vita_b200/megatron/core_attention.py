"""Megatron-core operator surface of the attention hot path (boundaries B1 and B3).

B1 - `B200DotProductAttention`: a `core_attention` module with the constructor / forward the
reference's layer specs expect (long_vita_megatron/core/models/gpt/gpt_layer_specs.py:38-44 puts
`TEDotProductAttention` there; vit_layer_specs.py:87-91 the ViT twin).  Megatron 0.7 builds it as
`cls(config, layer_number, attn_mask_type, attention_type)` and calls
`forward(query, key, value, attention_mask, attn_mask_type=, packed_seq_params=)` with
query [sq, b, np, hn], key/value [sk, b, ng, hn] (GQA not expanded) and expects [sq, b, np*hn]
(long_vita_megatron/core/transformer/dot_product_attention.py:153, 176-184, 392).

B3 - `b200_dot_product_attention_forward_wrapper`: a function whose name ends in `wrapper`, so the
reference's patch registry treats it as a decorator over the original
`DotProductAttention.forward` (patch_utils.py:46-53, 62-63):
    MindSpeedPatchesManager.register_patch(
        'megatron.core.transformer.dot_product_attention.DotProductAttention.forward',
        b200_dot_product_attention_forward_wrapper)
registered after `import long_vita_megatron.megatron_adaptor` (megatron_adaptor.py:21-22).

Megatron-core is NOT importable in the build container (un-vendored submodule), so
nothing here imports it: the mask type is duck-typed by name and process-group lookups go through
`parallel_state` only when context parallelism is on.  tests/test_gpu_surfaces.py (kernels) and
tests/test_surfaces_host.py (host logic, CPU) drive these objects with a stub that reproduces the
ModuleSpec / build_module calling convention.
"""
from __future__ import annotations

import math
from functools import wraps
from typing import Optional

import torch

from .. import ops


def _is_causal(attn_mask_type) -> bool:
    """AttnMaskType.causal / padding_causal -> True; no_mask / padding -> False (duck-typed by name)."""
    name = getattr(attn_mask_type, "name", str(attn_mask_type)).lower()
    return "causal" in name


def _attention_sbhd(query, key, value, causal: bool, scale: float, cp_ctx=None):
    sq, b, np_, hn = query.shape
    if cp_ctx is not None:
        if b != 1:
            raise AssertionError("context-parallel attention supports micro-batch 1 (the reference's long-context setting)")
        if torch.is_grad_enabled() and (query.requires_grad or key.requires_grad or value.requires_grad):
            from ..cp import cp_attention      # training: fused exchange forward + lv_attn_bwd / reduce-scatter backward

            out = cp_attention(query[:, 0], key[:, 0], value[:, 0], cp_ctx, scale)
        else:
            out = cp_ctx.attention_separate(query[:, 0], key[:, 0], value[:, 0], scale=scale)   # [sq, np*hn]
        return out.view(sq, 1, np_ * hn)
    if torch.is_grad_enabled() and (query.requires_grad or key.requires_grad or value.requires_grad):
        # training: differentiable path (lv_attn_bwd through torch.autograd.Function)
        out = ops.attention(query.permute(1, 0, 2, 3), key.permute(1, 0, 2, 3), value.permute(1, 0, 2, 3),
                            causal=causal, scale=scale)                       # [b, sq, np, hn]
        return out.permute(1, 0, 2, 3).reshape(sq, b, np_ * hn)
    out = ops.attention_fwd(query, key, value, causal=causal, scale=scale, layout="sbhd")
    return out.reshape(sq, b, np_ * hn)


class B200DotProductAttention(torch.nn.Module):
    """Drop-in for the `core_attention` ModuleSpec slot."""

    def __init__(self, config, layer_number: int, attn_mask_type, attention_type: str = "self",
                 attention_dropout: Optional[float] = None):
        super().__init__()
        self.config = config
        self.layer_number = max(1, layer_number)
        self.attn_mask_type = attn_mask_type
        self.attention_type = attention_type
        p = config.attention_dropout if attention_dropout is None else attention_dropout
        if p not in (0, 0.0, None):
            raise ValueError("attention dropout is not supported by the fused kernel (the reference trains with 0.0)")
        kv_channels = getattr(config, "kv_channels", None) or config.hidden_size // config.num_attention_heads
        self.hidden_size_per_attention_head = kv_channels
        self.softmax_scale = 1.0 / math.sqrt(kv_channels)
        if getattr(config, "apply_query_key_layer_scaling", False):
            # Megatron divides by layer_number and multiplies back inside the fp32 softmax: a no-op here
            pass
        self.cp_size = int(getattr(config, "context_parallel_size", 1) or 1)
        self._cp_ctx = None

    def _cp(self, query, key):
        if self.cp_size <= 1 or not _is_causal(self.attn_mask_type):
            return None
        sq, _, np_, hn = query.shape
        if self._cp_ctx is None or self._cp_ctx.S != sq * self.cp_size:
            # keyed on the sequence length of THIS call: a later micro-batch of another length gets its own buffers
            from megatron.core import parallel_state as mpu   # only reached inside a Megatron job

            from ..cp import CPContext

            self._cp_ctx = CPContext.shared(mpu.get_context_parallel_group(), sq * self.cp_size, np_, key.shape[2], hn,
                                            query.device, fused_qkv=False)
        return self._cp_ctx

    def forward(self, query, key, value, attention_mask=None, attn_mask_type=None, packed_seq_params=None):
        assert packed_seq_params is None, (
            "Packed sequence is not supported by B200DotProductAttention."     # same contract as the reference,
        )                                                                     # dot_product_attention.py:156-159
        mask_type = self.attn_mask_type if attn_mask_type is None else attn_mask_type
        causal = _is_causal(mask_type)
        return _attention_sbhd(query, key, value, causal, self.softmax_scale, self._cp(query, key))


def b200_dot_product_attention_forward_wrapper(fn):
    """Decorator-style patch for `DotProductAttention.forward` (see module docstring)."""

    @wraps(fn)
    def wrapper(self, query, key, value, attention_mask, attn_mask_type=None, packed_seq_params=None):
        assert packed_seq_params is None, (
            "Packed sequence is not supported by DotProductAttention."
            "Please use TEDotProductAttention instead."
        )
        if not (query.is_cuda and query.dtype == torch.bfloat16):
            return fn(self, query, key, value, attention_mask, attn_mask_type, packed_seq_params)
        mask_type = getattr(self, "attn_mask_type", None) if attn_mask_type is None else attn_mask_type
        hn = query.shape[-1]
        scale = getattr(self, "softmax_scale", None) or 1.0 / math.sqrt(hn)
        return _attention_sbhd(query, key, value, _is_causal(mask_type), scale)

    return wrapper


def register_b200_patches(patches_manager=None) -> None:
    """Register + apply the attention patch through the reference's own registry
    (long_vita_megatron/patch_utils.py:105-118).  Call after importing
    long_vita_megatron.megatron_adaptor."""
    if patches_manager is None:
        from long_vita_megatron.patch_utils import MindSpeedPatchesManager as patches_manager  # noqa: N813
    patches_manager.register_patch(
        "megatron.core.transformer.dot_product_attention.DotProductAttention.forward",
        b200_dot_product_attention_forward_wrapper,
    )
    patches_manager.apply_patches()
