"""Megatron-free harness reproducing the calling convention of the operator surface.

Megatron-LM core_r0.7.0 is an empty submodule in the reference and not installable here (no
network), so the claim "drops into pretrain_long_vita.py unchanged" cannot be executed in this
container.  This stub reproduces exactly the parts of the convention the hot path touches, as cited
in SURVEY.md 8b, so that tests exercise the same call shapes either way:

* `ModuleSpec(module=cls, params=..., submodules=...)` + `build_module(spec, **kw)`
  (gpt_layer_specs.py:32-55 builds `SelfAttentionSubmodules(core_attention=<cls>)`; Megatron
  instantiates it as `cls(config=, layer_number=, attn_mask_type=, attention_type=)`),
* `AttnMaskType` with the members the specs use (`causal`, `no_mask`, `padding`),
* a `TransformerConfig` namespace with the fields the attention module reads,
* the reference's own patch-registry semantics (patch_utils.py:38-71): a function whose name ends
  in `wrapper` / `decorator` decorates the original attribute.
"""
from __future__ import annotations

import enum
from dataclasses import dataclass, field
from typing import Any, Optional


class AttnMaskType(enum.Enum):
    padding = 1
    causal = 2
    no_mask = 3
    padding_causal = 4


@dataclass
class TransformerConfig:
    hidden_size: int = 5120
    num_attention_heads: int = 40
    num_query_groups: int = 8
    kv_channels: Optional[int] = None
    attention_dropout: float = 0.0
    context_parallel_size: int = 1
    apply_query_key_layer_scaling: bool = False


@dataclass
class ModuleSpec:
    module: Any
    params: dict = field(default_factory=dict)
    submodules: Any = None


def build_module(spec, *args, **kwargs):
    if isinstance(spec, ModuleSpec):
        kw = dict(spec.params)
        kw.update(kwargs)
        if spec.submodules is not None:
            kw["submodules"] = spec.submodules
        return spec.module(*args, **kw)
    return spec(*args, **kwargs)


class DotProductAttention:
    """Stand-in for megatron.core.transformer.dot_product_attention.DotProductAttention - only the
    attributes the reference's wrapper reads (dot_product_attention.py:170-184, 331-332)."""

    def __init__(self, config: TransformerConfig, layer_number: int, attn_mask_type, attention_type="self"):
        self.config = config
        self.attn_mask_type = attn_mask_type
        self.num_attention_heads_per_partition = config.num_attention_heads
        self.num_query_groups_per_partition = config.num_query_groups
        self.hidden_size_per_attention_head = config.kv_channels or config.hidden_size // config.num_attention_heads
        self.softmax_scale = None

    def forward(self, query, key, value, attention_mask, attn_mask_type=None, packed_seq_params=None):
        raise RuntimeError("the un-patched eager Megatron attention must not run on the hot path")


def apply_reference_style_patch(owner, attr: str, replacement) -> None:
    """patch_utils.Patch.apply_patch semantics: `*wrapper` / `*decorator` functions receive the
    original and return the new attribute; anything else replaces it."""
    orig = getattr(owner, attr)
    name = getattr(replacement, "__name__", "")
    if name.endswith(("wrapper", "decorator")):
        setattr(owner, attr, replacement(orig))
    else:
        setattr(owner, attr, replacement)
