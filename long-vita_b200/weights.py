"""Synthetic ("random-init") weights with the reference's HF state-dict names and shapes.

There is no network for checkpoints, so benches and tests materialise weights from seeds:
every Linear / Embedding ~ N(0, initializer_range^2) (config_14B.json:17,40;
resampler_projector.py:62-66), biases 0, norm weights 1, layer-scale = initializer_factor,
class / position embeddings ~ N(0, 1) (modeling_intern_vit.py:75-86).  Each layer is drawn from
`seed + layer_index` so a CPU oracle and the GPU build can materialise identical tensors layer by
layer without holding the whole 14B model.  `perturb=True` additionally randomises biases, norm
weights and layer-scales so parity tests exercise those paths.

Names follow `LongVITAForCausalLM` (modeling_long_vita.py:57-72, 229-236): `model.layers.N...`,
`model.vision_model...`, `model.vision_projection...`, `lm_head.weight`.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional

import torch

from .config import LongVITAConfig

VIT_SEED_OFFSET = 100_000
GLOBAL_SEED_OFFSET = 200_000


def _gen(seed: int, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def _normal(shape, std, g, device, dtype):
    return (torch.randn(shape, generator=g, device=device, dtype=torch.float32) * std).to(dtype)


def llm_layer_weights(cfg: LongVITAConfig, i: int, seed: int = 1234, device="cpu", dtype=torch.bfloat16,
                      perturb: bool = False) -> Dict[str, torch.Tensor]:
    g = _gen(seed + i, device)
    std = cfg.initializer_range
    H, I = cfg.hidden_size, cfg.intermediate_size
    p = f"model.layers.{i}."
    w = {
        p + "self_attn.q_proj.weight": _normal((cfg.q_size, H), std, g, device, dtype),
        p + "self_attn.k_proj.weight": _normal((cfg.kv_size, H), std, g, device, dtype),
        p + "self_attn.v_proj.weight": _normal((cfg.kv_size, H), std, g, device, dtype),
        p + "self_attn.o_proj.weight": _normal((H, cfg.q_size), std, g, device, dtype),
        p + "mlp.gate_proj.weight": _normal((I, H), std, g, device, dtype),
        p + "mlp.up_proj.weight": _normal((I, H), std, g, device, dtype),
        p + "mlp.down_proj.weight": _normal((H, I), std, g, device, dtype),
    }
    for name, n in (("q_proj", cfg.q_size), ("k_proj", cfg.kv_size), ("v_proj", cfg.kv_size)):
        w[p + f"self_attn.{name}.bias"] = (_normal((n,), 0.5, g, device, dtype) if perturb
                                           else torch.zeros(n, device=device, dtype=dtype))
    for name in ("input_layernorm", "post_attention_layernorm"):
        w[p + name + ".weight"] = ((1 + _normal((H,), 0.1, g, device, torch.float32)).to(dtype) if perturb
                                   else torch.ones(H, device=device, dtype=dtype))
    return w


def vit_layer_weights(cfg: LongVITAConfig, i: int, seed: int = 1234, device="cpu", dtype=torch.bfloat16,
                      perturb: bool = False) -> Dict[str, torch.Tensor]:
    v = cfg.visual
    g = _gen(seed + VIT_SEED_OFFSET + i, device)
    std = cfg.initializer_range
    C, I = v.hidden_size, v.intermediate_size
    p = f"model.vision_model.encoder.layers.{i}."

    def bias(n):
        return _normal((n,), 0.1, g, device, dtype) if perturb else torch.zeros(n, device=device, dtype=dtype)

    def ones(n, jitter=0.1):
        if perturb:
            return (1 + _normal((n,), jitter, g, device, torch.float32)).to(dtype)
        return torch.ones(n, device=device, dtype=dtype)

    return {
        p + "attn.qkv.weight": _normal((3 * C, C), std, g, device, dtype),
        p + "attn.qkv.bias": bias(3 * C),
        p + "attn.proj.weight": _normal((C, C), std, g, device, dtype),
        p + "attn.proj.bias": bias(C),
        p + "mlp.fc1.weight": _normal((I, C), std, g, device, dtype),
        p + "mlp.fc1.bias": bias(I),
        p + "mlp.fc2.weight": _normal((C, I), std, g, device, dtype),
        p + "mlp.fc2.bias": bias(C),
        p + "norm1.weight": ones(C),
        p + "norm1.bias": bias(C),
        p + "norm2.weight": ones(C),
        p + "norm2.bias": bias(C),
        p + "ls1": (ones(C) * v.initializer_factor).to(dtype),
        p + "ls2": (ones(C) * v.initializer_factor).to(dtype),
    }


def global_weights(cfg: LongVITAConfig, seed: int = 1234, device="cpu", dtype=torch.bfloat16, perturb: bool = False,
                   with_lm: bool = True, with_vit: bool = True) -> Dict[str, torch.Tensor]:
    v = cfg.visual
    g = _gen(seed + GLOBAL_SEED_OFFSET, device)
    std = cfg.initializer_range
    w: Dict[str, torch.Tensor] = {}
    if with_vit:
        C = v.hidden_size
        pin = C * int(1 / v.downsample_ratio) ** 2
        w["model.vision_model.embeddings.class_embedding"] = _normal((1, 1, C), 1.0, g, device, dtype)
        w["model.vision_model.embeddings.position_embedding"] = _normal((1, v.num_patches + 1, C), 1.0, g, device, dtype)
        w["model.vision_model.embeddings.patch_embedding.weight"] = _normal((C, 3, v.patch_size, v.patch_size), std, g, device, dtype)
        w["model.vision_model.embeddings.patch_embedding.bias"] = (
            _normal((C,), 0.1, g, device, dtype) if perturb else torch.zeros(C, device=device, dtype=dtype))
        w["model.vision_projection.pre_proj_layernorm.weight"] = (
            (1 + _normal((pin,), 0.1, g, device, torch.float32)).to(dtype) if perturb else torch.ones(pin, device=device, dtype=dtype))
        w["model.vision_projection.pre_proj_layernorm.bias"] = (
            _normal((pin,), 0.1, g, device, dtype) if perturb else torch.zeros(pin, device=device, dtype=dtype))
        w["model.vision_projection.mlp.0.weight"] = _normal((C, pin), std, g, device, dtype)
        w["model.vision_projection.mlp.2.weight"] = _normal((cfg.hidden_size, C), std, g, device, dtype)
    if with_lm:
        w["model.embed_tokens.weight"] = _normal((cfg.vocab_size, cfg.hidden_size), std, g, device, dtype)
        w["model.norm.weight"] = ((1 + _normal((cfg.hidden_size,), 0.1, g, device, torch.float32)).to(dtype) if perturb
                                  else torch.ones(cfg.hidden_size, device=device, dtype=dtype))
        w["lm_head.weight"] = _normal((cfg.vocab_size, cfg.hidden_size), std, g, device, dtype)
    return w


def synthetic_state_dict(cfg: LongVITAConfig, seed: int = 1234, device="cpu", dtype=torch.bfloat16,
                         perturb: bool = False, llm_layers: Optional[Iterable[int]] = None,
                         vit_layers: Optional[Iterable[int]] = None) -> Dict[str, torch.Tensor]:
    """Whole (or partial) model.  Drawing on `device` directly keeps 14B bf16 (29.5 GB) off the host."""
    w = global_weights(cfg, seed, device, dtype, perturb)
    for i in (range(cfg.visual.num_hidden_layers) if vit_layers is None else vit_layers):
        w.update(vit_layer_weights(cfg, i, seed, device, dtype, perturb))
    for i in (range(cfg.num_hidden_layers) if llm_layers is None else llm_layers):
        w.update(llm_layer_weights(cfg, i, seed, device, dtype, perturb))
    return w
