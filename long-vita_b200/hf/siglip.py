"""SigLIP-400M vision tower on the same kernels.

Follows long_vita_megatron/core/models/vision/siglip_vit_model.py:165-228 (conv patch embedding,
learned position embedding, no class token) and :29-86 (pre-LN block without layer scale), with the
geometry of pretrain_long_vita.py:268-307: 27 layers, hidden 1152, 16 heads x 72, FFN 4304,
tanh-GELU, LayerNorm, qkv / linear bias.  head_dim 72 is not a tensor-core tile size, so q, k, v are
zero-padded to 128 per head (exact: the padded dot products and output columns are zero) and the
d=128 instantiation of the fused attention kernel runs with scale 72^-0.5.

Parameter names are Megatron's for this model: conv1.*, position_embeddings.weight,
decoder.layers.N.{input_layernorm, self_attention.linear_qkv, self_attention.linear_proj,
pre_mlp_layernorm, mlp.linear_fc1, mlp.linear_fc2}.  `linear_qkv` rows are interleaved per head
[head, (q, k, v), hn] as Megatron stores them (tools/hf2mcore_long_vita.py:397-414).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import torch
import torch.nn.functional as F

from .. import ops


@dataclass(frozen=True)
class SigLIPConfig:
    hidden_size: int = 1152
    ffn_hidden_size: int = 4304
    num_layers: int = 27
    num_attention_heads: int = 16
    kv_channels: int = 72
    image_size: int = 448
    patch_dim: int = 14
    layernorm_epsilon: float = 1e-6

    @property
    def num_patches(self) -> int:
        return (self.image_size // self.patch_dim) ** 2


class SigLIPViTModel:
    def __init__(self, cfg: SigLIPConfig, w: Dict[str, torch.Tensor], prefix: str = ""):
        self.cfg = cfg
        self.patch_w = ops.pad_patch_weight(w[prefix + "conv1.weight"])
        self.patch_b = w[prefix + "conv1.bias"]
        self.pos = w[prefix + "position_embeddings.weight"]
        self.layers = []
        for i in range(cfg.num_layers):
            p = f"{prefix}decoder.layers.{i}."
            self.layers.append({k[len(p):]: t for k, t in w.items() if k.startswith(p)})

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        c = self.cfg
        n = x.shape[0]
        H, hn, C = c.num_attention_heads, c.kv_channels, c.hidden_size
        h = ops.patch_embed(x, self.patch_w, self.patch_b, None, self.pos, c.patch_dim)    # [n, P, C], no cls
        S = h.shape[1]
        h = h.view(n * S, C)
        for L in self.layers:
            y = ops.layernorm(h, L["input_layernorm.weight"], L["input_layernorm.bias"], c.layernorm_epsilon)
            qkv = ops.linear(y, L["self_attention.linear_qkv.weight"], L["self_attention.linear_qkv.bias"])
            qkv = F.pad(qkv.view(n, S, H, 3, hn), (0, 128 - hn))          # [n, S, H, 3, 128], zero columns
            att = ops.attention_fwd(qkv[:, :, :, 0], qkv[:, :, :, 1], qkv[:, :, :, 2], causal=False, scale=hn ** -0.5)
            att = att[..., :hn].reshape(n * S, H * hn)
            o = ops.linear(att, L["self_attention.linear_proj.weight"], L["self_attention.linear_proj.bias"])
            h = ops.ls_residual(h, o)
            y = ops.layernorm(h, L["pre_mlp_layernorm.weight"], L["pre_mlp_layernorm.bias"], c.layernorm_epsilon)
            f = ops.linear(y, L["mlp.linear_fc1.weight"], L["mlp.linear_fc1.bias"], act="gelu_tanh")
            f = ops.linear(f, L["mlp.linear_fc2.weight"], L["mlp.linear_fc2.bias"])
            h = ops.ls_residual(h, f)
        return h.view(n, S, C)

    __call__ = forward
