"""CPU oracle - whole-model restatement of the reference's HF forward.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows long_vita/models/long_vita_qwen2_intern/modeling_long_vita.py:74-221 (composition:
vision tower :91, drop cls :97, projector :98, embed :138, index_put :143-147, position_ids =
cache_position :151-158, decoder loop :175-204, final norm :209) and :238-327 (lm_head over the
last num_logits_to_keep rows :311).  The vision tower restates modeling_intern_vit.py:96-108
(embeddings), :144-161 (`_naive_attn`, the CPU path of InternAttention), :217-229 (encoder layer)
and resampler_projector.py:26-46.  The decoder layer restates transformers' Qwen2DecoderLayer /
Qwen2Attention / Qwen2MLP / Qwen2RMSNorm / apply_rotary_pos_emb (un-vendored third party,
`transformers>=4.48.3` per requirements.txt:13; the reference file itself no longer imports under
the installed transformers 5.5.0 - SURVEY.md 8c).

Pinning: tests/test_oracle_pinning.py checks `long_vita_forward` (every decoder layer's input, the final
normed state and the logits) against outputs of the reference's OWN `LongVITAForCausalLM.forward`, executed
from /root/reference in the build container (oracle/ref_loader.py: three API shims for the installed
transformers 5.5, none touching arithmetic) and committed as tests/golden/ref_long_vita_tiny.pt - agreement
5e-7 relative in fp32; `vit_forward` / `projector_forward` against golden outputs of the reference's own
InternVisionModel / ResamplerProjector; and `decoder_layer` against the installed
transformers.Qwen2DecoderLayer on identical weights (tests/golden/make_golden.py generates the fixtures).

Every function takes a flat state dict with the reference's HF parameter names and computes in
the dtype of its inputs (fp32 for the accumulating oracle, bf16 to reproduce the eager rounding
sequence the reference would execute on CPU).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import ops as O


# ------------------------------------------------------------------------------------------------
# vision tower + projector
# ------------------------------------------------------------------------------------------------
def vit_layer(x, w, p, num_heads, eps):
    """x + ls1 * attn(norm1(x)); x + ls2 * mlp(norm2(x))  (modeling_intern_vit.py:217-229)."""
    B, N, C = x.shape
    d = C // num_heads
    h = F.layer_norm(x, (C,), w[p + "norm1.weight"], w[p + "norm1.bias"], eps)
    qkv = F.linear(h, w[p + "attn.qkv.weight"], w.get(p + "attn.qkv.bias"))
    qkv = qkv.reshape(B, N, 3, num_heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    attn = ((q * (d ** -0.5)) @ k.transpose(-2, -1)).softmax(dim=-1)
    a = (attn @ v).transpose(1, 2).reshape(B, N, C)
    a = F.linear(a, w[p + "attn.proj.weight"], w[p + "attn.proj.bias"])
    x = x + a * w[p + "ls1"]
    h = F.layer_norm(x, (C,), w[p + "norm2.weight"], w[p + "norm2.bias"], eps)
    f = F.linear(h, w[p + "mlp.fc1.weight"], w[p + "mlp.fc1.bias"])
    f = F.gelu(f)
    f = F.linear(f, w[p + "mlp.fc2.weight"], w[p + "mlp.fc2.bias"])
    return x + f * w[p + "ls2"]


def vit_forward(cfg, w: Dict[str, torch.Tensor], images: torch.Tensor, prefix="model.vision_model.",
                num_layers: Optional[int] = None):
    v = cfg.visual
    e = prefix + "embeddings."
    x = O.patch_embed(images, w[e + "patch_embedding.weight"], w[e + "patch_embedding.bias"],
                      w[e + "class_embedding"], w[e + "position_embedding"])
    n_layers = v.num_hidden_layers if num_layers is None else num_layers
    for i in range(n_layers):
        x = vit_layer(x, w, f"{prefix}encoder.layers.{i}.", v.num_attention_heads, v.layer_norm_eps)
    return x


def projector_forward(cfg, w, vit_tokens_no_cls, prefix="model.vision_projection."):
    """resampler_projector.py:26-34 on [n, hw*hw, C] (cls already dropped)."""
    v = cfg.visual
    n, _, C = vit_tokens_no_cls.shape
    x = O.pixel_shuffle_half(vit_tokens_no_cls.reshape(n, v.grid, v.grid, C)).reshape(n, -1, 4 * C)
    x = F.layer_norm(x, (4 * C,), w[prefix + "pre_proj_layernorm.weight"], w[prefix + "pre_proj_layernorm.bias"],
                     v.pre_proj_ln_eps)
    x = F.gelu(F.linear(x, w[prefix + "mlp.0.weight"]))
    return F.linear(x, w[prefix + "mlp.2.weight"])


# ------------------------------------------------------------------------------------------------
# Qwen2 decoder
# ------------------------------------------------------------------------------------------------
def decoder_layer(cfg, w, i: int, x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor,
                  q_pos: Optional[torch.Tensor] = None, kv_pos: Optional[torch.Tensor] = None, attention_fn=None):
    """One Qwen2 decoder layer on x [s, H] (batch 1), cos/sin [s, head_dim] in x.dtype.  `attention_fn` (default
    `oracle.ops.attention`) lets bench.py's CPU leg time the attention of one kv group apart from the token-wise part."""
    p = f"model.layers.{i}."
    s = x.shape[0]
    hq, hkv, d = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    h = O.rmsnorm(x, w[p + "input_layernorm.weight"], cfg.rms_norm_eps)
    q = F.linear(h, w[p + "self_attn.q_proj.weight"], w[p + "self_attn.q_proj.bias"]).view(s, hq, d)
    k = F.linear(h, w[p + "self_attn.k_proj.weight"], w[p + "self_attn.k_proj.bias"]).view(s, hkv, d)
    v = F.linear(h, w[p + "self_attn.v_proj.weight"], w[p + "self_attn.v_proj.bias"]).view(s, hkv, d)
    q, k = O.rope_apply(q, cos, sin), O.rope_apply(k, cos, sin)
    att, _ = (attention_fn or O.attention)(q[None], k[None], v[None], causal=True, q_pos=q_pos, kv_pos=kv_pos)
    att = att[0].to(x.dtype).reshape(s, hq * d)
    x = x + F.linear(att, w[p + "self_attn.o_proj.weight"])
    h = O.rmsnorm(x, w[p + "post_attention_layernorm.weight"], cfg.rms_norm_eps)
    g = F.linear(h, w[p + "mlp.gate_proj.weight"])
    u = F.linear(h, w[p + "mlp.up_proj.weight"])
    return x + F.linear(F.silu(g) * u, w[p + "mlp.down_proj.weight"])


def long_vita_forward(cfg, w, input_ids: torch.Tensor, images: Optional[torch.Tensor] = None,
                      image_indices: Optional[torch.Tensor] = None, num_logits_to_keep: int = 0,
                      position_ids: Optional[torch.Tensor] = None, num_layers: Optional[int] = None,
                      return_hidden: bool = False):
    """LongVITAForCausalLM.forward, prefill, batch 1.  Returns logits [1, keep, vocab] (and the
    per-layer hidden states [s, H] when return_hidden)."""
    dtype = w["model.embed_tokens.weight"].dtype
    b, s = input_ids.shape
    assert b == 1
    x = w["model.embed_tokens.weight"][input_ids.view(-1)].clone()
    if images is not None:
        vit = vit_forward(cfg, w, images.to(dtype))
        feat = projector_forward(cfg, w, vit[:, 1:, :])
        ib, isq = image_indices.unbind(dim=0)
        x[ib.reshape(-1) * s + isq.reshape(-1)] = feat.reshape(-1, feat.shape[-1])
    if position_ids is None:
        position_ids = torch.arange(s).unsqueeze(0)
    inv = O.rope_inv_freq(cfg.head_dim, cfg.rope_theta)
    cos, sin = O.rope_tables(position_ids.view(-1), inv, dtype)
    hidden = [x]
    for i in range(cfg.num_hidden_layers if num_layers is None else num_layers):
        x = decoder_layer(cfg, w, i, x, cos, sin)
        hidden.append(x)
    h = O.rmsnorm(x, w["model.norm.weight"], cfg.rms_norm_eps)
    sel = h[-num_logits_to_keep:] if num_logits_to_keep else h
    logits = F.linear(sel, w["lm_head.weight"]).unsqueeze(0)
    return (logits, hidden, h) if return_hidden else logits


def cast_weights(w: Dict[str, torch.Tensor], dtype) -> Dict[str, torch.Tensor]:
    return {k: v.to(dtype) for k, v in w.items()}


def siglip_forward(cfg, w: Dict[str, torch.Tensor], images: torch.Tensor, prefix: str = ""):
    """SigLIPViTModel.forward (long_vita_megatron/core/models/vision/siglip_vit_model.py:165-228) with the
    block of :29-86 and the geometry of pretrain_long_vita.py:268-307.  Pinned (tests/test_oracle_pinning.py) against
    the HF SigLIP vision encoder the Megatron weights are converted from (transformers SiglipVisionModel, last hidden
    state before post_layernorm - the Megatron model omits that norm, siglip_vit_model.py:141-142).  `cfg` has hidden_size,
    num_attention_heads, kv_channels, num_layers, patch_dim, layernorm_epsilon.  linear_qkv rows are
    Megatron's per-head interleave [head, (q, k, v), hn]."""
    H, hn, C = cfg.num_attention_heads, cfg.kv_channels, cfg.hidden_size
    x = F.conv2d(images, w[prefix + "conv1.weight"], w[prefix + "conv1.bias"], stride=cfg.patch_dim)
    x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
    x = x + w[prefix + "position_embeddings.weight"].to(x.dtype)[None]
    n, S, _ = x.shape
    for i in range(cfg.num_layers):
        p = f"{prefix}decoder.layers.{i}."
        y = F.layer_norm(x, (C,), w[p + "input_layernorm.weight"], w[p + "input_layernorm.bias"], cfg.layernorm_epsilon)
        qkv = F.linear(y, w[p + "self_attention.linear_qkv.weight"], w[p + "self_attention.linear_qkv.bias"])
        qkv = qkv.view(n, S, H, 3, hn)
        att, _ = O.attention(qkv[:, :, :, 0], qkv[:, :, :, 1], qkv[:, :, :, 2], causal=False, scale=hn ** -0.5)
        att = att.to(x.dtype).reshape(n, S, H * hn)
        x = x + F.linear(att, w[p + "self_attention.linear_proj.weight"], w[p + "self_attention.linear_proj.bias"])
        y = F.layer_norm(x, (C,), w[p + "pre_mlp_layernorm.weight"], w[p + "pre_mlp_layernorm.bias"], cfg.layernorm_epsilon)
        f = F.gelu(F.linear(y, w[p + "mlp.linear_fc1.weight"], w[p + "mlp.linear_fc1.bias"]), approximate="tanh")
        x = x + F.linear(f, w[p + "mlp.linear_fc2.weight"], w[p + "mlp.linear_fc2.bias"])
    return x
