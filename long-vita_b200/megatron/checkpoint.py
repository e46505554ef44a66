"""Megatron-core <-> HF state-dict layouts of Long-VITA (host-side, one time at load).

The reference converts between the two with tools/hf2mcore_long_vita.py; the mcore -> HF direction
is `convert_checkpoint_from_megatron_to_transformers` (:373-510), which this module restates as pure
index permutations so a Megatron (TE-spec) checkpoint can feed the kernels directly:

* ViT `self_attention.linear_qkv` rows are per-head interleaved [head, (q, k, v), hn]; HF's
  `attn.qkv` rows are [(q, k, v), head, hn] (:397-414, :441, :450).
* LLM `self_attention.linear_qkv` rows are grouped [ng, (np/ng q heads, k, v), hn]
  (:488-498); HF has separate q/k/v projections.
* `mlp.linear_fc1.weight = cat(gate, up)` (:502-504).
* `vit.position_embeddings.weight` is [1025, C] (HF: [1, 1025, C], :417), `vit.class_token` is the
  HF `class_embedding`, `vit.conv1` the HF `patch_embedding` (:418-420).
* projector: `pre_proj_layernorm`, `vision_projection.encoder.linear_fc1/fc2` <-> HF
  `pre_proj_layernorm`, `mlp.0`, `mlp.2` (:466-474).

Everything here is bit-exact data movement (tests/test_megatron_model_host.py round-trips it).
"""
from __future__ import annotations

from typing import Dict

import torch

from ..config import LongVITAConfig

VIT = "external_feature_model.vit."
EFM = "external_feature_model."


def vit_qkv_index(num_heads: int, hn: int) -> torch.Tensor:
    """`indices` of hf2mcore_long_vita.py:397-414: hf_rows = mcore_rows[indices]."""
    head = torch.arange(num_heads).view(1, num_heads, 1)
    part = torch.arange(3).view(3, 1, 1)
    lane = torch.arange(hn).view(1, 1, hn)
    return (head * 3 * hn + part * hn + lane).reshape(-1)


def mcore_to_hf(sd: Dict[str, torch.Tensor], cfg: LongVITAConfig) -> Dict[str, torch.Tensor]:
    """Megatron-core (TE-spec names) state dict -> HF `LongVITAForCausalLM` names."""
    v = cfg.visual
    out: Dict[str, torch.Tensor] = {}
    # ---- vision tower --------------------------------------------------------------------------
    if VIT + "class_token" in sd:
        e = "model.vision_model.embeddings."
        out[e + "position_embedding"] = sd[VIT + "position_embeddings.weight"].unsqueeze(0)
        out[e + "class_embedding"] = sd[VIT + "class_token"].reshape(1, 1, -1)
        out[e + "patch_embedding.weight"] = sd[VIT + "conv1.weight"]
        out[e + "patch_embedding.bias"] = sd[VIT + "conv1.bias"]
        idx = vit_qkv_index(v.num_attention_heads, v.head_dim)
        for i in range(v.num_hidden_layers):
            m = f"{VIT}decoder.layers.{i}."
            if m + "ls1" not in sd:
                break
            h = f"model.vision_model.encoder.layers.{i}."
            out[h + "norm1.weight"] = sd[m + "input_layernorm.weight"]
            out[h + "norm1.bias"] = sd[m + "input_layernorm.bias"]
            out[h + "norm2.weight"] = sd[m + "pre_mlp_layernorm.weight"]
            out[h + "norm2.bias"] = sd[m + "pre_mlp_layernorm.bias"]
            out[h + "ls1"] = sd[m + "ls1"]
            out[h + "ls2"] = sd[m + "ls2"]
            dev = sd[m + "self_attention.linear_qkv.weight"].device
            out[h + "attn.qkv.weight"] = sd[m + "self_attention.linear_qkv.weight"][idx.to(dev)]
            out[h + "attn.qkv.bias"] = sd[m + "self_attention.linear_qkv.bias"][idx.to(dev)]
            out[h + "attn.proj.weight"] = sd[m + "self_attention.linear_proj.weight"]
            out[h + "attn.proj.bias"] = sd[m + "self_attention.linear_proj.bias"]
            for a, b in (("linear_fc1", "fc1"), ("linear_fc2", "fc2")):
                out[h + f"mlp.{b}.weight"] = sd[m + f"mlp.{a}.weight"]
                out[h + f"mlp.{b}.bias"] = sd[m + f"mlp.{a}.bias"]
        p = "model.vision_projection."
        out[p + "pre_proj_layernorm.weight"] = sd[EFM + "pre_proj_layernorm.weight"]
        out[p + "pre_proj_layernorm.bias"] = sd[EFM + "pre_proj_layernorm.bias"]
        out[p + "mlp.0.weight"] = sd[EFM + "vision_projection.encoder.linear_fc1.weight"]
        out[p + "mlp.2.weight"] = sd[EFM + "vision_projection.encoder.linear_fc2.weight"]
    # ---- language model ------------------------------------------------------------------------
    ng, np_, hn, H = cfg.num_key_value_heads, cfg.num_attention_heads, cfg.head_dim, cfg.hidden_size
    g = np_ // ng
    if "embedding.word_embeddings.weight" in sd:
        out["model.embed_tokens.weight"] = sd["embedding.word_embeddings.weight"]
    for i in range(cfg.num_hidden_layers):
        m = f"decoder.layers.{i}."
        if m + "self_attention.linear_qkv.weight" not in sd:
            continue
        h = f"model.layers.{i}."
        out[h + "input_layernorm.weight"] = sd[m + "self_attention.linear_qkv.layer_norm_weight"]
        w = sd[m + "self_attention.linear_qkv.weight"].view(ng, g + 2, hn, H)
        out[h + "self_attn.q_proj.weight"] = w[:, :g].reshape(np_ * hn, H)
        out[h + "self_attn.k_proj.weight"] = w[:, g].reshape(ng * hn, H)
        out[h + "self_attn.v_proj.weight"] = w[:, g + 1].reshape(ng * hn, H)
        b = sd[m + "self_attention.linear_qkv.bias"].view(ng, g + 2, hn)
        out[h + "self_attn.q_proj.bias"] = b[:, :g].reshape(-1)
        out[h + "self_attn.k_proj.bias"] = b[:, g].reshape(-1)
        out[h + "self_attn.v_proj.bias"] = b[:, g + 1].reshape(-1)
        out[h + "self_attn.o_proj.weight"] = sd[m + "self_attention.linear_proj.weight"]
        fc1 = sd[m + "mlp.linear_fc1.weight"]
        out[h + "mlp.gate_proj.weight"] = fc1[: cfg.intermediate_size]
        out[h + "mlp.up_proj.weight"] = fc1[cfg.intermediate_size :]
        out[h + "mlp.down_proj.weight"] = sd[m + "mlp.linear_fc2.weight"]
        out[h + "post_attention_layernorm.weight"] = sd[m + "mlp.linear_fc1.layer_norm_weight"]
    if "decoder.final_layernorm.weight" in sd:
        out["model.norm.weight"] = sd["decoder.final_layernorm.weight"]
    if "output_layer.weight" in sd:
        out["lm_head.weight"] = sd["output_layer.weight"]
    elif "embedding.word_embeddings.weight" in sd:      # tied head (share_embeddings_and_output_weights)
        out["lm_head.weight"] = sd["embedding.word_embeddings.weight"]
    return out


def hf_to_mcore(sd: Dict[str, torch.Tensor], cfg: LongVITAConfig) -> Dict[str, torch.Tensor]:
    """Algebraic inverse of `mcore_to_hf`.  (The LLM half equals the script's own HF -> mcore
    function, :590-613; its vision half there targets an older Qwen2-VL-style tower and is stale, so
    the ViT / projector half is inverted from the live mcore -> HF direction instead.)"""
    v = cfg.visual
    out: Dict[str, torch.Tensor] = {}
    e = "model.vision_model.embeddings."
    if e + "class_embedding" in sd:
        out[VIT + "position_embeddings.weight"] = sd[e + "position_embedding"].squeeze(0)
        out[VIT + "class_token"] = sd[e + "class_embedding"]
        out[VIT + "conv1.weight"] = sd[e + "patch_embedding.weight"]
        out[VIT + "conv1.bias"] = sd[e + "patch_embedding.bias"]
        inv = torch.argsort(vit_qkv_index(v.num_attention_heads, v.head_dim))
        for i in range(v.num_hidden_layers):
            h = f"model.vision_model.encoder.layers.{i}."
            if h + "ls1" not in sd:
                break
            m = f"{VIT}decoder.layers.{i}."
            out[m + "input_layernorm.weight"] = sd[h + "norm1.weight"]
            out[m + "input_layernorm.bias"] = sd[h + "norm1.bias"]
            out[m + "pre_mlp_layernorm.weight"] = sd[h + "norm2.weight"]
            out[m + "pre_mlp_layernorm.bias"] = sd[h + "norm2.bias"]
            out[m + "ls1"] = sd[h + "ls1"]
            out[m + "ls2"] = sd[h + "ls2"]
            dev = sd[h + "attn.qkv.weight"].device
            out[m + "self_attention.linear_qkv.weight"] = sd[h + "attn.qkv.weight"][inv.to(dev)]
            out[m + "self_attention.linear_qkv.bias"] = sd[h + "attn.qkv.bias"][inv.to(dev)]
            out[m + "self_attention.linear_proj.weight"] = sd[h + "attn.proj.weight"]
            out[m + "self_attention.linear_proj.bias"] = sd[h + "attn.proj.bias"]
            for a, b in (("linear_fc1", "fc1"), ("linear_fc2", "fc2")):
                out[m + f"mlp.{a}.weight"] = sd[h + f"mlp.{b}.weight"]
                out[m + f"mlp.{a}.bias"] = sd[h + f"mlp.{b}.bias"]
        p = "model.vision_projection."
        out[EFM + "pre_proj_layernorm.weight"] = sd[p + "pre_proj_layernorm.weight"]
        out[EFM + "pre_proj_layernorm.bias"] = sd[p + "pre_proj_layernorm.bias"]
        out[EFM + "vision_projection.encoder.linear_fc1.weight"] = sd[p + "mlp.0.weight"]
        out[EFM + "vision_projection.encoder.linear_fc2.weight"] = sd[p + "mlp.2.weight"]
    ng, np_, hn, H = cfg.num_key_value_heads, cfg.num_attention_heads, cfg.head_dim, cfg.hidden_size
    g = np_ // ng
    if "model.embed_tokens.weight" in sd:
        out["embedding.word_embeddings.weight"] = sd["model.embed_tokens.weight"]
    for i in range(cfg.num_hidden_layers):
        h = f"model.layers.{i}."
        if h + "self_attn.q_proj.weight" not in sd:
            continue
        m = f"decoder.layers.{i}."
        out[m + "self_attention.linear_qkv.layer_norm_weight"] = sd[h + "input_layernorm.weight"]
        q = sd[h + "self_attn.q_proj.weight"].view(ng, g, hn, H)
        k = sd[h + "self_attn.k_proj.weight"].view(ng, 1, hn, H)
        vv = sd[h + "self_attn.v_proj.weight"].view(ng, 1, hn, H)
        out[m + "self_attention.linear_qkv.weight"] = torch.cat([q, k, vv], dim=1).reshape((np_ + 2 * ng) * hn, H)
        qb = sd[h + "self_attn.q_proj.bias"].view(ng, g, hn)
        kb = sd[h + "self_attn.k_proj.bias"].view(ng, 1, hn)
        vb = sd[h + "self_attn.v_proj.bias"].view(ng, 1, hn)
        out[m + "self_attention.linear_qkv.bias"] = torch.cat([qb, kb, vb], dim=1).reshape(-1)
        out[m + "self_attention.linear_proj.weight"] = sd[h + "self_attn.o_proj.weight"]
        out[m + "mlp.linear_fc1.weight"] = torch.cat([sd[h + "mlp.gate_proj.weight"], sd[h + "mlp.up_proj.weight"]], dim=0)
        out[m + "mlp.linear_fc2.weight"] = sd[h + "mlp.down_proj.weight"]
        out[m + "mlp.linear_fc1.layer_norm_weight"] = sd[h + "post_attention_layernorm.weight"]
    if "model.norm.weight" in sd:
        out["decoder.final_layernorm.weight"] = sd["model.norm.weight"]
    if "lm_head.weight" in sd:
        out["output_layer.weight"] = sd["lm_head.weight"]
    return out
