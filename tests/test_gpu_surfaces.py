"""GPU tests of the reference-facing operator surfaces: Megatron `core_attention` slot and patch
wrapper (through the Megatron-free stub), HF AttentionInterface function, InternAttention.inner_attn
replacement."""
import math

import pytest
import torch

from oracle import ops as O
from tests.util import randn_bf16, rel_fro, seeded

pytestmark = pytest.mark.gpu


def _excess(out, ref):
    e, f = rel_fro(out, ref), rel_fro(ref.to(torch.bfloat16), ref)
    return math.sqrt(max(e * e - f * f, 0.0))


def _qkv_sbhd(s, b, np_, ng, hn, seed):
    g = seeded(seed)
    return randn_bf16((s, b, np_, hn), g), randn_bf16((s, b, ng, hn), g), randn_bf16((s, b, ng, hn), g)


def test_core_attention_slot_llm_and_vit(lib_built):
    from long_vita_b200.megatron import stub
    from long_vita_b200.megatron.core_attention import B200DotProductAttention

    cfg = stub.TransformerConfig(hidden_size=1280, num_attention_heads=10, num_query_groups=2)
    spec = stub.ModuleSpec(module=B200DotProductAttention)
    attn = stub.build_module(spec, config=cfg, layer_number=1, attn_mask_type=stub.AttnMaskType.causal,
                             attention_type="self")
    q, k, v = _qkv_sbhd(700, 2, 10, 2, 128, 1)
    out = attn(q.cuda(), k.cuda(), v.cuda(), None, attn_mask_type=stub.AttnMaskType.causal, packed_seq_params=None)
    assert out.shape == (700, 2, 1280)
    ref, _ = O.attention(q.permute(1, 0, 2, 3), k.permute(1, 0, 2, 3), v.permute(1, 0, 2, 3), causal=True)
    assert _excess(out.view(700, 2, 10, 128).permute(1, 0, 2, 3), ref) < 2e-3
    # ViT slot: no_mask, 16 heads x 64
    vcfg = stub.TransformerConfig(hidden_size=1024, num_attention_heads=16, num_query_groups=16)
    vattn = B200DotProductAttention(vcfg, 3, stub.AttnMaskType.no_mask, "self")
    q, k, v = _qkv_sbhd(1025, 2, 16, 16, 64, 2)
    out = vattn(q.cuda(), k.cuda(), v.cuda(), None)
    ref, _ = O.attention(q.permute(1, 0, 2, 3), k.permute(1, 0, 2, 3), v.permute(1, 0, 2, 3), causal=False)
    assert _excess(out.view(1025, 2, 16, 64).permute(1, 0, 2, 3), ref) < 2e-3
    with pytest.raises(AssertionError):
        vattn(q.cuda(), k.cuda(), v.cuda(), None, packed_seq_params=object())


def test_patch_registry_wrapper(lib_built):
    from long_vita_b200.megatron import stub
    from long_vita_b200.megatron.core_attention import b200_dot_product_attention_forward_wrapper

    class Patched(stub.DotProductAttention):
        pass

    stub.apply_reference_style_patch(Patched, "forward", b200_dot_product_attention_forward_wrapper)
    cfg = stub.TransformerConfig(hidden_size=640, num_attention_heads=5, num_query_groups=1)
    mod = Patched(cfg, 1, stub.AttnMaskType.causal)
    q, k, v = _qkv_sbhd(384, 1, 5, 1, 128, 3)
    out = mod.forward(q.cuda(), k.cuda(), v.cuda(), None, stub.AttnMaskType.causal, None)
    ref, _ = O.attention(q.permute(1, 0, 2, 3), k.permute(1, 0, 2, 3), v.permute(1, 0, 2, 3), causal=True)
    assert out.shape == (384, 1, 640)
    assert _excess(out.view(384, 1, 5, 128).permute(1, 0, 2, 3), ref) < 2e-3


def test_hf_attention_interface_function(lib_built):
    from long_vita_b200.hf import attention_interface as AI

    name = AI.register()
    from transformers import AttentionInterface

    assert name in AttentionInterface()._global_mapping or name in dict(AttentionInterface._global_mapping)
    g = seeded(4)
    q, k, v = randn_bf16((2, 10, 300, 128), g), randn_bf16((2, 2, 300, 128), g), randn_bf16((2, 2, 300, 128), g)
    out, w = AI.b200_attention_forward(None, q.cuda(), k.cuda(), v.cuda(), None, dropout=0.0, scaling=128 ** -0.5,
                                       is_causal=True)
    assert w is None and out.shape == (2, 300, 10, 128)
    ref, _ = O.attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), causal=True)
    assert _excess(out, ref) < 2e-3


def test_intern_inner_attn_replacement(lib_built):
    from long_vita_b200.hf.attention_interface import B200FlashAttention

    g = seeded(5)
    qkv = randn_bf16((3, 1025, 3, 16, 64), g)
    out, w = B200FlashAttention(attention_dropout=0.0)(qkv.cuda(), key_padding_mask=None, need_weights=False, causal=False)
    assert w is None and out.shape == (3, 1025, 16, 64)
    ref, _ = O.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], causal=False)
    assert _excess(out, ref) < 2e-3
    with pytest.raises(AssertionError):
        B200FlashAttention()(qkv.float().cuda())
