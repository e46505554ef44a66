"""CPU tests of the host logic of the reference-facing operator surfaces (B1, B3, B4):
constructor / forward conventions, layout handling (sbhd, bhsd), mask-type duck typing, argument
guards and the patch-registry semantics - with the operator wrappers replaced by the CPU oracle
(tests/hostlogic.py).  The same surfaces run against the real kernels in tests/test_gpu_surfaces.py."""
import pytest
import torch

from long_vita_b200.megatron import stub
from oracle import ops as O
from tests.hostlogic import oracle_ops
from tests.util import randn_bf16, rel_fro, seeded


def _qkv_sbhd(s, b, np_, ng, hn, seed):
    g = seeded(seed)
    return randn_bf16((s, b, np_, hn), g), randn_bf16((s, b, ng, hn), g), randn_bf16((s, b, ng, hn), g)


def _ref_sbhd(q, k, v, causal, scale=None):
    ref, _ = O.attention(q.permute(1, 0, 2, 3), k.permute(1, 0, 2, 3), v.permute(1, 0, 2, 3), causal=causal, scale=scale)
    return ref.permute(1, 0, 2, 3).reshape(q.shape[0], q.shape[1], -1)          # [sq, b, np*hn]


def test_core_attention_slot_conventions():
    from long_vita_b200.megatron.core_attention import B200DotProductAttention

    cfg = stub.TransformerConfig(hidden_size=320, num_attention_heads=10, num_query_groups=2)
    attn = stub.build_module(stub.ModuleSpec(module=B200DotProductAttention), config=cfg, layer_number=1,
                             attn_mask_type=stub.AttnMaskType.causal, attention_type="self")
    assert attn.hidden_size_per_attention_head == 32 and abs(attn.softmax_scale - 32 ** -0.5) < 1e-12
    q, k, v = _qkv_sbhd(96, 2, 10, 2, 32, 1)
    with oracle_ops():
        out = attn(q, k, v, None, attn_mask_type=stub.AttnMaskType.causal, packed_seq_params=None)
        # the mask type given at call time wins over the constructor's (dot_product_attention.py:153)
        out_full = attn(q, k, v, None, attn_mask_type=stub.AttnMaskType.no_mask)
        out_default = attn(q, k, v, None)
    assert out.shape == (96, 2, 320)
    assert rel_fro(out, _ref_sbhd(q, k, v, True)) < 5e-3
    assert rel_fro(out_full, _ref_sbhd(q, k, v, False)) < 5e-3
    assert torch.equal(out_default, out)
    with pytest.raises(AssertionError):
        attn(q, k, v, None, packed_seq_params=object())
    with pytest.raises(ValueError):
        B200DotProductAttention(stub.TransformerConfig(attention_dropout=0.1), 1, stub.AttnMaskType.causal)


def test_core_attention_slot_training_branch_returns_gradients_in_megatron_layout():
    """With gradients enabled the slot takes the differentiable path (lv_attn_bwd behind ops.attention): dq / dk / dv
    must come back in Megatron's [s, b, heads, hn] layout through the permutes around the kernel call, GQA-summed."""
    from long_vita_b200.megatron.core_attention import B200DotProductAttention

    cfg = stub.TransformerConfig(hidden_size=256, num_attention_heads=8, num_query_groups=2)
    attn = B200DotProductAttention(cfg, 1, stub.AttnMaskType.causal)
    q, k, v = (t.requires_grad_(True) for t in _qkv_sbhd(80, 2, 8, 2, 32, 5))
    d_out = randn_bf16((80, 2, 256), seeded(6))
    with oracle_ops():
        out = attn(q, k, v, None)
        assert out.requires_grad and out.shape == (80, 2, 256)
        out.backward(d_out)
    qf, kf, vf = (t.detach().float().requires_grad_(True) for t in (q, k, v))
    ref, _ = O.attention(qf.permute(1, 0, 2, 3), kf.permute(1, 0, 2, 3), vf.permute(1, 0, 2, 3), causal=True)
    ref.permute(1, 0, 2, 3).reshape(80, 2, 256).backward(d_out.float())
    assert rel_fro(out, ref.permute(1, 0, 2, 3).reshape(80, 2, 256)) < 5e-3
    for got, want, name in ((q.grad, qf.grad, "dq"), (k.grad, kf.grad, "dk"), (v.grad, vf.grad, "dv")):
        assert got.shape == want.shape and rel_fro(got, want) < 1e-2, name
    with oracle_ops(), torch.no_grad():                 # inference: the plain forward, no graph
        assert not attn(q, k, v, None).requires_grad


def test_patch_registry_wrapper_semantics():
    """patch_utils.py:46-53: a function named *wrapper decorates the original attribute.  bf16 CUDA inputs
    take the fused path; anything else falls through to the original eager forward (here: the stub's, which
    raises) - the same split the reference's wrapper makes with `use_flash_attn`."""
    from long_vita_b200.megatron.core_attention import b200_dot_product_attention_forward_wrapper

    class Patched(stub.DotProductAttention):
        pass

    assert b200_dot_product_attention_forward_wrapper.__name__.endswith("wrapper")
    stub.apply_reference_style_patch(Patched, "forward", b200_dot_product_attention_forward_wrapper)
    mod = Patched(stub.TransformerConfig(hidden_size=160, num_attention_heads=5, num_query_groups=1), 1,
                  stub.AttnMaskType.causal)
    q, k, v = _qkv_sbhd(64, 1, 5, 1, 32, 3)
    with pytest.raises(RuntimeError, match="eager Megatron attention"):
        mod.forward(q, k, v, None, stub.AttnMaskType.causal, None)          # CPU tensors -> original forward
    with pytest.raises(AssertionError):
        mod.forward(q, k, v, None, stub.AttnMaskType.causal, object())


def test_hf_attention_interface_function_layout_and_guards():
    from long_vita_b200.hf import attention_interface as AI

    g = seeded(4)
    q, k, v = randn_bf16((2, 10, 80, 32), g), randn_bf16((2, 2, 80, 32), g), randn_bf16((2, 2, 80, 32), g)
    with oracle_ops():
        out, w = AI.b200_attention_forward(None, q, k, v, None, dropout=0.0, scaling=0.2, is_causal=True)
        mod = type("M", (), {"is_causal": True})()
        out2, _ = AI.b200_attention_forward(mod, q, k, v, None, scaling=0.2)           # is_causal from the module
        dec, _ = AI.b200_attention_forward(mod, q[:, :, -1:], k, v, None, scaling=0.2)  # 1 query row: not causal
    assert w is None and out.shape == (2, 80, 10, 32)                                  # [b, s, h, d]
    ref, _ = O.attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), causal=True, scale=0.2)
    assert rel_fro(out, ref) < 5e-3 and torch.equal(out, out2)
    assert rel_fro(dec[:, 0], ref[:, -1]) < 5e-3
    with pytest.raises(NotImplementedError):
        AI.b200_attention_forward(None, q, k, v, None, dropout=0.1)
    with pytest.raises(NotImplementedError):
        AI.b200_attention_forward(None, q, k, v, None, sliding_window=16)
    with pytest.raises(NotImplementedError):
        AI.b200_attention_forward(None, q, k, v, torch.tensor([[1] * 79 + [0], [1] * 80]))
    # 4-D masks (what transformers builds for sdpa / eager): pure causal passes, a padded key column is refused
    causal4 = torch.tril(torch.ones(80, 80, dtype=torch.bool))[None, None].expand(2, 1, 80, 80)
    additive = torch.zeros(2, 1, 80, 80).masked_fill(~causal4, float("-inf"))
    with oracle_ops():
        o_bool, _ = AI.b200_attention_forward(None, q, k, v, causal4, scaling=0.2, is_causal=True)
        o_add, _ = AI.b200_attention_forward(None, q, k, v, additive, scaling=0.2, is_causal=True)
    assert torch.equal(o_bool, out) and torch.equal(o_add, out)
    padded = causal4.clone()
    padded[1, :, :, 0] = False
    with pytest.raises(NotImplementedError, match="4-D attention mask"):
        AI.b200_attention_forward(None, q, k, v, padded, scaling=0.2, is_causal=True)
    with pytest.raises(NotImplementedError, match="4-D attention mask"):
        AI.b200_attention_forward(None, q, k, v, additive.masked_fill(~padded, float("-inf")), scaling=0.2, is_causal=True)
    from transformers import AttentionInterface

    name = AI.register()
    assert name == "b200_fa" and name in AttentionInterface._global_mapping


def test_intern_inner_attn_asserts_match_the_reference():
    """flash_attention.py:41-42: `assert qkv.dtype in [fp16, bf16]` and `assert qkv.is_cuda`."""
    from long_vita_b200.hf.attention_interface import B200FlashAttention

    m = B200FlashAttention(attention_dropout=0.0)
    with pytest.raises(AssertionError):
        m(torch.zeros(1, 8, 3, 2, 64))                       # fp32
    with pytest.raises(AssertionError):
        m(torch.zeros(1, 8, 3, 2, 64, dtype=torch.bfloat16))  # CPU tensor
    with pytest.raises(AssertionError):
        m(torch.zeros(1, 8, 3, 2, 64, dtype=torch.bfloat16), need_weights=True)
