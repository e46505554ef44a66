"""CPU tests of the host logic of the HF surface: `LongVITAForCausalLM.forward`
argument handling - image scatter, `num_logits_to_keep`, `inputs_embeds`, explicit `position_ids`,
`labels` -> shifted loss, `output_hidden_states`, tuple return, guards - with the operator wrappers
replaced by the CPU oracle (tests/hostlogic.py).  Kernel parity of the same model: tests/test_gpu_model.py."""
import pytest
import torch

from long_vita_b200.config import LongVITAConfig
from long_vita_b200.hf.modeling import LongVITAForCausalLM
from long_vita_b200.weights import synthetic_state_dict
from oracle import model as OM
from tests.hostlogic import oracle_ops
from tests.util import rel_fro


@pytest.fixture(scope="module")
def setup():
    cfg = LongVITAConfig.tiny(layers=2, vit_layers=1)
    w = synthetic_state_dict(cfg, seed=31, dtype=torch.bfloat16, perturb=True)
    return cfg, w, LongVITAForCausalLM(cfg, w), OM.cast_weights(w, torch.float32)


def _inputs(cfg, s=290, seed=2):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, cfg.vocab_size, (1, s), generator=g)
    images = torch.randn(1, 3, 448, 448, generator=g).to(torch.bfloat16)
    idx_s = torch.arange(11, 11 + 256).unsqueeze(0)
    return ids, images, torch.stack([torch.zeros_like(idx_s), idx_s])


def test_forward_with_images_keeps_last_rows_and_hidden_states(setup):
    cfg, w, model, w32 = setup
    ids, images, idx = _inputs(cfg)
    with oracle_ops():
        out = model(input_ids=ids, images=images, image_indices=idx, num_logits_to_keep=3, output_hidden_states=True)
        tup = model(input_ids=ids, images=images, image_indices=idx, num_logits_to_keep=3, return_dict=False)
    logits, hidden, h_final = OM.long_vita_forward(cfg, w32, ids, images.float(), idx, num_logits_to_keep=3, return_hidden=True)
    assert out.logits.shape == (1, 3, cfg.vocab_size) and out.loss is None and out.past_key_values is None
    assert rel_fro(out.logits[0], logits[0]) < 1.5e-2
    assert len(out.hidden_states) == cfg.num_hidden_layers + 1
    for hg, hr in zip(out.hidden_states[:-1], hidden):
        assert rel_fro(hg[0], hr) < 1e-2
    assert rel_fro(out.hidden_states[-1][0], h_final) < 1e-2           # the last entry is the normed state
    assert isinstance(tup, tuple) and torch.equal(tup[0], out.logits)


def test_inputs_embeds_position_ids_and_labels(setup):
    cfg, w, model, w32 = setup
    ids, _, _ = _inputs(cfg, s=96)
    labels = ids.clone()
    labels[0, :10] = -100
    with oracle_ops():
        a = model(input_ids=ids, labels=labels)
        emb = w["model.embed_tokens.weight"][ids[0]].unsqueeze(0)
        b = model(inputs_embeds=emb)
        shifted = model(input_ids=ids, position_ids=(torch.arange(96) + 5).unsqueeze(0))
    assert torch.equal(a.logits, b.logits)
    ref = OM.long_vita_forward(cfg, w32, ids)
    assert rel_fro(a.logits[0], ref[0]) < 1.5e-2
    want = torch.nn.functional.cross_entropy(a.logits[0, :-1].float(), labels[0, 1:], ignore_index=-100)
    assert torch.allclose(a.loss, want)
    ref_shift = OM.long_vita_forward(cfg, w32, ids, position_ids=(torch.arange(96) + 5).unsqueeze(0))
    assert rel_fro(shifted.logits[0], ref_shift[0]) < 1.5e-2
    assert not torch.equal(shifted.logits, a.logits)                   # RoPE saw the explicit positions


def test_guards_raise_before_any_kernel_call(setup):
    cfg, w, model, _ = setup
    ids = torch.zeros(1, 8, dtype=torch.long)
    with pytest.raises(ValueError):
        model(input_ids=None)
    with pytest.raises(ValueError):
        model(input_ids=ids, inputs_embeds=torch.zeros(1, 8, cfg.hidden_size))
    for kw in ({"output_attentions": True}, {"past_key_values": [1]},
               {"attention_mask": torch.tensor([[1, 1, 1, 1, 1, 1, 0, 0]])}):
        with pytest.raises(NotImplementedError):
            model(input_ids=ids, **kw)
    with pytest.raises(NotImplementedError):
        model(input_ids=torch.zeros(2, 8, dtype=torch.long))


def test_product_composition_against_the_references_own_forward():
    """The product's HF-surface model (host logic; kernels replaced by the oracle) against the committed outputs
    of the reference's own LongVITAForCausalLM.forward (tests/golden/ref_long_vita_tiny.pt)."""
    import os
    import sys

    gold_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, gold_dir)
    from make_golden import long_vita_inputs

    gold = torch.load(os.path.join(gold_dir, "ref_long_vita_tiny.pt"))
    cfg = LongVITAConfig.tiny(layers=2, vit_layers=1)
    w = synthetic_state_dict(cfg, seed=gold["seed"], dtype=torch.float32, perturb=True)
    ids, images, idx = long_vita_inputs(cfg, gold["seed"])
    model = LongVITAForCausalLM(cfg, {k: v.to(torch.bfloat16) for k, v in w.items()})
    with oracle_ops():
        out = model(input_ids=ids, images=images.to(torch.bfloat16), image_indices=idx, num_logits_to_keep=8,
                    output_hidden_states=True)
    # the fixture is fp32 with fp32 weights; bf16 weights + activations alone give ~6.6e-3
    assert rel_fro(out.logits[0], gold["logits_last8"]) < 2e-2
    assert torch.equal(out.logits[0].float().argmax(-1), gold["logits_last8"].argmax(-1))
    rows = gold["rows"]
    for li in range(cfg.num_hidden_layers + 1):
        assert rel_fro(out.hidden_states[li][0][rows], gold["hidden_rows"][li]) < 1.5e-2, li
    # labels -> loss: the reference's ForCausalLMLoss (shift by one, mean over the labels that are not -100)
    from make_golden import golden_labels

    with oracle_ops():
        loss = model(input_ids=ids, images=images.to(torch.bfloat16), image_indices=idx, labels=golden_labels(ids)).loss
    assert abs(float(loss) - float(gold["loss"])) < 2e-2 * float(gold["loss"]), (float(loss), float(gold["loss"]))




def test_kv_cache_decode_equals_full_forward(setup):
    """8f-2: prefill with use_cache, then single-token steps and a 3-token chunk through the cache must give
    the logits of a full forward over the extended sequence (same operators per token; only the attention's
    split / merge order differs)."""
    cfg, w, model, w32 = setup
    ids, images, idx = _inputs(cfg, s=290)
    extra = torch.randint(0, cfg.vocab_size, (1, 5), generator=torch.Generator().manual_seed(77))
    with oracle_ops():
        full = model(input_ids=torch.cat([ids, extra], dim=1), images=images, image_indices=idx).logits   # [1, 295, V]
        out = model(input_ids=ids, images=images, image_indices=idx, use_cache=True, num_logits_to_keep=1, max_cache_len=400)
        cache = out.past_key_values
        assert len(cache) == 290 and cache.get_seq_length() == 290 and cache.capacity == 400
        assert rel_fro(out.logits[0, -1], full[0, 289]) < 4e-3
        for i in range(2):                                     # one token at a time -> ops.attention_decode
            out = model(input_ids=extra[:, i : i + 1], past_key_values=cache, use_cache=True, images=images, image_indices=idx)
            assert out.past_key_values is cache and len(cache) == 291 + i
            # the split partial outputs are bf16 (one extra rounding before the merge; with n_splits = 1 the
            # step is bit-identical to the full forward): 6e-3 on the logits of this 2-layer model
            assert rel_fro(out.logits[0, 0], full[0, 290 + i]) < 1e-2, i
        out = model(input_ids=extra[:, 2:], past_key_values=cache, use_cache=True)      # 3 tokens: chunked prefill
        assert len(cache) == 295 and out.logits.shape == (1, 3, cfg.vocab_size)
        assert rel_fro(out.logits[0], full[0, 292:]) < 1e-2
        with pytest.raises(RuntimeError, match="overflow"):
            model(input_ids=torch.zeros(1, 200, dtype=torch.long), past_key_values=cache, use_cache=True)


def test_attention_decode_splits_and_merges(setup):
    """Flash-decoding composition: GQA heads packed as query rows, key range split into batch entries plus a
    ragged remainder, log-sum-exp merge - against one plain attention over the same keys."""
    from oracle import ops as O

    g = torch.Generator().manual_seed(3)
    hq, hkv, d, L = 10, 2, 32, 700
    q = torch.randn(hq, d, generator=g).to(torch.bfloat16)
    kc = torch.randn(1024, hkv, d, generator=g).to(torch.bfloat16)
    vc = torch.randn(1024, hkv, d, generator=g).to(torch.bfloat16)
    ref, ref_lse = O.attention(q[None, None], kc[None, :L], vc[None, :L], causal=False)
    with oracle_ops() as ops:
        for n_splits in (1, 3, 18):
            out, lse = ops.attention_decode(q, kc, vc, L, n_splits=n_splits, return_lse=True)
            assert out.shape == (hq, d) and lse.shape == (hq,)
            assert rel_fro(out, ref[0, 0]) < 6e-3, n_splits
            assert torch.allclose(lse, ref_lse[0, :, 0], atol=1e-4), n_splits
        with pytest.raises(ValueError):
            ops.attention_decode(q, kc, vc, 2000)


def test_generate_greedy_matches_re_prefill(setup):
    cfg, w, model, _ = setup
    ids, images, idx = _inputs(cfg, s=290)
    with oracle_ops():
        gen = model.generate_greedy(ids, images, idx, max_new_tokens=3)
        # the reference's loop: feed everything again, take the last row
        seq = ids
        want = []
        for _ in range(3):
            tok = model(input_ids=seq, images=images, image_indices=idx, num_logits_to_keep=1).logits[0, -1].float().argmax()
            want.append(int(tok))
            seq = torch.cat([seq, tok.view(1, 1)], dim=1)
    assert gen.shape == (1, 3) and gen[0].tolist() == want


def test_kv_cache_decode_against_the_references_own_incremental_decoding():
    """The build's decode path (prefill with use_cache, then single-token steps through ops.attention_decode; kernels
    replaced by the oracle) against the committed logits of the reference's own DynamicCache decoding
    (tests/golden/ref_long_vita_decode.pt)."""
    import os
    import sys

    gold_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, gold_dir)
    from make_golden import decode_new_tokens, long_vita_inputs

    dec = torch.load(os.path.join(gold_dir, "ref_long_vita_decode.pt"))
    cfg = LongVITAConfig.tiny(layers=2, vit_layers=1)
    w = synthetic_state_dict(cfg, seed=dec["seed"], dtype=torch.float32, perturb=True)
    ids, images, idx = long_vita_inputs(cfg, dec["seed"])
    new = decode_new_tokens(cfg)
    model = LongVITAForCausalLM(cfg, {k: v.to(torch.bfloat16) for k, v in w.items()})
    with oracle_ops():
        out = model(input_ids=ids, images=images.to(torch.bfloat16), image_indices=idx, use_cache=True, num_logits_to_keep=1,
                    max_cache_len=ids.shape[1] + 8)
        assert rel_fro(out.logits[0, -1], dec["prefill_last"]) < 2e-2
        cache = out.past_key_values
        for i in range(new.shape[1]):
            o = model(input_ids=new[:, i : i + 1], past_key_values=cache, use_cache=True, images=images.to(torch.bfloat16),
                      image_indices=idx)
            assert rel_fro(o.logits[0, 0], dec["steps"][i]) < 2e-2, (i, rel_fro(o.logits[0, 0], dec["steps"][i]))
            assert int(o.logits[0, 0].float().argmax()) == int(dec["steps"][i].argmax())
        assert len(cache) == dec["cache_len"]


# ---- the surface tools/inference_long_vita.py:811-868 uses: nn.Module, from_pretrained, generation_config, generate ----
def test_model_is_a_module_with_the_reference_state_dict(setup):
    cfg, w, model, _ = setup
    assert isinstance(model, torch.nn.Module) and isinstance(model.model, torch.nn.Module)
    assert model.eval() is model and model.dtype == torch.bfloat16 and model.device.type == "cpu"
    sd = model.state_dict()
    assert set(sd) == set(w), (set(sd) ^ set(w))
    for k_, t in w.items():                       # fused / interleaved / padded buffers un-fuse to the HF tensors bit for bit
        assert sd[k_].shape == t.shape and torch.equal(sd[k_], t), k_


def test_module_movement_reaches_the_plain_weight_tensors(setup):
    """`.to()` / `.cuda()` go through nn.Module._apply; the weights are plain (fused) tensors, not Parameters, so the
    holder classes forward the conversion to every tensor they keep - and nothing is shared with the original."""
    import copy

    cfg, w, model, _ = setup
    m2 = copy.deepcopy(model).to(torch.float64)
    assert m2.dtype == torch.float64 and m2.model.embed_tokens.dtype == torch.float64
    assert m2.model.layers[0].wqkv.dtype == torch.float64 and m2.model.inv_freq.dtype == torch.float64
    sd2 = m2.state_dict()
    assert set(sd2) == set(w)
    for k_, t in w.items():
        assert sd2[k_].dtype == torch.float64 and torch.equal(sd2[k_], t.double()), k_
    assert model.dtype == torch.bfloat16 and model.model.layers[0].wqkv.dtype == torch.bfloat16      # the original is untouched
    assert list(model.parameters()) == []          # nothing for an optimizer to pick up by accident


def test_from_pretrained_reads_an_hf_checkpoint_directory(setup, tmp_path):
    import dataclasses
    import json

    from safetensors.torch import save_file

    cfg, w, model, _ = setup
    c = {f.name: getattr(cfg, f.name) for f in dataclasses.fields(cfg) if f.name != "visual"}
    c["visual"] = dataclasses.asdict(cfg.visual)
    c["architectures"] = ["LongVITAForCausalLM"]
    (tmp_path / "config.json").write_text(json.dumps(c))
    names = sorted(w)
    save_file({k_: w[k_].contiguous() for k_ in names[::2]}, str(tmp_path / "model-00001-of-00002.safetensors"))
    save_file({k_: w[k_].contiguous() for k_ in names[1::2]}, str(tmp_path / "model-00002-of-00002.safetensors"))
    m2 = LongVITAForCausalLM.from_pretrained(str(tmp_path), trust_remote_code=True, device_map="cpu", torch_dtype=torch.bfloat16,
                                             attn_implementation="flash_attention_2")
    assert m2.config == cfg
    sd = m2.state_dict()
    assert all(torch.equal(sd[k_], w[k_]) for k_ in w)
    with pytest.raises(NotImplementedError):
        LongVITAForCausalLM.from_pretrained(str(tmp_path), torch_dtype=torch.float16, device_map="cpu")


def test_generate_follows_the_inference_script_contract(setup):
    cfg, w, model, w32 = setup
    ids, images, idx = _inputs(cfg, s=290)
    model.generation_config.max_new_tokens = 4          # the script mutates model.generation_config in place (:821-826)
    model.generation_config.do_sample = False
    with oracle_ops():
        out = model.generate(inputs=ids, images=images, image_indices=idx)
        first = model(input_ids=ids, images=images, image_indices=idx, num_logits_to_keep=1).logits[0, -1].float().argmax()
        stop = model.generate(inputs=ids, images=images, image_indices=idx, eos_token_id=[int(out[0, 290]), 7])
    assert out.shape == (1, 294) and torch.equal(out[:, :290], ids)          # prompt + new tokens, like GenerationMixin
    assert int(out[0, 290]) == int(first)
    assert stop.shape == (1, 291)                                            # stops AT the eos token (kept, as in HF)
    # greedy continuation equals re-running the full forward on the grown sequence (cache == no cache)
    seq = ids
    with oracle_ops():
        for i in range(3):
            nxt = model(input_ids=seq, images=images, image_indices=idx, num_logits_to_keep=1).logits[0, -1].float().argmax()
            assert int(nxt) == int(out[0, 290 + i])
            seq = torch.cat([seq, nxt.view(1, 1)], dim=1)
    with pytest.raises(NotImplementedError):
        model.generate(inputs=ids, do_sample=True)
    model.generation_config.max_new_tokens = 1024


def test_siglip_tower_composition_head_dim_72_padding_and_interleaved_qkv():
    """a9 host logic (the GPU twin is tests/test_gpu_model.py::test_siglip_tower_matches_oracle): Megatron's per-head
    interleaved linear_qkv rows, the zero padding of head_dim 72 to the kernel's 128 with scale 72^-0.5, no class token,
    tanh-GELU - composed from the operator wrappers, against oracle.model.siglip_forward (pinned to transformers'
    SiglipVisionModel in tests/test_oracle_pinning.py)."""
    from long_vita_b200.hf.siglip import SigLIPConfig, SigLIPViTModel

    cfg = SigLIPConfig(hidden_size=144, ffn_hidden_size=304, num_layers=2, num_attention_heads=2, kv_channels=72, image_size=112)
    g = torch.Generator().manual_seed(21)

    def rn(*shape, std=0.05):
        return (torch.randn(*shape, generator=g) * std).to(torch.bfloat16)

    C, I = cfg.hidden_size, cfg.ffn_hidden_size
    w = {"conv1.weight": rn(C, 3, 14, 14), "conv1.bias": rn(C), "position_embeddings.weight": rn(cfg.num_patches, C, std=1.0)}
    for i in range(cfg.num_layers):
        p = f"decoder.layers.{i}."
        w.update({
            p + "input_layernorm.weight": (1 + rn(C)).to(torch.bfloat16), p + "input_layernorm.bias": rn(C),
            p + "self_attention.linear_qkv.weight": rn(3 * C, C), p + "self_attention.linear_qkv.bias": rn(3 * C),
            p + "self_attention.linear_proj.weight": rn(C, C), p + "self_attention.linear_proj.bias": rn(C),
            p + "pre_mlp_layernorm.weight": (1 + rn(C)).to(torch.bfloat16), p + "pre_mlp_layernorm.bias": rn(C),
            p + "mlp.linear_fc1.weight": rn(I, C), p + "mlp.linear_fc1.bias": rn(I),
            p + "mlp.linear_fc2.weight": rn(C, I), p + "mlp.linear_fc2.bias": rn(C),
        })
    images = torch.randn(2, 3, 112, 112, generator=g).to(torch.bfloat16)
    with oracle_ops():
        out = SigLIPViTModel(cfg, w)(images)
    ref = OM.siglip_forward(cfg, OM.cast_weights(w, torch.float32), images.float())
    assert out.shape == (2, 64, C)
    assert rel_fro(out, ref) < 6e-3, rel_fro(out, ref)
