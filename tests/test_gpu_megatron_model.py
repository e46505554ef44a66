"""GPU parity of the Megatron composition and of the fused training tail (8f-3):
`B200GPTVLModel.forward` (gpt_vl_model.py:233-416 mirrored: external_inputs modes, logit_mask, labels -> per-token
loss, inference_params overrides) over a Megatron-core state dict, on the device, against the CPU oracle of the same
model; the chunked LM-head + cross-entropy kernels against fp32 cross_entropy and its autograd."""
import types

import pytest
import torch

from long_vita_b200.config import LongVITAConfig
from long_vita_b200.megatron import checkpoint as ck
from long_vita_b200.weights import synthetic_state_dict
from oracle import model as OM
from tests.util import rel_fro

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(lib_built):
    from long_vita_b200.megatron.gpt_vl_model import B200GPTVLModel

    cfg = LongVITAConfig.tiny(layers=2, vit_layers=1)
    hf = synthetic_state_dict(cfg, seed=77, dtype=torch.bfloat16, perturb=True)
    mc = {k: v.cuda() for k, v in ck.hf_to_mcore(hf, cfg).items()}
    return cfg, hf, B200GPTVLModel(cfg, mc)


def _inputs(cfg, s=300, seed=5):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, cfg.vocab_size, (1, s), generator=g)
    images = torch.randn(1, 3, 448, 448, generator=g).to(torch.bfloat16)
    idx_s = torch.arange(7, 7 + 256).unsqueeze(0)
    return ids, images, torch.stack([torch.zeros_like(idx_s), idx_s])


def _oracle(cfg, hf, ids, images, idx):
    return OM.long_vita_forward(cfg, OM.cast_weights(hf, torch.float32), ids, images.float(), idx)[0]


def test_gptvl_forward_on_the_device(setup):
    cfg, hf, model = setup
    ids, images, idx = _inputs(cfg)
    s = ids.shape[1]
    pos = torch.arange(s).unsqueeze(0).cuda()
    ref = _oracle(cfg, hf, ids, images, idx)
    ext = {"images": images.cuda(), "indices": idx.cuda()}
    a = model(ids.cuda(), pos, None, external_inputs=ext)
    assert a.shape == (1, s, cfg.vocab_size) and rel_fro(a[0], ref) < 1.5e-2
    # random-init tiny model: near-ties in the logits flip a few argmax positions under bf16 (measured 0.967 agreement)
    assert float((a[0].float().argmax(-1).cpu() == ref.argmax(-1)).float().mean()) > 0.93
    # the three embedding-merge modes agree bit for bit (language_model_embedding.py:102-134)
    b = model(ids.cuda(), pos, None, external_inputs={"images": images.cuda(), "pre_len": 7})
    src = (torch.zeros(256, dtype=torch.long).cuda(), torch.arange(256).cuda())
    tgt = (torch.zeros(256, dtype=torch.long).cuda(), torch.arange(7, 7 + 256).cuda())
    c = model(ids.cuda(), pos, None, external_inputs={"images": images.cuda(), "src_indices": src, "tgt_indices": tgt})
    assert torch.equal(a, b) and torch.equal(a, c)
    # logit mask through inference_params (generation.py:141-165) and labels -> per-token loss, fused and un-fused
    mask = torch.zeros(1, s, dtype=torch.bool)
    mask[0, 280:] = True
    ip = types.SimpleNamespace(external_inputs=ext, key_value_memory_dict={}, logit_mask=mask.cuda(), use_kv_cache=False)
    lg = model(ids.cuda(), pos, None, inference_params=ip)
    assert lg.shape == (1, 20, cfg.vocab_size) and torch.equal(lg[0], a[0, 280:])
    labels = torch.randint(0, cfg.vocab_size, (1, s), generator=torch.Generator().manual_seed(9))
    loss = model(ids.cuda(), pos, None, labels=labels.cuda(), external_inputs=ext, logit_mask=mask.cuda())
    want = torch.nn.functional.cross_entropy(lg[0].float(), labels[0, 280:].cuda(), reduction="none")
    assert loss.shape == (1, 20) and loss.dtype == torch.float32
    assert torch.allclose(loss[0], want, rtol=1e-5, atol=1e-4), float((loss[0] - want).abs().max())
    model.fused_loss = False
    loss_u = model(ids.cuda(), pos, None, labels=labels.cuda(), external_inputs=ext, logit_mask=mask.cuda())
    model.fused_loss = True
    assert torch.equal(loss_u[0], want)
    ref_loss = torch.nn.functional.cross_entropy(ref[280:], labels[0, 280:], reduction="none")
    assert rel_fro(loss[0], ref_loss) < 2e-2


@pytest.mark.parametrize("M,V,chunk", [(37, 2048, 512), (300, 152064, 16384), (1, 1000, 16384)])
def test_fused_lm_head_cross_entropy(lib_built, M, V, chunk):
    """8f-3: loss and gradients of the chunked LM head + cross-entropy vs fp32 cross_entropy over materialised logits."""
    from long_vita_b200 import ops

    g = torch.Generator().manual_seed(M + V)
    s, c = M + 11, 640
    h = torch.randn(s, 1, c, generator=g).to(torch.bfloat16)
    w = (torch.randn(V, c, generator=g) * 0.05).to(torch.bfloat16)
    mask = torch.zeros(1, s, dtype=torch.bool)
    mask[0, torch.randperm(s, generator=g)[:M]] = True
    labels = torch.randint(0, V, (1, M), generator=g)
    up = torch.randn(1, M, generator=g)
    hd, wd = h.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    loss = ops.masked_lm_head_ce(hd, wd, mask.cuda(), labels.cuda(), vocab_chunk=chunk)
    loss.backward(up.cuda())
    hf, wf = h.float().requires_grad_(True), w.float().requires_grad_(True)
    sel = hf[mask[0]][:, 0]
    logits_bf = (sel @ wf.t()).to(torch.bfloat16).float()           # the reference's dtypes: bf16 logits, fp32 loss
    ref = torch.nn.functional.cross_entropy(logits_bf, labels[0], reduction="none")
    assert torch.allclose(loss[0].cpu(), ref, rtol=2e-3, atol=2e-3), float((loss[0].cpu() - ref).abs().max())
    torch.nn.functional.cross_entropy(sel @ wf.t(), labels[0], reduction="none").backward(up[0])
    assert rel_fro(hd.grad, hf.grad) < 1e-2, rel_fro(hd.grad, hf.grad)
    assert rel_fro(wd.grad, wf.grad) < 1e-2, rel_fro(wd.grad, wf.grad)
    assert torch.equal(hd.grad.cpu()[~mask[0]], torch.zeros(s - M, 1, c, dtype=torch.bfloat16))
    # against the un-fused product path (masked GEMM -> logits -> torch CE): same bf16 logits
    lg = ops.masked_linear(hd.detach(), wd.detach(), mask.cuda())
    want = torch.nn.functional.cross_entropy(lg[:, 0].float(), labels[0].cuda(), reduction="none")
    assert torch.allclose(loss[0].detach(), want, rtol=1e-5, atol=1e-4)


def test_hf_generate_on_the_device(lib_built):
    """tools/inference_long_vita.py:868: model.generate(inputs=, images=, image_indices=) with host inputs."""
    from long_vita_b200.hf.modeling import LongVITAForCausalLM

    cfg = LongVITAConfig.tiny(layers=2, vit_layers=1)
    w = synthetic_state_dict(cfg, seed=5, dtype=torch.bfloat16, perturb=True)
    model = LongVITAForCausalLM(cfg, {k: v.cuda() for k, v in w.items()}).eval()
    ids, images, idx = _inputs(cfg, s=290)
    model.generation_config.max_new_tokens = 5
    out = model.generate(inputs=ids, images=images, image_indices=idx)          # CPU inputs, as in the script
    assert out.shape == (1, 295) and out.device.type == "cuda" and torch.equal(out[:, :290].cpu(), ids)
    seq = ids.cuda()
    for i in range(5):       # greedy + cache == greedy by re-running the whole forward (what the reference's Megatron loop does)
        nxt = model(input_ids=seq, images=images.cuda(), image_indices=idx.cuda(), num_logits_to_keep=1).logits[0, -1].float().argmax()
        assert int(nxt) == int(out[0, 290 + i]), i
        seq = torch.cat([seq, nxt.view(1, 1)], dim=1)
    sd = model.state_dict()
    assert all(torch.equal(sd[k].cpu(), w[k]) for k in w)
