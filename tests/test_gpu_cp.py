"""Multi-GPU parity of the fused context-parallel attention (in-kernel K/V exchange over NVLink)
and of the sharded prefill: cp = N ranks must reproduce the single-device result (the invariant
TE / ring-flash-attn test for their ring).  Needs >= 2 GPUs: run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_cp.py -m gpu`."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm())


def _worker(rank, world, port):
    import math

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from long_vita_b200 import cp as CP
        from long_vita_b200 import ops
        from long_vita_b200.config import LongVITAConfig
        from long_vita_b200.hf.modeling import LongVITAForCausalLM
        from long_vita_b200.synthetic import build_prompt, synthetic_frames
        from long_vita_b200.weights import synthetic_state_dict
        from oracle import ops as O

        # ---- kernel level: CPContext.attention over three epochs (both buffer parities) ----
        hq, hkv, d = 10, 2, 128
        S = 2 * world * 384          # chunk = 384: a multiple of 128 but not of 256
        ctx = CP.CPContext(dist.group.WORLD, S, hq, hkv, d, dev)
        own = CP.zigzag_index(S, world, rank)
        for epoch in range(3):
            g = torch.Generator().manual_seed(100 + epoch)        # same tensors on every rank
            qkv = torch.randn(S, (hq + 2 * hkv) * d, generator=g).to(torch.bfloat16)
            ctx.qkv_buffer().copy_(qkv[own].to(dev))
            out = ctx.attention()
            q = qkv[:, : hq * d].view(1, S, hq, d)
            k = qkv[:, hq * d : (hq + hkv) * d].view(1, S, hkv, d)
            v = qkv[:, (hq + hkv) * d :].view(1, S, hkv, d)
            ref, _ = O.attention(q, k, v, causal=True)
            ref_l = ref[0, own].reshape(own.numel(), hq * d)
            e = _rel(out, ref_l)
            e_floor = _rel(ref_l.to(torch.bfloat16), ref_l)
            assert math.sqrt(max(e * e - e_floor * e_floor, 0.0)) < 2e-3, (rank, epoch, e, e_floor)
        dist.barrier()

        # ---- the 14B K/V geometry (8 kv heads x 128 = 4 KB K|V rows): the copier's one-row-per-pass fast path ----
        hq2, hkv2 = 40, 8
        S2 = 2 * world * 128
        ctx2 = CP.CPContext(dist.group.WORLD, S2, hq2, hkv2, d, dev)
        own2 = CP.zigzag_index(S2, world, rank)
        for epoch in range(2):
            g = torch.Generator().manual_seed(300 + epoch)
            qkv = torch.randn(S2, (hq2 + 2 * hkv2) * d, generator=g).to(torch.bfloat16)
            ctx2.qkv_buffer().copy_(qkv[own2].to(dev))
            out = ctx2.attention()
            q = qkv[:, : hq2 * d].view(1, S2, hq2, d)
            k = qkv[:, hq2 * d : (hq2 + hkv2) * d].view(1, S2, hkv2, d)
            v = qkv[:, (hq2 + hkv2) * d :].view(1, S2, hkv2, d)
            ref, _ = O.attention(q, k, v, causal=True)
            ref_l = ref[0, own2].reshape(own2.numel(), hq2 * d)
            e = _rel(out, ref_l)
            e_floor = _rel(ref_l.to(torch.bfloat16), ref_l)
            assert math.sqrt(max(e * e - e_floor * e_floor, 0.0)) < 2e-3, (rank, epoch, e, e_floor)
        dist.barrier()
        # release the peer-mapped buffers (collective): mappings closed on every rank, then the allocations freed
        free0 = torch.cuda.mem_get_info(dev)[0]
        ctx.close()
        ctx2.close()
        ctx.close()      # idempotent
        assert ctx.base is None and torch.cuda.mem_get_info(dev)[0] >= free0

        # ---- whole sharded prefill vs the single-device forward ----
        cfg = LongVITAConfig.tiny(layers=3, vit_layers=1)
        w = synthetic_state_dict(cfg, seed=11, dtype=torch.bfloat16, perturb=True)
        model = LongVITAForCausalLM(cfg, {k_: t.to(dev) for k_, t in w.items()})
        ids, idx = build_prompt(cfg, 5, n_text=30, pad_multiple=2 * world * 128, seed=3)
        images = synthetic_frames(cfg, 5, seed=3)
        single = model(input_ids=ids.to(dev), images=images.to(dev), image_indices=idx.to(dev), num_logits_to_keep=1).logits
        runner = CP.ContextParallelRunner(model, dist.group.WORLD)
        for _ in range(2):   # second call re-uses the context (epochs continue)
            sharded = runner.forward(ids.to(dev), images.to(dev), idx.to(dev))
            assert _rel(sharded, single) < 1e-2, (rank, _rel(sharded, single))
            assert int(sharded.float().argmax()) == int(single.float().argmax())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_context_parallel_matches_single_device(lib_built, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    mp.spawn(_worker, args=(world, _free_port()), nprocs=world, join=True)


def _bwd_worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from long_vita_b200 import cp as CP
        from oracle import ops as O

        hq, hkv, d = 10, 2, 128
        S = 2 * world * 256
        ctx = CP.CPContext(dist.group.WORLD, S, hq, hkv, d, dev, fused_qkv=False)
        own = CP.zigzag_index(S, world, rank)
        for step in range(2):                                   # two steps: both buffer parities, epochs continue
            g = torch.Generator().manual_seed(500 + step)       # same tensors on every rank
            q = torch.randn(S, hq, d, generator=g).to(torch.bfloat16)
            k = torch.randn(S, hkv, d, generator=g).to(torch.bfloat16)
            v = torch.randn(S, hkv, d, generator=g).to(torch.bfloat16)
            d_out = torch.randn(S, hq * d, generator=g).to(torch.bfloat16)
            ql, kl, vl = (t[own].to(dev).requires_grad_(True) for t in (q, k, v))
            out = CP.cp_attention(ql, kl, vl, ctx)
            out.backward(d_out[own].to(dev))
            dq, dk, dv = O.attention_grads(q[None], k[None], v[None], d_out.view(1, S, hq, d), causal=True)
            for name, got, ref in (("dq", ql.grad, dq[0][own]), ("dk", kl.grad, dk[0][own]), ("dv", vl.grad, dv[0][own])):
                e = _rel(got, ref)
                # single-GPU backward tolerance (tests/test_gpu_attention_bwd.py) plus one bf16 rounding of the
                # partial dK/dV before the reduce-scatter
                assert e < 8e-3, (rank, step, name, e)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_context_parallel_backward_matches_single_device(lib_built, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    mp.spawn(_bwd_worker, args=(world, _free_port()), nprocs=world, join=True)


def _decode_worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from long_vita_b200 import cp as CP
        from long_vita_b200.config import LongVITAConfig
        from long_vita_b200.hf.modeling import LongVITAForCausalLM
        from long_vita_b200.synthetic import build_prompt, synthetic_frames
        from long_vita_b200.weights import synthetic_state_dict

        cfg = LongVITAConfig.tiny(layers=3, vit_layers=1)
        w = synthetic_state_dict(cfg, seed=11, dtype=torch.bfloat16, perturb=True)
        model = LongVITAForCausalLM(cfg, {k_: t.to(dev) for k_, t in w.items()})
        ids, idx = build_prompt(cfg, 3, n_text=30, pad_multiple=2 * world * 128, seed=3)
        images = synthetic_frames(cfg, 3, seed=3)
        S = ids.shape[1]
        new = torch.randint(0, cfg.vocab_size, (4,), generator=torch.Generator().manual_seed(6))
        full = model(input_ids=torch.cat([ids, new.view(1, 4)], dim=1).to(dev), images=images.to(dev),
                     image_indices=idx.to(dev)).logits
        runner = CP.ContextParallelRunner(model, dist.group.WORLD)
        first = runner.forward(ids.to(dev), images.to(dev), idx.to(dev), use_cache=True, max_new_tokens=16)
        assert _rel(first[0, 0], full[0, S - 1]) < 1.5e-2
        for i in range(4):
            lg = runner.decode(new[i].to(dev))
            assert _rel(lg[0, 0], full[0, S + i]) < 2e-2, (rank, i, _rel(lg[0, 0], full[0, S + i]))
            assert int(lg[0, 0].float().argmax()) == int(full[0, S + i].float().argmax())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_context_parallel_sharded_cache_decode(lib_built, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    mp.spawn(_decode_worker, args=(world, _free_port()), nprocs=world, join=True)


def _fault_worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LV_CP_TIMEOUT_MS="1500")
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import time

        from long_vita_b200 import cp as CP

        hq, hkv, d = 10, 2, 128
        S = 2 * world * 128
        ctx = CP.CPContext(dist.group.WORLD, S, hq, hkv, d, dev)
        ctx.qkv_buffer().normal_()
        out = ctx.attention()            # healthy epoch 0 on both ranks
        ctx.check()
        if rank == 0:
            ctx.qkv_buffer().normal_()
            t0 = time.time()
            ctx.attention()              # rank 1 never runs this epoch: the peer wait must time out, not hang
            with pytest.raises(RuntimeError, match="timed out"):
                ctx.check()
            assert time.time() - t0 < 30.0
            with pytest.raises(RuntimeError, match="code -4"):     # LV_ESTATE, sticky
                ctx.check()
        else:
            time.sleep(6.0)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_a_missing_peer_times_out_with_an_error_instead_of_hanging(lib_built):
    """Bounded in-kernel waits (LV_CP_TIMEOUT_MS): a rank whose peer never publishes its K/V rows gets LV_ESTATE."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mp.spawn(_fault_worker, args=(2, _free_port()), nprocs=2, join=True)
