"""Context-parallel attention backward, host logic on CPU (world_size 2 over gloo).

`cp.CPBackwardMixin` composes the backward from: all-gather of K/V + zig-zag -> global re-order, the
single-GPU backward on the two local query segments, re-order + reduce-scatter of the partial
dK/dV.  Here the kernels are replaced by the CPU oracle (tests/hostlogic.py) and the CUDA-IPC context
by a gloo stand-in that keeps the mixin, so what is tested is exactly the index / collective logic:
the sharded gradients, gathered and un-permuted, must equal the unsharded oracle gradients.
Kernel parity of the same composition is the `-m gpu` test in tests/test_gpu_cp.py.
"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import rel_fro

S, HQ, HKV, D = 512, 4, 2, 32


def _tensors():
    g = torch.Generator().manual_seed(11)
    mk = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16)   # noqa: E731
    return mk(S, HQ, D), mk(S, HKV, D), mk(S, HKV, D), mk(S, HQ * D)


def _worker(rank, world, port, path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from long_vita_b200 import cp as CP
        from oracle import ops as O
        from tests.hostlogic import oracle_ops

        class GlooCP(CP.CPBackwardMixin):
            def __init__(self):
                self.group, self.cp, self.rank = dist.group.WORLD, world, rank
                self.S, self.T, self.hq, self.hkv, self.d = S, S // world, HQ, HKV, D

            def attention_separate(self, q, k, v, out=None, scale=None, return_lse=False):
                K, V = self.gather_kv(k, v)                       # the forward's exchange, over gloo
                o, lse = O.attention(q[None], K[None], V[None], causal=True, scale=scale,
                                     q_pos=CP.zigzag_index(S, world, rank), kv_pos=torch.arange(S))
                o = o[0].to(torch.bfloat16).reshape(self.T, HQ * D)
                return (o, lse) if return_lse else o

        torch.set_num_threads(2)
        q, k, v, d_out = _tensors()
        own = CP.zigzag_index(S, world, rank)
        ql, kl, vl = (t[own].clone().requires_grad_(True) for t in (q, k, v))
        with oracle_ops():
            out = CP.cp_attention(ql, kl, vl, GlooCP())
            out.backward(d_out[own])
        res = {}
        for name, t in (("out", out.detach()), ("dq", ql.grad), ("dk", kl.grad), ("dv", vl.grad)):
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t.contiguous())
            res[name] = torch.cat(parts)[CP.zigzag_unpermute_index(S, world)]
        if rank == 0:
            torch.save(res, path)
    finally:
        dist.destroy_process_group()


def test_two_rank_cp_backward_equals_unsharded(tmp_path):
    from oracle import ops as O

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    path = str(tmp_path / "grads.pt")
    mp.spawn(_worker, args=(2, port, path), nprocs=2, join=True)
    got = torch.load(path)
    q, k, v, d_out = _tensors()
    ref_out, _ = O.attention(q[None], k[None], v[None], causal=True)
    dq, dk, dv = O.attention_grads(q[None], k[None], v[None], d_out.view(1, S, HQ, D), causal=True)
    assert rel_fro(got["out"].view(S, HQ, D), ref_out[0]) < 4e-3
    # bf16 roundings: out / partial dK,dV / the reduce-scatter sum
    assert rel_fro(got["dq"], dq[0]) < 6e-3, rel_fro(got["dq"], dq[0])
    assert rel_fro(got["dk"], dk[0]) < 6e-3, rel_fro(got["dk"], dk[0])
    assert rel_fro(got["dv"], dv[0]) < 6e-3, rel_fro(got["dv"], dv[0])
