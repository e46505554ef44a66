"""Attention forward + backward step time for BASELINE config 5 (128K training step, CP over N GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 \
        tools/bench_train_attn.py [--seq 131072] [--iters 5]

Every rank holds its zig-zag shard of one sequence (q [T, 40, 128], k / v [T, 8, 128], T = S / N) and runs
`cp.cp_attention` forward (fused in-kernel K/V exchange) + backward (K/V all-gather, `lv_attn_bwd` on the local
segments, dK/dV reduce-scatter).  Time = max over ranks of the CUDA-event time; FLOPs are the algorithmic causal
counts of SURVEY.md 8d (forward 4 Hq d S(S+1)/2, backward 2.5x).  N = 1 runs the single-GPU `ops.attention`.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seq", type=int, default=131072)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--cp", type=int, default=0, help="context-parallel group size (default: all ranks); ranks / cp = "
                    "data-parallel replicas, each with its own sequence (BASELINE config 5: 8 ranks, --cp 4 -> DP2 x CP4)")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from long_vita_b200 import cp as CP
    from long_vita_b200 import ops

    hq, hkv, d = 40, 8, 128
    cp = a.cp if a.cp else max(world, 1)
    assert world % cp == 0
    dp = world // cp
    group = None
    if world > 1:
        # consecutive ranks form a CP group (Megatron's rank order: CP varies fastest inside a DP replica)
        for r0 in range(0, world, cp):
            g_ = dist.new_group(list(range(r0, r0 + cp)))
            if r0 <= rank < r0 + cp:
                group = g_
    S = a.seq // (2 * cp * 128) * (2 * cp * 128)
    T = S // cp
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    q = torch.randn(T, hq, d, device=dev, generator=g).to(torch.bfloat16).requires_grad_(True)
    k = torch.randn(T, hkv, d, device=dev, generator=g).to(torch.bfloat16).requires_grad_(True)
    v = torch.randn(T, hkv, d, device=dev, generator=g).to(torch.bfloat16).requires_grad_(True)
    d_out = torch.randn(T, hq * d, device=dev, generator=g).to(torch.bfloat16)
    ctx = CP.CPContext(group, S, hq, hkv, d, dev, fused_qkv=False) if cp > 1 else None

    def step():
        for t in (q, k, v):
            t.grad = None
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        if ctx is not None:
            out = CP.cp_attention(q, k, v, ctx)
        else:
            out = ops.attention(q.unsqueeze(0), k.unsqueeze(0), v.unsqueeze(0), causal=True).reshape(T, hq * d)
        e[1].record()
        out.backward(d_out)
        e[2].record()
        torch.cuda.synchronize()
        return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])

    for _ in range(2):
        step()
    fw, bw = [], []
    for _ in range(a.iters):
        if world > 1:
            dist.barrier()
        f, b = step()
        fw.append(f)
        bw.append(b)
    t = torch.tensor([sorted(fw)[len(fw) // 2], sorted(bw)[len(bw) // 2]], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        f_flops = 4.0 * hq * d * (S * (S + 1) / 2)
        fwd_ms, bwd_ms = float(t[0]), float(t[1])
        print(json.dumps({"what": "attention fwd+bwd, one layer, causal 40:8x128", "seq": S, "n_gpus": world,
                          "parallelism": f"dp{dp} x cp{cp}", "sequences_per_step": dp,
                          "fwd_ms": fwd_ms, "bwd_ms": bwd_ms,
                          "fwd_tflops_per_gpu": f_flops / fwd_ms / 1e9 / cp,
                          "bwd_tflops_per_gpu": 2.5 * f_flops / bwd_ms / 1e9 / cp,
                          "step_ms_48_layers": 48 * (fwd_ms + bwd_ms),
                          "tokens_per_s_attention_only_48_layers": dp * S / (48 * (fwd_ms + bwd_ms) / 1e3)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
