"""Context-parallel attention of one layer: the fused in-kernel exchange (`lv_attn_cp_fwd`) next to the comparator the
reference's stack stands for - `ring_flash_attn.zigzag_ring_flash_attn_func` (flash-attn 2.8 kernels + NCCL P2P ring, the
same zig-zag schedule as TransformerEngine's `AttnFuncWithCP` behind gpt_layer_specs.py:35-45).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_cp_compare.py [--seq 18432 131072]

Every rank holds its zig-zag shard (q [T, 40, 128], k / v [T, 8, 128]); time = max over ranks of the CUDA-event median.
The third-party package's `__init__` does not import under the installed transformers (an HF adapter), so the two
modules the comparator needs are loaded from their files; nothing of it is used by the product.
"""
import argparse
import importlib.util
import json
import os
import sys
import types

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load_zigzag_ring():
    import site

    for sp in site.getsitepackages():
        base = os.path.join(sp, "ring_flash_attn")
        if os.path.isdir(base):
            pkg = types.ModuleType("ring_flash_attn")
            pkg.__path__ = [base]
            sys.modules["ring_flash_attn"] = pkg
            mods = {}
            for name in ("utils", "zigzag_ring_flash_attn"):
                spec = importlib.util.spec_from_file_location("ring_flash_attn." + name, os.path.join(base, name + ".py"))
                m = importlib.util.module_from_spec(spec)
                sys.modules["ring_flash_attn." + name] = m
                spec.loader.exec_module(m)
                mods[name] = m
            return mods["zigzag_ring_flash_attn"].zigzag_ring_flash_attn_func
    return None


def median_ms(fn, iters, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = torch.tensor([sorted(ts)[len(ts) // 2]], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seq", type=int, nargs="+", default=[18432, 131072])
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from long_vita_b200 import cp as CP

    ring = load_zigzag_ring()
    hq, hkv, d = 40, 8, 128
    for seq in a.seq:
        S = seq // (2 * world * 128) * (2 * world * 128)
        T = S // world
        g = torch.Generator(device=dev).manual_seed(7 + rank)
        q = torch.randn(T, hq, d, device=dev, generator=g).to(torch.bfloat16)
        k = torch.randn(T, hkv, d, device=dev, generator=g).to(torch.bfloat16)
        v = torch.randn(T, hkv, d, device=dev, generator=g).to(torch.bfloat16)
        ctx = CP.CPContext(dist.group.WORLD, S, hq, hkv, d, dev, fused_qkv=False)
        ours = median_ms(lambda: ctx.attention_separate(q, k, v), a.iters)
        theirs = None
        if ring is not None:
            qb, kb, vb = q.unsqueeze(0), k.unsqueeze(0), v.unsqueeze(0)
            with torch.no_grad():
                theirs = median_ms(lambda: ring(qb, kb, vb, causal=True, group=dist.group.WORLD), a.iters)
                if S <= 32768:      # same zig-zag layout: the two results must agree (two bf16 flash attentions)
                    o1 = ctx.attention_separate(q, k, v).view(T, hq, d).float()
                    o2 = ring(qb, kb, vb, causal=True, group=dist.group.WORLD)[0].float()
                    diff = float((o1 - o2).norm() / o2.norm())
                else:
                    diff = None
        ctx.close()
        if rank == 0:
            flops = 4.0 * hq * d * (S * (S + 1) / 2) / world
            print(json.dumps({"what": "one layer of causal context-parallel attention, 40:8 x 128", "seq": S, "n_gpus": world,
                              "lv_attn_cp_fwd_ms": ours, "lv_tflops_per_gpu": flops / ours / 1e9,
                              "zigzag_ring_flash_attn_ms": theirs,
                              "ring_tflops_per_gpu": None if theirs is None else flops / theirs / 1e9,
                              "speedup": None if theirs is None else theirs / ours,
                              "rel_diff_between_the_two_results": diff}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
