#!/usr/bin/env python
"""Headline benchmark: tokens/sec of one Long-VITA prefill forward (ViT tower + projector + 48-layer
14B decoder + masked LM head) on synthetic frames, through the HF `forward()` surface.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--frames F]

A "step" is one full prefill of the configured prompt.  N = 1: BASELINE.json configs[1]
("Long-VITA-16K bf16 single B200, 64 synthetic frames -> 16K tokens").  Launched under torchrun
for N > 1 (one rank per GPU); see DESIGN.md for what each N runs.  Rank 0 prints ONE JSON line.

`--impl reference` times the reference's own CPU implementation of the path - the oracle port
of the HF forward (oracle/model.py; the reference itself cannot be imported here) -
on the host cores with all threads, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "prefill_tokens_per_sec"
UNIT = "tokens/s"


T0 = time.time()


def log(msg: str) -> None:
    sys.stderr.write(f"[bench +{time.time() - T0:7.1f}s] {msg}\n")
    sys.stderr.flush()


def host_threads() -> int:
    """Threads the CPU legs may use: the cores this process is allowed on (cgroup / affinity aware)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(n, 64))


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:  # noqa: BLE001
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU baseline (oracle port of the reference HF forward), bounded sample + extrapolation
# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(cfg, total_tokens: int, n_frames: int, threads: int, sample_tokens: int = 2048):
    """Time the oracle on a bounded sample and extrapolate to one full prefill of `total_tokens`.

    Sample: ONE of the 48 decoder layers on `sample_tokens` tokens (attention timed separately from
    the token-wise part), ONE frame through 2 of the 24 ViT layers + the projector, and the LM head
    on one row.  Extrapolation: token-wise cost scales linearly with S, attention quadratically,
    ViT linearly with frames and layers.  Returns (tokens_per_sec, seconds_measured, description)."""
    import torch

    from long_vita_b200.weights import global_weights, llm_layer_weights, vit_layer_weights
    from oracle import model as OM
    from oracle import ops as O

    torch.set_num_threads(threads)
    S = min(sample_tokens, total_tokens)
    w = llm_layer_weights(cfg, 0, 1234, "cpu", torch.float32)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(S, cfg.hidden_size, generator=g)
    pos = torch.arange(S)
    cos, sin = O.rope_tables(pos, O.rope_inv_freq(cfg.head_dim, cfg.rope_theta), torch.float32)
    t0 = time.perf_counter()
    with torch.no_grad():
        OM.decoder_layer(cfg, w, 0, x, cos, sin)
    t_layer = time.perf_counter() - t0
    # attention alone (same shapes) to split the quadratic part
    q = torch.randn(1, S, cfg.num_attention_heads, cfg.head_dim, generator=g)
    k = torch.randn(1, S, cfg.num_key_value_heads, cfg.head_dim, generator=g)
    t0 = time.perf_counter()
    with torch.no_grad():
        O.attention(q, k, k, causal=True)
    t_attn = time.perf_counter() - t0
    t_tok = max(t_layer - t_attn, 1e-6)
    del w
    # vision: one frame, two layers
    v = cfg.visual
    wv = global_weights(cfg, 1234, "cpu", torch.float32, with_lm=False)
    for i in range(2):
        wv.update(vit_layer_weights(cfg, i, 1234, "cpu", torch.float32))
    img = torch.randn(1, 3, v.image_size, v.image_size, generator=g)
    t0 = time.perf_counter()
    with torch.no_grad():
        vit = OM.vit_forward(cfg, wv, img, num_layers=2)
        OM.projector_forward(cfg, wv, vit[:, 1:, :])
    t_vit2 = time.perf_counter() - t0
    measured = t_layer + t_attn + t_vit2
    L = cfg.num_hidden_layers
    ratio = total_tokens / S
    t_full = L * (t_tok * ratio + t_attn * ratio * ratio) + n_frames * t_vit2 * (v.num_hidden_layers / 2.0)
    desc = (f"oracle port (fp32, {threads} threads): 1/{L} decoder layers on {S} tokens (attention timed apart), "
            f"1/{n_frames} frames through 2/{v.num_hidden_layers} ViT layers + projector; extrapolated "
            f"(token-wise ~S, attention ~S^2, ViT ~frames x layers) to {total_tokens} tokens")
    return total_tokens / t_full, measured, desc


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=64, help="synthetic frames (64 -> 16K, 512 -> 128K, 4096 -> 1M)")
    ap.add_argument("--text", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--long-run", action="store_true",
                    help="minutes-per-step configs (1M tokens): honour --warmup < 3 and skip the separate e2e pass; "
                         "the printed line is then marked as outside the timing contract")
    ap.add_argument("--layers", type=int, default=None, help="debug: fewer decoder layers (number is then INVALID)")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: everything libraries print meanwhile (e.g. NCCL's
    # "NCCL version ..." banner, written to fd 1 from C) is routed to stderr until the line is emitted
    sys.stdout.flush()
    _stdout_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(_stdout_fd, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from long_vita_b200.config import LongVITAConfig
    from long_vita_b200.synthetic import build_prompt

    cfg = LongVITAConfig.long_vita_14b()
    cp = max(world, 1)
    # the SAME prompt at every N: length padded (with text tokens) to a multiple of 2*8*128 so that it
    # shards zig-zag over 1, 2, 4 or 8 ranks in 128-token units
    ids, image_indices = build_prompt(cfg, args.frames, args.text, pad_multiple=2048)
    S = ids.shape[1]
    workload = (f"Long-VITA-{'16K' if args.frames == 64 else str(S)} prefill, {args.frames} synthetic frames "
                f"({args.frames * 256} visual + {2 * args.frames} delimiter tokens) + text, padded to {S} tokens")
    config = {"workload": workload, "frames": args.frames, "tokens": S, "layers": cfg.num_hidden_layers,
              "parallelism": f"cp{cp}" if cp > 1 else "single",
              "l2": "per-step working set (29.5 GB weights + activations) far exceeds the 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return
        threads = host_threads()
        vals, meas = [], 0.0
        for i in range(args.warmup + args.steps):
            v, m, desc = cpu_reference_sample(cfg, S, args.frames, threads)
            if i >= args.warmup:
                vals.append(v)
                meas += m
        value = sum(vals) / len(vals)
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * S / value,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    import torch
    import torch.distributed as dist

    from long_vita_b200 import ops
    from long_vita_b200.hf.modeling import LongVITAForCausalLM
    from long_vita_b200.synthetic import synthetic_frames

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    log("materialising random-init weights on the GPU")
    model = LongVITAForCausalLM.from_synthetic(cfg, seed=1234, device=dev, num_layers=args.layers)
    log("weights ready")
    if world > 1:
        from long_vita_b200 import cp as cpmod

        runner = cpmod.ContextParallelRunner(model, dist.group.WORLD)
    else:
        runner = None

    # host-side inputs in pinned memory (the e2e region copies them every step)
    # --long-run (no e2e pass): frames are drawn on the device, no pinned host copy of ~5 GB per rank
    images_h = None if args.long_run else synthetic_frames(cfg, args.frames, pin=True)
    ids_h = ids.pin_memory()
    idx_h = image_indices.pin_memory()
    h2d = (0 if images_h is None else images_h.numel() * 2) + ids_h.numel() * 8 + idx_h.numel() * 8
    logits_h = torch.empty((1, 1, cfg.vocab_size), dtype=torch.bfloat16).pin_memory()
    d2h = logits_h.numel() * 2

    def forward_resident(images_d, ids_d, idx_d):
        if runner is not None:
            return runner.forward(ids_d, images_d, idx_d)
        return model(input_ids=ids_d, images=images_d, image_indices=idx_d, num_logits_to_keep=1).logits

    def step_e2e():
        images_d = images_h.to(dev, non_blocking=True)
        ids_d = ids_h.to(dev, non_blocking=True)
        idx_d = idx_h.to(dev, non_blocking=True)
        logits = forward_resident(images_d, ids_d, idx_d)
        logits_h.copy_(logits.view(1, 1, -1), non_blocking=True)

    images_d = synthetic_frames(cfg, args.frames, device=dev) if images_h is None else images_h.to(dev)
    ids_d = ids_h.to(dev)
    idx_d = idx_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    n_warm = args.warmup if args.long_run else max(args.warmup, 3)
    for i in range(n_warm):
        forward_resident(images_d, ids_d, idx_d)
        torch.cuda.synchronize()
        log(f"warm-up step {i} done")
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    timer = ops.KernelTimer()
    ops.set_kernel_timer(timer)
    n0 = ops.launch_count()
    ms_resident = timed(lambda: forward_resident(images_d, ids_d, idx_d), args.steps)
    log(f"timed region done: {ms_resident / args.steps:.1f} ms/step")
    launches = ops.launch_count() - n0
    ops.set_kernel_timer(None)
    torch.cuda.synchronize()
    ksum = timer.summary()
    # e2e: host buffers, H2D + forward + D2H inside the timed region
    if args.long_run:
        ms_e2e = None
    else:
        step_e2e()
        ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk, pk_kind = peaks()
    ms_step = ms_resident / args.steps
    value = S / (ms_step / 1000.0)
    e2e_value = S / (ms_e2e / args.steps / 1000.0) if ms_e2e is not None else None

    def roof(kind):
        d = ksum.get(kind)
        if not d or d["ms"] <= 0:
            return None
        ach = d["flops"] / (d["ms"] / 1000.0) / 1e12
        peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        return {"bound": "tensor", "kernel": kind, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "peak_source": f"{pk_kind} sustained cuBLAS bf16 (kernel timed inside a long step)",
                "launches_per_step": d["launches"] / args.steps, "avg_launch_ms": d["ms"] / d["launches"],
                "share_of_step": d["ms"] / ms_resident, "traffic": None}

    dominant = max(ksum, key=lambda k_: ksum[k_]["ms"]) if ksum else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": n_warm,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic", "config": config, "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "roofline": roof(dominant) if dominant else None,
        "roofline_attn": roof("attn_fwd"),
    }
    if args.long_run:
        line["note"] = "--long-run: fewer than 3 warm-up steps and no separate e2e pass (minutes per step)"
    if args.layers is not None:
        line["INVALID"] = f"debug run with {args.layers} decoder layers"
    if world == 1 and not args.no_cpu_baseline:
        cores = host_threads()
        log(f"cpu baseline on {cores} threads")
        v, m, desc = cpu_reference_sample(cfg, S, args.frames, cores)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
                                "seconds_measured": m}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
