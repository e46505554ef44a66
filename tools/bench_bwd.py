"""Attention backward timing (dQ, dK, dV; algorithmic FLOPs = 2.5x forward) next to flash-attn 2.8."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from long_vita_b200 import ops
import flash_attn

def timeit(fn, iters=5, warmup=2):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]

for name, s, hq, hkv, d, causal in [("llm8k", 8192, 40, 8, 128, True), ("llm16k", 16384, 40, 8, 128, True)]:
    q = torch.randn(1, s, hq, d, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(1, s, hkv, d, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(1, s, hkv, d, device="cuda", dtype=torch.bfloat16)
    do = torch.randn_like(q)
    out, lse = ops.attention_fwd(q, k, v, causal=causal, return_lse=True)
    flops = 2.5 * 4.0 * hq * d * (s * (s + 1) / 2)
    ms = timeit(lambda: ops.attention_bwd(do, q, k, v, out, lse, causal=causal))
    qf, kf, vf = (t.clone().requires_grad_(True) for t in (q, k, v))
    of = flash_attn.flash_attn_func(qf, kf, vf, causal=causal)
    fms = timeit(lambda: torch.autograd.grad(of, (qf, kf, vf), do, retain_graph=True))
    print(json.dumps({"kernel": "attn_bwd", "shape": name, "ms": ms, "tflops": flops / ms / 1e9, "flash_attn2_ms": fms,
                      "flash_attn2_tflops": flops / fms / 1e9}), flush=True)
