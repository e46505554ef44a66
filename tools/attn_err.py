"""Print the attention error decomposition (vs fp32 oracle) for the P-format in effect."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from long_vita_b200 import ops
from oracle import ops as O

def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm())

for (sq, hq, hkv, d, causal) in [(512, 8, 2, 128, True), (1025, 16, 16, 64, False), (2048, 40, 8, 128, True)]:
    g = torch.Generator().manual_seed(sq)
    q = torch.randn(1, sq, hq, d, generator=g).bfloat16(); k = torch.randn(1, sq, hkv, d, generator=g).bfloat16(); v = torch.randn(1, sq, hkv, d, generator=g).bfloat16()
    ref, lse_ref = O.attention(q, k, v, causal=causal)
    out, lse = ops.attention_fwd(q.cuda(), k.cuda(), v.cuda(), causal=causal, return_lse=True)
    et, ef = rel(out, ref), rel(ref.bfloat16(), ref)
    line = {"P": "bf16", "shape": [sq, hq, hkv, d, causal], "e_total": et, "e_floor": ef,
            "excess": math.sqrt(max(et * et - ef * ef, 0)), "lse_abs": float((lse.cpu() - lse_ref).abs().max())}
    try:
        import flash_attn
        fa = flash_attn.flash_attn_func(q.cuda(), k.cuda(), v.cuda(), causal=causal)
        e = rel(fa, ref); line["flash_attn_excess"] = math.sqrt(max(e * e - ef * ef, 0))
    except Exception as ex:
        line["flash_attn"] = str(ex)[:80]
    print(line, flush=True)
