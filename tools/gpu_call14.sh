#!/bin/bash
# 1 GPU, ~2 minutes: ncu --set full of the final attention forward kernel (16K LLM shape and the ViT shape).
mkdir -p gpurun_out
T="timeout -k 5"
$T 200 python long-vita_b200/build.py > gpurun_out/build.log 2>&1 || { tail -5 gpurun_out/build.log; exit 1; }
cat > /tmp/attn16k.py <<'PY'
import torch, sys
sys.path.insert(0, '.')
from long_vita_b200 import ops
q = torch.randn(1, 16384, 40, 128, device='cuda', dtype=torch.bfloat16)
k = torch.randn(1, 16384, 8, 128, device='cuda', dtype=torch.bfloat16)
v = torch.randn(1, 16384, 8, 128, device='cuda', dtype=torch.bfloat16)
for _ in range(3):
    ops.attention_fwd(q, k, v, causal=True)
qv = torch.randn(64, 1025, 16, 64, device='cuda', dtype=torch.bfloat16)
for _ in range(3):
    ops.attention_fwd(qv, qv, qv, causal=False)
torch.cuda.synchronize()
PY
$T 120 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 2 -c 1 -f -o gpurun_out/r2_attn16k_final python /tmp/attn16k.py > gpurun_out/ncu_attn_final.log 2>&1
echo "== ncu attn 16k exit $?"; tail -1 gpurun_out/ncu_attn_final.log
$T 120 ncu --set full --clock-control none -k regex:attn_fwd -s 5 -c 1 -f -o gpurun_out/r2_attn_vit_final python /tmp/attn16k.py > gpurun_out/ncu_attn_vit_final.log 2>&1
echo "== ncu attn vit exit $?"; tail -1 gpurun_out/ncu_attn_vit_final.log
