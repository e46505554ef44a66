#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for grp in elementwise attention; do
  timeout 600 python -m pytest tests/test_gpu_${grp}.py -m gpu -q --timeout 90 --timeout-method=thread > gpurun_out/test_${grp}.log 2>&1
  echo "== ${grp}: exit $?"; tail -n 6 gpurun_out/test_${grp}.log
done
timeout 600 python tools/bench_kernels.py > gpurun_out/bench_kernels.log 2>&1
echo "== bench exit $?"; cat gpurun_out/bench_kernels.log | tail -40
# ncu: one attention launch at 16K and one GEMM launch
cat > /tmp/ncu_attn.py <<'PY'
import torch, sys
sys.path.insert(0, '.')
from long_vita_b200 import ops
q = torch.randn(1, 16384, 40, 128, device='cuda', dtype=torch.bfloat16)
k = torch.randn(1, 16384, 8, 128, device='cuda', dtype=torch.bfloat16)
v = torch.randn(1, 16384, 8, 128, device='cuda', dtype=torch.bfloat16)
for _ in range(2): ops.attention_fwd(q, k, v, causal=True)
x = torch.randn(16384, 5120, device='cuda', dtype=torch.bfloat16); w = torch.randn(7168, 5120, device='cuda', dtype=torch.bfloat16)
for _ in range(2): ops.linear(x, w)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd|gemm_bf16" -s 2 -c 2 -o gpurun_out/r1_attn_gemm -f python /tmp/ncu_attn.py > gpurun_out/ncu.log 2>&1
echo "== ncu exit $?"; tail -5 gpurun_out/ncu.log
