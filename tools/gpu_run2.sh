#!/bin/bash
# every step under its own hard timeout; the sum stays below the gpurun limit
mkdir -p gpurun_out
T="timeout -k 5"
$T 120 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
$T 120 python tools/attn_err.py > gpurun_out/attn_err.log 2>&1; echo "== attn_err exit $?"; tail -4 gpurun_out/attn_err.log
for grp in attention model surfaces elementwise gemm; do
  $T 300 python -m pytest tests/test_gpu_${grp}.py -m gpu -q --timeout 60 --timeout-method=thread > gpurun_out/test_${grp}.log 2>&1
  echo "== ${grp}: exit $?"; tail -n 8 gpurun_out/test_${grp}.log
done
$T 500 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "== bench exit $?"; tail -12 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
cat > /tmp/ncu_attn.py <<'PY'
import torch, sys
sys.path.insert(0, '.')
from long_vita_b200 import ops
q = torch.randn(1, 16384, 40, 128, device='cuda', dtype=torch.bfloat16)
k = torch.randn(1, 16384, 8, 128, device='cuda', dtype=torch.bfloat16)
v = torch.randn(1, 16384, 8, 128, device='cuda', dtype=torch.bfloat16)
for _ in range(3): ops.attention_fwd(q, k, v, causal=True)
torch.cuda.synchronize()
PY
$T 240 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 2 -c 1 -o gpurun_out/r1_attn16k -f python /tmp/ncu_attn.py > gpurun_out/ncu.log 2>&1
echo "== ncu exit $?"; tail -3 gpurun_out/ncu.log
