#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 5"
$T 120 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
$T 200 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_attention.py -m gpu -q --timeout 60 --timeout-method=thread > gpurun_out/test_a.log 2>&1
echo "== elementwise+attention(v1): exit $?"; tail -n 5 gpurun_out/test_a.log
LV_ATTN_VERSION=2 $T 200 python -m pytest tests/test_gpu_attention.py -m gpu -q --timeout 60 --timeout-method=thread > gpurun_out/test_b.log 2>&1
echo "== attention(v2): exit $?"; tail -n 5 gpurun_out/test_b.log
$T 200 python tools/bench_kernels.py --only elem --out gpurun_out/kernels_elem.json > gpurun_out/bench_elem.log 2>&1
echo "== bench elem exit $?"; cat gpurun_out/bench_elem.log | cut -c1-200
LV_ATTN_VERSION=2 $T 200 python tools/bench_kernels.py --only attn --quick --out gpurun_out/kernels_attn_v2b.json > gpurun_out/bench_attn_v2b.log 2>&1
echo "== bench attn v2 exit $?"; cat gpurun_out/bench_attn_v2b.log | cut -c1-160
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
LV_ATTN_VERSION=2 $T 240 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 2 -c 1 -o gpurun_out/r1_attn16k_v2 -f python /tmp/ncu_attn.py > gpurun_out/ncu2.log 2>&1
echo "== ncu exit $?"; tail -2 gpurun_out/ncu2.log
$T 400 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "== bench exit $?"; tail -2 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json | cut -c1-1800
