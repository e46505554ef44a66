#!/bin/bash
# 1 GPU: lean softmax loop of the attention forward (parity on the watchdog build, speed, ncu), new GPU tests, bench line.
mkdir -p gpurun_out
T="timeout -k 5"
LV_WATCHDOG=1 $T 300 python long-vita_b200/build.py > gpurun_out/build_watchdog.log 2>&1 || { tail -20 gpurun_out/build_watchdog.log; exit 1; }
$T 400 python -m pytest tests/test_gpu_attention.py tests/test_gpu_attention_long.py -m gpu -q -x --timeout 200 --timeout-method=thread > gpurun_out/c6_test_attn.log 2>&1
echo "== attention parity (watchdog build) exit $?"; tail -n 5 gpurun_out/c6_test_attn.log
LV_ATTN_TURNS=0 $T 300 python -m pytest tests/test_gpu_attention.py -m gpu -q -x --timeout 200 --timeout-method=thread > gpurun_out/c6_test_attn_t0.log 2>&1
echo "== attention parity turns=0 exit $?"; tail -n 3 gpurun_out/c6_test_attn_t0.log
LV_WATCHDOG=0 $T 300 python long-vita_b200/build.py > gpurun_out/build_release.log 2>&1 || { tail -20 gpurun_out/build_release.log; exit 1; }
$T 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread --deselect tests/test_gpu_cp.py --deselect tests/test_gpu_attention_long.py -rf > gpurun_out/c6_test_all.log 2>&1
echo "== all other 1-GPU tests exit $?"; tail -n 8 gpurun_out/c6_test_all.log
for TU in 1 0; do
  LV_ATTN_TURNS=$TU $T 200 python tools/bench_kernels.py --only attn --out gpurun_out/c6_attn_t$TU.json > gpurun_out/c6_attn_t$TU.log 2>&1
  echo "== attn turns=$TU exit $?"; cut -c1-125 gpurun_out/c6_attn_t$TU.log | tail -n 5
done
LV_ATTN_VERSION=2 $T 200 python tools/bench_kernels.py --only attn --quick --out gpurun_out/c6_attn_v2.json > gpurun_out/c6_attn_v2.log 2>&1
echo "== attn v2 exit $?"; cut -c1-125 gpurun_out/c6_attn_v2.log | tail -n 4
cat > /tmp/attn16k.py <<'PY'
import torch, sys
sys.path.insert(0, '.')
from long_vita_b200 import ops
q = torch.randn(1, 16384, 40, 128, device='cuda', dtype=torch.bfloat16)
k = torch.randn(1, 16384, 8, 128, device='cuda', dtype=torch.bfloat16)
v = torch.randn(1, 16384, 8, 128, device='cuda', dtype=torch.bfloat16)
for _ in range(3):
    ops.attention_fwd(q, k, v, causal=True)
torch.cuda.synchronize()
PY
$T 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 2 -c 1 -f -o gpurun_out/r2_attn16k_lean python /tmp/attn16k.py > gpurun_out/ncu_attn_lean.log 2>&1
echo "== ncu attn lean exit $?"; tail -2 gpurun_out/ncu_attn_lean.log
$T 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c6_bench_n1.json 2> gpurun_out/c6_bench_n1.err
echo "== bench exit $?"; tail -3 gpurun_out/c6_bench_n1.err; cut -c1-300 gpurun_out/c6_bench_n1.json
# backward kernel: speed + one ncu capture of each pass (source view for the next optimisation step)
$T 200 python tools/bench_bwd.py > gpurun_out/c6_bwd.log 2>&1
echo "== bwd exit $?"; tail -n 3 gpurun_out/c6_bwd.log
cat > /tmp/bwd16k.py <<'PY'
import torch, sys
sys.path.insert(0, '.')
from long_vita_b200 import ops
S = 16384
q = torch.randn(1, S, 40, 128, device='cuda', dtype=torch.bfloat16)
k = torch.randn(1, S, 8, 128, device='cuda', dtype=torch.bfloat16)
v = torch.randn(1, S, 8, 128, device='cuda', dtype=torch.bfloat16)
out, lse = ops.attention_fwd(q, k, v, causal=True, return_lse=True)
do = torch.randn_like(out)
for _ in range(2):
    ops.attention_bwd(do, q, k, v, out, lse, causal=True)
torch.cuda.synchronize()
PY
$T 500 ncu --set full --clock-control none --import-source on -k regex:attn_bwd2 -s 2 -c 2 -f -o gpurun_out/r2_bwd16k python /tmp/bwd16k.py > gpurun_out/ncu_bwd.log 2>&1
echo "== ncu bwd exit $?"; tail -2 gpurun_out/ncu_bwd.log
# DRAM traffic of the dominant kernel inside the bench step (roofline.traffic): the QKV / O / gate|up / down GEMMs of one decoder layer
$T 600 ncu --set full --clock-control none -k regex:gemm_bf16 -s 103 -c 8 -f -o gpurun_out/r2_bench_gemm python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-attn-probe > gpurun_out/ncu_bench_gemm.log 2>&1
echo "== ncu bench gemm exit $?"; tail -2 gpurun_out/ncu_bench_gemm.log
