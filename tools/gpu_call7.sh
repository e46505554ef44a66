#!/bin/bash
# 1 GPU: packed-fp32 softmax + Q double buffering + templated GEMM: parity (watchdog build), speed, ncu, bench line.
mkdir -p gpurun_out
T="timeout -k 5"
LV_WATCHDOG=1 $T 300 python long-vita_b200/build.py > gpurun_out/build_watchdog.log 2>&1 || { tail -20 gpurun_out/build_watchdog.log; exit 1; }
$T 500 python -m pytest tests/test_gpu_attention.py tests/test_gpu_attention_long.py tests/test_gpu_gemm.py tests/test_gpu_preprocess.py -m gpu -q -x --timeout 200 --timeout-method=thread > gpurun_out/c7_test_attn.log 2>&1
echo "== attention / gemm / preprocess parity (watchdog build) exit $?"; tail -n 5 gpurun_out/c7_test_attn.log
LV_GEMM_BN=128 $T 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x --timeout 200 --timeout-method=thread > gpurun_out/c7_test_gemm128.log 2>&1
echo "== gemm parity with 128-wide tiles exit $?"; tail -n 3 gpurun_out/c7_test_gemm128.log
LV_WATCHDOG=0 $T 300 python long-vita_b200/build.py > gpurun_out/build_release.log 2>&1 || { tail -20 gpurun_out/build_release.log; exit 1; }
$T 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread --deselect tests/test_gpu_cp.py --deselect tests/test_gpu_attention_long.py -rf > gpurun_out/c7_test_all.log 2>&1
echo "== all other 1-GPU tests exit $?"; tail -n 6 gpurun_out/c7_test_all.log
$T 200 python tools/bench_kernels.py --only attn --out gpurun_out/c7_attn.json > gpurun_out/c7_attn.log 2>&1
echo "== attn exit $?"; cut -c1-125 gpurun_out/c7_attn.log | tail -n 5
LV_ATTN_TURNS=1 $T 200 python tools/bench_kernels.py --only attn --quick --out gpurun_out/c7_attn_t1.json > gpurun_out/c7_attn_t1.log 2>&1
echo "== attn turns=1 exit $?"; cut -c1-125 gpurun_out/c7_attn_t1.log | tail -n 4
$T 200 python tools/bench_kernels.py --only attn --quick --out gpurun_out/c7_attn_b.json > gpurun_out/c7_attn_b.log 2>&1
echo "== attn (repeat) exit $?"; cut -c1-125 gpurun_out/c7_attn_b.log | tail -n 4
for BN in 256 128; do
  LV_GEMM_BN=$BN $T 200 python tools/bench_kernels.py --only gemm --out gpurun_out/c7_gemm_bn$BN.json > gpurun_out/c7_gemm_bn$BN.log 2>&1
  echo "== gemm BN=$BN exit $?"; cut -c1-150 gpurun_out/c7_gemm_bn$BN.log | tail -n 14
done
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
$T 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 2 -c 1 -f -o gpurun_out/r2_attn16k_f32x2 python /tmp/attn16k.py > gpurun_out/ncu_attn_f32x2.log 2>&1
echo "== ncu attn 16k exit $?"; tail -1 gpurun_out/ncu_attn_f32x2.log
$T 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 5 -c 1 -f -o gpurun_out/r2_attn_vit python /tmp/attn16k.py > gpurun_out/ncu_attn_vit.log 2>&1
echo "== ncu attn vit exit $?"; tail -1 gpurun_out/ncu_attn_vit.log
$T 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c7_bench_n1.json 2> gpurun_out/c7_bench_n1.err
echo "== bench exit $?"; tail -3 gpurun_out/c7_bench_n1.err; cut -c1-300 gpurun_out/c7_bench_n1.json
