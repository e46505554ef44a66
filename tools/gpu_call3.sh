#!/bin/bash
# 1 GPU: long-length parity, unit micro-benchmarks, attention variants, ncu of the attention kernel, full bench line.
mkdir -p gpurun_out
T="timeout -k 5"
$T 300 python long-vita_b200/build.py > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
$T 60 ./tools/ubench > gpurun_out/ubench.log 2>&1; echo "== ubench exit $?"; cat gpurun_out/ubench.log
$T 600 python -m pytest tests/test_gpu_attention_long.py -m gpu -q --timeout 300 --timeout-method=thread -rf --durations=8 > gpurun_out/test_long.log 2>&1
echo "== long parity exit $?"; tail -n 25 gpurun_out/test_long.log
for cfg in "1 0" "3 0" "3 1"; do
  set -- $cfg
  LV_ATTN_VERSION=$1 LV_ATTN_POLY=$2 $T 200 python tools/bench_kernels.py --only attn --quick --out gpurun_out/c3_attn_v$1_p$2.json > gpurun_out/c3_attn_v$1_p$2.log 2>&1
  echo "== attn v$1 poly$2 exit $?"; cut -c1-120 gpurun_out/c3_attn_v$1_p$2.log | tail -n 4
done
for O in 0 1; do
  LV_ATTN_ORDER=$O $T 200 python tools/bench_kernels.py --only attn --out gpurun_out/c3_attn_order$O.json > gpurun_out/c3_attn_order$O.log 2>&1
  echo "== attn order=$O exit $?"; cut -c1-120 gpurun_out/c3_attn_order$O.log | tail -n 5
done
# ncu: the attention kernel at 16K (v1 and v3), full set with source
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
for V in 1 3; do
  LV_ATTN_VERSION=$V $T 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 2 -c 1 -f -o gpurun_out/r2_attn16k_v$V python /tmp/attn16k.py > gpurun_out/ncu_attn_v$V.log 2>&1
  echo "== ncu attn v$V exit $?"; tail -2 gpurun_out/ncu_attn_v$V.log
done
$T 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c3_bench_n1.json 2> gpurun_out/c3_bench_n1.err
echo "== bench exit $?"; tail -4 gpurun_out/c3_bench_n1.err; cut -c1-400 gpurun_out/c3_bench_n1.json
$T 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c3_bench_ref.json 2> gpurun_out/c3_bench_ref.err
echo "== reference arm exit $?"; tail -3 gpurun_out/c3_bench_ref.err; cut -c1-300 gpurun_out/c3_bench_ref.json
