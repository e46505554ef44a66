#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 5"
$T 120 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
$T 200 python -m pytest tests/test_gpu_attention.py -m gpu -q --timeout 60 --timeout-method=thread > gpurun_out/test_attention_v2.log 2>&1
echo "== attention v2: exit $?"; tail -n 12 gpurun_out/test_attention_v2.log
$T 200 python -m pytest tests/test_gpu_model.py tests/test_gpu_surfaces.py -m gpu -q --timeout 60 --timeout-method=thread > gpurun_out/test_model_v2.log 2>&1
echo "== model+surfaces v2: exit $?"; tail -n 6 gpurun_out/test_model_v2.log
$T 300 python tools/bench_kernels.py --only attn --out gpurun_out/kernels_attn_v2.json > gpurun_out/bench_attn_v2.log 2>&1
echo "== bench attn v2 exit $?"; cat gpurun_out/bench_attn_v2.log | cut -c1-400
LV_ATTN_VERSION=1 $T 200 python tools/bench_kernels.py --only attn --quick --out gpurun_out/kernels_attn_v1.json > gpurun_out/bench_attn_v1.log 2>&1
echo "== bench attn v1 exit $?"; cat gpurun_out/bench_attn_v1.log | cut -c1-200
$T 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_v2.json 2> gpurun_out/bench_n1_v2.err
echo "== bench exit $?"; tail -3 gpurun_out/bench_n1_v2.err; cat gpurun_out/bench_n1_v2.json | cut -c1-1500
