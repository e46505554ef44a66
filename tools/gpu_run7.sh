#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 5"
$T 120 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
LV_ATTN_VERSION=3 $T 200 python -m pytest tests/test_gpu_attention.py -m gpu -q --timeout 60 --timeout-method=thread > gpurun_out/test_qh.log 2>&1
echo "== attention(QH): exit $?"; tail -n 4 gpurun_out/test_qh.log
LV_ATTN_VERSION=3 $T 200 python tools/bench_kernels.py --only attn --quick --out gpurun_out/kernels_attn_qh.json > gpurun_out/bench_attn_qh.log 2>&1
echo "== bench attn QH exit $?"; cat gpurun_out/bench_attn_qh.log | cut -c1-170
$T 300 python -m pytest tests/test_gpu_surfaces.py tests/test_gpu_gemm.py -m gpu -q --timeout 60 --timeout-method=thread > gpurun_out/test_c.log 2>&1
echo "== surfaces+gemm: exit $?"; tail -n 6 gpurun_out/test_c.log | cut -c1-300
# launch list (every launch with its device time) + full sections for one launch of each kernel
$T 300 ncu --set full --clock-control none --import-source on -s 13 -c 12 -o gpurun_out/r1_all_kernels -f python tools/ncu_all.py > gpurun_out/ncu_all.log 2>&1
echo "== ncu exit $?"; tail -2 gpurun_out/ncu_all.log
$T 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --layers 4 > gpurun_out/launches_bench.log 2>&1
echo "== launch list exit $?"; wc -l gpurun_out/launches.csv
