#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 5"
$T 120 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
$T 300 ncu --set full --clock-control none --import-source on -k regex:"lv::" -s 13 -c 12 -o gpurun_out/r1_all_kernels -f python tools/ncu_all.py > gpurun_out/ncu_all.log 2>&1
echo "== ncu all exit $?"; tail -2 gpurun_out/ncu_all.log
$T 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lv::" -s 2124 -c 708 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
echo "== launch list exit $?"; wc -l gpurun_out/launches.csv
$T 100 python -m pytest tests/test_gpu_gemm.py -m gpu -q --timeout 60 --timeout-method=thread -k masked > gpurun_out/test_d.log 2>&1
echo "== masked lm head: exit $?"; tail -n 3 gpurun_out/test_d.log | cut -c1-300
