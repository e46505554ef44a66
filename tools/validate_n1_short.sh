#!/bin/bash
# 1 GPU, last call of the round: backward after the packed interior path (parity + speed), whole suite, bench line.
mkdir -p gpurun_out
T="timeout -k 5"
LV_WATCHDOG=1 $T 300 python long-vita_b200/build.py > gpurun_out/build_wd.log 2>&1 || { tail -5 gpurun_out/build_wd.log; exit 1; }
$T 150 python -m pytest tests/test_gpu_attention_bwd.py tests/test_gpu_surfaces.py -m gpu -q -x --timeout 60 --timeout-method=thread > gpurun_out/c13_test_bwd.log 2>&1
RC=$?; echo "== backward parity (watchdog build) exit $RC"; grep -h "lv watchdog" gpurun_out/c13_test_bwd.log | sort | uniq -c | head -5; tail -n 3 gpurun_out/c13_test_bwd.log
LV_WATCHDOG=0 $T 300 python long-vita_b200/build.py > gpurun_out/build_release.log 2>&1 || { tail -5 gpurun_out/build_release.log; exit 1; }
$T 100 python tools/bench_bwd.py > gpurun_out/c13_bwd.log 2>&1
echo "== bwd exit $?"; cut -c1-220 gpurun_out/c13_bwd.log | tail -n 2
LV_BWD_VERSION=1 $T 100 python tools/bench_bwd.py > gpurun_out/c13_bwd_v1.log 2>&1
echo "== bwd v1 exit $?"; cut -c1-120 gpurun_out/c13_bwd_v1.log | tail -n 1
$T 500 python -m pytest tests -m gpu -q -x --timeout 150 --timeout-method=thread --deselect tests/test_gpu_cp.py -rf > gpurun_out/c13_test_all.log 2>&1
echo "== all 1-GPU tests exit $?"; tail -n 4 gpurun_out/c13_test_all.log
$T 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c13_bench_n1.json 2> gpurun_out/c13_bench_n1.err
echo "== bench exit $?"; tail -2 gpurun_out/c13_bench_n1.err; cut -c1-200 gpurun_out/c13_bench_n1.json
$T 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c13_smoke.log 2>&1
echo "== smoke exit $?"; tail -1 gpurun_out/c13_smoke.log
