#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 5"
$T 120 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
$T 300 python -m pytest tests/test_gpu_attention_bwd.py -m gpu -q --timeout 60 --timeout-method=thread > gpurun_out/test_bwd.log 2>&1
echo "== attention bwd: exit $?"; tail -n 14 gpurun_out/test_bwd.log | cut -c1-300
$T 200 python tools/bench_bwd.py > gpurun_out/bench_bwd.log 2>&1; echo "== bench bwd exit $?"; tail -6 gpurun_out/bench_bwd.log | cut -c1-300
