#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 5"
$T 120 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
$T 300 python -m pytest tests/test_gpu_attention_bwd.py -m gpu -q --timeout 60 --timeout-method=thread > gpurun_out/test_bwd.log 2>&1
echo "== attention bwd: exit $?"; tail -n 30 gpurun_out/test_bwd.log | cut -c1-300
$T 400 python bench.py --frames 512 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_128k.json 2> gpurun_out/bench_n1_128k.err
echo "== bench 128K exit $?"; tail -2 gpurun_out/bench_n1_128k.err; cat gpurun_out/bench_n1_128k.json | cut -c1-1800
