#!/bin/bash
# 2 GPUs: context-parallel tests at 2 ranks (forward, backward, sharded decode, missing peer) + bench N=2 with the fp32 parity probe.
mkdir -p gpurun_out
T="timeout -k 5"
$T 300 python long-vita_b200/build.py > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
$T 240 python -m pytest tests/test_gpu_cp.py -m gpu -q -x --timeout 100 --timeout-method=thread -k "2 or missing" -rf > gpurun_out/c9_test_cp.log 2>&1
echo "== cp tests (2 ranks) exit $?"; tail -n 6 gpurun_out/c9_test_cp.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$T 150 $TR --nproc-per-node 2 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c9_bench_n2.json 2> gpurun_out/c9_bench_n2.err
echo "== bench N=2 exit $?"; grep -h "parity\|timed region" gpurun_out/c9_bench_n2.err | head -2; cut -c1-250 gpurun_out/c9_bench_n2.json
LV_CP_ORDER=1 $T 150 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/c9_bench_n2_ring.json 2> gpurun_out/c9_bench_n2_ring.err
echo "== bench N=2 ring order exit $?"; grep -h "parity\|timed region" gpurun_out/c9_bench_n2_ring.err | head -2; cut -c1-250 gpurun_out/c9_bench_n2_ring.json
