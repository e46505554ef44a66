#!/bin/bash
# Second GPU call of round 2 (2 GPUs): context-parallel forward (regression) and the new CP backward.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/r2_validate_cp.sh'
mkdir -p gpurun_out
T="timeout -k 5"
LV_WATCHDOG=1 $T 300 python long-vita_b200/build.py > gpurun_out/build_watchdog.log 2>&1 || { tail -20 gpurun_out/build_watchdog.log; exit 1; }
$T 500 python -m pytest tests/test_gpu_cp.py -m gpu -q -x --timeout 200 --timeout-method=thread > gpurun_out/test_cp.log 2>&1
echo "== cp tests exit $?"; tail -n 30 gpurun_out/test_cp.log
LV_WATCHDOG=0 $T 300 python long-vita_b200/build.py > gpurun_out/build_release.log 2>&1 || { tail -20 gpurun_out/build_release.log; exit 1; }
NG=$(nvidia-smi -L | wc -l)
$T 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 2 --warmup 3 > gpurun_out/r2_bench_n${NG}.json 2> gpurun_out/r2_bench_n${NG}.err
echo "== bench N=$NG exit $?"; tail -3 gpurun_out/r2_bench_n${NG}.err; cut -c1-600 gpurun_out/r2_bench_n${NG}.json
