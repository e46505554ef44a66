#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 5"
$T 120 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
$T 400 python -m pytest tests/test_gpu_cp.py -m gpu -q -x --timeout 180 --timeout-method=thread > gpurun_out/test_cp.log 2>&1
echo "== cp test exit $?"; tail -n 30 gpurun_out/test_cp.log
NG=$(nvidia-smi -L | wc -l)
$T 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 2 --warmup 3 > gpurun_out/bench_n${NG}.json 2> gpurun_out/bench_n${NG}.err
echo "== bench N=$NG exit $?"; tail -5 gpurun_out/bench_n${NG}.err; cat gpurun_out/bench_n${NG}.json
