#!/bin/bash
# 8-GPU confirmation: CP parity test (2/4/8 ranks), then the prefill bench at 18K / 128K / 1M tokens.
mkdir -p gpurun_out
T="timeout -k 5"
NG=$(nvidia-smi -L | wc -l)
$T 120 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
$T 300 python -m pytest tests/test_gpu_cp.py -m gpu -q --timeout 200 --timeout-method=thread > gpurun_out/test_cp_n${NG}.log 2>&1
echo "== cp test exit $?"; tail -n 5 gpurun_out/test_cp_n${NG}.log | cut -c1-300
run() {  # name, timeout, bench args...
  local name=$1; local to=$2; shift 2
  $T $to python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NG "$@" > gpurun_out/bench_n${NG}_${name}.json 2> gpurun_out/bench_n${NG}_${name}.err
  echo "== bench $name exit $?"; grep "bench +" gpurun_out/bench_n${NG}_${name}.err | tail -2; cat gpurun_out/bench_n${NG}_${name}.json | cut -c1-900
}
run 16k 200 --steps 3 --warmup 3
run 128k 300 --frames 512 --steps 2 --warmup 3
run 1m 600 --frames 4096 --steps 1 --warmup 1 --long-run
