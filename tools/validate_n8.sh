#!/bin/bash
# 8 GPUs, final lines of the round: bench 18K and 128K over 8 ranks (in-run fp32 parity probe), the exchange kernel next to
# zigzag_ring_flash_attn_func at 8 ranks, bench 18K over 4 ranks.
mkdir -p gpurun_out
T="timeout -k 5"
$T 300 python long-vita_b200/build.py > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$T 120 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/c12_bench_n8_18k.json 2> gpurun_out/c12_bench_n8_18k.err
echo "== bench N=8 18K exit $?"; grep -h "parity\|timed region" gpurun_out/c12_bench_n8_18k.err | head -2 | cut -c1-250; cut -c1-200 gpurun_out/c12_bench_n8_18k.json
$T 120 $TR --nproc-per-node 8 --master-port 29521 tools/bench_cp_compare.py --seq 18432 131072 --iters 3 > gpurun_out/c12_cp_compare_n8.json 2> gpurun_out/c12_cp_compare_n8.err
echo "== cp comparator N=8 exit $?"; cat gpurun_out/c12_cp_compare_n8.json
$T 150 $TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --steps 2 --warmup 3 --frames 512 > gpurun_out/c12_bench_n8_128k.json 2> gpurun_out/c12_bench_n8_128k.err
echo "== bench N=8 128K exit $?"; grep -h "parity\|timed region" gpurun_out/c12_bench_n8_128k.err | head -2 | cut -c1-250; cut -c1-200 gpurun_out/c12_bench_n8_128k.json
$T 100 $TR --nproc-per-node 4 --master-port 29513 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/c12_bench_n4_18k.json 2> gpurun_out/c12_bench_n4_18k.err
echo "== bench N=4 18K exit $?"; cut -c1-200 gpurun_out/c12_bench_n4_18k.json
