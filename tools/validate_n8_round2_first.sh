#!/bin/bash
# 8 GPUs, tightly bounded: bench at 18K (the SCALE configuration) first, then context-parallel parity at 4 / 8 ranks,
# 128K on 8 and 4 GPUs, config 5 (DP2 x CP4) attention step.
mkdir -p gpurun_out
T="timeout -k 5"
$T 300 python long-vita_b200/build.py > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$T 150 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/c5_bench_n8_18k.json 2> gpurun_out/c5_bench_n8_18k.err
echo "== bench N=8 18K exit $?"; grep -h "parity\|timed region" gpurun_out/c5_bench_n8_18k.err | head -3; cut -c1-250 gpurun_out/c5_bench_n8_18k.json
LV_CP_ORDER=0 $T 150 $TR --nproc-per-node 8 --master-port 29515 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/c5_bench_n8_18k_order0.json 2> gpurun_out/c5_bench_n8_18k_order0.err
echo "== bench N=8 18K, global (non-ring) visiting order exit $?"; cut -c1-250 gpurun_out/c5_bench_n8_18k_order0.json
$T 240 python -m pytest tests/test_gpu_cp.py -m gpu -q -x --timeout 100 --timeout-method=thread -k "4 or 8 or missing" -rf > gpurun_out/c5_test_cp.log 2>&1
echo "== cp tests (4, 8 ranks, missing peer) exit $?"; tail -n 8 gpurun_out/c5_test_cp.log
$T 200 $TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --steps 2 --warmup 3 --frames 512 > gpurun_out/c5_bench_n8_128k.json 2> gpurun_out/c5_bench_n8_128k.err
echo "== bench N=8 128K exit $?"; grep -h "parity\|timed region" gpurun_out/c5_bench_n8_128k.err | head -3; cut -c1-250 gpurun_out/c5_bench_n8_128k.json
$T 200 $TR --nproc-per-node 4 --master-port 29513 bench.py --gpus 4 --steps 2 --warmup 3 --frames 512 > gpurun_out/c5_bench_n4_128k.json 2> gpurun_out/c5_bench_n4_128k.err
echo "== bench N=4 128K (BASELINE config 3) exit $?"; grep -h "parity\|timed region" gpurun_out/c5_bench_n4_128k.err | head -3; cut -c1-250 gpurun_out/c5_bench_n4_128k.json
$T 150 $TR --nproc-per-node 8 --master-port 29514 tools/bench_train_attn.py --seq 131072 --cp 4 --iters 3 > gpurun_out/c5_train_attn_dp2cp4.json 2> gpurun_out/c5_train_attn_dp2cp4.err
echo "== config 5 attention fwd+bwd, DP2 x CP4 exit $?"; tail -2 gpurun_out/c5_train_attn_dp2cp4.err; cat gpurun_out/c5_train_attn_dp2cp4.json
