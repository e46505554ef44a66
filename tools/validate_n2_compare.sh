#!/bin/bash
# 2 GPUs: the fused exchange kernel next to zigzag_ring_flash_attn_func (flash-attn 2.8 + NCCL ring) on one layer.
mkdir -p gpurun_out
T="timeout -k 5"
$T 300 python long-vita_b200/build.py > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$T 200 $TR --nproc-per-node 2 --master-port 29521 tools/bench_cp_compare.py --seq 18432 131072 > gpurun_out/c11_cp_compare_n2.json 2> gpurun_out/c11_cp_compare_n2.err
echo "== cp comparator N=2 exit $?"; tail -3 gpurun_out/c11_cp_compare_n2.err | cut -c1-300; cat gpurun_out/c11_cp_compare_n2.json
