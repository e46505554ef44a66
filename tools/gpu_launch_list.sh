#!/bin/bash
# One timed bench.py step under ncu's duration-only pass (the launch list profiles/README.md cites).  Bounded: 170 s.
mkdir -p gpurun_out
K='regex:(attn_|gemm_bf16|gemv_bf16|rmsnorm|layernorm|rope_|swiglu|bias_gelu|ls_residual|pixel_shuffle|row_copy|im2col|add_cls|ce_acc|ce_grad|decode_merge|pre_)'
timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 1980 -c 660 --csv \
  --log-file gpurun_out/r2_launch_list.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-attn-probe \
  > gpurun_out/r2_launch_list_bench.log 2>&1
echo "rc=$? rows=$(grep -c '^"' gpurun_out/r2_launch_list.csv 2>/dev/null)"
