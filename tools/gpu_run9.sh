#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 5"
KRE='regex:attn_|gemm_bf16|rmsnorm|layernorm|rope_|swiglu|ls_residual|pixel_shuffle|row_copy|im2col|add_cls|bias_gelu'
$T 120 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
$T 200 python -m pytest tests/test_gpu_attention.py tests/test_gpu_gemm.py tests/test_gpu_model.py tests/test_gpu_surfaces.py -m gpu -q --timeout 60 --timeout-method=thread > gpurun_out/test_a.log 2>&1
echo "== attention: exit $?"; tail -n 4 gpurun_out/test_a.log
$T 200 python tools/bench_kernels.py --only attn --quick --out gpurun_out/kernels_attn_sched.json > gpurun_out/bench_attn_sched.log 2>&1
echo "== bench attn exit $?"; cat gpurun_out/bench_attn_sched.log | cut -c1-170
$T 300 ncu --set full --clock-control none --import-source on -k "$KRE" -s 13 -c 12 -o gpurun_out/r1_all_kernels -f python tools/ncu_all.py > gpurun_out/ncu_all.log 2>&1
echo "== ncu all exit $?"; tail -2 gpurun_out/ncu_all.log
$T 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -s 1980 -c 660 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
echo "== launch list exit $?"; wc -l gpurun_out/launches.csv
$T 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "== bench exit $?"; cat gpurun_out/bench_n1.json | cut -c1-400
