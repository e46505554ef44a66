#!/bin/bash
# 1 GPU: half hand-off A/B + parity on the watchdog build, two-Q-buffer ViT A/B, whole GPU suite, bench (with CPU baseline),
# 128K single-GPU prefill (scaling denominator), ncu DRAM traffic of the GEMMs inside the bench step.
mkdir -p gpurun_out
T="timeout -k 5"
LV_WATCHDOG=1 $T 300 python long-vita_b200/build.py > gpurun_out/build_wd.log 2>&1 || { tail -5 gpurun_out/build_wd.log; exit 1; }
$T 200 python -m pytest tests/test_gpu_attention.py -m gpu -q -x --timeout 60 --timeout-method=thread > gpurun_out/c10_test_wd.log 2>&1
RC=$?; echo "== attention parity, half hand-off, watchdog build exit $RC"; grep -h "lv watchdog" gpurun_out/c10_test_wd.log | sort | uniq -c | head -5; tail -n 3 gpurun_out/c10_test_wd.log
LV_WATCHDOG=0 $T 300 python long-vita_b200/build.py > gpurun_out/build_release.log 2>&1 || { tail -5 gpurun_out/build_release.log; exit 1; }
if [ $RC -ne 0 ]; then export LV_ATTN_HALF=0; echo "!! half hand-off failed parity: continuing with LV_ATTN_HALF=0"; fi
$T 500 python -m pytest tests -m gpu -q -x --timeout 150 --timeout-method=thread --deselect tests/test_gpu_cp.py -rf > gpurun_out/c10_test_all.log 2>&1
echo "== all 1-GPU tests exit $?"; tail -n 5 gpurun_out/c10_test_all.log
for H in 1 0; do
  LV_ATTN_HALF=$H $T 120 python tools/bench_kernels.py --only attn --out gpurun_out/c10_attn_h$H.json > gpurun_out/c10_attn_h$H.log 2>&1
  echo "== attn half=$H exit $?"; cut -c1-125 gpurun_out/c10_attn_h$H.log | tail -n 5
done
$T 100 python tools/bench_bwd.py > gpurun_out/c10_bwd.log 2>&1
echo "== bwd exit $?"; cut -c1-220 gpurun_out/c10_bwd.log | tail -n 2
$T 300 python bench.py --steps 5 --warmup 3 > gpurun_out/c10_bench_n1.json 2> gpurun_out/c10_bench_n1.err
echo "== bench exit $?"; tail -3 gpurun_out/c10_bench_n1.err; cut -c1-200 gpurun_out/c10_bench_n1.json
$T 300 python bench.py --steps 2 --warmup 3 --frames 512 --no-cpu-baseline --no-attn-probe > gpurun_out/c10_bench_n1_128k.json 2> gpurun_out/c10_bench_n1_128k.err
echo "== bench 128K N=1 exit $?"; tail -2 gpurun_out/c10_bench_n1_128k.err; cut -c1-200 gpurun_out/c10_bench_n1_128k.json
$T 300 ncu --set full --clock-control none -k regex:gemm_bf16 -s 103 -c 8 -f -o gpurun_out/r2_bench_gemm_final python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-attn-probe > gpurun_out/ncu_bench_gemm_final.log 2>&1
echo "== ncu bench gemm exit $?"; tail -1 gpurun_out/ncu_bench_gemm_final.log
LV_WATCHDOG=0 LV_EXTRA_DEFINES="-DLV_ATTN_QBUF64=2" $T 300 python long-vita_b200/build.py > gpurun_out/build_qbuf2.log 2>&1 && {
  $T 120 python -m pytest tests/test_gpu_attention.py -m gpu -q -x -k "many_items or 1025 or layouts" --timeout 60 --timeout-method=thread > gpurun_out/c10_test_qbuf2.log 2>&1
  echo "== two-Q-buffer parity exit $?"; tail -n 2 gpurun_out/c10_test_qbuf2.log
  $T 100 python tools/bench_kernels.py --only attn --quick --out gpurun_out/c10_attn_qbuf2.json > gpurun_out/c10_attn_qbuf2.log 2>&1
  echo "== attn two-Q-buffer exit $?"; cut -c1-125 gpurun_out/c10_attn_qbuf2.log | tail -n 4
}
