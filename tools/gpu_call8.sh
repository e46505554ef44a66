#!/bin/bash
# 1 GPU, every step tightly bounded: (a) where does the two-Q-buffer path of head_dim 64 hang (watchdog build of that
# variant, one test), (b) default build: parity, polynomial-exp share A/B, backward, GEMM, bench.
mkdir -p gpurun_out
T="timeout -k 5"
LV_WATCHDOG=1 LV_EXTRA_DEFINES="-DLV_ATTN_QBUF64=2" $T 300 python long-vita_b200/build.py > gpurun_out/build_qbuf2.log 2>&1 || { tail -5 gpurun_out/build_qbuf2.log; exit 1; }
$T 90 python -m pytest tests/test_gpu_attention.py -m gpu -q -x -k "many_items" --timeout 60 --timeout-method=thread > gpurun_out/c8_qbuf2.log 2>&1
echo "== two-Q-buffer variant, many items (watchdog) exit $?"; grep -h "lv watchdog" gpurun_out/c8_qbuf2.log | sort | uniq -c | head -12; tail -n 3 gpurun_out/c8_qbuf2.log
LV_WATCHDOG=1 LV_EXTRA_DEFINES="" $T 300 python long-vita_b200/build.py > gpurun_out/build_wd.log 2>&1 || { tail -5 gpurun_out/build_wd.log; exit 1; }
$T 90 python -m pytest tests/test_gpu_attention.py -m gpu -q -x -k "many_items or 1025" --timeout 60 --timeout-method=thread > gpurun_out/c8_qbuf1.log 2>&1
echo "== one-Q-buffer (default) variant, many items (watchdog) exit $?"; grep -h "lv watchdog" gpurun_out/c8_qbuf1.log | sort | uniq -c | head -12; tail -n 3 gpurun_out/c8_qbuf1.log
LV_WATCHDOG=0 LV_EXTRA_DEFINES="" $T 300 python long-vita_b200/build.py > gpurun_out/build_release.log 2>&1 || { tail -5 gpurun_out/build_release.log; exit 1; }
$T 240 python -m pytest tests/test_gpu_attention.py tests/test_gpu_attention_bwd.py tests/test_gpu_gemm.py tests/test_gpu_surfaces.py -m gpu -q -x --timeout 60 --timeout-method=thread > gpurun_out/c8_test_a.log 2>&1
RC=$?; echo "== parity (default build) exit $RC"; tail -n 4 gpurun_out/c8_test_a.log
[ $RC -ne 0 ] && exit 1
LV_ATTN_POLY=4 $T 120 python -m pytest tests/test_gpu_attention.py -m gpu -q -x --timeout 60 --timeout-method=thread > gpurun_out/c8_test_p4.log 2>&1
echo "== attention parity poly=4 exit $?"; tail -n 3 gpurun_out/c8_test_p4.log
$T 400 python -m pytest tests -m gpu -q -x --timeout 120 --timeout-method=thread --deselect tests/test_gpu_cp.py --deselect tests/test_gpu_attention_long.py -rf > gpurun_out/c8_test_all.log 2>&1
echo "== all other 1-GPU tests exit $?"; tail -n 5 gpurun_out/c8_test_all.log
for P in 0 2 3 4; do
  LV_ATTN_POLY=$P $T 100 python tools/bench_kernels.py --only attn --quick --out gpurun_out/c8_attn_p$P.json > gpurun_out/c8_attn_p$P.log 2>&1
  echo "== attn poly=$P exit $?"; cut -c1-125 gpurun_out/c8_attn_p$P.log | tail -n 4
done
$T 100 python tools/bench_bwd.py > gpurun_out/c8_bwd.log 2>&1
echo "== bwd exit $?"; cut -c1-220 gpurun_out/c8_bwd.log | tail -n 2
for GM in 0 16; do
  LV_GEMM_GM=$GM $T 100 python tools/bench_kernels.py --only gemm --quick --out gpurun_out/c8_gemm_gm$GM.json > gpurun_out/c8_gemm_gm$GM.log 2>&1
  echo "== gemm GM=$GM (0 = adaptive) exit $?"; cut -c1-150 gpurun_out/c8_gemm_gm$GM.log | head -n 4
done
for P in 0 4; do
  LV_ATTN_POLY=$P $T 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c8_bench_p$P.json 2> gpurun_out/c8_bench_p$P.err
  echo "== bench poly=$P exit $?"; tail -2 gpurun_out/c8_bench_p$P.err; cut -c1-200 gpurun_out/c8_bench_p$P.json
done
