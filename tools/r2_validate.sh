#!/bin/bash
# First GPU call of round 2: validate the r2-prep kernel changes (warp-uniform MMA issue) on ONE GPU.
#   gpurun --timeout 1800 -- 'bash tools/r2_validate.sh'
# Every step runs under its own timeout; a hang in a new kernel costs at most that step.
mkdir -p gpurun_out
T="timeout -k 5"
# Parity runs on the WATCHDOG build (csrc/ptx.cuh: a wait that spins ~1 s prints which barrier of which warp is
# stuck and traps - a protocol bug costs one error line, not a hung GPU); ship it prebuilt
# (`LV_WATCHDOG=1 python long-vita_b200/build.py` before gpurun) or let this line build it (~40 s of box time).
LV_WATCHDOG=1 $T 300 python long-vita_b200/build.py > gpurun_out/build_watchdog.log 2>&1 || { tail -20 gpurun_out/build_watchdog.log; exit 1; }
# 1. smallest possible smoke of each touched kernel first (fast fail)
$T 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "== smoke exit $?"; tail -n 3 gpurun_out/smoke.log
# 2. parity suites of the touched kernels
$T 300 python -m pytest tests/test_gpu_attention.py tests/test_gpu_gemm.py tests/test_gpu_attention_bwd.py -m gpu -q -x --timeout 90 --timeout-method=thread > gpurun_out/test_kernels.log 2>&1
echo "== kernel parity exit $?"; tail -n 6 gpurun_out/test_kernels.log
$T 600 python -m pytest tests -m gpu -q --timeout 120 --timeout-method=thread --deselect tests/test_gpu_cp.py -rf > gpurun_out/test_all.log 2>&1   # no -x: see every failing feature in one call
echo "== all 1-GPU tests exit $?"; tail -n 25 gpurun_out/test_all.log
# 2b. the double-buffered-S kernel (v2) shares the new issue path: parity before it is timed
LV_ATTN_VERSION=2 $T 200 python -m pytest tests/test_gpu_attention.py -m gpu -q -x --timeout 90 --timeout-method=thread > gpurun_out/test_attn_v2.log 2>&1
echo "== attention parity v2: exit $?"; tail -n 4 gpurun_out/test_attn_v2.log
# 2c. the pipelined backward kernel (LV_BWD_VERSION=2): parity on the watchdog build
LV_BWD_VERSION=2 $T 240 python -m pytest tests/test_gpu_attention_bwd.py -m gpu -q -x --timeout 90 --timeout-method=thread > gpurun_out/test_bwd_v2.log 2>&1
echo "== backward v2 parity: exit $?"; tail -n 6 gpurun_out/test_bwd_v2.log
# ---- timing runs use the release build (no printf / trap code in the wait loops) ----
LV_WATCHDOG=0 $T 300 python long-vita_b200/build.py > gpurun_out/build_release.log 2>&1 || { tail -20 gpurun_out/build_release.log; exit 1; }
# 3. speed: attention v1 / v2 / v3, GEMM, backward
for V in 1 2 3; do
  LV_ATTN_VERSION=$V $T 200 python tools/bench_kernels.py --only attn --quick --out gpurun_out/r2_attn_v$V.json > gpurun_out/r2_attn_v$V.log 2>&1
  echo "== attn v$V exit $?"; cut -c1-170 gpurun_out/r2_attn_v$V.log | tail -n 8
done
# 3b. polynomial exp2 on the FMA pipe for every 4th softmax element (LV_ATTN_POLY=1): parity, then speed
LV_ATTN_POLY=1 $T 200 python -m pytest tests/test_gpu_attention.py -m gpu -q -x --timeout 90 --timeout-method=thread > gpurun_out/test_attn_poly.log 2>&1
echo "== attention parity with poly exp2: exit $?"; tail -n 4 gpurun_out/test_attn_poly.log
for V in 1 2; do
  LV_ATTN_POLY=1 LV_ATTN_VERSION=$V $T 200 python tools/bench_kernels.py --only attn --quick --out gpurun_out/r2_attn_poly_v$V.json > gpurun_out/r2_attn_poly_v$V.log 2>&1
  echo "== attn poly v$V exit $?"; cut -c1-170 gpurun_out/r2_attn_poly_v$V.log | tail -n 8
done
$T 200 python tools/bench_kernels.py --only gemm --quick --out gpurun_out/r2_gemm.json > gpurun_out/r2_gemm.log 2>&1
echo "== gemm exit $?"; cut -c1-170 gpurun_out/r2_gemm.log | tail -n 8
LV_GEMV=0 $T 200 python tools/bench_kernels.py --only gemm --quick --out gpurun_out/r2_gemm_nogemv.json > gpurun_out/r2_gemm_nogemv.log 2>&1
echo "== gemm (LM head through the tensor-core kernel, LV_GEMV=0) exit $?"; grep -i "lm\|152064" gpurun_out/r2_gemm_nogemv.log | cut -c1-170 | tail -n 3
$T 200 python tools/bench_bwd.py > gpurun_out/r2_bwd.log 2>&1
echo "== bwd exit $?"; tail -n 6 gpurun_out/r2_bwd.log
LV_BWD_VERSION=2 $T 200 python tools/bench_bwd.py > gpurun_out/r2_bwd_v2.log 2>&1
echo "== bwd v2 exit $?"; tail -n 6 gpurun_out/r2_bwd_v2.log
LV_GEMM_GM=32 $T 200 python tools/bench_kernels.py --only gemm --quick --out gpurun_out/r2_gemm_gm32.json > gpurun_out/r2_gemm_gm32.log 2>&1
echo "== gemm with 32-M-block rasterisation groups (LV_GEMM_GM=32) exit $?"; cut -c1-170 gpurun_out/r2_gemm_gm32.log | tail -n 8
# 3c. work-item order A/B at the model's own shape: serpentine (default) vs plain round-robin, in the bench step
LV_ATTN_SCHED=0 $T 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1_sched0.json 2> gpurun_out/r2_bench_n1_sched0.err
echo "== bench with LV_ATTN_SCHED=0 exit $?"; cut -c1-300 gpurun_out/r2_bench_n1_sched0.json
# 4. the headline
$T 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "== bench exit $?"; cut -c1-600 gpurun_out/r2_bench_n1.json
# 5. K/V-cache decode (8f-2): ms per token at the 18K context next to the re-prefill the reference does
$T 300 python tools/bench_decode.py --tokens 16 > gpurun_out/r2_decode.json 2> gpurun_out/r2_decode.err
echo "== decode exit $?"; tail -2 gpurun_out/r2_decode.err; cut -c1-600 gpurun_out/r2_decode.json
