#!/bin/bash
# One-call GPU diagnostic: each test group runs in its own process under a timeout so that a hung
# kernel in one group does not hide the results of the others.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for grp in elementwise gemm attention; do
  timeout 600 python -m pytest tests/test_gpu_${grp}.py -m gpu -q --timeout 90 --timeout-method=thread > gpurun_out/test_${grp}.log 2>&1
  echo "== ${grp}: exit $?"; tail -n 40 gpurun_out/test_${grp}.log
done
