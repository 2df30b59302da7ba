#!/bin/bash
# One gpurun call: GPU tests + smoke + short bench, logs into gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest gpu" ; timeout 900 python -m pytest tests -q -m gpu -rA -s -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -60 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log; tail -5 gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py --steps ${STEPS:-3} --warmup ${WARMUP:-3} ${BENCH_ARGS:-} > gpurun_out/bench.log 2>&1; echo "bench exit $?" | tee -a gpurun_out/bench.log; tail -3 gpurun_out/bench.log
