#!/bin/bash
# Round 2, call 1: MMA issue-cost probe, the GPU test suite with the new parity tests, default bench.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_1_smi.txt 2>&1
timeout 120 scripts/mma_probe all > gpurun_out/mma_probe_all.txt 2>&1; echo "probe all exit $?"
timeout 60 scripts/mma_probe cg2 > gpurun_out/mma_probe_cg2.txt 2>&1; echo "probe cg2 exit $?"
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-300
