#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hubert_gpu.py tests/test_whisper_gpu.py -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_hubert.log 2>&1; echo "pytest exit $?"; grep -E "hubert|features|projected|embedded|layer0|encoded|units|pred_vec|passed|failed|Error|assert" gpurun_out/pytest_hubert.log | tail -30
SVCB_DUMP_KERNELS=1 timeout 600 python bench.py --workload hubert > gpurun_out/bench_hubert.log 2> gpurun_out/bench_hubert.err; echo "bench exit $?"; tail -1 gpurun_out/bench_hubert.log | cut -c1-1500; tail -3 gpurun_out/bench_hubert.err
