#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_whisper_gpu.py -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_whisper.log 2>&1; echo "pytest exit $?"; grep -E "attention_tc|rel-l2|passed|failed|Error|error" gpurun_out/pytest_whisper.log | tail -30
SVCB_DUMP_KERNELS=1 timeout 600 python bench.py --workload whisper > gpurun_out/bench_whisper.log 2> gpurun_out/bench_whisper.err; echo "bench exit $?"; tail -1 gpurun_out/bench_whisper.log | cut -c1-600; tail -25 gpurun_out/bench_whisper.err
