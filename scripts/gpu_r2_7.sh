#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -s -k "batch_engine" > gpurun_out/pytest_eng.log 2>&1; echo "pytest engine exit $?"; tail -3 gpurun_out/pytest_eng.log
timeout 900 python scripts/bench_config5.py --utterances 1000 > gpurun_out/config5_1gpu.json 2> gpurun_out/config5_1gpu.err; echo "config5 exit $?"; tail -1 gpurun_out/config5_1gpu.json | cut -c1-600; tail -3 gpurun_out/config5_1gpu.err
