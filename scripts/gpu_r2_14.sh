#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gen.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gen.log
SVCB_DUMP_KERNELS=1 timeout 600 python bench.py --no-subconfigs > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-300; tail -3 gpurun_out/bench.err
