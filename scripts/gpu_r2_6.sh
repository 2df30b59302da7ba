#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -s -k "amp_s2d" > gpurun_out/pytest_s2d.log 2>&1; echo "pytest s2d exit $?"; grep -E "amp_s2d_link C|passed|failed|rror" gpurun_out/pytest_s2d.log | head -30
timeout 300 python scripts/trace_s2d.py > gpurun_out/trace_s2d.txt 2>&1; echo "trace exit $?"
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -s -k "tensor_core or chunk_loop" > gpurun_out/pytest_par.log 2>&1; echo "pytest parity exit $?"; grep -E "passed|failed" gpurun_out/pytest_par.log | head
SVCB_DUMP_KERNELS=1 timeout 600 python bench.py --no-subconfigs > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-300
