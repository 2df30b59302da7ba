#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -s -k "rel_attention" > gpurun_out/pytest_ra.log 2>&1; echo "pytest rel_attention exit $?"; grep -E "rel_attention T|passed|failed|rror" gpurun_out/pytest_ra.log | head -20
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -s -k "tensor_core or stage_taps or stage_entry" > gpurun_out/pytest_par.log 2>&1; echo "pytest parity exit $?"; grep -E "wave max-abs|passed|failed" gpurun_out/pytest_par.log | head -20
SVCB_DUMP_KERNELS=1 timeout 600 python bench.py --no-subconfigs > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-300; tail -3 gpurun_out/bench.err
