#!/bin/bash
# Round 2, call 2: straight-line MMA probe, the revised large-alpha test, the new bench.py (sub-configs), reference arm.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 120 scripts/mma_probe p2 > gpurun_out/mma_probe_p2.txt 2>&1; echo "probe p2 exit $?"
timeout 60 scripts/mma_probe n > gpurun_out/mma_probe_n.txt 2>&1; echo "probe n exit $?"
timeout 60 scripts/mma_probe align > gpurun_out/mma_probe_align.txt 2>&1; echo "probe align exit $?"
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -s -k "large_snake or golden_full" > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu2.log
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref exit $?"; tail -1 gpurun_out/bench_ref.log | cut -c1-300
