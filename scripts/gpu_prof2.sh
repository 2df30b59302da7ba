#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -s -k "tensor_core" 2>&1 | grep -E "wave max-abs|passed|failed|Error" | tail -8
BA="--batch 4 --steps 1 --warmup 1 --no-cpu-baseline --no-roofline --precision 3"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:amp_block_fused -s 4 -c 2 -o gpurun_out/r01_amp_block -f python bench.py $BA > gpurun_out/ncu_ab.log 2>&1; echo "ab exit $?"
