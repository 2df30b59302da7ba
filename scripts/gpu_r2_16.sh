#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hubert_gpu.py -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_hubert.log 2>&1; echo "pytest exit $?"; grep -E "rel-l2|max-abs|passed|failed|Error|error|assert" gpurun_out/pytest_hubert.log | tail -30
