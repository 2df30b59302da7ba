#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python scripts/trace_s2d.py > gpurun_out/trace_s2d.txt 2>&1; echo "trace exit $?"
