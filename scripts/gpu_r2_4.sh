#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:amp_s2d_link -c 6 -o gpurun_out/r02b_s2d -f python scripts/prof_s2d.py > gpurun_out/ncu_s2d.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/ncu_s2d.log
