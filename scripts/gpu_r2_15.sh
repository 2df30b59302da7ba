#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
BA="--steps 1 --warmup 1 --no-cpu-baseline --no-roofline --no-subconfigs"
timeout 900 ncu --set full --clock-control none -k regex:conv_tc_kernel -s 30 -c 14 -o /tmp/r02g_conv_tc -f python bench.py $BA > gpurun_out/ncu_conv_tc.log 2>&1; echo "ncu exit $?"
ncu -i /tmp/r02g_conv_tc.ncu-rep --page raw --csv > gpurun_out/r02g_conv_tc_raw.csv 2>/dev/null
tail -2 gpurun_out/ncu_conv_tc.log
