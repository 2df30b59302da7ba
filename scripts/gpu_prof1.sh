#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
BA="--batch 8 --steps 1 --warmup 1 --no-cpu-baseline --no-roofline --precision 3"
# launch list (every kernel, device time) of one step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches.csv python bench.py $BA > gpurun_out/ncu_list.log 2>&1; echo "list exit $?"
# full captures: a few launches of the three top kernels (stage-2/3 shapes: skip early launches)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:snake_pack -s 60 -c 2 -o gpurun_out/r01_snake_pack -f python bench.py $BA > gpurun_out/ncu_sp.log 2>&1; echo "sp exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:amp_conv_tc -s 10 -c 1 -o gpurun_out/r01_amp_conv_s0 -f python bench.py $BA > gpurun_out/ncu_ac0.log 2>&1; echo "ac0 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:amp_conv_tc -s 100 -c 1 -o gpurun_out/r01_amp_conv_s2 -f python bench.py $BA > gpurun_out/ncu_ac2.log 2>&1; echo "ac2 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:amp_conv_tc -s 200 -c 1 -o gpurun_out/r01_amp_conv_s4 -f python bench.py $BA > gpurun_out/ncu_ac4.log 2>&1; echo "ac4 exit $?"
ls -la gpurun_out/*.ncu-rep
