#!/bin/bash
# Round-end re-validation after the tcgen05 issue-loop change: full GPU tests, smoke(), default bench,
# reference arm, refreshed launch list + amp_conv_tc capture.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-200
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref exit $?"
BA="--steps 1 --warmup 1 --no-cpu-baseline --no-roofline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --print-units base --csv --log-file gpurun_out/r01d_launches.csv python bench.py $BA > gpurun_out/ncu_list.log 2>&1; echo "list exit $?"
cap() { name=$1; regex=$2; skip=$3; cnt=$4; timeout 900 ncu --set full --clock-control none -k regex:$regex -s $skip -c $cnt -o /tmp/$name -f python bench.py $BA > gpurun_out/ncu_$name.log 2>&1; echo "$name exit $?"; ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null; }
cap r01d_amp_conv_c40 amp_conv_tc_kernel 90 18
