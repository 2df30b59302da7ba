#!/bin/bash
# Round-end validation + profile set on one box: full GPU tests, smoke(), default bench (with the CPU
# baseline leg), reference arm, PPG-extractor bench, the ncu launch list of one bench step and
# --set full captures of the top kernels (reports stay on the box; CSV extracts come back).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref exit $?"; tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout 600 python bench.py --workload whisper --steps 3 --warmup 2 > gpurun_out/bench_whisper.log 2>&1; echo "whisper exit $?"; tail -1 gpurun_out/bench_whisper.log | cut -c1-300
BA="--steps 1 --warmup 1 --no-cpu-baseline --no-roofline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --print-units base --csv --log-file gpurun_out/r01c_launches.csv python bench.py $BA > gpurun_out/ncu_list.log 2>&1; echo "list exit $?"
cap() { name=$1; regex=$2; skip=$3; cnt=$4; timeout 900 ncu --set full --clock-control none -k regex:$regex -s $skip -c $cnt -o /tmp/$name -f python bench.py $BA > gpurun_out/ncu_$name.log 2>&1; echo "$name exit $?"; ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null; }
cap r01c_amp_block amp_block_fused 6 6
cap r01c_amp_conv_c40 amp_conv_tc_kernel 90 6
cap r01c_snake_pack snake_pack3 90 2
cap r01c_rel_attention rel_attention 6 1
ls -la gpurun_out/ | grep -E "r01c|bench|pytest|smoke"
