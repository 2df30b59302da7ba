#!/bin/bash
# Round-1 profile set: launch list of one step + full captures of the top kernels (precision 3).
# Reports stay on the box; only CSV extracts come back (gpurun_out is capped at 64 MiB).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_whisper_gpu.py -q -m gpu -p no:cacheprovider -s 2>&1 | grep -E "whisper D|passed|failed|Error" | tail -8
python bench.py --workload whisper --steps 3 --warmup 2 > gpurun_out/bench_whisper.log 2>&1; tail -1 gpurun_out/bench_whisper.log | cut -c1-900
BA="--batch 8 --steps 1 --warmup 1 --no-cpu-baseline --no-roofline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --print-units base --csv --log-file gpurun_out/r01b_launches.csv python bench.py $BA > gpurun_out/ncu_list.log 2>&1; echo "list exit $?"
cap() { name=$1; regex=$2; skip=$3; cnt=$4; shift 4; timeout 900 ncu --set full --clock-control none -k regex:$regex -s $skip -c $cnt -o /tmp/$name -f "$@" > gpurun_out/ncu_$name.log 2>&1; echo "$name exit $?"; ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null; }
cap r01b_amp_block amp_block_fused 6 6 python bench.py $BA
cap r01b_amp_conv amp_conv_tc 162 9 python bench.py $BA
cap r01b_amp_conv_c40 amp_conv_tc 234 3 python bench.py $BA
cap r01b_snake_pack snake_pack 234 2 python bench.py $BA
cap r01b_conv_tc "conv_tc_kernel" 204 12 python bench.py $BA
cap r01b_whisper_gemm gemm_tc 20 4 python bench.py --workload whisper --steps 1 --warmup 1
ls -la gpurun_out/ | tail -12
