#!/bin/bash
# tcgen05 bring-up: each risky piece in its own process under its own timeout.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/$name.log; tail -${TAILN:-25} gpurun_out/$name.log; }
run t_selftest python -m pytest tests/test_ops_gpu.py -q -m gpu -rA -s -p no:cacheprovider -k tcgen05
run t_ampconv python -m pytest tests/test_ops_gpu.py -q -m gpu -rA -s -p no:cacheprovider -k amp_conv_tc
run t_modes python -m pytest tests/test_parity_gpu.py -q -m gpu -rA -s -p no:cacheprovider -k "tensor_core or stage_taps"
TMO=600 run bench_fp32 python bench.py --steps 3 --warmup 3 --no-cpu-baseline
TMO=600 run bench_x3 python bench.py --steps 3 --warmup 3 --precision 3 --no-cpu-baseline
TMO=600 run bench_bf16 python bench.py --steps 3 --warmup 3 --precision 1 --no-cpu-baseline
for t in 4 8 16 32; do SVCB_CPU_THREADS=$t TAILN=2 TMO=200 run cpu_$t python bench.py --impl reference --steps 2 --warmup 1; done
