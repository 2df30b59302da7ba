#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/$name.log; tail -${TAILN:-25} gpurun_out/$name.log | cut -c1-400; }
run t_ampconv python -m pytest tests/test_ops_gpu.py -q -m gpu -rA -s -p no:cacheprovider -k amp_conv_tc
run t_convtc python -m pytest tests/test_ops_gpu.py -q -m gpu -rA -s -p no:cacheprovider -k "conv_tc and not amp"
run t_modes python -m pytest tests/test_parity_gpu.py -q -m gpu -rA -s -p no:cacheprovider -k "tensor_core"
TAILN=3 TMO=600 run bench_x3 python bench.py --steps 3 --warmup 3 --precision 3 --no-cpu-baseline
TAILN=3 TMO=600 run bench_bf16 python bench.py --steps 3 --warmup 3 --precision 1 --no-cpu-baseline
