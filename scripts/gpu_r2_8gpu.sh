#!/bin/bash
# 8-GPU run of BASELINE configs[4] (batch CLI engine, 1000 synthetic 10 s utterances)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "gpus visible: $N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/bench_config5.py --utterances 1000 > gpurun_out/config5_${N}gpu.json 2> gpurun_out/config5_${N}gpu.err; echo "config5 x$N exit $?"; grep '^{' gpurun_out/config5_${N}gpu.json | cut -c1-500; tail -2 gpurun_out/config5_${N}gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --no-subconfigs > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err; echo "bench x$N exit $?"; tail -1 gpurun_out/bench_${N}gpu.log | cut -c1-400; tail -2 gpurun_out/bench_${N}gpu.err
