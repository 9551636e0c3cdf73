#!/bin/bash
# ncu full capture of one settled step at a chosen env count: N=$1
set -x
export MJB_BENCH_ENVS=${1:-16384} MJB_BENCH_SETTLE=300
python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_env -s 304 -c 1 -o gpurun_out/prof -f \
    python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/plain2.log | cut -c1-300
