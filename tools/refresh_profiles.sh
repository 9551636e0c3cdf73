#!/bin/bash
# after tools/gpu_round.sh: turn what came back in gpurun_out/ into the tracked summaries under profiles/
set -e
cd "$(dirname "$0")/.."
python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/r02_ncu_full_k_env_c2_4096.txt 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_lite.ncu-rep > profiles/r02_ncu_full_k_lite_65536.txt 2>/dev/null
python tools/ncu_lines.py gpurun_out/prof.ncu-rep > profiles/r02_ncu_lines_k_env_c2_4096.txt 2>/dev/null
python tools/make_counts.py C2=gpurun_out/prof.ncu-rep:4096 > /dev/null
cp gpurun_out/launches.csv profiles/r02_launches_c2_4096.csv
cp gpurun_out/bench.json profiles/r02_bench_n1.json
cp gpurun_out/r02_lite.json profiles/r02_lite.json
python -c "import json,bench; c=json.load(open('profiles/r02_counts.json')); print('counts sha', c['csrc_sha16'], 'sources sha', bench.csrc_sha16())"
