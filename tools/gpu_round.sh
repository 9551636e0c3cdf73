#!/bin/bash
# One GPU visit: gpu tests, smoke, bench (all configs), ncu launch list, ncu full capture of the step kernel and of the
# tile kernel of the literal step.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
python bench.py --steps 100 --warmup 10 ${BENCH_FLAGS} 2> gpurun_out/bench.err | tee gpurun_out/bench.json
python tools/bench_lite.py 2>&1 | tail -6 | tee gpurun_out/lite.log
if [ "${SKIP_NCU}" != "1" ]; then
export MJB_BENCH_SETTLE=300
python bench.py --steps 4 --warmup 3 --no-cpu --no-configs > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 290 -c 40 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu --no-configs > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_env -s 304 -c 1 -o gpurun_out/prof -f \
    python bench.py --steps 4 --warmup 3 --no-cpu --no-configs > gpurun_out/ncu_full.log 2>&1
python tools/lite_prof.py > gpurun_out/lite_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_lite -s 6 -c 1 -o gpurun_out/prof_lite -f \
    python tools/lite_prof.py > gpurun_out/ncu_lite.log 2>&1
fi
ls -la gpurun_out
