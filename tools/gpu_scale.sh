#!/bin/bash
# scaling run of bench.py on one 8-GPU box: N = 1, 2, 4, 8 (one JSON line per N in gpurun_out/scale_nN.json).  The headline
# config is weak-scaled (4096 envs per GPU); the configs sub-record carries the strong-scaled ones (65536 / 32768 TOTAL envs).
mkdir -p gpurun_out
STEPS=${STEPS:-100}
timeout 300 python bench.py --gpus 1 --steps $STEPS --warmup 10 --no-cpu > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
for N in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
    bench.py --gpus $N --steps $STEPS --warmup 10 --no-cpu > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
done
for N in 1 2 4 8; do tail -1 gpurun_out/scale_n$N.json | cut -c1-300; done
