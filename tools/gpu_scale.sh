#!/bin/bash
# weak-scaling run of bench.py on one 8-GPU box: N = 1, 2, 4, 8 (one line per N in gpurun_out/scale_nN.json)
mkdir -p gpurun_out
timeout 200 python bench.py --gpus 1 --steps 300 --warmup 20 --no-cpu > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
for N in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
    bench.py --gpus $N --steps 300 --warmup 20 --no-cpu > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
done
for N in 1 2 4 8; do tail -1 gpurun_out/scale_n$N.json | cut -c1-400; done
