#!/bin/bash
# 8-GPU mosaic (configs[4]) with work-balanced row slabs; also N=4 and N=2 for the scaling curve.
set -u
mkdir -p gpurun_out/n8m
for n in 8 4; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29620+n)) \
    bench.py --gpus $n --config mosaic --steps 30 --warmup 3 > gpurun_out/n8m/mosaic_n$n.json 2> gpurun_out/n8m/mosaic_n$n.err
  echo "n=$n rc=$?"; tail -c 600 gpurun_out/n8m/mosaic_n$n.json
done
