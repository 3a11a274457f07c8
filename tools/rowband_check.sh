#!/bin/bash
# 2-GPU check of the banded forward: bench.py --partition rowband at both geometries (needs gpurun --gpus 2)
set -x
mkdir -p gpurun_out/rowband
for hw in 64x1800 128x2048; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus ${NG:-2} --partition rowband --hw $hw --steps 50 --warmup 5 > gpurun_out/rowband/band_${hw}_n${NG:-2}.json 2> gpurun_out/rowband/band_${hw}_n${NG:-2}.err
  tail -c 1500 gpurun_out/rowband/band_${hw}_n${NG:-2}.json; tail -3 gpurun_out/rowband/band_${hw}_n${NG:-2}.err
done
