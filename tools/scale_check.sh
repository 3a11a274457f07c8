#!/bin/bash
# On a box with 8 GPUs: the bench line at N = 2, 4, 8 in both multi-GPU modes (frame pairs at 64x1800, row bands of one
# pair at 128x2048; BAND_ONLY=1: only the latter) -> gpurun_out/scale/.  Every run is bounded.
mkdir -p gpurun_out/scale
for n in ${NS:-2 4 8}; do
  [ -n "${BAND_ONLY:-}" ] || timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus $n --steps 200 --warmup 10 --no-cpu > gpurun_out/scale/pairs_n$n.json 2> gpurun_out/scale/pairs_n$n.err
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29552 \
    bench.py --gpus $n --partition rowband --hw 128x2048 --steps 50 --warmup 5 > gpurun_out/scale/band128_n$n.json 2> gpurun_out/scale/band128_n$n.err
  python - <<PY
import json
for f in ("pairs_n$n", "band128_n$n"):
    try:
        d = json.loads([l for l in open("gpurun_out/scale/%s.json" % f) if l.startswith("{")][-1])
        c = d["config"]
        print(f, "value %.0f e2e %.0f ms/step %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]),
              c.get("single_gpu_ms_per_forward"), c.get("pose_max_abs_diff_vs_single_gpu"), c["parallelism"][:90])
    except Exception as e:
        print(f, "FAILED", e)
PY
done
