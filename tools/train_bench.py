"""Config 4 of BASELINE.json: full forward + backward + Adam step in training mode, B frame pairs per GPU.

    python tools/train_bench.py --batch 8 --steps 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/train_bench.py --batch 8 --steps 5

Data parallel over frame pairs (weak scaling): every rank steps its own B pairs, the gradients are averaged
with one flat NCCL all-reduce per step.  Timed on the device, max over ranks.  Prints one JSON line."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import elo_b200 as elo  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--hw", default="64x1800")
    ap.add_argument("--graph", type=int, default=1, help="1: the whole step replayed as one CUDA graph; 0: eager launches")
    a = ap.parse_args()
    H, W = (int(v) for v in a.hw.split("x"))
    npts = 150000 if H * W <= 150000 else 300000
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    tg = elo.train_graph
    tp = tg.TrainableParams(elo.params.init_params(0), dev)
    tr = tg.Trainer(tp, batch_size=a.batch * world, H_input=H, W_input=W, process_group=group, use_graph=bool(a.graph))
    pc, T = elo.synth.synth_batch(a.batch, H, W, npts, seed0=rank * a.batch)
    pc, T = pc.to(dev), T.to(dev)
    perms = elo.params.make_perms(rank)
    for _ in range(a.warmup):
        tr.step(pc, T, perms=perms)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.reset_peak_memory_stats()
    e0.record()
    for _ in range(a.steps):
        loss = tr.step(pc, T, perms=perms)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    if world > 1:
        ms = elo.dist.max_over_ranks(ms, device=dev)
    if rank == 0:
        print(json.dumps({"metric": "frame-pairs/s, training step (forward + backward + Adam), batch-stat BN",
                          "value": a.batch * world / (ms * 1e-3), "unit": "frame-pairs/s", "n_gpus": world,
                          "ms_per_step": ms, "batch_per_gpu": a.batch, "hw": a.hw, "steps": a.steps, "warmup": a.warmup,
                          "loss": float(loss), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                          "scaling": "weak", "step": "one CUDA graph" if a.graph else "eager launches", "collective": "one flat gradient all-reduce per step" if world > 1 else None}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
