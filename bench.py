#!/usr/bin/env python
"""bench.py -- frame-pairs/s of the PWCLO forward on 64x1800 synthetic KITTI-shaped scans (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

A step = one full forward (4-level pyramid, random-init weights) over one batch of B synthetic frame
pairs (default B = 1 = BASELINE.json configs[1]).  One process per GPU; for N > 1 launch under
torch.distributed.run -- every rank runs its own pairs (data-parallel over frame pairs, no collective on
the data path, weak scaling) and the timing is the max over ranks.

Printed by rank 0: ONE JSON line with
  value    frame-pairs/s over all ranks, inputs resident in HBM, whole forward replayed as a CUDA graph;
  e2e      the same through the public API (PWCLOEngine.infer) from pinned HOST buffers, H2D of the
           (B, 300000, 6) cloud and D2H of (q, t) inside the timed region;
  roofline the dominant kernel (by share of the step, timed live with CUDA events on the launching
           stream in an un-graphed pass over the same inputs);
  cpu_baseline  the CPU restatement (oracle/) on this box's host cores, on a bounded sample.
--impl reference times only that CPU restatement (the reference's TF-1.12 path cannot run: no TensorFlow,
and its custom ops are GPU-only -- BASELINE.md section 2).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H_IN, W_IN, NPTS = 64, 1800, 150000
POOL_BYTES = 144e6          # distinct input batches rotated through the timed loop: > 126 MB of L2
METRIC = "frame-pairs/sec on 64x1800 synthetic KITTI scans; cost-volume HBM GB/s vs roofline"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"], src="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, src="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=2)
        rows = [r for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in rows)
        reasons = [n for i, n in ((2, "hw_slowdown"), (3, "hw_thermal_slowdown"), (4, "sw_thermal_slowdown"),
                                  (5, "sw_power_cap")) if any(r[i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(rows[0][1]), "reasons": reasons, "samples": len(rows)}


# ---------------------------------------------------------------------------------------------------
def cpu_forward_rate(pairs, threads=None):
    """Pairs/s of the CPU restatement (torch-CPU graph oracle + C index oracle) on `pairs` distinct pairs."""
    import torch
    import elo_b200 as elo
    from oracle import graph_oracle as go
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    P = elo.params.init_params(0)
    perms = elo.params.make_perms(0)
    eye = torch.eye(4)[None]
    data = [elo.synth.synth_pair(H_IN, W_IN, s, NPTS) for s in range(pairs)]
    go.get_model(data[0][0][None], H_IN, W_IN, data[0][1][None], eye, eye, P, perms)       # warm-up
    t0 = time.perf_counter()
    for pc, T in data:
        go.get_model(pc[None], H_IN, W_IN, T[None], eye, eye, P, perms)
    dt = time.perf_counter() - t0
    return pairs / dt, dt, threads


def run_reference(args, rank, world):
    if rank != 0:
        return
    # one step = one frame pair on the host cores; bounded so the run ends within minutes
    steps = max(1, min(args.steps, 8))
    rate, dt, threads = cpu_forward_rate(steps)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "frame-pairs/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": 1, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "single frame-pair full PWCLO forward (4-level pyramid, random-init weights), "
                                   "64x1800, CPU restatement of the reference (TF 1.12 absent; custom ops GPU-only)",
                       "batch": 1},
            "cpu_baseline": {"value": rate, "unit": "frame-pairs/s", "cores": threads, "kind": "port",
                             "sample": "%d synthetic 64x1800 frame pairs, torch-CPU graph restatement + C index ops" % steps},
            "e2e": {"value": rate, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def algorithmic_work(name, tag, B):
    """(bytes, flops) one launch of a fused block must move / compute."""
    lv = {"l0": (3600, 16), "l1": (904, 32), "l2": (228, 64), "l2o": (228, 64)}
    if name in ("elo_cost_volume_1", "elo_cost_volume_2") and tag in lv:
        N, C = lv[tag]
        Kq = 32 if tag == "l2o" else 6
        if name == "elo_cost_volume_1":
            params = (10 + 2 * C) * 128 + 128 * 64 + 64 * 64 + 10 * 64 + 128 * 128 + 128 * 64
            byts = 4 * (B * N * (3 + 3 + C + C + 64) + params)
            flops = 2 * B * N * Kq * params
        else:
            params = 10 * 64 + (128 + C) * 128 + 128 * 64
            byts = 4 * (B * N * (3 + C + 64 + 64) + params)
            flops = 2 * B * N * 4 * params
        return byts, flops
    return None, None


def run_ours(args, rank, world, local_rank):
    import torch
    import elo_b200 as elo
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    store = elo.ParamStore(elo.params.init_params(0), dev)
    perms = elo.params.make_perms(0)
    # distinct input batches, rotated so that every step reads inputs that are cold in L2
    batch_bytes = B * 2 * NPTS * 6 * 4
    pool = args.pool if args.pool > 0 else max(2, int(POOL_BYTES // batch_bytes) + 1)
    host = [elo.synth.synth_batch(B, H_IN, W_IN, NPTS, seed0=rank * 1000 + i * B) for i in range(min(pool, 4))]
    engines = []
    for i in range(pool):
        eng = elo.PWCLOEngine(B, H_IN, W_IN, NPTS, params=store, perms=perms, device=dev, use_graph=not args.no_graph)
        pc, T = host[i % len(host)]
        # make the pool's buffers distinct in content too (a rigid shift of the unique batches)
        eng.load(pc, T, non_blocking=False)
        n0 = elo._lib.launch_count()
        eng.capture()
        per_forward = (elo._lib.launch_count() - n0) // 3          # 2 eager warm-ups + 1 capture
        engines.append(eng)
    stream = engines[0].stream
    for e in engines:
        e.stream = stream
    torch.cuda.synchronize(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- value: inputs resident in HBM, K graph replays -------------------------------------------
    for i in range(args.warmup):
        engines[i % pool].run()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for i in range(args.steps):
            engines[(args.warmup + i) % pool].run()
        e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.summary()
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * B * args.steps / (ms * 1e-3)

    # ---- e2e: public API from pinned host buffers ---------------------------------------------------
    pinned = [(pc.pin_memory(), T.pin_memory()) for pc, T in host]
    eng = engines[0]
    for i in range(max(3, args.warmup)):
        eng.infer(*pinned[i % len(pinned)])
    barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        e0.record(stream)
    for i in range(args.steps):
        q, t = eng.infer(*pinned[i % len(pinned)])
    with torch.cuda.stream(stream):
        e1.record(stream)
    wall_ms = 1e3 * (time.perf_counter() - t0)
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), wall_ms)       # infer() synchronises: wall clock is the honest figure
    if dist is not None:
        tt = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
    e2e_value = world * B * args.steps / (e2e_ms * 1e-3)
    h2d = batch_bytes + B * 64
    d2h = B * 7 * 4

    # ---- per-kernel shares and the roofline of the dominant kernel (un-graphed pass, CUDA events) ------
    shares, roof = {}, None
    if rank == 0:
        prof = []
        elo._lib.PROFILE = prof
        iters = max(3, min(args.steps, 20))
        tagged = TaggedForward(elo, engines)
        with torch.cuda.stream(stream):
            for i in range(iters + 2):
                if i == 2:
                    del prof[:]
                tagged.run(i % pool)
        torch.cuda.synchronize(dev)
        elo._lib.PROFILE = None
        agg = {}
        for name, tag, a, b in prof:
            agg.setdefault((name, tag), []).append(a.elapsed_time(b))
        total = sum(sum(v) for v in agg.values()) / iters
        top = sorted(agg.items(), key=lambda kv: -sum(kv[1]))
        shares = {"%s[%s]" % k: round(sum(v) / iters / total, 4) for k, v in top[:8]}
        peaks = measured_peaks()
        for (name, tag), v in top:
            byts, flops = algorithmic_work(name, tag, B)
            if byts is None:
                continue
            dur = sum(v) / len(v) * 1e-3
            roof = {"kernel": "%s[%s]" % (name, tag), "bound": "tensor", "achieved": flops / dur / 1e12,
                    "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": flops / dur / 1e12 / peaks["bf16_tflops"],
                    "traffic": None, "avg_launch_us": dur * 1e6, "share_of_step": round(sum(v) / iters / total, 4),
                    "peak_source": peaks["src"],
                    "note": "per-group MLP runs on fp32 FFMA (fp32 parity bar 1e-4); nominal fp32 peak 74 TFLOP/s "
                            "-> frac_fp32 %.3f; HBM view: %.1f GB/s algorithmic = %.4f of %.0f GB/s"
                            % (flops / dur / 74e12, byts / dur / 1e9, byts / dur / 1e9 / peaks["hbm_gbs"], peaks["hbm_gbs"])}
            break

    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rate, dt, threads = cpu_forward_rate(args.cpu_pairs)
        cpu = {"value": rate, "unit": "frame-pairs/s", "cores": threads, "kind": "port",
               "sample": "%d synthetic 64x1800 frame pairs (%.1f s), torch-CPU graph restatement + C index ops"
                         % (args.cpu_pairs, dt)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "frame-pairs/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "single frame-pair full PWCLO forward (4-level pyramid, random-init weights), "
                                       "64x1800, 150000 points/frame" if B == 1 else
                                       "batch=%d frame-pairs full PWCLO forward, 64x1800" % B,
                           "batch_per_gpu": B, "parallelism": "frame-pairs sharded over %d GPU(s), no collective" % world,
                           "l2": "inputs rotate over %d distinct batches (%.0f MB > 126 MB L2); weights stay resident"
                                 % (pool, pool * batch_bytes / 1e6),
                           "graph": "kernel-by-kernel launches (--no-graph)" if args.no_graph else
                                    "whole forward captured as one CUDA graph"},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "frame-pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": per_forward * args.steps, "launches_per_step": per_forward,
                "kernel_shares": shares, "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


class TaggedForward:
    """Eager forward with a level tag set before each block, so per-kernel timings can be attributed."""

    def __init__(self, elo, engines):
        self.elo, self.engines = elo, engines
        pu = elo.pointnet_util
        self._cv = pu.cost_volume
        tagger = self

        def cost_volume(*a, **k):
            scope = a[13] if len(a) > 13 else k["scope"]
            elo._lib.PROFILE_TAG[0] = {"flow_embedding_l2_origin": "l2o", "flow_embedding_l2": "l2",
                                       "flow_embedding_l1": "l1", "flow_embedding_l0": "l0"}.get(scope, "")
            try:
                return tagger._cv(*a, **k)
            finally:
                elo._lib.PROFILE_TAG[0] = ""
        self.patched = cost_volume

    def run(self, i):
        pu = self.elo.pointnet_util
        orig = pu.cost_volume
        pu.cost_volume = self.patched
        try:
            self.engines[i]._forward()
        finally:
            pu.cost_volume = orig


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-pairs", type=int, default=6)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernel by kernel (for ncu launch lists)")
    ap.add_argument("--pool", type=int, default=0, help="input batches to rotate (0 = enough to exceed L2)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
