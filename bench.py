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


def profiled_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` from the committed ncu --set full
    capture (profiles/traffic_r*.json, written by the profiling pass; B = 1), or None."""
    import glob
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "traffic_r*.json"))):
        try:
            best = json.load(open(path)).get("dram_bytes_per_launch", {}).get(kernel, best)
        except Exception:
            pass
    return best


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
            self.stop_flag.wait(0.05)

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
LEVELS = {"l0": (3600, 16), "l1": (904, 32), "l2": (228, 64), "l2o": (228, 64)}


def algorithmic_work(name, tag, B):
    """(bytes, flops) one launch of a fused block must move / compute: inputs read
    once, outputs written once, weights once; 2 * rows * Cin * Cout per layer (the 3xTF32 split is NOT
    counted three times)."""
    mac = lambda dims: sum(a * b for a, b in zip(dims[:-1], dims[1:]))
    if name == "elo_cost_volume_1" and tag in LEVELS:
        N, C = LEVELS[tag]
        Kq = 32 if tag == "l2o" else 6
        params = mac([10 + 2 * C, 128, 64, 64]) + 10 * 64 + mac([128, 128, 64])
        return 4 * (B * N * (3 + 3 + C + C + 64 + Kq) + params), 2 * B * N * Kq * params
    if name == "elo_cost_volume_2" and tag in LEVELS:
        N, C = LEVELS[tag]
        params = 10 * 64 + mac([128 + C, 128, 64])
        return 4 * (B * N * (3 + C + 64 + 64 + 4) + params), 2 * B * N * 4 * params
    if name == "elo_group_mlp_max":
        if tag == "sa3":       # sa1/layer3, both frames: 116 centres on the 4x57 grid, K = 16
            params = mac([67, 64, 64, 128])
            return 4 * (2 * B * (228 * 67 + 116 * (128 + 16)) + params), 2 * 2 * B * 116 * 16 * params
        if tag == "l3":        # new_layer3
            params = mac([67, 128, 64, 64])
            return 4 * (B * (228 * 67 + 116 * (64 + 16)) + params), 2 * B * 116 * 16 * params
        if tag in LEVELS:      # the level's two set-upconvs, first half (K = 8)
            N, _ = LEVELS[tag]
            params = mac([67, 128, 64])
            coarse = {"l0": 904, "l1": 228, "l2": 116}[tag]
            return 4 * 2 * (B * (N * (3 + 64 + 8) + coarse * 67) + params), 2 * 2 * B * N * 8 * params
    if name == "elo_row_mlp":
        if tag == "l3":
            params = mac([192, 128, 64])
            return 4 * (B * 116 * (192 + 64) + params), 2 * B * 116 * params
        if tag in LEVELS:      # second half of both up-convs chained into both predictors
            N, C = LEVELS[tag]
            params = mac([64 + C, 128, 64]) + mac([C + 128, 128, 64])
            return 4 * 2 * (B * N * (64 + C + 64 + 64) + params), 2 * 2 * B * N * params
    return None, None


def index_op_roofline(elo, dev, peaks, iters=20):
    """BASELINE.json configs[0]: fused_conv_select_k on one 64x1800 frame, K = 16, window 7x25, every
    pixel a query, all four outputs of the reference op (194.5 MB: HBM-write bound).  Called through the
    C ABI with pre-allocated outputs (what the reference's op wrapper hands its Launcher), CUDA events
    on the launching stream; the 194 MB of outputs exceed L2 every launch."""
    import torch
    H, W, K, kH, kW = 64, 1800, 16, 7, 25
    N, kt = H * W, kH * kW
    xyz = elo.synth.synth_scan(H, W, seed=0)[None].to(dev)
    idx = elo.synth.hw_index(1, H, W, dev)
    rhw = torch.randperm(kt, generator=torch.Generator().manual_seed(0)).to(torch.int32).to(dev)
    o_idx = torch.empty((1, N, K, 3), dtype=torch.int32, device=dev)
    o_mask = torch.empty((1, N, K, 1), dtype=torch.float32, device=dev)
    o_valid = torch.empty((1, N, kt, 1), dtype=torch.float32, device=dev)
    o_vdis = torch.empty((1, N, kt, 1), dtype=torch.float32, device=dev)
    byts = 4 * (2 * 3 * N + 2 * N + kt + 3 * N * K + N * K + 2 * N * kt)
    lib = elo._lib.lib()

    def call():
        rc = lib.elo_fused_conv_select_k(1, H, W, N, kH, kW, K, 0, 1000.0, 1, 1, xyz.data_ptr(), xyz.data_ptr(),
                                         idx.data_ptr(), rhw.data_ptr(), o_idx.data_ptr(), o_valid.data_ptr(),
                                         o_vdis.data_ptr(), o_mask.data_ptr(), H, W,
                                         torch.cuda.current_stream().cuda_stream)
        elo._lib.check(rc, "elo_fused_conv_select_k")

    for _ in range(3):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize(dev)
    dur = e0.elapsed_time(e1) * 1e-3 / iters
    return {"kernel": "fused_conv_tiled_kernel<select, 17, 160> (64x1800, K=16, 7x25)", "bound": "hbm",
            "achieved": byts / dur / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": byts / dur / 1e9 / peaks["hbm_gbs"],
            "traffic": profiled_traffic("elo_fused_conv_select_k[config1]"), "avg_launch_us": dur * 1e6,
            "algorithmic_bytes": byts, "peak_source": peaks["src"],
            "note": "all four outputs of the reference op, pre-allocated; outputs (194 MB) exceed L2 every launch"}


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
    elo._lib.set_mlp_engine(1 if args.engine == "tc" else 0)
    policy = args.tile_policy if args.tile_policy >= 0 else (1 if args.streams > 1 else 0)
    elo._lib.set_tile_policy(policy)
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
    # One forward of a single frame pair is a chain of ~40 dependent single-wave kernels that leaves most of
    # the 148 SMs idle; --streams S keeps S independent forwards (S different frame pairs, each its own
    # captured graph, input buffer and scratch) in flight on S streams.  Every step is still one complete
    # forward of one batch; the timed region is bracketed by events on `stream`, which all S streams fork
    # from and join back into.
    S = max(1, args.streams)
    lanes = [stream] + [torch.cuda.Stream(dev) for _ in range(S - 1)]

    def run_steps(first, count):
        if S == 1:
            for i in range(count):
                engines[(first + i) % pool].run()
            return
        fork = torch.cuda.Event()
        fork.record(stream)
        for ln in lanes[1:]:
            ln.wait_event(fork)
        for i in range(count):
            eng = engines[(first + i) % pool]
            eng.stream = lanes[i % S]
            eng.run()
        for ln in lanes[1:]:
            join = torch.cuda.Event()
            join.record(ln)
            stream.wait_event(join)

    def timed(first, count):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            run_steps(first, count)
            e1.record(stream)
        barrier()
        return e0.elapsed_time(e1)

    with torch.cuda.stream(stream):
        run_steps(0, args.warmup)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms = timed(args.warmup, args.steps)
    clocks = sampler.summary()
    serial_ms = None
    if S > 1:                      # the same steps one after the other on one stream: the latency of a forward
        S_keep, S = S, 1
        for e in engines:
            e.stream = stream
        serial_ms = timed(args.warmup, args.steps) / args.steps
        S = S_keep
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * B * args.steps / (ms * 1e-3)

    # ---- e2e: public API from pinned host buffers ---------------------------------------------------
    # (a) PWCLOPipeline.run: every batch is uploaded from pinned host memory and its (q, t) read back;
    #     the copies overlap the neighbouring batches' compute (two input buffers, copy stream);
    # (b) PWCLOEngine.infer: fully synchronous per batch (upload -> forward -> read back -> host wakes).
    pinned = [(pc.pin_memory(), T.pin_memory()) for pc, T in host]
    h2d = batch_bytes + B * 64
    d2h = B * 7 * 4
    pipe = elo.PWCLOPipeline(B, H_IN, W_IN, NPTS, params=store, perms=perms, device=dev, streams=S)
    feed = lambda n: (pinned[i % len(pinned)] for i in range(n))
    for _ in pipe.run(feed(max(3, args.warmup))):
        pass
    barrier()
    t0 = time.perf_counter()
    nres = sum(1 for _ in pipe.run(feed(args.steps)))
    torch.cuda.synchronize(dev)
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    assert nres == args.steps
    barrier()
    eng = engines[0]
    for i in range(3):
        eng.infer(*pinned[i % len(pinned)])
    t0 = time.perf_counter()
    for i in range(args.steps):
        q, t = eng.infer(*pinned[i % len(pinned)])
    sync_ms = 1e3 * (time.perf_counter() - t0)
    barrier()
    if dist is not None:
        tt = torch.tensor([e2e_ms, sync_ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms, sync_ms = float(tt[0].item()), float(tt[1].item())
    e2e_value = world * B * args.steps / (e2e_ms * 1e-3)
    sync_value = world * B * args.steps / (sync_ms * 1e-3)

    # ---- per-kernel shares and the roofline of the dominant kernel (un-graphed pass, CUDA events) ------
    shares, roof, roof_index = {}, None, None
    if rank == 0:
        prof = []
        elo._lib.PROFILE = prof
        iters = max(3, min(args.steps, 20))
        with torch.cuda.stream(stream):
            for i in range(iters + 2):
                if i == 2:
                    del prof[:]
                engines[i % pool]._forward()
        torch.cuda.synchronize(dev)
        elo._lib.PROFILE = None
        agg = {}
        for name, tag, a, b in prof:
            agg.setdefault((name, tag), []).append(a.elapsed_time(b))
        total = sum(sum(v) for v in agg.values()) / iters
        top = sorted(agg.items(), key=lambda kv: -sum(kv[1]))
        shares = {"%s[%s]" % k: round(sum(v) / iters / total, 4) for k, v in top[:8]}
        if args.kernel_times:
            for (name, tag), v in sorted(agg.items(), key=lambda kv: kv[0][1] + kv[0][0]):
                sys.stderr.write("%-28s %-6s n=%d  %8.2f us/launch\n" % (name, tag, len(v) // iters, sum(v) / len(v) * 1e3))
            sys.stderr.write("sum of kernel times per forward (un-graphed, event-timed): %.1f us\n" % (total * 1e3))
        peaks = measured_peaks()
        engine = "tcgen05 tf32x3" if elo._lib.mlp_engine() == 1 else "fp32 FFMA"
        for (name, tag), v in top:
            byts, flops = algorithmic_work(name, tag, B)
            if byts is None:
                continue
            dur = sum(v) / len(v) * 1e-3
            tf = flops / dur / 1e12
            roof = {"kernel": "%s[%s]" % (name, tag), "bound": "tensor", "achieved": tf,
                    "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": tf / peaks["bf16_tflops"],
                    "traffic": profiled_traffic("%s[%s]" % (name, tag)) if B == 1 else None,
                    "avg_launch_us": dur * 1e6, "share_of_step": round(sum(v) / iters / total, 4),
                    "peak_source": peaks["src"], "algorithmic_flops": flops, "algorithmic_bytes": byts,
                    "note": "per-group MLP on %s; algorithmic FLOPs (the three tf32 partial products of the "
                            "fp32-grade split are counted once, so 1/6 of the bf16 peak is this design's ceiling); "
                            "HBM view: %.1f GB/s algorithmic = %.4f of %.0f GB/s -- the fused block is compute/"
                            "latency-bound, not HBM-bound" % (engine, byts / dur / 1e9,
                                                              byts / dur / 1e9 / peaks["hbm_gbs"], peaks["hbm_gbs"])}
            break
        roof_index = index_op_roofline(elo, dev, peaks)

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
                                    "whole forward captured as one CUDA graph",
                           "streams": "%d independent forwards in flight on %d CUDA streams (each step = one complete "
                                      "forward of one batch; --streams 1 runs them back to back)" % (S, S),
                           "serial_ms_per_forward": serial_ms,
                           "tile_policy": "throughput (full 128-row tiles)" if policy == 1 else "latency (small calls spread over all SMs)",
                           "pdl": bool(elo._lib.lib().elo_get_pdl())},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "frame-pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms / args.steps, "api": "PWCLOPipeline.run (pinned host batches in, (q,t) out; "
                        "copies overlap neighbouring batches; %d forwards in flight)" % S, "synchronous_infer": {"value": sync_value,
                                                                                       "ms_per_step": sync_ms / args.steps}},
                "gpu_launches": per_forward * args.steps, "launches_per_step": per_forward,
                "kernel_shares": shares, "roofline": roof, "roofline_index_op": roof_index, "cpu_baseline": cpu,
                "mlp_engine": "tcgen05 tf32x3" if elo._lib.mlp_engine() == 1 else "fp32 FFMA"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-pairs", type=int, default=6)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--engine", default="tc", choices=["tc", "ffma"], help="MLP engine: tcgen05 3xTF32 or fp32 FFMA")
    ap.add_argument("--kernel-times", action="store_true", help="print every kernel's average time to stderr")
    ap.add_argument("--no-graph", action="store_true", help="launch kernel by kernel (for ncu launch lists)")
    ap.add_argument("--pool", type=int, default=0, help="input batches to rotate (0 = enough to exceed L2)")
    ap.add_argument("--tile-policy", type=int, default=-1, help="tensor-core tiles: 0 latency, 1 throughput, -1 by --streams")
    ap.add_argument("--streams", type=int, default=12, help="independent forwards kept in flight (1 = back to back)")
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
