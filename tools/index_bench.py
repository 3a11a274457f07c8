"""BASELINE.json configs[0] timed alone: fused_conv_select_k / random_k on one 64x1800 frame, every pixel a
query, through the C ABI with pre-allocated outputs (CUDA events, L2 flushed by the 194 MB of outputs).
Usage: python tools/index_bench.py [iters]   -> one JSON line per (kernel, op, window)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import elo_b200 as elo


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    dev = torch.device("cuda:0")
    H, W, K = 64, 1800, 16
    xyz = elo.synth.synth_scan(H, W, seed=0)[None].to(dev)
    idx = elo.synth.hw_index(1, H, W, dev)
    lib = elo._lib.lib()
    peaks = 6524.9
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
        peaks = float(peaks.get("hbm_gbs") or peaks.get("hbm_copy_gbs") or 6524.9)
    except Exception:
        peaks = 6524.9
    for select, kH, kW, dist in [(True, 7, 25, 1000.0), (True, 11, 41, 1000.0), (True, 5, 15, 1000.0),
                                 (False, 9, 15, 0.5), (False, 7, 25, 1000.0)]:
        kt = kH * kW
        rhw = torch.randperm(kt, generator=torch.Generator().manual_seed(0)).to(torch.int32).to(dev)
        N = H * W
        o_idx = torch.empty((1, N, K, 3), dtype=torch.int32, device=dev)
        o_mask = torch.empty((1, N, K, 1), dtype=torch.float32, device=dev)
        o_valid = torch.empty((1, N, kt, 1), dtype=torch.float32, device=dev)
        o_vdis = torch.empty((1, N, kt, 1), dtype=torch.float32, device=dev)
        byts = 4 * (2 * 3 * N + 2 * N + kt + 3 * N * K + N * K + 2 * N * kt)
        fn = lib.elo_fused_conv_select_k if select else lib.elo_fused_conv_random_k
        for which, name in ((1, "tiled"), (2, "warp")):
            elo._lib.set_index_kernel(which)
            for full in (True, False):
                def call():
                    rc = fn(1, H, W, N, kH, kW, K, 0, dist, 1, 1, xyz.data_ptr(), xyz.data_ptr(), idx.data_ptr(),
                            rhw.data_ptr(), o_idx.data_ptr(), o_valid.data_ptr() if full else None,
                            o_vdis.data_ptr() if full else None, o_mask.data_ptr(), H, W,
                            torch.cuda.current_stream().cuda_stream)
                    assert rc == 0, elo._lib.last_error()
                for _ in range(3):
                    call()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(iters):
                    call()
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / iters
                b = byts if full else byts - 4 * 2 * N * kt
                print(json.dumps({"kernel": name, "op": "select_k" if select else "random_k", "window": [kH, kW], "K": K,
                                  "outputs": "all four" if full else "idx + mask only", "us": round(us, 1),
                                  "algorithmic_MB": round(b / 1e6, 1), "GBps": round(b / us / 1e3, 1),
                                  "frac_of_hbm_peak": round(b / us / 1e3 / peaks, 3)}), flush=True)
    elo._lib.set_index_kernel(0)


if __name__ == "__main__":
    main()
