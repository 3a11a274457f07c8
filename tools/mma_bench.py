"""Cycles per tcgen05.mma.kind::tf32 (128 x N x 8) on this GPU, A operand from TMEM (TS) or shared memory (SS)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import importlib
import elo_b200 as elo
lib = ctypes.CDLL(importlib.import_module("efficientlo-net_b200.build").build_test_lib())     # tests/csrc hooks
lib.elo_tc_mma_bench.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
out = torch.zeros(1, dtype=torch.int64, device="cuda")
for ts in (1, 2):
    for N, nacc in ((64, 1), (64, 2), (64, 3), (128, 1), (128, 2), (128, 3), (32, 1), (32, 3)):
        res = []
        for iters in (96, 960):
            lib.elo_tc_mma_bench(N, iters, ts, nacc, out.data_ptr(), None)
            torch.cuda.synchronize()
            res.append((iters, int(out.item())))
        per = (res[1][1] - res[0][1]) / (res[1][0] - res[0][0])
        print("ts=%d N=%3d nacc=%d  -> %.1f cycles / MMA  (%.0f MAC/clk)" % (ts, N, nacc, per, 128 * N * 8 / per))
