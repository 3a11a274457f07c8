"""Phase timeline of the tile-staged index kernel from %globaltimer stamps taken by thread 0 of every CTA (and lane 0
of the store warp).  Needs the experiment build of the library:

    nvcc ... -DELO_TILED_TS  (tools/build_tiled_ts.sh -> tools/micro/libelo_b200_ts.so)
    ELO_B200_LIB=tools/micro/libelo_b200_ts.so python tools/tiled_timeline.py [kH kW] [select|random]
"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import elo_b200 as elo

dev = torch.device("cuda:0")
H, W, K = 64, 1800, 16
kH, kW = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (7, 25)
mode = sys.argv[3] if len(sys.argv) > 3 else "select"
kt, N = kH * kW, H * W
xyz = elo.synth.synth_scan(H, W, seed=0)[None].to(dev)
idx = elo.synth.hw_index(1, H, W, dev)
rhw = torch.randperm(kt, generator=torch.Generator().manual_seed(0)).to(torch.int32).to(dev)
o_idx = torch.empty((1, N, K, 3), dtype=torch.int32, device=dev)
o_mask = torch.empty((1, N, K, 1), dtype=torch.float32, device=dev)
o_valid = torch.empty((1, N, kt, 1), dtype=torch.float32, device=dev)
o_vdis = torch.empty((1, N, kt, 1), dtype=torch.float32, device=dev)
lib = elo._lib.lib()
elo._lib.set_index_kernel(1)
fn = lib.elo_fused_conv_select_k if mode == "select" else lib.elo_fused_conv_random_k


def call():
    rc = fn(1, H, W, N, kH, kW, K, 0, 1000.0, 1, 1, xyz.data_ptr(), xyz.data_ptr(), idx.data_ptr(), rhw.data_ptr(),
            o_idx.data_ptr(), o_valid.data_ptr(), o_vdis.data_ptr(), o_mask.data_ptr(), H, W,
            torch.cuda.current_stream().cuda_stream)
    assert rc == 0


for _ in range(5):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    call()
e1.record()
torch.cuda.synchronize()
print("%s %dx%d: %.1f us per launch (50 back to back)" % (mode, kH, kW, e0.elapsed_time(e1) * 1e3 / 50))
n = 1024 * 8
buf = (ctypes.c_ulonglong * n)()
lib.elo_debug_tiled_ts.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.elo_debug_tiled_ts(buf, n) == 0
ts = np.frombuffer(buf, dtype=np.uint64).reshape(1024, 8).astype(np.int64)
ts = ts[ts[:, 0] > 0]
ts = ts[ts[:, 6] >= ts[:, 0]]
t0 = ts[:, 0].min()
names = ["entry", "geometry done", "tile staged", "walk done (warp 0)", "idx/mask written", "count rows written (warp 0)",
         "exit (warp 0)", "store warp done"]
print("CTAs %d; times in us after the first CTA's entry: min / median / max over CTAs" % len(ts))
for k, nm in enumerate(names):
    col = ts[:, k]
    col = col[col >= t0]
    if len(col) == 0:
        print("  %-28s -" % nm)
        continue
    d = (col - t0) / 1e3
    print("  %-28s %6.1f / %6.1f / %6.1f" % (nm, d.min(), np.median(d), d.max()))
