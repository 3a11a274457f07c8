"""One configs[0]-style call timed alone: python tools/index_one.py [kH kW] [iters] -> us per launch (all four outputs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import elo_b200 as elo
dev = torch.device("cuda:0")
H, W, K = 64, 1800, 16
kH, kW = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (7, 25)
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 30
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
def call():
    rc = lib.elo_fused_conv_select_k(1, H, W, N, kH, kW, K, 0, 1000.0, 1, 1, xyz.data_ptr(), xyz.data_ptr(), idx.data_ptr(),
                                     rhw.data_ptr(), o_idx.data_ptr(), o_valid.data_ptr(), o_vdis.data_ptr(), o_mask.data_ptr(),
                                     H, W, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
for _ in range(3):
    call()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(iters):
    call()
e1.record()
torch.cuda.synchronize()
print("%dx%d dbg=%s: %.1f us" % (kH, kW, os.environ.get("ELO_TILED_DBG", "0"), e0.elapsed_time(e1) * 1e3 / iters))
