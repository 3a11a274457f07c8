"""BASELINE.json configs[0]: fused_conv_select_k on one 64x1800 frame (ncu / timing target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import elo_b200 as elo
dev = torch.device("cuda:0")
H, W, K = 64, 1800, 16
kH, kW = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (7, 25)
xyz = elo.synth.synth_scan(H, W, seed=0)[None].to(dev)
idx = elo.synth.hw_index(1, H, W, dev)
rhw = torch.randperm(kH * kW, generator=torch.Generator().manual_seed(0)).to(torch.int32).to(dev)
for rep in range(4):
    out = elo.fused_conv_select_k(xyz, xyz, idx, rhw, H, W, H * W, kH, kW, K, 0, 1000.0, 1, 1)
torch.cuda.synchronize()
print(int(out[3].sum().item()))
