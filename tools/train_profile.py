"""Where a training step's GPU time goes: torch profiler over two eager steps, top kernels by device time.
    python tools/train_profile.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
import elo_b200 as elo

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
tg = elo.train_graph
tp = tg.TrainableParams(elo.params.init_params(0), dev)
tr = tg.Trainer(tp, batch_size=B)
pc, T = elo.synth.synth_batch(B, 64, 1800, 150000)
pc, T = pc.to(dev), T.to(dev)
perms = elo.params.make_perms(0)
for _ in range(2):
    tr.step(pc, T, perms=perms)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        tr.step(pc, T, perms=perms)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=30, max_name_column_width=70))
