"""One eager (un-graphed) forward at B=1 -- the target for compute-sanitizer / ncu runs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import elo_b200 as elo  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda:0")
store = elo.ParamStore(elo.params.init_params(0), dev)
perms = {k: v.to(dev) for k, v in elo.params.make_perms(0).items()}
pc, T = elo.synth.synth_batch(B)
pc, T = pc.to(dev), T.to(dev)
for _ in range(reps):
    out = elo.get_model(pc, 64, 1800, T, None, None, False, params=store, perms=perms)
torch.cuda.synchronize()
print("q", out[0].cpu().tolist(), "t", out[1].cpu().tolist())
