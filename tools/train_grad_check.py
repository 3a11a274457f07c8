"""Per-parameter gradient comparison: training-mode graph on the GPU vs the torch-CPU restatement."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import elo_b200 as elo  # noqa: E402
from oracle import graph_oracle as go  # noqa: E402

tg = elo.train_graph
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dt = torch.float64 if "--f64" in sys.argv else torch.float32
P = elo.params.init_params(0)
perms = elo.params.make_perms(0)
pc, T = elo.synth.synth_batch(B, 64, 1800, 150000)
eye = torch.eye(4).expand(B, 4, 4).contiguous()
Pc = {k: v.clone().to(dt).requires_grad_(not k.endswith(("moving_mean", "moving_variance"))) for k, v in P.items()}
keep_c, keep_g = {}, {}
with go.training(bn_decay=0.5):
    out_c = go.get_model(pc, 64, 1800, T, eye, eye, Pc, perms, dtype=dt, keep=keep_c)
loss_c = go.get_loss(*out_c[:8], out_c[9], out_c[10], Pc["w_x"], Pc["w_q"])
for k in keep_c.values():
    if k.requires_grad:
        k.retain_grad()
loss_c.backward()
tp = tg.TrainableParams(P, "cuda")
out_g = tg.get_model(pc.cuda(), 64, 1800, T.cuda(), None, None, tp, bn_decay=0.5, perms=perms, dropout=0.0, keep=keep_g)
loss_g = elo.get_loss(*out_g[:8], out_g[9], out_g[10], tp["w_x"], tp["w_q"])
for k in keep_g.values():
    if k.requires_grad:
        k.retain_grad()
loss_g.backward()
print("loss", float(loss_c), float(loss_g))


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


print("---- intermediates: value rel err, grad rel err")
for k in keep_g:
    if k in keep_c and keep_g[k].shape == keep_c[k].shape:
        gv = rel(keep_g[k], keep_c[k])
        gg = rel(keep_g[k].grad, keep_c[k].grad) if keep_g[k].grad is not None and keep_c[k].grad is not None else float("nan")
        print("%-28s %.2e %.2e" % (k, gv, gg))
print("---- parameters")
for n, p in tp.named_parameters():
    w = Pc[n].grad
    if w is None or float(w.norm()) < 1e-7:
        continue
    print("%-60s |g| %.3e rel %.2e" % (n, float(w.norm()), rel(p.grad, w)))
