import sys, threading, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import elo_b200 as elo
from test_rowband_gpu import ThreadBand
cuda = torch.device("cuda:0")
H, W, N = 64, 1800, 150000
P = elo.params.init_params(0); perms = elo.params.make_perms(0)
pc, T = elo.synth.synth_batch(1, H, W, N, seed0=3); pc, T = pc.to(cuda), T.to(cuda)
want = elo.get_model(pc, H, W, T, None, None, False, params=elo.ParamStore(P, cuda), perms=perms)
torch.cuda.synchronize()
def trial(world, skip):
    shared, barrier = [None]*world, threading.Barrier(world)
    res, errs = [None]*world, []
    def run(rank):
        try:
            band = ThreadBand(elo, rank, world, shared, barrier, skip=skip)
            out = elo.get_model(pc, H, W, T, None, None, False, params=elo.ParamStore(P, cuda), perms=perms, band=band)
            torch.cuda.synchronize(); res[rank] = out
        except Exception as e:
            errs.append(e); barrier.abort()
    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in th]; [t.join() for t in th]
    if errs: return "ERR %s" % errs[0]
    return max(float((g - w).abs().max()) for out in res for g, w in zip(out[:8], want[:8]))
for world in (2, 4, 8):
    for skip in ((), ("layer0",), ("l0",), ("l1",), ("l2",), ("l0", "l1", "l2"), ("layer0", "l1", "l2"), ("layer0", "l0", "l2"), ("layer0", "l0", "l1")):
        print(world, skip, trial(world, skip), flush=True)
