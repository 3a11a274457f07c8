"""Phase timeline (ns) of CTA 0 of a tensor-core cost-volume stage-1 launch at a given level."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import elo_b200 as elo
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
dev = torch.device("cuda:0")
lib = elo._lib.lib()
lib.elo_set_time_log.argtypes = [ctypes.c_void_p]
store = elo.ParamStore(elo.params.init_params(0), dev)
perms = elo.params.make_perms(0)
g = torch.Generator().manual_seed(0)
for name, (H, W, C, kq, nq) in {"l0": (16, 225, 16, (11, 41), 6), "l2": (4, 57, 64, (5, 15), 6)}.items():
    xyz1 = (torch.randn(1, H, W, 3, generator=g) * 5).to(dev)
    xyz2 = (torch.randn(1, H, W, 3, generator=g) * 5).to(dev)
    f1 = torch.randn(1, H, W, C, generator=g).to(dev)
    f2 = torch.randn(1, H, W, C, generator=g).to(dev)
    scope = "flow_embedding_%s" % name
    tlog = torch.zeros(64, dtype=torch.int64, device=dev)
    orig_call = elo._lib.call
    state = {"on": False}

    def call(fn, desc, device):
        lib.elo_set_time_log(tlog.data_ptr() if (state["on"] and fn == WHICH) else None)
        orig_call(fn, desc, device)
        lib.elo_set_time_log(None)
    elo._lib.call = call
    WHICH = sys.argv[1] if len(sys.argv) > 1 else "elo_cost_volume_1"
    with elo.use_store(store):
        for rep in range(3):
            state["on"] = rep == 2
            out = elo.cost_volume(xyz1, xyz2, f1, f2, [3, 5], list(kq), 4, nq, 1.0, [128, 64, 64], [128, 64], False, None, scope)
            torch.cuda.synchronize()
    elo._lib.call = orig_call
    t = tlog.cpu().tolist()
    t0 = t[0]
    # stamps: 0 start, 1 init done, 2 nbr loaded, 3 gathered, 4 A loaded, 8+2l done(l) seen, 9+2l epilogue l over,
    # 5 logits staged, 6 pooled, 7 finished; MMA thread: 2+2l a_ready(l) seen, 3+2l layer l issued
    print(name, "compute:", [(i, x - t0) for i, x in enumerate(t[:32]) if x])
    print(name, "mma    :", [(i, x - t0) for i, x in enumerate(t[32:]) if x])
