"""Pose-tolerance sweep (BASELINE.json configs[4]): regressed SE(3) pose of the CUDA path -- tcgen05 3xTF32 engine and
fp32 FFMA engine -- against the torch-CPU restatement in fp32 AND fp64, at 64x1800 and 128x2048.

    python tools/pose_tolerance.py [--pairs 2]
Prints one JSON line per (size, engine): per pyramid level max |q - q_ref| and max |t - t_ref| / |t_ref|."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import elo_b200 as elo  # noqa: E402
from oracle import graph_oracle as go  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=2)
    ap.add_argument("--sizes", nargs="+", default=["64x1800", "128x2048"])
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    P = elo.params.init_params(0)
    perms = elo.params.make_perms(0)
    for size in a.sizes:
        H, W = (int(v) for v in size.split("x"))
        npts = 150000 if H * W <= 150000 else 300000
        pc, T = elo.synth.synth_batch(a.pairs, H, W, npts)
        eye = torch.eye(4).expand(a.pairs, 4, 4).contiguous()
        refs = {}
        for name, dt in (("fp32", torch.float32), ("fp64", torch.float64)):
            Pd = {k: v.to(dt) for k, v in P.items()}
            refs[name] = [o.double() for o in go.get_model(pc, H, W, T, eye, eye, Pd, perms, dtype=dt)[:8]]
        drift = {}
        for lvl in range(4):
            q32, t32, q64, t64 = refs["fp32"][2 * lvl], refs["fp32"][2 * lvl + 1], refs["fp64"][2 * lvl], refs["fp64"][2 * lvl + 1]
            drift["l%d" % lvl] = {"dq": float((q32 - q64).norm(dim=-1).max()),
                                  "dt_rel": float(((t32 - t64).norm(dim=-1) / t64.norm(dim=-1)).max())}
        print(json.dumps({"size": size, "pairs": a.pairs, "what": "restatement fp32 vs fp64 (the yardstick)", "levels": drift}))
        for engine, ename in ((1, "tcgen05 tf32x3"), (0, "fp32 FFMA")):
            elo._lib.set_mlp_engine(engine)
            store = elo.ParamStore(P, dev)
            out = elo.get_model(pc.to(dev), H, W, T.to(dev), None, None, False, params=store, perms=perms)
            torch.cuda.synchronize()
            got = [o.cpu().double() for o in out[:8]]
            rec = {"size": size, "pairs": a.pairs, "engine": ename, "levels": {}}
            for lvl in range(4):
                q, t = got[2 * lvl], got[2 * lvl + 1]
                rec["levels"]["l%d" % lvl] = {
                    "vs_" + name: {"dq": float((q - refs[name][2 * lvl]).norm(dim=-1).max()),
                                   "dt_rel": float(((t - refs[name][2 * lvl + 1]).norm(dim=-1) /
                                                    refs[name][2 * lvl + 1].norm(dim=-1)).max())}
                    for name in ("fp32", "fp64")}
            print(json.dumps(rec))
        elo._lib.set_mlp_engine(1)


if __name__ == "__main__":
    main()
