"""Writes tests/golden/index_golden.npz: inputs and outputs of the REFERENCE's own kernel bodies
(oracle/_ref/libref_cpu.so, compiled by oracle/Makefile from /root/reference/tf_ops/*/fused_conv_g.cu)
on small seeded cases.  Run in the build container (needs /root/reference):

    make -C oracle ref && python tests/golden/make_index_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cases  # noqa: E402
from oracle import index_oracle as io  # noqa: E402


def main():
    out = {}
    picked = []
    # randomised small cases of both ops (float and tie-heavy integer coordinates) ...
    for seed in range(1000, 1400):
        rng = np.random.default_rng(seed)
        c = cases.random_case(rng)
        if c["npoints"] * c["kernel_size_H"] * c["kernel_size_W"] > 6000:
            continue
        picked.append(c)
        if len(picked) == 10:
            break
    # ... and two real call signatures at batch 1 (coarsest levels, to stay small)
    rng = np.random.default_rng(77)
    picked.append(cases.site_case(rng, cases.MODEL_SITES[6]))   # flow_embedding_l2 select-K 5x15 K=6
    picked.append(cases.site_case(rng, cases.MODEL_SITES[3]))   # sa1/layer3 random-K 5x9 K=16
    for i, c in enumerate(picked):
        res = cases.call(io.ref_cpu, c)
        p = "c%d_" % i
        out[p + "mode"] = np.array(c["mode"])
        for k in ("xyz1", "xyz2", "idx_n2", "random_hw"):
            out[p + k] = c[k]
        out[p + "ints"] = np.array([c[k] for k in ("H", "W", "npoints", "kernel_size_H", "kernel_size_W",
                                                   "K", "flag_copy", "stride_h", "stride_w")], np.int32)
        out[p + "distance"] = np.array(c["distance"], np.float32)
        for name, r in zip(cases.OUT_NAMES, res):
            out[p + name] = r
    out["n_cases"] = np.array(len(picked))
    path = os.path.join(HERE, "index_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes,", len(picked), "cases")


if __name__ == "__main__":
    main()
