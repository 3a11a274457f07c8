"""Converts the reference's shipped checkpoint (/root/reference/pretrained_model) into
oracle/_ref/pretrained_params.npz -- a derived artefact in the git-ignored _ref directory, so that the GPU
box (which has no /root/reference) can run the parity / behaviour tests with the real trained weights.
TEST INFRASTRUCTURE; run by __graft_entry__.build() where /root/reference exists."""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
PREFIX = "/root/reference/pretrained_model/pretrained_model.ckpt"
OUT = os.path.join(HERE, "_ref", "pretrained_params.npz")


def main():
    if not os.path.exists(PREFIX + ".index"):
        print("no reference checkpoint here; keeping", OUT)
        return
    ck = importlib.import_module("efficientlo-net_b200.tf_checkpoint")
    P, step = ck.load_reference_checkpoint(PREFIX)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez(OUT, __global_step__=np.array(step), **{k.replace("/", "|"): v.numpy() for k, v in P.items()})
    print(OUT, len(P), "tensors, global step", step)


def load():
    """{name: torch tensor} from the converted file, or None if it has not been built."""
    import torch
    if not os.path.exists(OUT):
        return None
    z = np.load(OUT)
    return {k.replace("|", "/"): torch.from_numpy(z[k]) for k in z.files if k != "__global_step__"}


if __name__ == "__main__":
    main()
