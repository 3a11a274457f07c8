"""The reference's trained weights as a behavioural pin.

The graph blocks have no TensorFlow oracle (parity unpinned).  What the reference does
ship is its trained checkpoint: if the CPU restatement (and the CUDA path) of set-conv, cost volume, warp,
re-projection, embedding mask and pose composition are faithful, the network must regress the true
ego-motion of a synthetic frame pair it has never seen; any wrong channel order, warp direction, BN fold or
quaternion convention destroys that."""
import importlib
import os

import pytest
import torch

from oracle import graph_oracle as go
from oracle import make_ref_params

REF = "/root/reference/pretrained_model/pretrained_model.ckpt"
H, W, N = 64, 1800, 150000


def pose_errors(out, T):
    t_gt = T[:, :3, 3]
    t_err = (out[1].detach().cpu() - t_gt).norm(dim=-1)
    q_gt = out[9].detach().cpu()
    q_err = (out[0].detach().cpu() - q_gt).norm(dim=-1)
    return t_err, q_err, t_gt.norm(dim=-1)


@pytest.mark.skipif(not os.path.exists(REF + ".index"), reason="reference checkpoint not present")
def test_checkpoint_matches_inventory_and_regresses_motion_on_cpu(elo):
    ck = importlib.import_module("efficientlo-net_b200.tf_checkpoint")
    P, step = ck.load_reference_checkpoint(REF)
    mine = elo.params.init_params(0)
    assert step == 945400
    assert set(P) == set(mine) and all(tuple(P[k].shape) == tuple(mine[k].shape) for k in mine)
    assert elo.params.num_parameters(P) == 899134            # 899 135 trainable in the graph minus the step counter
    assert abs(float(P["w_x"]) + 3.67) < 0.01 and abs(float(P["w_q"]) + 6.41) < 0.01
    pc, T = elo.synth.synth_batch(1, H, W, N, seed0=0)
    eye = torch.eye(4)[None]
    out = go.get_model(pc, H, W, T, eye, eye, P, elo.params.make_perms(0))
    t_err, q_err, t_norm = pose_errors(out, T)
    # 0.68 m of forward motion, 12 mrad of yaw: recovered to a few centimetres / a few mrad
    assert float(t_err) < 0.06 * float(t_norm) + 0.02, (t_err, t_norm)
    assert float(q_err) < 5e-3, q_err


@pytest.mark.gpu
def test_trained_weights_on_gpu_match_oracle_and_regress_motion(elo, cuda):
    P = make_ref_params.load()
    if P is None:
        pytest.skip("oracle/_ref/pretrained_params.npz not built (needs /root/reference at build time)")
    perms = elo.params.make_perms(3)
    pc, T = elo.synth.synth_batch(2, H, W, N, seed0=1)
    eye = torch.eye(4).expand(2, 4, 4).contiguous()
    want = go.get_model(pc, H, W, T, eye, eye, P, perms)
    got = elo.get_model(pc.to(cuda), H, W, T.to(cuda), None, None, False, params=elo.ParamStore(P, cuda), perms=perms)
    torch.cuda.synchronize()
    for name, g, w in zip("l0_q l0_t l1_q l1_t l2_q l2_t l3_q l3_t".split(), got, want):
        err = (g.cpu().double() - w.double()).abs().max().item()
        assert err <= 1e-4 * w.abs().max().item() + 5e-5, "%s differs by %.3g" % (name, err)
    t_err, q_err, t_norm = pose_errors(got, T)
    assert bool((t_err < 0.06 * t_norm + 0.02).all()) and bool((q_err < 5e-3).all()), (t_err, q_err)


@pytest.mark.skipif(not os.path.exists(REF + ".index"), reason="reference checkpoint not present")
def test_eval_kitti_cli_loads_the_tf_checkpoint_from_directory_or_prefix(capsys):
    """tools/eval_kitti.py --checkpoint takes the checkpoint directory (its `checkpoint` file names a file that is
    not shipped, so the lone *.index decides) or the prefix; both load all 560 tensors."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    import eval_kitti
    for ck in (os.path.dirname(REF), REF):
        eval_kitti.main(["--data_root", "/nonexistent", "--checkpoint", ck, "--check_only"])
        assert "560 tensors, 914046 values" in capsys.readouterr().out      # weights + BN moving statistics
    with pytest.raises(FileNotFoundError):
        eval_kitti.load_params("/nonexistent/dir")
