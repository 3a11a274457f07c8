"""GPU: the training-mode graph (batch-statistics batch norm + autograd, train_graph.py) against the
torch-CPU restatement run in its `training` mode: forward outputs, the moving averages the step
assigns, the loss and the gradient of every parameter tensor.  Dropout is random in the reference
(tf.layers.dropout) and therefore disabled on both sides here; it is checked separately."""
import pytest
import torch

from oracle import graph_oracle as go

pytestmark = pytest.mark.gpu
H_IN, W_IN, NPTS = 64, 1800, 150000


@pytest.fixture(scope="module")
def pair(elo, cuda):
    P = elo.params.init_params(0)
    perms = elo.params.make_perms(0)
    B = 2
    pc, T = elo.synth.synth_batch(B, H_IN, W_IN, NPTS)
    eye = torch.eye(4).expand(B, 4, 4).contiguous()
    # --- oracle, training mode, autograd on CPU
    Pc = {k: v.clone().requires_grad_(not k.endswith(("moving_mean", "moving_variance"))) for k, v in P.items()}
    w_x = torch.tensor(0.0, requires_grad=True)
    w_q = torch.tensor(-2.5, requires_grad=True)
    with go.training(bn_decay=0.5) as moving:
        out_c = go.get_model(pc, H_IN, W_IN, T, eye, eye, Pc, perms)
    loss_c = go.get_loss(*out_c[:8], out_c[9], out_c[10], w_x, w_q)
    loss_c.backward()
    # the same in fp64: the yardstick for the gradients (see test_gradients_match)
    P64 = {k: v.double().requires_grad_(not k.endswith(("moving_mean", "moving_variance"))) for k, v in P.items()}
    w64 = (torch.tensor(0.0, dtype=torch.float64, requires_grad=True), torch.tensor(-2.5, dtype=torch.float64, requires_grad=True))
    with go.training(bn_decay=0.5):
        out64 = go.get_model(pc, H_IN, W_IN, T, eye, eye, P64, perms, dtype=torch.float64)
    go.get_loss(*out64[:8], out64[9], out64[10], *w64).backward()
    # --- product, GPU
    from importlib import import_module
    tg = import_module("efficientlo-net_b200.train_graph")
    tp = tg.TrainableParams(P, cuda)
    api_out = elo.get_model(pc.to(cuda), H_IN, W_IN, T.to(cuda), None, None, True, bn_decay=0.5, params=tp,
                            perms=perms)
    # get_model(is_training=True) uses the reference's dropout 0.5; for the comparison run it with none
    tp2 = tg.TrainableParams(P, cuda)
    out_g = tg.get_model(pc.to(cuda), H_IN, W_IN, T.to(cuda), None, None, tp2, bn_decay=0.5, perms=perms, dropout=0.0)
    loss_g = elo.get_loss(*out_g[:8], out_g[9], out_g[10], tp2["w_x"], tp2["w_q"])
    loss_g.backward()
    return dict(P=P, Pc=Pc, moving=moving, out_c=out_c, loss_c=loss_c, w_c=(w_x, w_q), tp=tp2, out_g=out_g,
                loss_g=loss_g, tg=tg, api_out=api_out, P64=P64, w64=w64)


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_forward_outputs_match(pair):
    names = ["l0_q", "l0_t", "l1_q", "l1_t", "l2_q", "l2_t", "l3_q", "l3_t", "l0_xyz_f1", "q_gt", "t_gt"]
    for n, g, c in zip(names, pair["out_g"], pair["out_c"]):
        g, c = g.detach().cpu(), c.detach()
        assert g.shape == c.shape, n
        assert torch.allclose(g, c, rtol=1e-4, atol=2e-5), "%s: max |err| %.3g" % (n, float((g - c).abs().max()))
    assert abs(float(pair["loss_g"]) - float(pair["loss_c"])) <= 1e-4 * abs(float(pair["loss_c"]))


def test_reference_signature_with_dropout(pair):
    out = pair["api_out"]
    assert len(out) == 11 and out[0].requires_grad
    assert all(bool(torch.isfinite(o).all()) for o in out)
    # dropout acts on the 256-wide pose feature of every level: the poses differ from the dropout-free run
    assert not torch.equal(out[6], pair["out_g"][6])


def test_moving_averages_match(pair):
    tp, moving = pair["tp"], pair["moving"]
    assert len(moving) == 2 * 89                            # every batch-norm scope (101 sites, 12 shared by the two frames)
    worst = 0.0
    for name, want in moving.items():
        got = tp[name].detach().cpu()
        assert not torch.equal(got, pair["P"][name]), name + " was not updated"
        # fp32 noise grows through the chain of batch-stat normalisations (deep pyramid layers divide by small
        # standard deviations), hence looser than the 1e-4 of the outputs
        assert torch.allclose(got, want, rtol=1e-3, atol=1e-4), name
        worst = max(worst, rel(got, want))
    assert worst < 1e-3
    # a 456-row layer: the Bessel correction (n/(n-1) = 1.0022) of the variance handed to the moving average is visible
    name = "l3_costvolume_predict_ww/conv_predictor0/bn/moving_variance"
    assert rel(tp[name], moving[name]) < 5e-4


def test_gradients_match(pair):
    """At this random initialisation the gradient is ill-conditioned in fp32: the restatement's OWN fp32 gradient
    is up to ~2e-2 (relative, per tensor) away from its fp64 gradient in the deepest layers, 1e-4 at the heads.  So
    the yardstick is the fp64 gradient, and the GPU graph has to be as close to it as the fp32 restatement is
    (factor 3 + a 1e-3 floor), tensor by tensor."""
    tp, Pc, P64 = pair["tp"], pair["Pc"], pair["P64"]
    checked, worst = 0, 0.0
    for name, p in tp.named_parameters():
        if name in ("w_x", "w_q"):
            i = 0 if name == "w_x" else 1
            cpu32, want = pair["w_c"][i].grad, pair["w64"][i].grad
        else:
            cpu32, want = Pc[name].grad, P64[name].grad
        assert p.grad is not None and want is not None, name
        if name.endswith("/biases") and name.replace("/biases", "/bn/gamma") in P64:
            # a bias in front of a batch norm: its true gradient is exactly zero, what is left is rounding noise
            assert float(p.grad.norm()) < 1e-3, name
            continue
        budget = 3 * rel(cpu32, want) + 1e-3
        r = rel(p.grad, want)
        worst = max(worst, r)
        assert r < budget, "%s: relative gradient error %.3g, fp32 restatement %.3g (|g| %.3g)" % (
            name, r, rel(cpu32, want), float(want.norm()))
        checked += 1
    assert checked > 250 and worst < 0.1


def test_dropout_scales_and_zeroes(pair, cuda):
    tg = pair["tg"]
    net = tg._Net(pair["tp"], None, 0.5, torch.Generator(device=cuda).manual_seed(1))
    x = torch.ones(4, 1, 256, device=cuda)
    y = net.drop(x)
    vals = set(y.unique().tolist())
    assert vals == {0.0, 2.0}
    assert 0.35 < float((y == 0).float().mean()) < 0.65


def test_trainer_reduces_loss_and_exports(elo, cuda, pair):
    tg = pair["tg"]
    torch.manual_seed(0)
    B = 2
    pc, T = elo.synth.synth_batch(B, H_IN, W_IN, NPTS)
    tp = tg.TrainableParams(pair["P"], cuda)
    tr = tg.Trainer(tp, batch_size=B, dropout=0.0, base_lr=3e-4)
    perms = elo.params.make_perms(3)
    losses = [float(tr.step(pc.to(cuda), T.to(cuda), perms=perms)) for _ in range(8)]
    assert all(l == l for l in losses)
    # a random-init network on one repeated batch: the trajectory is bumpy (the loss carries the learnable
    # uncertainty weights), but it gets below where it started within a few steps
    assert min(losses[2:]) < losses[0], losses
    assert tr.batch == 8
    # the trained weights drop straight into the fused inference path
    store = elo.ParamStore(tp.export(), cuda)
    out = elo.get_model(pc.to(cuda), H_IN, W_IN, T.to(cuda), None, None, False, params=store, perms=perms)
    assert all(bool(torch.isfinite(o).all()) for o in out[:8])


def test_training_needs_trainable_params(elo, cuda, pair):
    pc, T = elo.synth.synth_batch(1, H_IN, W_IN, NPTS)
    with pytest.raises(TypeError):
        elo.get_model(pc.to(cuda), H_IN, W_IN, T.to(cuda), None, None, True, params=elo.ParamStore(pair["P"], cuda))


def test_graphed_trainer_follows_the_eager_trainer(elo, cuda, pair):
    """Trainer(use_graph=True): forward + loss + backward + Adam replayed as ONE CUDA graph.  From the same start
    (dropout off, same scan orders) the first loss is the eager trainer's to the last bit -- the warm-up passes of
    the capture leave parameters, moving averages and Adam slots untouched -- and after two updates losses and
    parameters agree to rounding (capturable Adam orders its arithmetic differently; float atomics of the scatter
    backward reorder sums; with a random-init network at lr 1e-3 later steps amplify that, so they are not compared)."""
    tg = pair["tg"]
    B = 2
    pc, T = elo.synth.synth_batch(B, H_IN, W_IN, NPTS)
    pc, T = pc.to(cuda), T.to(cuda)
    perms = elo.params.make_perms(3)
    runs = []
    for use_graph in (False, True):
        tp = tg.TrainableParams(pair["P"], cuda)
        tr = tg.Trainer(tp, batch_size=B, dropout=0.0, use_graph=use_graph)
        losses = [float(tr.step(pc, T, perms=perms))]
        after_one = tp.export()
        losses.append(float(tr.step(pc, T, perms=perms)))
        torch.cuda.synchronize()
        assert tr.batch == 2
        runs.append((losses, after_one))
    (le, pe), (lg, pg) = runs
    assert abs(le[0] - lg[0]) <= 2e-6 * abs(le[0]), (le, lg)
    assert abs(le[1] - lg[1]) <= 1e-4 * abs(le[1]) + 1e-5, (le, lg)
    assert le[1] != le[0]
    # after ONE update: the batch-norm moving averages (functions of the initial weights only) agree to rounding, the
    # weights to Adam's step size (its update of a weight with a noise-level gradient is +-lr whatever the gradient)
    for name in ("sa1/layer0/conv0/bn/moving_mean", "flow_embedding_l0/CV_0/bn/moving_variance",
                 "l0_costvolume_predict/conv_predictor1/bn/moving_variance"):
        assert torch.allclose(pe[name], pg[name], rtol=1e-4, atol=1e-6), name
    w = "sa1/layer0/conv0/weights"
    assert float((pe[w] - pg[w]).abs().max()) <= 2.5e-3 and not torch.equal(pg[w], pair["P"][w])
