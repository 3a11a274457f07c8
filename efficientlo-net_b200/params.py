"""Parameter inventory of the PWCLO network, under the reference's TensorFlow variable names.

The reference creates variables implicitly inside ``tf.variable_scope``s (utils/tf_util.py:120-185,
52-115, 512-531); the shipped checkpoint lists them (383 trainable variables,
899 135 parameters).  Here they live in a flat ``dict name -> tensor``:

    <scope>/<layer>/weights            (Cin, Cout)   1x1 conv kernel, stored without the [1,1] dims
    <scope>/<layer>/biases             (Cout,)
    <scope>/<layer>/bn/{gamma,beta,moving_mean,moving_variance}   (Cout,)   conv2d layers only
    w_x, w_q                           scalars of the loss (main.py:151-152)

``fold_bn`` turns a conv2d layer into the (W', b') the inference kernels consume
(y = relu(x W' + b'), W' = W s, b' = (b - mean) s + beta, s = gamma / sqrt(var + 1e-3)).
"""
import math
from collections import OrderedDict

import torch

BN_EPS = 1e-3   # tf.contrib.layers.batch_norm default, FusedBatchNorm attr in the shipped graph


def mlp_layers(scope, names, cin, widths):
    """[(tf scope of the layer, cin, cout)] for a chain of 1x1 convs."""
    out = []
    for name, cout in zip(names, widths):
        out.append(("%s/%s" % (scope, name), cin, cout))
        cin = cout
    return out


def conv2d_layers():
    """Every conv2d(+BN+ReLU) layer of the model as (tf scope, Cin, Cout) -- pwclo_model.py:117-401."""
    L = []
    feat = [3, 16, 32, 64, 128]
    # siamese feature pyramid (weights shared by both frames, pwclo_model.py:143)
    for i, widths in enumerate([(8, 8, 16), (16, 16, 32), (32, 32, 64), (64, 64, 128)]):
        L += mlp_layers("sa1/layer%d" % i, ["conv0", "conv1", "conv2"], 3 + feat[i], widths)
    for scope, C in (("flow_embedding_l2_origin", 64), ("flow_embedding_l2", 64),
                     ("flow_embedding_l1", 32), ("flow_embedding_l0", 16)):
        L += mlp_layers(scope, ["CV_0", "CV_1", "CV_2"], 10 + 2 * C, (128, 64, 64))
        L += mlp_layers(scope, ["CV_xyz"], 10, (64,))
        L += mlp_layers(scope, ["sum_CV_0", "sum_CV_1"], 128, (128, 64))
        L += mlp_layers(scope, ["sum_xyz_encoding"], 10, (64,))
        L += mlp_layers(scope, ["sum_cost_volume_0", "sum_cost_volume_1"], 128 + C, (128, 64))
    L += mlp_layers("new_layer3", ["conv0", "conv1", "conv2"], 3 + 64, (128, 64, 64))
    L += mlp_layers("l3_costvolume_predict_ww", ["conv_predictor0", "conv_predictor1"], 128 + 64, (128, 64))
    for lvl, C in ((2, 64), (1, 32), (0, 16)):
        for kind in ("w", "costvolume"):
            scope = "up_sa_layer_layer_l%d%s" % (lvl, kind)
            L += mlp_layers(scope, ["up_1_0", "up_1_1"], 3 + 64, (128, 64))
            L += mlp_layers(scope, ["up_2_0", "up_2_1"], 64 + C, (128, 64))
        for scope in ("l%d_costvolume_predict" % lvl, "l%d_w_predict" % lvl):
            L += mlp_layers(scope, ["conv_predictor0", "conv_predictor1"], C + 64 + 64, (128, 64))
    return L


def conv1d_layers():
    """The pose heads: conv1d without BN or activation (pwclo_model.py:197-205, 264-275, ...)."""
    L = []
    for lvl in (3, 2, 1, 0):
        L.append(("l%d_big" % lvl, 64, 256))
        suffix = "coarse" if lvl == 3 else "det"
        L.append(("l%d_q_%s" % (lvl, suffix), 256, 4))
        L.append(("l%d_t_%s" % (lvl, suffix), 256, 3))
    return L


def init_params(seed=0, randomize=True, dtype=torch.float32, pose_near_identity=True):
    """Random-init weights of the reference architecture.

    Convolution kernels follow utils/tf_util.py:42 (Xavier uniform).  With ``randomize`` the biases
    and batch-norm statistics get non-trivial values (as a trained checkpoint has) so that BN
    folding is actually exercised; otherwise they take the reference's initial values
    (bias 0, gamma 1, beta 0, moving mean 0, moving variance 1).  ``pose_near_identity`` shrinks the
    12 pose-head layers and biases the quaternion heads towards (1,0,0,0), so that an untrained
    network regresses small motions (as a trained one does on KITTI) instead of arbitrary rotations
    that would fold the warped cloud onto a few cells of the re-projection."""
    g = torch.Generator().manual_seed(seed)
    P = OrderedDict()

    def xavier(cin, cout):
        lim = math.sqrt(6.0 / (cin + cout))
        return ((torch.rand(cin, cout, generator=g, dtype=torch.float64) * 2 - 1) * lim).to(dtype)

    def vec(n, lo, hi):
        return (torch.rand(n, generator=g, dtype=torch.float64) * (hi - lo) + lo).to(dtype)

    for scope, cin, cout in conv2d_layers():
        P[scope + "/weights"] = xavier(cin, cout)
        if randomize:
            P[scope + "/biases"] = vec(cout, -0.1, 0.1)
            P[scope + "/bn/gamma"] = vec(cout, 0.7, 1.3)
            P[scope + "/bn/beta"] = vec(cout, -0.1, 0.1)
            P[scope + "/bn/moving_mean"] = vec(cout, -0.2, 0.2)
            P[scope + "/bn/moving_variance"] = vec(cout, 0.5, 1.5)
        else:
            P[scope + "/biases"] = torch.zeros(cout, dtype=dtype)
            P[scope + "/bn/gamma"] = torch.ones(cout, dtype=dtype)
            P[scope + "/bn/beta"] = torch.zeros(cout, dtype=dtype)
            P[scope + "/bn/moving_mean"] = torch.zeros(cout, dtype=dtype)
            P[scope + "/bn/moving_variance"] = torch.ones(cout, dtype=dtype)
    for scope, cin, cout in conv1d_layers():
        P[scope + "/weights"] = xavier(cin, cout)
        P[scope + "/biases"] = vec(cout, -0.05, 0.05) if randomize else torch.zeros(cout, dtype=dtype)
        if pose_near_identity and cout in (3, 4):
            P[scope + "/weights"] = P[scope + "/weights"] * 0.02
            P[scope + "/biases"] = P[scope + "/biases"] * 0.2
            if cout == 4:
                P[scope + "/biases"][0] = 1.0
    P["w_x"] = torch.tensor(0.0, dtype=dtype)     # main.py:151
    P["w_q"] = torch.tensor(-2.5, dtype=dtype)    # main.py:152
    return P


def num_parameters(P):
    """Trainable parameter count (weights, biases, gamma, beta, w_x, w_q): 899 135 for the reference."""
    return sum(v.numel() for k, v in P.items() if not k.endswith(("moving_mean", "moving_variance")))


def fold_bn(P, scope):
    """(W', b') of one conv2d+BN layer for inference, in float64 then cast back."""
    w = P[scope + "/weights"].double()
    b = P[scope + "/biases"].double()
    s = P[scope + "/bn/gamma"].double() / torch.sqrt(P[scope + "/bn/moving_variance"].double() + BN_EPS)
    wf = w * s[None, :]
    bf = (b - P[scope + "/bn/moving_mean"].double()) * s + P[scope + "/bn/beta"].double()
    dt = P[scope + "/weights"].dtype
    return wf.to(dt), bf.to(dt)


# scan-order permutations: one per custom-op call site (the reference draws tf.random_shuffle at each
# call, utils/pointnet_util.py:45,104,193,270, so parity tests have to feed them explicitly)
def perm_sites():
    """[(site name, kernel_total)] in graph construction order."""
    S = []
    for f in ("f1", "f2"):
        S += [("sa1/layer0/" + f, 9 * 15), ("sa1/layer1/" + f, 7 * 11), ("sa1/layer2/" + f, 5 * 9),
              ("sa1/layer3/" + f, 5 * 9)]
    S += [("flow_embedding_l2_origin/q", 5 * 35), ("flow_embedding_l2_origin/p", 3 * 5), ("new_layer3", 5 * 9)]
    for lvl, kq in ((2, 5 * 15), (1, 7 * 25), (0, 11 * 41)):
        S += [("flow_embedding_l%d/q" % lvl, kq), ("flow_embedding_l%d/p" % lvl, 3 * 5),
              ("up_sa_layer_layer_l%dw" % lvl, 7 * 15), ("up_sa_layer_layer_l%dcostvolume" % lvl, 7 * 15)]
    return S


def make_perms(seed=0):
    g = torch.Generator().manual_seed(seed)
    return OrderedDict((name, torch.randperm(kt, generator=g).to(torch.int32)) for name, kt in perm_sites())
