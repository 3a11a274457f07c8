"""CPU: host logic of the training path -- schedules of main.py:120-138, the oracle's training-mode batch
norm against torch's own, and the flat gradient all-reduce on gloo with world size 2."""
import importlib
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import graph_oracle as go


def tg():
    return importlib.import_module("efficientlo-net_b200.train_graph")


def test_schedules():
    t = tg()
    assert t.get_learning_rate(0, 8) == 0.001
    assert t.get_learning_rate(24999, 8) == 0.001                 # 199 992 samples: still the first stair
    assert abs(t.get_learning_rate(25000, 8) - 0.0007) < 1e-12
    assert t.get_learning_rate(10 ** 7, 8) == 0.00001             # clipped
    assert t.get_bn_decay(0, 8) == 0.5
    assert t.get_bn_decay(25000, 8) == 0.75
    assert t.get_bn_decay(10 ** 7, 8) == 0.99


def test_oracle_training_batch_norm_is_torch_batch_norm():
    torch.manual_seed(0)
    x = torch.randn(4, 50, 8, 6)
    P = {"s/weights": torch.randn(6, 5), "s/biases": torch.randn(5), "s/bn/gamma": torch.rand(5) + 0.5,
         "s/bn/beta": torch.randn(5), "s/bn/moving_mean": torch.zeros(5), "s/bn/moving_variance": torch.ones(5)}
    with go.training(bn_decay=0.8) as moving:
        y = go.conv2d(x, P, "s")
    bn = torch.nn.BatchNorm1d(5, eps=1e-3, momentum=0.2)
    bn.weight.data, bn.bias.data = P["s/bn/gamma"].clone(), P["s/bn/beta"].clone()
    z = (x @ P["s/weights"] + P["s/biases"]).reshape(-1, 5)
    want = torch.relu(bn(z)).reshape(4, 50, 8, 5)
    assert torch.allclose(y, want, atol=1e-5)
    assert torch.allclose(moving["s/bn/moving_mean"], bn.running_mean, atol=1e-6)
    assert torch.allclose(moving["s/bn/moving_variance"], bn.running_var, atol=1e-6)     # Bessel-corrected, like TF


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a = torch.zeros(3, 2, requires_grad=True)
    b = torch.zeros(5, requires_grad=True)
    c = torch.zeros(1, requires_grad=True)              # never receives a gradient on rank 1
    a.grad = torch.full((3, 2), float(rank + 1))
    b.grad = torch.arange(5.0) * (rank + 1)
    if rank == 0:
        c.grad = torch.tensor([4.0])
    tg().all_reduce_gradients([a, b, c])
    results[rank] = (a.grad.clone(), b.grad.clone(), c.grad.clone())
    dist.destroy_process_group()


def test_gradient_all_reduce_gloo():
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), results), nprocs=2, join=True)
    for r in range(2):
        a, b, c = results[r]
        assert torch.equal(a, torch.full((3, 2), 1.5))
        assert torch.equal(b, torch.arange(5.0) * 1.5)
        assert torch.equal(c, torch.tensor([2.0]))


def test_slabbed_weight_gradient_is_the_plain_one():
    """train_graph._Linear: dW as a batched GEMM over row slabs + their sum equals autograd's x^T @ dy (fp32 rounding)."""
    t = tg()
    torch.manual_seed(0)
    for rows in (5, 4096, 10007, 70001):
        x = torch.randn(rows, 19, requires_grad=True)
        w = torch.randn(19, 16, requires_grad=True)
        b = torch.randn(16, requires_grad=True)
        up = torch.randn(rows, 16)
        (t._Linear.apply(x, w, b) * up).sum().backward()
        got = (x.grad.clone(), w.grad.clone(), b.grad.clone())
        x.grad = w.grad = b.grad = None
        ((x @ w + b) * up).sum().backward()
        for g, want in zip(got, (x.grad, w.grad, b.grad)):
            assert torch.allclose(g, want, rtol=1e-4, atol=1e-4 * float(want.abs().max()))
