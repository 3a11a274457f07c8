"""GPU: the tcgen05 (3xTF32) dense layer against a float64 matmul -- pins the TMEM / shared-memory
descriptor conventions of csrc/elo_tc.cuh.  The hook lives in tests/csrc/ (its own library, not the product's)."""
import ctypes
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N", [(16, 64), (32, 64), (48, 128), (64, 64), (128, 128), (144, 128), (192, 128), (128, 64)])
def test_tc_dense_matches_float64(elo, cuda, K, N):
    g = torch.Generator().manual_seed(K * 1000 + N)
    X = torch.randn(128, K, generator=g)
    W = torch.randn(K, N, generator=g) * 0.2
    b = torch.randn(N, generator=g) * 0.1
    Xd, Wd, bd = X.to(cuda), W.to(cuda), b.to(cuda)
    Y = torch.full((128, N), float("nan"), device=cuda)
    lib = ctypes.CDLL(importlib.import_module("efficientlo-net_b200.build").build_test_lib())
    lib.elo_tc_dense_test.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 3 + [ctypes.c_void_p]
    rc = lib.elo_tc_dense_test(Xd.data_ptr(), Wd.data_ptr(), bd.data_ptr(), Y.data_ptr(), K, N, 1,
                               torch.cuda.current_stream().cuda_stream)
    assert rc == 0, "elo_tc_dense_test rc=%d" % rc
    torch.cuda.synchronize()
    want = torch.relu(X.double() @ W.double() + b.double())
    err = (Y.cpu().double() - want).abs().max().item()
    scale = want.abs().max().item()
    print("K=%d N=%d max abs err %.3e (scale %.2f)" % (K, N, err, scale))
    assert err <= 5e-6 * scale + 1e-6, "tensor-core layer off by %.3e" % err
