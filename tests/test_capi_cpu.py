"""CPU: the C-ABI library builds, loads, and exports every symbol include/*.h declares.
No compute call is made here (no GPU in the build container)."""
import ctypes
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names += re.findall(r"\b(elo_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_declares_the_reference_seam():
    names = declared_symbols()
    assert "elo_fused_conv_select_k" in names and "elo_fused_conv_random_k" in names


def test_library_exports_every_declared_symbol(elo):
    from importlib import import_module
    build = import_module("efficientlo-net_b200.build")
    path = build.build()
    assert os.path.exists(path)
    handle = ctypes.CDLL(path)
    missing = [n for n in declared_symbols() if not hasattr(handle, n)]
    assert not missing, "declared in include/ but not exported: %s" % missing
    assert elo._lib.lib().elo_version() >= 100
    # every bound signature refers to an exported function
    for name in elo._lib.SIGNATURES:
        assert hasattr(handle, name), name


def test_product_path_refuses_cpu_tensors(elo):
    import torch
    z = torch.zeros(1, 4, 7, 3)
    with pytest.raises(elo._lib.EloError):
        elo.fused_conv_select_k(z, z, torch.zeros(1, 2, 2, dtype=torch.int32), torch.arange(5, dtype=torch.int32),
                                4, 7, 2, 1, 5, 8, 0, 200.0, 1, 1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "efficientlo-net_b200")
    for path in glob.glob(os.path.join(pkg, "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
            text = open(path).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), path
            assert "libelo_oracle" not in text and "oracle/_ref" not in text, path


def test_zero_widened_chain_computes_the_same_function(elo):
    """packing.folded_chain(pad_hidden=64): the 32-32-64 set-conv of pyramid layer 2 runs on the 64-/128-wide GEMM engines
    with its hidden layers widened by zero columns (zero bias) and the next layer by matching zero rows -- the chain of
    ReLU(x W + b) layers must give the same outputs, bit for bit where the added terms are exact zeros."""
    import torch
    P = elo.params.init_params(0)
    scopes = ["sa1/layer2/conv%d" % j for j in range(3)]
    plain = elo.packing.folded_chain(P, scopes)
    wide = elo.packing.folded_chain(P, scopes, pad_hidden=64)
    assert [tuple(w.shape) for w, _ in plain] == [(35, 32), (32, 32), (32, 64)]
    assert [tuple(w.shape) for w, _ in wide] == [(35, 64), (64, 64), (64, 64)]
    x = torch.randn(257, 35, generator=torch.Generator().manual_seed(0))

    def run(chain):
        y = x.double()
        for w, b in chain:
            y = torch.relu(y @ w.double() + b.double())
        return y

    assert torch.equal(run(plain), run(wide)[:, :64])
    # the widened stream is what the engines take: 64-wide layers only
    assert elo.packing.pack_stream(P, scopes, pad_hidden=64).numel() % 64 == 0
