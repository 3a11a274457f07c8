"""CPU: pin the C restatement (oracle/fused_conv_oracle.c) of the two index ops.

1. known answers derived from the reference's own __main__ demo
   (tf_ops/2d_conv_select_k/fused_conv_select_k.py:93-145);
2. the unstable-tie case of the selection sort;
3. bit-exact agreement with the reference's kernel bodies compiled as host C++
   (oracle/_ref/libref_cpu.so) on randomised cases, the model's real call signatures, and
4. the committed golden vectors (tests/golden/index_golden.npz, written by
   tests/golden/make_index_golden.py from the reference's own code).
"""
import os

import numpy as np
import pytest

import cases
from oracle import index_oracle as io

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "index_golden.npz")
needs_ref = pytest.mark.skipif(not io.have_ref_cpu(), reason="oracle/_ref/libref_cpu.so not built")

IMPLS = [io.port] + ([io.ref_cpu] if io.have_ref_cpu() else [])


def demo_inputs():
    H, W = 4, 7
    xyz = np.tile(np.arange(H * W, dtype=np.float32).reshape(1, H, W, 1), (1, 1, 1, 3))
    idx = np.array([[[0, 0], [0, 1]]], np.int32)
    return H, W, xyz, idx


@pytest.mark.parametrize("impl", IMPLS, ids=lambda f: f.__name__)
@pytest.mark.parametrize("mode,rhw,cols", [
    ("select", [0, 1, 2, 3, 4], [1, 2, 3, 6]),
    ("select", [3, 0, 4, 2, 1], [1, 2, 3, 6]),
    ("random", [0, 1, 2, 3, 4], [6, 1, 2, 3]),
    ("random", [3, 0, 4, 2, 1], [2, 6, 3, 1]),
])
def test_demo_known_answers(impl, mode, rhw, cols):
    H, W, xyz, idx = demo_inputs()
    sel, valid, vdis, mask = impl(mode, xyz, xyz, idx, np.array(rhw, np.int32), H, W, 2, 1, 5, 8, 0, 200.0, 1, 1)
    assert sel.shape == (1, 2, 8, 3) and valid.shape == (1, 2, 5, 1) and mask.shape == (1, 2, 8, 1)
    # query (0,0) is the point (0,0,0): invalid centre, everything stays zero
    assert not sel[0, 0].any() and not valid[0, 0].any() and not vdis[0, 0].any() and not mask[0, 0].any()
    # query (0,1): columns {6 (wrap), 0 (empty), 1, 2, 3}
    assert sel[0, 1, :, 2].tolist() == cols + [0, 0, 0, 0]
    assert sel[0, 1, :, 0].tolist() == [0] * 8 and sel[0, 1, :, 1].tolist() == [0] * 8
    assert mask[0, 1, :, 0].tolist() == [1, 1, 1, 1, 0, 0, 0, 0]
    assert valid[0, 1, :, 0].tolist() == [1, 1, 1, 1, 0]
    assert vdis[0, 1, :, 0].tolist() == [1, 1, 1, 1, 0]


@pytest.mark.parametrize("impl", IMPLS, ids=lambda f: f.__name__)
def test_select_k_tie_order_is_the_swap_order(impl):
    """Scan order cols 0,2,1 gives Dist=[1,1,1e-10]; the swap of step 0 moves col 0 behind col 2."""
    W = 8
    xyz = np.zeros((1, 1, W, 3), np.float32)
    xyz[0, 0, :, 0] = [11, 10, 9, 50, 60, 70, 80, 90]
    xyz[0, 0, :, 1] = 1
    idx = np.array([[[0, 1]]], np.int32)
    sel, _, _, mask = impl("select", xyz, xyz, idx, np.array([0, 2, 1], np.int32), 1, W, 1, 1, 3, 3, 0, 1000.0, 1, 1)
    assert sel[0, 0, :, 2].tolist() == [1, 2, 0]
    assert mask[0, 0, :, 0].tolist() == [1, 1, 1]


@pytest.mark.parametrize("impl", IMPLS, ids=lambda f: f.__name__)
def test_flag_copy_quirks(impl):
    H, W = 3, 9
    rng = np.random.default_rng(5)
    xyz1 = cases.range_image(rng, 1, H, W, holes=0.0)
    far = xyz1 + 500.0     # nothing within distance
    idx = cases.all_cells(1, H, W)
    rhw = rng.permutation(9).astype(np.int32)
    # select-K: flag_copy with no in-range neighbour -> index (b,0,0) with mask ONE (reference :180-192)
    sel, _, vdis, mask = impl("select", xyz1, far, idx, rhw, H, W, H * W, 3, 3, 4, 1, 1.0, 1, 1)
    assert not sel.any() and mask.min() == 1.0 and not vdis.any()
    # random-K: the copy only happens at the first accepted neighbour -> all zero
    sel, _, _, mask = impl("random", xyz1, far, idx, rhw, H, W, H * W, 3, 3, 4, 1, 1.0, 1, 1)
    assert not sel.any() and not mask.any()
    # random-K with neighbours: tail slots repeat the first accepted one
    sel, _, vdis, mask = impl("random", xyz1, xyz1, idx, rhw[:1] * 0, H, W, H * W, 1, 1, 4, 1, 1.0, 1, 1)
    assert mask.min() == 1.0 and vdis[..., 0].sum() == H * W
    assert (sel[:, :, 1:, :] == sel[:, :, :1, :]).all()


@needs_ref
@pytest.mark.parametrize("seed", range(160))
def test_port_equals_reference_body_random(seed):
    rng = np.random.default_rng(seed)
    case = cases.random_case(rng)
    cases.assert_same(cases.call(io.port, case), cases.call(io.ref_cpu, case), "seed %d" % seed)


@needs_ref
@pytest.mark.parametrize("site", cases.MODEL_SITES, ids=lambda s: s[0])
def test_port_equals_reference_body_model_sites(site):
    rng = np.random.default_rng(abs(hash(site[0])) % (2 ** 31))
    case = cases.site_case(rng, site, B=2)
    a = cases.call(io.port, case, nthreads=4)
    b = cases.call(io.ref_cpu, case, block_threads=16, omp_threads=4)
    cases.assert_same(a, b, site[0])
    assert a[3].sum() > 0


def test_port_threads_agree():
    rng = np.random.default_rng(11)
    case = cases.site_case(rng, cases.MODEL_SITES[8], B=2)
    cases.assert_same(cases.call(io.port, case, nthreads=1), cases.call(io.port, case, nthreads=8))


def test_golden_vectors():
    g = np.load(GOLDEN)
    n = int(g["n_cases"])
    assert n >= 8
    for i in range(n):
        p = "c%d_" % i
        case = dict(mode=str(g[p + "mode"]), xyz1=g[p + "xyz1"], xyz2=g[p + "xyz2"], idx_n2=g[p + "idx_n2"],
                    random_hw=g[p + "random_hw"])
        for k, v in zip(("H", "W", "npoints", "kernel_size_H", "kernel_size_W", "K", "flag_copy",
                         "stride_h", "stride_w"), g[p + "ints"].tolist()):
            case[k] = int(v)
        case["distance"] = float(g[p + "distance"])
        got = cases.call(io.port, case)
        want = tuple(g[p + name] for name in cases.OUT_NAMES)
        cases.assert_same(got, want, "golden %d" % i)
