"""fused_conv_select_k / fused_conv_random_k -- the reference's two custom ops, on B200.

Mirrors the Python stubs of the reference (same 14 positional arguments, same 4-tuple result):
  tf_ops/2d_conv_select_k/fused_conv_select_k.py:14-29
  tf_ops/2d_conv_random_k/fused_conv_random_k.py:14-29
and the checks of the C++ op wrapper they load (tf_ops/*/fused_conv.cpp:77-123); InvalidArgument
becomes ValueError.  The work is done by hand-written sm_100a kernels behind the C ABI
(include/elo_b200.h); tensors must live on a CUDA device -- there is no CPU path.
"""
import math

import torch

from . import _lib


def _check_inputs(name, xyz1, xyz2, idx_n2, random_hw, npoints, kernel_size_H, kernel_size_W,
                  stride_h, stride_w):
    for t, label in ((xyz1, "xyz1"), (xyz2, "xyz2"), (idx_n2, "idx_n2"), (random_hw, "random_hw")):
        if not isinstance(t, torch.Tensor):
            raise TypeError("%s: %s must be a torch.Tensor" % (name, label))
        if not t.is_cuda:
            raise _lib.EloError("%s: %s is on %s; this op only runs on CUDA (no CPU fallback)"
                                % (name, label, t.device))
    if xyz1.dim() != 4 or xyz1.shape[3] != 3:                       # fused_conv.cpp:107
        raise ValueError("%s expects (batch_size, H, W, 3) xyz1 shape." % name)
    if stride_h <= 0 or stride_w <= 0:                              # fused_conv.cpp:96-100
        raise ValueError("FusedConv expects positive stride_h / stride_w")
    B, H, W = xyz1.shape[0], xyz1.shape[1], xyz1.shape[2]
    H2 = math.ceil(H / float(stride_h))                             # fused_conv.cpp:114-115
    if xyz2.dim() != 4 or xyz2.shape[1] != H2 or xyz2.shape[3] != 3 or xyz2.shape[0] != B:
        raise ValueError("%s expects (batch_size, H/stride_h, W/stride_w, 3) xyz2 shape." % name)
    if idx_n2.dim() != 3 or idx_n2.shape[2] != 2 or idx_n2.shape[1] != npoints or idx_n2.shape[0] != B:
        raise ValueError("FusedConv expects (batch_size, npoints, 2) idx_n2 shape.")   # :120
    if random_hw.dim() != 1 or random_hw.shape[0] != kernel_size_H * kernel_size_W:
        raise ValueError("FusedConv expects (kernel_size_h * kernel_size_w) random_hw shape.")  # :123
    return B, H, W, xyz2.shape[1], xyz2.shape[2]


def _run(select, xyz1, xyz2, idx_n2, random_hw, H, W, npoints, kernel_size_H, kernel_size_W, K,
         flag_copy, distance, stride_h, stride_w, want_valid=True):
    name = "FusedConvSelectK" if select else "FusedConvRandomK"
    npoints, kernel_size_H, kernel_size_W, K = int(npoints), int(kernel_size_H), int(kernel_size_W), int(K)
    flag_copy, stride_h, stride_w = int(flag_copy), int(stride_h), int(stride_w)
    B, H_in, W_in, h2, w2 = _check_inputs(name, xyz1, xyz2, idx_n2, random_hw, npoints,
                                          kernel_size_H, kernel_size_W, stride_h, stride_w)
    # the H / W attrs of the op are declared but never read by the reference kernel wrapper
    # (fused_conv.cpp:110-111 takes them from xyz1's shape); same here.
    if K <= 0:
        raise ValueError("FusedConv expects positive K")
    dev = xyz1.device
    xyz1 = xyz1.contiguous().float()
    xyz2 = xyz2.contiguous().float()
    idx_n2 = idx_n2.contiguous().to(torch.int32)
    random_hw = random_hw.contiguous().to(torch.int32)
    kt = kernel_size_H * kernel_size_W
    idx = torch.empty((B, npoints, K, 3), dtype=torch.int32, device=dev)
    mask = torch.empty((B, npoints, K, 1), dtype=torch.float32, device=dev)
    valid = vdis = None
    if want_valid:
        valid = torch.empty((B, npoints, kt, 1), dtype=torch.float32, device=dev)
        vdis = torch.empty((B, npoints, kt, 1), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        fn = _lib.lib().elo_fused_conv_select_k if select else _lib.lib().elo_fused_conv_random_k
        rc = fn(B, H_in, W_in, npoints, kernel_size_H, kernel_size_W, K, flag_copy, float(distance),
                stride_h, stride_w, xyz1.data_ptr(), xyz2.data_ptr(), idx_n2.data_ptr(),
                random_hw.data_ptr(), idx.data_ptr(), _lib.ptr(valid), _lib.ptr(vdis),
                mask.data_ptr(), h2, w2, _lib.stream_ptr(dev))
    _lib.check(rc, name)
    return idx, valid, vdis, mask


class _FusedConv(torch.autograd.Function):
    """Index ops carry no gradient in the reference (no RegisterGradient; masks are wrapped in
    tf.stop_gradient, utils/pointnet_util.py:54-55): all four outputs are non-differentiable."""

    @staticmethod
    def forward(ctx, select, xyz1, xyz2, idx_n2, random_hw, H, W, npoints, kernel_size_H,
                kernel_size_W, K, flag_copy, distance, stride_h, stride_w):
        outs = _run(select, xyz1, xyz2, idx_n2, random_hw, H, W, npoints, kernel_size_H,
                    kernel_size_W, K, flag_copy, distance, stride_h, stride_w)
        ctx.mark_non_differentiable(*outs)
        return outs

    @staticmethod
    def backward(ctx, *grads):
        return (None,) * 15


def fused_conv_select_k(xyz1, xyz2, idx_n2, random_hw, H, W, npoints, kernel_size_H, kernel_size_W,
                        K, flag_copy, distance, stride_h, stride_w):
    """K nearest (3-D distance) non-empty pixels of xyz2 inside a kH x kW window (cylindrical in W)
    around each query pixel.  Returns (selected_bhw_idx int32 (B,n,K,3), valid_idx f32 (B,n,kH*kW,1),
    valid_in_dis_idx f32 (B,n,kH*kW,1), selected_mask f32 (B,n,K,1)) -- fused_conv.cpp:127-136."""
    return _FusedConv.apply(True, xyz1, xyz2, idx_n2, random_hw, H, W, npoints, kernel_size_H,
                            kernel_size_W, K, flag_copy, distance, stride_h, stride_w)


def fused_conv_random_k(xyz1, xyz2, idx_n2, random_hw, H, W, npoints, kernel_size_H, kernel_size_W,
                        K, flag_copy, distance, stride_h, stride_w):
    """First K in-range non-empty pixels of xyz2 in the (shuffled) scan order `random_hw` of the
    kH x kW window around each query pixel.  Same outputs as fused_conv_select_k."""
    return _FusedConv.apply(False, xyz1, xyz2, idx_n2, random_hw, H, W, npoints, kernel_size_H,
                            kernel_size_W, K, flag_copy, distance, stride_h, stride_w)


def fused_conv_indices(select, xyz1, xyz2, idx_n2, random_hw, kernel_size_H, kernel_size_W, K,
                       flag_copy, distance, stride_h, stride_w):
    """Same kernels without the two kt-wide count outputs nobody in the model reads
    (utils/pointnet_util.py:49,106,197,272 discard them).  Returns (selected_bhw_idx, selected_mask)."""
    idx, _, _, mask = _run(select, xyz1, xyz2, idx_n2, random_hw, xyz1.shape[1], xyz1.shape[2],
                           idx_n2.shape[1], kernel_size_H, kernel_size_W, K, flag_copy, distance,
                           stride_h, stride_w, want_valid=False)
    return idx, mask
