"""ctypes doors onto the index-op checkers (TEST INFRASTRUCTURE -- see oracle/__init__.py).

``port``     : oracle/libelo_oracle.so  -- C restatement, oracle/fused_conv_oracle.c
``ref_cpu``  : oracle/_ref/libref_cpu.so -- reference kernel bodies as host C++ (oracle/ref_host_shim.h)
``ref_gpu``  : oracle/_ref/libref_gpu.so -- reference .cu compiled unmodified for sm_100a (device pointers)

All three take the reference op's 14 arguments (tf_ops/2d_conv_select_k/fused_conv_select_k.py:14)
and return its 4 outputs with the real extents of fused_conv.cpp:127-136.
"""
import ctypes
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_PORT = os.path.join(HERE, "libelo_oracle.so")
_REF_CPU = os.path.join(HERE, "_ref", "libref_cpu.so")
_REF_GPU = os.path.join(HERE, "_ref", "libref_gpu.so")

_c_int, _c_float, _c_void_p = ctypes.c_int, ctypes.c_float, ctypes.c_void_p
_COMMON = [_c_int] * 8 + [_c_float, _c_int, _c_int] + [_c_void_p] * 8 + [_c_int, _c_int]

_libs = {}


def build(ref=True):
    """(Re)build the checkers with oracle/Makefile.  `ref` targets are skipped without /root/reference."""
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    if ref and os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def _load(path):
    if path not in _libs:
        if not os.path.exists(path):
            if path == _PORT:
                build(ref=False)
            else:
                raise FileNotFoundError(path + " (run `make -C oracle ref` where /root/reference exists)")
        _libs[path] = ctypes.CDLL(path)
    return _libs[path]


def have_ref_cpu():
    return os.path.exists(_REF_CPU)


def have_ref_gpu():
    return os.path.exists(_REF_GPU)


def small_hw(H, W, stride_h, stride_w):
    """fused_conv.cpp:114-115"""
    return math.ceil(H / float(stride_h)), math.ceil(W / float(stride_w))


def _prep(xyz1, xyz2, idx_n2, random_hw):
    xyz1 = np.ascontiguousarray(xyz1, dtype=np.float32)
    xyz2 = np.ascontiguousarray(xyz2, dtype=np.float32)
    idx_n2 = np.ascontiguousarray(idx_n2, dtype=np.int32)
    random_hw = np.ascontiguousarray(random_hw, dtype=np.int32)
    return xyz1, xyz2, idx_n2, random_hw


def _alloc(B, npoints, kt, K, fill=None):
    mk = (lambda s, d: np.empty(s, d)) if fill is None else (lambda s, d: np.full(s, fill, d))
    return (mk((B, npoints, K, 3), np.int32), mk((B, npoints, kt, 1), np.float32),
            mk((B, npoints, kt, 1), np.float32), mk((B, npoints, K, 1), np.float32))


def _p(a):
    return a.ctypes.data_as(_c_void_p)


def port(mode, xyz1, xyz2, idx_n2, random_hw, H, W, npoints, kernel_size_H, kernel_size_W, K,
         flag_copy, distance, stride_h, stride_w, nthreads=1):
    """mode: 'select' | 'random'.  This repo's C restatement."""
    lib = _load(_PORT)
    fn = lib.elo_oracle_fused_conv_mt
    fn.argtypes = [_c_int] + _COMMON + [_c_int]
    fn.restype = _c_int
    xyz1, xyz2, idx_n2, random_hw = _prep(xyz1, xyz2, idx_n2, random_hw)
    B = xyz1.shape[0]
    h2, w2 = xyz2.shape[1], xyz2.shape[2]
    outs = _alloc(B, npoints, kernel_size_H * kernel_size_W, K)
    rc = fn(0 if mode == "select" else 1, B, H, W, npoints, kernel_size_H, kernel_size_W, K,
            flag_copy, distance, stride_h, stride_w, _p(xyz1), _p(xyz2), _p(idx_n2), _p(random_hw),
            _p(outs[0]), _p(outs[1]), _p(outs[2]), _p(outs[3]), h2, w2, nthreads)
    if rc:
        raise RuntimeError("elo_oracle_fused_conv rc=%d" % rc)
    return outs


def ref_cpu(mode, xyz1, xyz2, idx_n2, random_hw, H, W, npoints, kernel_size_H, kernel_size_W, K,
            flag_copy, distance, stride_h, stride_w, block_threads=1, omp_threads=1):
    """The reference's own kernel body, host-compiled; <<<B, block_threads>>> emulated by loops."""
    lib = _load(_REF_CPU)
    fn = lib.ref_cpu_select_k if mode == "select" else lib.ref_cpu_random_k
    fn.argtypes = _COMMON + [_c_int, _c_int]
    fn.restype = _c_int
    xyz1, xyz2, idx_n2, random_hw = _prep(xyz1, xyz2, idx_n2, random_hw)
    B = xyz1.shape[0]
    h2, w2 = xyz2.shape[1], xyz2.shape[2]
    outs = _alloc(B, npoints, kernel_size_H * kernel_size_W, K)
    fn(B, H, W, npoints, kernel_size_H, kernel_size_W, K, flag_copy, distance, stride_h, stride_w,
       _p(xyz1), _p(xyz2), _p(idx_n2), _p(random_hw), _p(outs[0]), _p(outs[1]), _p(outs[2]),
       _p(outs[3]), h2, w2, block_threads, omp_threads)
    return outs


def ref_gpu(mode, xyz1, xyz2, idx_n2, random_hw, H, W, npoints, kernel_size_H, kernel_size_W, K,
            flag_copy, distance, stride_h, stride_w, outs=None, sync=True):
    """The reference's own CUDA kernel on the current device (torch CUDA tensors in and out).
    Launches on the legacy default stream exactly as the reference does, so synchronise around it."""
    import torch
    lib = _load(_REF_GPU)
    fn = lib.ref_gpu_select_k if mode == "select" else lib.ref_gpu_random_k
    fn.argtypes = _COMMON + [_c_int]
    fn.restype = _c_int
    xyz1 = xyz1.contiguous().float()
    xyz2 = xyz2.contiguous().float()
    idx_n2 = idx_n2.contiguous().int()
    random_hw = random_hw.contiguous().int()
    B = xyz1.shape[0]
    h2, w2 = xyz2.shape[1], xyz2.shape[2]
    kt = kernel_size_H * kernel_size_W
    dev = xyz1.device
    if outs is None:
        outs = (torch.empty((B, npoints, K, 3), dtype=torch.int32, device=dev),
                torch.empty((B, npoints, kt, 1), dtype=torch.float32, device=dev),
                torch.empty((B, npoints, kt, 1), dtype=torch.float32, device=dev),
                torch.empty((B, npoints, K, 1), dtype=torch.float32, device=dev))
    if sync:
        torch.cuda.synchronize(dev)
    rc = fn(B, H, W, npoints, kernel_size_H, kernel_size_W, K, flag_copy, distance, stride_h,
            stride_w, xyz1.data_ptr(), xyz2.data_ptr(), idx_n2.data_ptr(), random_hw.data_ptr(),
            outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), outs[3].data_ptr(), h2, w2, 1)
    if rc:
        raise RuntimeError("reference launcher: cudaError %d" % rc)
    if sync:
        torch.cuda.synchronize(dev)
    return outs
