"""ctypes binding of libelo_b200.so (the C ABI declared in include/elo_b200.h).

There is no fallback: if the CUDA library is missing or a call fails, an exception is raised.
"""
import ctypes
import os

from .build import LIB

# ELO_B200_LIB: another build of the same library (A/B timing of two builds on one box)
LIB = os.environ.get("ELO_B200_LIB") or LIB

_c_int, _c_float, _c_void_p = ctypes.c_int, ctypes.c_float, ctypes.c_void_p

_c_ll, _c_uint = ctypes.c_longlong, ctypes.c_uint


# ---- descriptor structs, field for field as in include/elo_b200.h --------------------------------
class Window(ctypes.Structure):
    _fields_ = [("kernel_size_H", _c_int), ("kernel_size_W", _c_int), ("K", _c_int), ("distance", _c_float),
                ("stride_h", _c_int), ("stride_w", _c_int), ("small_h", _c_int), ("small_w", _c_int),
                ("random_hw", _c_void_p)]


class Queries(ctypes.Structure):
    _fields_ = [("H", _c_int), ("W", _c_int), ("out_h", _c_int), ("out_w", _c_int),
                ("q_stride_h", _c_int), ("q_stride_w", _c_int)]


class GroupMlpDesc(ctypes.Structure):
    _fields_ = [("batch_size", _c_int), ("queries", Queries), ("nsets", _c_int), ("set_batch_offset", _c_int * 2),
                ("window", Window * 2), ("feat_channels", _c_int), ("num_layers", _c_int), ("cout", _c_int * 3),
                ("xyz1", _c_void_p), ("xyz2", _c_void_p), ("feat2", _c_void_p * 2), ("weights", _c_void_p * 2),
                ("out", _c_void_p * 2), ("dbg_nbr", _c_void_p * 2), ("nbr", _c_void_p * 2),
                ("query_begin", _c_ll), ("query_end", _c_ll)]


class SearchDesc(ctypes.Structure):
    _fields_ = [("select", _c_int), ("batch_size", _c_int), ("queries", Queries), ("window", Window),
                ("xyz1", _c_void_p), ("xyz2", _c_void_p), ("out_nbr", _c_void_p),
                ("query_begin", _c_ll), ("query_end", _c_ll)]


class CostVolumeDesc(ctypes.Structure):
    _fields_ = [("batch_size", _c_int), ("H", _c_int), ("W", _c_int), ("C", _c_int),
                ("window_q", Window), ("window_p", Window),
                ("xyz1", _c_void_p), ("xyz2", _c_void_p), ("f1", _c_void_p), ("f2", _c_void_p),
                ("weights_1", _c_void_p), ("weights_2", _c_void_p), ("stage1_out", _c_void_p), ("out", _c_void_p),
                ("dbg_nbr_q", _c_void_p), ("dbg_nbr_p", _c_void_p), ("nbr_q", _c_void_p), ("nbr_p", _c_void_p),
                ("query_begin", _c_ll), ("query_end", _c_ll)]


class RowMlpPhase(ctypes.Structure):
    _fields_ = [("num_sources", _c_int), ("channels", _c_int * 3), ("from_previous", _c_int * 3),
                ("num_layers", _c_int), ("cout", _c_int * 3), ("src", (_c_void_p * 3) * 2)]


class RowMlpDesc(ctypes.Structure):
    _fields_ = [("rows", _c_ll), ("nsets", _c_int), ("num_phases", _c_int), ("phase", RowMlpPhase * 2),
                ("weights", _c_void_p * 2), ("out", _c_void_p * 2), ("out_phase0", _c_void_p * 2)]


class ProjectDesc(ctypes.Structure):
    _fields_ = [("batch_size", _c_int), ("num_points", _c_int), ("H", _c_int), ("W", _c_int), ("C", _c_int),
                ("mode", _c_int), ("points", _c_void_p), ("point_stride", _c_ll), ("batch_stride", _c_ll),
                ("inner_batch", _c_int), ("outer_stride", _c_ll), ("feat", _c_void_p),
                ("T", _c_void_p), ("q", _c_void_p), ("t", _c_void_p),
                ("pi", _c_float), ("az_res", _c_float), ("v_res", _c_float), ("v_off", _c_float),
                ("cellmin", _c_void_p), ("state", _c_void_p), ("out_xyz", _c_void_p), ("out_feat", _c_void_p), ("out_points", _c_void_p),
                ("T_apply", _c_void_p), ("out_cell", _c_void_p), ("point_keys", _c_void_p)]


class PoseHeadDesc(ctypes.Structure):
    _fields_ = [("batch_size", _c_int), ("num_points", _c_int), ("num_slices", _c_int), ("has_coarse", _c_int),
                ("feature", _c_void_p), ("weight", _c_void_p), ("xyz", _c_void_p),
                ("w_big", _c_void_p), ("b_big", _c_void_p), ("w_q", _c_void_p), ("b_q", _c_void_p),
                ("w_t", _c_void_p), ("b_t", _c_void_p), ("q_coarse", _c_void_p), ("t_coarse", _c_void_p),
                ("partial", _c_void_p), ("counter", _c_void_p),
                ("q_out", _c_void_p), ("t_out", _c_void_p), ("q_norm_out", _c_void_p), ("pooled_out", _c_void_p)]


# name -> argtypes; every function returns int (0 = ok), see include/elo_b200.h
_FUSED_CONV = [_c_int] * 8 + [_c_float, _c_int, _c_int] + [_c_void_p] * 8 + [_c_int, _c_int, _c_void_p]
SIGNATURES = {
    "elo_fused_conv_select_k": _FUSED_CONV,
    "elo_fused_conv_random_k": _FUSED_CONV,
    "elo_multi_search": [ctypes.POINTER(SearchDesc), _c_int, _c_void_p],
    "elo_group_mlp_max": [ctypes.POINTER(GroupMlpDesc), _c_void_p],
    "elo_set_conv_small": [ctypes.POINTER(GroupMlpDesc), _c_void_p],
    "elo_cost_volume_1": [ctypes.POINTER(CostVolumeDesc), _c_void_p],
    "elo_cost_volume_2": [ctypes.POINTER(CostVolumeDesc), _c_void_p],
    "elo_row_mlp": [ctypes.POINTER(RowMlpDesc), _c_void_p],
    "elo_project": [ctypes.POINTER(ProjectDesc), _c_void_p],
    "elo_pose_head": [ctypes.POINTER(PoseHeadDesc), _c_void_p],
    "elo_pyramid_xyz": [_c_int, _c_int, _c_int, ctypes.POINTER(_c_int), ctypes.POINTER(_c_int),
                        ctypes.POINTER(_c_int), ctypes.POINTER(_c_int), _c_void_p, ctypes.POINTER(_c_void_p), _c_void_p],
    "elo_gt_pose": [_c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p],
}

_lib = None


class EloError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise EloError(
                "%s is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                "There is no CPU or PyTorch fallback for this path." % LIB)
        handle = ctypes.CDLL(LIB)
        handle.elo_last_error.restype = ctypes.c_char_p
        handle.elo_last_error.argtypes = []
        handle.elo_version.restype = _c_int
        handle.elo_launch_count.restype = _c_ll
        handle.elo_launch_count.argtypes = []
        handle.elo_set_mlp_engine.argtypes = [_c_int]
        handle.elo_set_mlp_engine.restype = _c_int
        handle.elo_get_mlp_engine.argtypes = []
        handle.elo_get_mlp_engine.restype = _c_int
        handle.elo_set_index_kernel.argtypes = [_c_int]
        handle.elo_set_index_kernel.restype = _c_int
        handle.elo_get_index_kernel.argtypes = []
        handle.elo_get_index_kernel.restype = _c_int
        handle.elo_set_tile_policy.argtypes = [_c_int]
        handle.elo_set_tile_policy.restype = _c_int
        handle.elo_get_tile_policy.argtypes = []
        handle.elo_get_tile_policy.restype = _c_int
        handle.elo_set_pdl.argtypes = [_c_int]
        handle.elo_set_pdl.restype = _c_int
        handle.elo_get_pdl.argtypes = []
        handle.elo_get_pdl.restype = _c_int
        for name, argtypes in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = _c_int
        _lib = handle
    return _lib


def check(rc, what):
    """Map the C ABI's status to the exceptions the reference op raises (InvalidArgument -> ValueError)."""
    if rc == 0:
        return
    msg = lib().elo_last_error().decode("utf-8", "replace")
    if rc == -1:
        raise ValueError("%s: %s" % (what, msg))
    raise EloError("%s failed (status %d): %s" % (what, rc, msg))


def stream_ptr(device):
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


# bench.py's per-kernel timing: when this is a list, every descriptor call is bracketed by CUDA events
# on the launching stream and (name, tag, start, end) is appended.
PROFILE = None
PROFILE_TAG = [""]


def mlp_engine():
    """1 = tcgen05 tensor cores (3xTF32), 0 = fp32 FFMA."""
    return int(lib().elo_get_mlp_engine())


def set_mlp_engine(engine):
    check(lib().elo_set_mlp_engine(int(engine)), "elo_set_mlp_engine")


def set_index_kernel(which):
    """0: choose by query count, 1: tile-staged thread-per-query kernel, 2: one warp per query."""
    check(lib().elo_set_index_kernel(int(which)), "elo_set_index_kernel")


def set_store_warp_min_cells(min_cells):
    """Windows of at least this many cells take the select-K kernel with a store warp (0 = always)."""
    h = lib()
    h.elo_set_store_warp_min_cells.argtypes = [_c_int]
    h.elo_set_store_warp_min_cells.restype = _c_int
    check(h.elo_set_store_warp_min_cells(int(min_cells)), "elo_set_store_warp_min_cells")


def set_tile_staging(mode):
    """Tile staging of the tiled index kernel: 0 bulk copies (TMA engine), 1 plain loads."""
    h = lib()
    h.elo_set_tile_staging.argtypes = [_c_int]
    h.elo_set_tile_staging.restype = _c_int
    check(h.elo_set_tile_staging(int(mode)), "elo_set_tile_staging")


def set_tile_policy(policy):
    """0: latency (spread small calls over all SMs), 1: throughput (full 128-row tiles)."""
    check(lib().elo_set_tile_policy(int(policy)), "elo_set_tile_policy")


def get_tile_policy():
    h = lib()
    h.elo_get_tile_policy.restype = _c_int
    return int(h.elo_get_tile_policy())


def set_pdl(on):
    """Programmatic dependent launch between this library's kernels (default on)."""
    check(lib().elo_set_pdl(int(bool(on))), "elo_set_pdl")


def launch_count():
    return int(lib().elo_launch_count())


def call(name, desc, device):
    """Launch one descriptor-style entry point on torch's current stream of `device`."""
    import torch
    with torch.cuda.device(device):
        if PROFILE is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(torch.cuda.current_stream(device))
        rc = getattr(lib(), name)(ctypes.byref(desc), stream_ptr(device))
        if PROFILE is not None:
            e1.record(torch.cuda.current_stream(device))
            PROFILE.append((name, PROFILE_TAG[0], e0, e1))
    check(rc, name)


def require_cuda(name, *tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise EloError("%s: tensor on %s; this path only runs on CUDA (no CPU fallback)" % (name, t.device))
