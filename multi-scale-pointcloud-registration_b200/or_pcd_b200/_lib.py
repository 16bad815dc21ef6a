"""ctypes binding of ``csrc/liborpcd_b200.so`` (C ABI declared in ``include/orpcd_b200.h``).

There is NO CPU fallback: if the CUDA library is missing or no CUDA device is
present, the B200 blocks raise.  PyTorch is used only to own device memory and
streams; every kernel on the registration path is in the shared library.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ORPCD_B200_LIB") or os.path.join(os.path.dirname(_HERE), "csrc", "liborpcd_b200.so")

ORPCD_OK = 0
MAX_KNN = 512

_lib = None


class GicpParams(C.Structure):
    """Mirror of ``orpcd_gicp_params`` (include/orpcd_b200.h)."""

    _fields_ = [
        ("n_units", C.c_int),
        ("n_src", C.c_int),
        ("n_tgt", C.c_int),
        ("n_src_sets", C.c_int),
        ("n_tgt_sets", C.c_int),
        ("max_iterations", C.c_int),
        ("target_splits", C.c_int),
        ("poll_interval", C.c_int),
        ("profile", C.c_int),
        ("cov_mode", C.c_int),
        ("max_correspondence_distance", C.c_double),
        ("relative_fitness", C.c_double),
        ("relative_rmse", C.c_double),
        ("gicp_epsilon", C.c_double),
        ("sweep_q", C.c_int),
        ("sweep_tj", C.c_int),
        ("sweep_smem_pad", C.c_int),
        ("sweep_mode", C.c_int),
        ("src_rigid_copies", C.c_int),
        ("half_patch_from_round", C.c_int),
    ]


class FgrParams(C.Structure):
    """Mirror of ``orpcd_fgr_params`` (include/orpcd_b200.h)."""

    _fields_ = [
        ("n_units", C.c_int),
        ("n_src", C.c_int),
        ("n_tgt", C.c_int),
        ("n_feat_src", C.c_int),
        ("n_feat_tgt", C.c_int),
        ("iteration_number", C.c_int),
        ("decrease_mu", C.c_int),
        ("max_tuple_count", C.c_int),
        ("tuple_test", C.c_int),
        ("division_factor", C.c_double),
        ("tuple_scale", C.c_double),
        ("max_correspondence_distance", C.c_double),
    ]


class OrpcdError(RuntimeError):
    pass


def lib():
    """Load the CUDA library once; fail loudly when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OrpcdError(
            f"{LIB_PATH} is missing: build it with `make -C multi-scale-pointcloud-registration_b200/csrc` "
            "(or __graft_entry__.build()).  There is no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    vp, ip, dp, sz, ll = C.c_void_p, C.c_int, C.c_double, C.c_size_t, C.c_longlong
    L.orpcd_version.restype = ip
    L.orpcd_compiled_arch.restype = ip
    L.orpcd_error_string.restype = C.c_char_p
    L.orpcd_error_string.argtypes = [ip]
    L.orpcd_knn_exact.restype = ip
    L.orpcd_knn_exact.argtypes = [vp, ip, ll, vp, ip, ll, ip, ip, vp, vp, vp, sz, vp]
    L.orpcd_knn_exact_workspace_bytes.restype = sz
    L.orpcd_knn_exact_workspace_bytes.argtypes = [ip, ip, ip, ip]
    L.orpcd_knn_candidate_bound.restype = ip
    L.orpcd_knn_candidate_bound.argtypes = [vp, vp, ip, ip, vp, vp]
    L.orpcd_knn_rerank_workspace_bytes.restype = sz
    L.orpcd_knn_rerank_workspace_bytes.argtypes = [ip, ip]
    L.orpcd_knn_rerank.restype = ip
    L.orpcd_knn_rerank.argtypes = [vp, vp, ip, vp, ip, ll, ip, vp, ip, vp, vp, vp, sz, vp, vp, vp]
    L.orpcd_spatial_order_workspace_bytes.restype = sz
    L.orpcd_spatial_order_workspace_bytes.argtypes = [ip]
    L.orpcd_spatial_order.restype = ip
    L.orpcd_spatial_order.argtypes = [vp, ip, vp, sz, vp, vp]
    L.orpcd_normals_cov_from_knn.restype = ip
    L.orpcd_normals_cov_from_knn.argtypes = [vp, ip, ll, ip, ip, vp, vp, dp, dp, vp, vp, vp, vp, vp, vp, vp]
    L.orpcd_sor_workspace_bytes.restype = sz
    L.orpcd_sor_workspace_bytes.argtypes = [ip, ip]
    L.orpcd_sor_mask.restype = ip
    L.orpcd_sor_mask.argtypes = [vp, ip, ip, dp, vp, sz, vp, vp, vp, vp]
    L.orpcd_normals_workspace_bytes.restype = sz
    L.orpcd_normals_workspace_bytes.argtypes = [ip, ip, ip]
    L.orpcd_normals_cov.restype = ip
    L.orpcd_normals_cov.argtypes = [vp, ip, ll, ip, ip, dp, dp, vp, sz, vp, vp, vp, vp, vp, vp]
    L.orpcd_apply_rigid_batch.restype = ip
    L.orpcd_apply_rigid_batch.argtypes = [vp, ip, vp, vp, ip, vp, vp]
    L.orpcd_apply_scale_batch.restype = ip
    L.orpcd_apply_scale_batch.argtypes = [vp, ip, vp, ip, vp, vp]
    L.orpcd_gicp_batch_workspace_bytes.restype = sz
    L.orpcd_gicp_batch_workspace_bytes.argtypes = [C.POINTER(GicpParams)]
    L.orpcd_gicp_batch.restype = ip
    L.orpcd_gicp_batch.argtypes = [C.POINTER(GicpParams), vp, vp, vp, vp, vp, vp, vp, sz, vp, vp, vp, vp,
                                   C.POINTER(C.c_double), vp]
    L.orpcd_gicp_sweep_only.restype = ip
    L.orpcd_gicp_sweep_only.argtypes = [C.POINTER(GicpParams), vp, sz, vp]
    L.orpcd_farthest_point_sample.restype = ip
    L.orpcd_farthest_point_sample.argtypes = [vp, ip, ip, ip, vp, vp, vp]
    L.orpcd_voxel_downsample_workspace_bytes.restype = sz
    L.orpcd_voxel_downsample_workspace_bytes.argtypes = [ip]
    L.orpcd_voxel_downsample.restype = ip
    L.orpcd_voxel_downsample.argtypes = [vp, ip, dp, vp, sz, vp, vp, vp, vp]
    L.orpcd_icp_batch.restype = ip
    L.orpcd_icp_batch.argtypes = [C.POINTER(GicpParams), ip, vp, vp, vp, vp, vp, vp, vp, sz, vp, vp, vp, vp]
    L.orpcd_evaluate_batch.restype = ip
    L.orpcd_evaluate_batch.argtypes = [C.POINTER(GicpParams), vp, vp, vp, vp, vp, vp, sz, vp, vp, vp, vp]
    L.orpcd_fpfh.restype = ip
    L.orpcd_fpfh.argtypes = [vp, vp, ip, ll, ip, vp, vp, ip, ip, dp, vp, vp, vp]
    L.orpcd_feature_pack_bytes.restype = sz
    L.orpcd_feature_pack_bytes.argtypes = [ip]
    L.orpcd_feature_pack.restype = ip
    L.orpcd_feature_pack.argtypes = [vp, ip, ip, vp, vp, vp, vp]
    L.orpcd_feature_center.restype = ip
    L.orpcd_feature_center.argtypes = [vp, ip, vp, vp]
    L.orpcd_feature_match_workspace_bytes.restype = sz
    L.orpcd_feature_match_workspace_bytes.argtypes = [ip, ip]
    L.orpcd_feature_match.restype = ip
    L.orpcd_feature_match.argtypes = [vp, vp, ip, vp, vp, vp, ip, vp, vp, vp, ip, ip, vp, sz, vp, vp, vp, vp, vp, vp]
    L.orpcd_fgr_batch_workspace_bytes.restype = sz
    L.orpcd_fgr_batch_workspace_bytes.argtypes = [C.POINTER(FgrParams)]
    L.orpcd_fgr_batch.restype = ip
    L.orpcd_fgr_batch.argtypes = [C.POINTER(FgrParams), vp, vp, vp, vp, vp, vp, vp, vp, sz, vp, vp, vp]
    L.orpcd_fgr_tuple_keys.restype = ip
    L.orpcd_fgr_tuple_keys.argtypes = [vp, ip, vp, ip, vp, vp, ip, C.c_ulonglong, vp, vp]
    L.orpcd_ffma_burn.restype = ip
    L.orpcd_ffma_burn.argtypes = [ip, ip, vp, vp]
    L.orpcd_ffma2_burn.restype = ip
    L.orpcd_ffma2_burn.argtypes = [ip, ip, vp, vp]
    L.orpcd_dfma_burn.restype = ip
    L.orpcd_dfma_burn.argtypes = [ip, ip, vp, vp]
    L.orpcd_mix_burn.restype = ip
    L.orpcd_mix_burn.argtypes = [ip, ip, ip, vp, vp]
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != ORPCD_OK:
        raise OrpcdError(f"{what} failed: {lib().orpcd_error_string(rc).decode()} ({rc})")


_torch = None


def torch():
    """Import torch lazily and insist on a CUDA device (no CPU path exists)."""
    global _torch
    if _torch is None:
        import torch as _t

        if not _t.cuda.is_available():
            raise OrpcdError("or_pcd_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
        _torch = _t
    return _torch


_side = {}


def side_stream():
    """One extra CUDA stream per device for work that runs next to the current stream.  Shared by every optimizer
    instance: the caching allocator keeps a pool per stream, so a stream per instance would turn every new
    instance's temporaries into fresh cudaMalloc calls."""
    t = torch()
    dev = t.cuda.current_device()
    if dev not in _side:
        _side[dev] = t.cuda.Stream(device=dev)
    return _side[dev]


def stream_ptr() -> int:
    return int(torch().cuda.current_stream().cuda_stream)


def to_device(a: np.ndarray, dtype=np.float64):
    """Host ndarray -> contiguous CUDA tensor (H2D copy on the current stream).  Arrays that
    live in pinned memory (see ``pinned_empty``) are copied asynchronously."""
    t = torch()
    a = np.ascontiguousarray(a, dtype=dtype)
    h = t.from_numpy(a)
    return h.to("cuda", non_blocking=bool(h.is_pinned()))


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """numpy array backed by page-locked host memory (the tensor is kept alive by the array)."""
    t = torch()
    td = {np.float64: t.float64, np.float32: t.float32, np.int32: t.int32}[dtype]
    return t.empty(tuple(shape), dtype=td, pin_memory=True).numpy()   # the array keeps the tensor alive


def workspace(nbytes: int):
    t = torch()
    return t.empty(max(int(nbytes), 256), dtype=t.uint8, device="cuda")
