"""Thin tensor-level wrappers over the C ABI (include/orpcd_b200.h).

Inputs/outputs are CUDA ``torch`` tensors (memory ownership and streams only);
all arithmetic happens in ``csrc/liborpcd_b200.so``.  ``launch_counter`` tallies
the kernels this process has launched through the library.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

from . import _lib
from ._lib import FgrParams, GicpParams, check

launch_counter = {"kernels": 0}


def _t():
    return _lib.torch()


def _f64(x):
    t = _t()
    assert x.is_cuda and x.dtype == t.float64 and x.is_contiguous(), "expected contiguous CUDA float64 tensor"
    return x


def knn_exact(points, queries, k: int):
    """Exact kNN, (idx int32 [B,M,k], d2 float64 [B,M,k]); points [B,N,3] or [N,3]."""
    t = _t()
    squeeze = points.dim() == 2
    if squeeze:
        points, queries = points[None], queries[None]
    points, queries = _f64(points), _f64(queries)
    B, N, _ = points.shape
    M = queries.shape[1]
    idx = t.empty((B, M, k), dtype=t.int32, device="cuda")
    d2 = t.empty((B, M, k), dtype=t.float64, device="cuda")
    L = _lib.lib()
    nws = L.orpcd_knn_exact_workspace_bytes(N, M, B, k)
    ws = _lib.workspace(nws) if nws else None
    check(L.orpcd_knn_exact(points.data_ptr(), N, N * 3, queries.data_ptr(), M, M * 3, B, k, idx.data_ptr(),
                            d2.data_ptr(), ws.data_ptr() if nws else None, nws, _lib.stream_ptr()), "orpcd_knn_exact")
    launch_counter["kernels"] += 2 if nws else 1
    return (idx[0], d2[0]) if squeeze else (idx, d2)


def spatial_order(points):
    """int32 [N]: index of the point at each position of the cloud's Hilbert order."""
    t = _t()
    points = _f64(points)
    N = points.shape[0]
    L = _lib.lib()
    ws = _lib.workspace(L.orpcd_spatial_order_workspace_bytes(N))
    order = t.empty(N, dtype=t.int32, device="cuda")
    check(L.orpcd_spatial_order(points.data_ptr(), N, ws.data_ptr(), ws.numel(), order.data_ptr(), _lib.stream_ptr()),
          "orpcd_spatial_order")
    launch_counter["kernels"] += 7
    return order


def knn_candidates(base_points, kc: int):
    """Neighbour lists of a base cloud for later re-ranking: (cand_idx int32 [N,kc], bound float64 [N],
    (order int32 [N] = the cloud's spatial processing order, cand_d2 float64 [N,kc] = the candidates' base distances))."""
    t = _t()
    base_points = _f64(base_points)
    N = base_points.shape[0]
    idx, d2 = knn_exact(base_points, base_points, kc)
    bound = t.empty(N, dtype=t.float64, device="cuda")
    check(_lib.lib().orpcd_knn_candidate_bound(idx.data_ptr(), d2.data_ptr(), N, kc, bound.data_ptr(),
                                               _lib.stream_ptr()), "orpcd_knn_candidate_bound")
    launch_counter["kernels"] += 1
    return idx, bound, (spatial_order(base_points), d2)


def knn_rerank(cand_idx, bound, moved_points, factor, k: int, order=None):
    """Exact kNN of moved copies [B,N,3] of the base cloud from its candidate lists.
    factor: float64 [B] lower bound of moved_d2/base_d2.  order: optional (processing order, candidate base distances)
    as returned by knn_candidates (L1 locality of the gathers; early end of a point's scan) -- the results do not depend
    on it.  Returns (idx [B,N,k], d2 [B,N,k], device count of points that took the exact fallback)."""
    order, cand_d2 = order if order is not None else (None, None)
    t = _t()
    moved_points, factor = _f64(moved_points), _f64(factor)
    B, N, _ = moved_points.shape
    kc = cand_idx.shape[1]
    L = _lib.lib()
    ws = _lib.workspace(L.orpcd_knn_rerank_workspace_bytes(N, B))
    idx = t.empty((B, N, k), dtype=t.int32, device="cuda")
    d2 = t.empty((B, N, k), dtype=t.float64, device="cuda")
    check(L.orpcd_knn_rerank(cand_idx.data_ptr(), bound.data_ptr(), kc, moved_points.data_ptr(), N, N * 3, B,
                             factor.data_ptr(), k, idx.data_ptr(), d2.data_ptr(), ws.data_ptr(), ws.numel(),
                             order.data_ptr() if order is not None else None,
                             cand_d2.data_ptr() if cand_d2 is not None else None, _lib.stream_ptr()), "orpcd_knn_rerank")
    launch_counter["kernels"] += 2
    n_fixed = ws[:4].view(t.int32)   # device scalar: points that needed the exact fallback
    return idx, d2, n_fixed


def normals_cov_from_knn(points, nn_idx, nn_d2, radius: float = 0.0, gicp_eps: float = 1e-3, want_normals=False,
                         want_cov=False, want_gicp_cov=True, want_gicp_axis=False, order=None) -> Dict[str, object]:
    """As normals_cov, from neighbour lists [B,N,k] computed elsewhere."""
    t = _t()
    points = _f64(points)
    B, N, _ = points.shape
    knn = nn_idx.shape[2]
    nrm = t.empty((B, N, 3), dtype=t.float64, device="cuda") if want_normals else None
    cov = t.empty((B, N, 6), dtype=t.float64, device="cuda") if want_cov else None
    gc = t.empty((B, N, 6), dtype=t.float64, device="cuda") if want_gicp_cov else None
    ax = t.empty((B, N, 3), dtype=t.float64, device="cuda") if want_gicp_axis else None
    p = lambda x: x.data_ptr() if x is not None else None
    check(_lib.lib().orpcd_normals_cov_from_knn(points.data_ptr(), N, N * 3, B, knn, nn_idx.data_ptr(),
                                                nn_d2.data_ptr(), float(radius), float(gicp_eps), p(nrm), p(cov),
                                                p(gc), p(ax), None, order.data_ptr() if order is not None else None,
                                                _lib.stream_ptr()), "orpcd_normals_cov_from_knn")
    launch_counter["kernels"] += 1
    out = {}
    for k_, v in (("normal", nrm), ("cov", cov), ("gicp_cov", gc), ("gicp_axis", ax)):
        if v is not None:
            out[k_] = v
    return out


def sor_mask(points, nb_neighbors: int, std_ratio: float):
    """(mask uint8 [N], avg float64 [N], stats float64 [4])."""
    t = _t()
    points = _f64(points)
    N = points.shape[0]
    L = _lib.lib()
    ws = _lib.workspace(L.orpcd_sor_workspace_bytes(N, nb_neighbors))
    mask = t.empty(N, dtype=t.uint8, device="cuda")
    avg = t.empty(N, dtype=t.float64, device="cuda")
    stats = t.zeros(4, dtype=t.float64, device="cuda")
    check(L.orpcd_sor_mask(points.data_ptr(), N, nb_neighbors, float(std_ratio), ws.data_ptr(), ws.numel(),
                           mask.data_ptr(), avg.data_ptr(), stats.data_ptr(), _lib.stream_ptr()), "orpcd_sor_mask")
    launch_counter["kernels"] += 4
    return mask, avg, stats


def normals_cov(points, knn: int = 20, radius: float = 0.0, gicp_eps: float = 1e-3, want_normals=False,
                want_cov=False, want_gicp_cov=True, want_nn=False, want_gicp_axis=False) -> Dict[str, object]:
    """Normals / covariances / GICP covariances for clouds [B,N,3] (or [N,3])."""
    t = _t()
    squeeze = points.dim() == 2
    if squeeze:
        points = points[None]
    points = _f64(points)
    B, N, _ = points.shape
    L = _lib.lib()
    ws = _lib.workspace(L.orpcd_normals_workspace_bytes(N, B, knn))
    out = {}
    nrm = t.empty((B, N, 3), dtype=t.float64, device="cuda") if want_normals else None
    cov = t.empty((B, N, 6), dtype=t.float64, device="cuda") if want_cov else None
    gc = t.empty((B, N, 6), dtype=t.float64, device="cuda") if want_gicp_cov else None
    nn = t.empty((B, N, knn), dtype=t.int32, device="cuda") if want_nn else None
    ax = t.empty((B, N, 3), dtype=t.float64, device="cuda") if want_gicp_axis else None
    p = lambda x: x.data_ptr() if x is not None else None
    check(L.orpcd_normals_cov(points.data_ptr(), N, N * 3, B, knn, float(radius), float(gicp_eps), ws.data_ptr(),
                              ws.numel(), p(nrm), p(cov), p(gc), p(ax), p(nn), _lib.stream_ptr()), "orpcd_normals_cov")
    launch_counter["kernels"] += 2
    for k, v in (("normal", nrm), ("cov", cov), ("gicp_cov", gc), ("nn_idx", nn), ("gicp_axis", ax)):
        if v is not None:
            out[k] = v[0] if squeeze else v
    return out


def apply_scale_batch(points, scales):
    """points [N,3] * scales [S,3] -> [S,N,3]."""
    t = _t()
    points, scales = _f64(points), _f64(scales)
    N, S = points.shape[0], scales.shape[0]
    out = t.empty((S, N, 3), dtype=t.float64, device="cuda")
    check(_lib.lib().orpcd_apply_scale_batch(points.data_ptr(), N, scales.data_ptr(), S, out.data_ptr(),
                                             _lib.stream_ptr()), "orpcd_apply_scale_batch")
    launch_counter["kernels"] += 1
    return out


def apply_rigid_batch(points, rotations, translations):
    """points [N,3] @ rotations [U,3,3] + translations [U,3] -> [U,N,3] (row-vector convention)."""
    t = _t()
    points, rotations, translations = _f64(points), _f64(rotations), _f64(translations)
    N, U = points.shape[0], rotations.shape[0]
    out = t.empty((U, N, 3), dtype=t.float64, device="cuda")
    check(_lib.lib().orpcd_apply_rigid_batch(points.data_ptr(), N, rotations.data_ptr(), translations.data_ptr(), U,
                                             out.data_ptr(), _lib.stream_ptr()), "orpcd_apply_rigid_batch")
    launch_counter["kernels"] += 1
    return out


class GicpBatchResult:
    __slots__ = ("T", "stats", "status", "trace", "profile", "workspace", "params")


def gicp_batch(src_points, src_cov6, tgt_points, tgt_cov6, unit_src, unit_tgt, max_correspondence_distance=0.5,
               max_iterations=100, relative_fitness=1e-6, relative_rmse=1e-6, target_splits=0, poll_interval=0,
               profile=False, want_trace=False, ws=None, gicp_epsilon=1e-3, sweep_q=0, sweep_tj=0, sweep_smem_pad=0, sweep_mode=0,
               src_rigid_copies=False, half_patch_from_round=0) -> GicpBatchResult:
    """Batched GICP.  src_points [S,Ns,3], tgt_points [C,Nt,3], unit_* int32 [U]; the covariance
    inputs are either [.,N,6] covariance matrices (generic path) or [.,N,3] plane axes (closed-form
    whitening, ``cov_mode`` 1 of the C ABI) -- the last dimension selects the mode."""
    t = _t()
    src_points, src_cov6, tgt_points, tgt_cov6 = _f64(src_points), _f64(src_cov6), _f64(tgt_points), _f64(tgt_cov6)
    S, Ns, _ = src_points.shape
    Cn, Nt, _ = tgt_points.shape
    U = unit_src.shape[0]
    assert unit_src.dtype == t.int32 and unit_tgt.dtype == t.int32 and unit_src.is_cuda and unit_tgt.is_cuda
    assert src_cov6.shape[-1] == tgt_cov6.shape[-1] and src_cov6.shape[-1] in (3, 6)
    assert src_cov6.shape[:2] == (S, Ns) and tgt_cov6.shape[:2] == (Cn, Nt)
    cov_mode = 1 if src_cov6.shape[-1] == 3 else 0
    prm = GicpParams(U, Ns, Nt, S, Cn, int(max_iterations), int(target_splits), int(poll_interval), int(bool(profile)),
                     cov_mode, float(max_correspondence_distance), float(relative_fitness), float(relative_rmse),
                     float(gicp_epsilon), int(sweep_q), int(sweep_tj), int(sweep_smem_pad), int(sweep_mode),
                     int(bool(src_rigid_copies)), int(half_patch_from_round))
    L = _lib.lib()
    need = L.orpcd_gicp_batch_workspace_bytes(C.byref(prm))
    if need == 0:
        raise _lib.OrpcdError("orpcd_gicp_batch: invalid parameters")
    if ws is None or ws.numel() < need:
        ws = _lib.workspace(need)
    res = GicpBatchResult()
    res.T = t.empty((U, 4, 4), dtype=t.float64, device="cuda")
    res.stats = t.empty((U, 4), dtype=t.float64, device="cuda")
    res.status = t.empty(U, dtype=t.int32, device="cuda")
    res.trace = t.full((U, max_iterations + 1, 2), float("nan"), dtype=t.float64, device="cuda") if want_trace else None
    prof = (C.c_double * 8)()
    check(L.orpcd_gicp_batch(C.byref(prm), src_points.data_ptr(), src_cov6.data_ptr(), tgt_points.data_ptr(),
                             tgt_cov6.data_ptr(), unit_src.data_ptr(), unit_tgt.data_ptr(), ws.data_ptr(), ws.numel(),
                             res.T.data_ptr(), res.stats.data_ptr(), res.status.data_ptr(),
                             res.trace.data_ptr() if want_trace else None, prof, _lib.stream_ptr()),
          "orpcd_gicp_batch")
    res.profile = list(prof)
    res.workspace = ws
    res.params = prm
    launch_counter["kernels"] += int(prof[0])
    return res


def farthest_point_sample(points, first_index: int, sample_size: int):
    """Indices (int32 [sample_size]) of the farthest-point sample of points [N,3] starting at first_index."""
    t = _t()
    points = _f64(points)
    N = points.shape[0]
    out = t.empty(sample_size, dtype=t.int32, device="cuda")
    ws = t.empty(N, dtype=t.float64, device="cuda") if N > 16384 else None
    check(_lib.lib().orpcd_farthest_point_sample(points.data_ptr(), N, int(first_index), int(sample_size),
                                                 ws.data_ptr() if ws is not None else None, out.data_ptr(),
                                                 _lib.stream_ptr()), "orpcd_farthest_point_sample")
    launch_counter["kernels"] += 1
    return out


def voxel_downsample(points, voxel_size: float, ws=None):
    """(voxel averages [n_voxels,3] in lexicographic voxel order, first point index of each voxel, workspace)."""
    t = _t()
    points = _f64(points)
    N = points.shape[0]
    L = _lib.lib()
    need = L.orpcd_voxel_downsample_workspace_bytes(N)
    if ws is None or ws.numel() < need:
        ws = _lib.workspace(need)
    out = t.empty((N, 3), dtype=t.float64, device="cuda")
    first = t.empty(N, dtype=t.int32, device="cuda")
    count = t.zeros(1, dtype=t.int32, device="cuda")
    check(L.orpcd_voxel_downsample(points.data_ptr(), N, float(voxel_size), ws.data_ptr(), ws.numel(), out.data_ptr(),
                                   first.data_ptr(), count.data_ptr(), _lib.stream_ptr()), "orpcd_voxel_downsample")
    launch_counter["kernels"] += 8
    n = int(count.item())
    if n == -1:
        raise _lib.OrpcdError("voxel_size is too small.")    # Open3D: "[VoxelDownSample] voxel_size is too small."
    if n < 0:
        # not an Open3D error: a limit of this implementation (dense cell id and point index share a 64-bit sort word)
        raise _lib.OrpcdError("orpcd_voxel_downsample: the voxel grid has too many cells for the 64-bit sort key "
                              "(dim_x * dim_y * dim_z * n_points must stay below 2^63); scale the cloud (RadiusScaler) "
                              "or use a larger voxel_size.  Open3D itself would accept this input.")
    return out[:n], first[:n], ws


def icp_batch(src_points, tgt_points, unit_src, unit_tgt, init_transforms=None, tgt_normals=None,
              max_correspondence_distance=0.5, max_iterations=200, relative_fitness=1e-6, relative_rmse=1e-6,
              poll_interval=8):
    """Point-to-point (tgt_normals None) or point-to-plane ICP per unit.
    Returns (T [U,4,4] column convention, stats [U,4] = fitness, rmse, iterations, n_corr, status [U])."""
    t = _t()
    src_points, tgt_points = _f64(src_points), _f64(tgt_points)
    S, Ns, _ = src_points.shape
    Cn, Nt, _ = tgt_points.shape
    U = unit_src.shape[0]
    prm = GicpParams(U, Ns, Nt, S, Cn, int(max_iterations), 0, int(poll_interval), 0, 1,
                     float(max_correspondence_distance), float(relative_fitness), float(relative_rmse), 1e-3, 0, 0, 0, 0, 0, 0)
    L = _lib.lib()
    need = L.orpcd_gicp_batch_workspace_bytes(C.byref(prm))
    if need == 0:
        raise _lib.OrpcdError("orpcd_icp_batch: invalid parameters")
    ws = _lib.workspace(need)
    T = t.empty((U, 4, 4), dtype=t.float64, device="cuda")
    stats = t.empty((U, 4), dtype=t.float64, device="cuda")
    status = t.empty(U, dtype=t.int32, device="cuda")
    if init_transforms is not None:
        init_transforms = _f64(init_transforms)
        assert init_transforms.shape == (U, 4, 4)
    if tgt_normals is not None:
        tgt_normals = _f64(tgt_normals)
        assert tgt_normals.shape == tgt_points.shape
    check(L.orpcd_icp_batch(C.byref(prm), 1 if tgt_normals is not None else 0, src_points.data_ptr(),
                            tgt_points.data_ptr(), tgt_normals.data_ptr() if tgt_normals is not None else None,
                            unit_src.data_ptr(), unit_tgt.data_ptr(),
                            init_transforms.data_ptr() if init_transforms is not None else None, ws.data_ptr(),
                            ws.numel(), T.data_ptr(), stats.data_ptr(), status.data_ptr(), _lib.stream_ptr()),
          "orpcd_icp_batch")
    launch_counter["kernels"] += 9   # Hilbert sort + set-up kernels, ring set-up, one persistent launch for all rounds
    return T, stats, status


def evaluate_batch(src_points, tgt_points, unit_src, unit_tgt, transforms=None, max_correspondence_distance=0.5,
                   want_corr=False, ws=None):
    """EvaluateRegistration per unit: (stats [U,4] = fitness, rmse, 0, n_corr; status [U]; corr [U,Ns] | None; ws)."""
    t = _t()
    src_points, tgt_points = _f64(src_points), _f64(tgt_points)
    S, Ns, _ = src_points.shape
    Cn, Nt, _ = tgt_points.shape
    U = unit_src.shape[0]
    prm = GicpParams(U, Ns, Nt, S, Cn, 0, 0, 0, 0, 1, float(max_correspondence_distance), 1e-6, 1e-6, 1e-3, 0, 0, 0, 0, 0, 0)
    L = _lib.lib()
    need = L.orpcd_gicp_batch_workspace_bytes(C.byref(prm))
    if need == 0:
        raise _lib.OrpcdError("orpcd_evaluate_batch: invalid parameters")
    if ws is None or ws.numel() < need:
        ws = _lib.workspace(need)
    stats = t.empty((U, 4), dtype=t.float64, device="cuda")
    status = t.empty(U, dtype=t.int32, device="cuda")
    corr = t.empty((U, Ns), dtype=t.int32, device="cuda") if want_corr else None
    if transforms is not None:
        transforms = _f64(transforms)
        assert transforms.shape == (U, 4, 4)
    check(L.orpcd_evaluate_batch(C.byref(prm), src_points.data_ptr(), tgt_points.data_ptr(), unit_src.data_ptr(),
                                 unit_tgt.data_ptr(), transforms.data_ptr() if transforms is not None else None,
                                 ws.data_ptr(), ws.numel(), stats.data_ptr(), status.data_ptr(),
                                 corr.data_ptr() if want_corr else None, _lib.stream_ptr()), "orpcd_evaluate_batch")
    launch_counter["kernels"] += 7
    return stats, status, corr, ws


def fpfh(points, normals, nn_idx, nn_d2, knn: int, radius: float, want_spfh=False):
    """FPFH descriptors [B,N,33] float64 from neighbour lists [B,N,list_k] (self query, sorted)."""
    t = _t()
    points, normals, nn_d2 = _f64(points), _f64(normals), _f64(nn_d2)
    B, N, _ = points.shape
    assert nn_idx.dtype == t.int32 and nn_idx.is_contiguous() and nn_idx.shape[:2] == (B, N)
    list_k = nn_idx.shape[2]
    spfh = t.empty((B, N, 3, 12), dtype=t.float64, device="cuda")     # padded rows: 3 histograms x (11 bins + a zero)
    out = t.empty((B, N, 33), dtype=t.float64, device="cuda")
    check(_lib.lib().orpcd_fpfh(points.data_ptr(), normals.data_ptr(), N, N * 3, B, nn_idx.data_ptr(),
                                nn_d2.data_ptr(), list_k, int(knn), float(radius), spfh.data_ptr(), out.data_ptr(),
                                _lib.stream_ptr()), "orpcd_fpfh")
    launch_counter["kernels"] += 2
    return (out, spfh[..., :11].reshape(B, N, 33)) if want_spfh else out


class PackedFeatures:
    """Feature sets [S,N,33] float64 plus their TF32 hi/lo operand-tile images for the tcgen05 filter."""
    __slots__ = ("features", "packed", "maxnorm", "n", "n_sets", "center")


def feature_center(features):
    """Mean feature [33] of set 0: the centring vector shared by the two sides of a match."""
    t = _t()
    features = _f64(features)
    c = t.empty(33, dtype=t.float64, device="cuda")
    check(_lib.lib().orpcd_feature_center(features.data_ptr(), features.shape[1], c.data_ptr(), _lib.stream_ptr()),
          "orpcd_feature_center")
    launch_counter["kernels"] += 1
    return c


def feature_pack(features, center=None) -> PackedFeatures:
    """center: None (computed from set 0 of `features`), False (no centring) or the tensor of the other side."""
    t = _t()
    features = _f64(features)
    S, N, D = features.shape
    assert D == 33
    L = _lib.lib()
    pf = PackedFeatures()
    pf.features, pf.n, pf.n_sets = features, N, S
    pf.center = None if center is False else (feature_center(features) if center is None else center)
    pf.packed = t.empty(L.orpcd_feature_pack_bytes(N) * S, dtype=t.uint8, device="cuda")
    pf.maxnorm = t.empty(S, dtype=t.float32, device="cuda")
    check(L.orpcd_feature_pack(features.data_ptr(), N, S, pf.center.data_ptr() if pf.center is not None else None,
                               pf.packed.data_ptr(), pf.maxnorm.data_ptr(), _lib.stream_ptr()), "orpcd_feature_pack")
    launch_counter["kernels"] += 1
    return pf


def feature_match(query: PackedFeatures, db: PackedFeatures, pair_q, pair_db, mode: int = 0, want_d2=False,
                  want_debug=False, row_mask=None):
    """For pair p: nearest row of db set pair_db[p] for every row of query set pair_q[p]
    (rows with row_mask[p, row] == 0 are skipped and get -1).
    Returns (nn int32 [P,Nq], d2 | None, device int32[2] = (rows that needed the exact scan, chunks re-scored),
    debug | None)."""
    t = _t()
    P = pair_q.shape[0]
    assert pair_q.dtype == t.int32 and pair_db.dtype == t.int32 and pair_q.is_cuda and pair_db.is_cuda
    L = _lib.lib()
    if query.center is not db.center:
        raise ValueError("both sides of a feature match must be packed with the same centring vector")
    if row_mask is not None:
        assert row_mask.dtype == t.uint8 and row_mask.is_cuda and row_mask.is_contiguous() and \
            tuple(row_mask.shape) == (P, query.n)
    ws = _lib.workspace(L.orpcd_feature_match_workspace_bytes(query.n, P))
    nn = t.empty((P, query.n), dtype=t.int32, device="cuda")
    d2 = t.empty((P, query.n), dtype=t.float64, device="cuda") if want_d2 else None
    stats = t.zeros(2, dtype=t.int32, device="cuda")
    dbg = None
    if want_debug:
        dbg = t.full((P, query.n, (db.n + 127) // 128 * 128), float("nan"), dtype=t.float32, device="cuda")
    check(L.orpcd_feature_match(query.features.data_ptr(), query.packed.data_ptr(), query.n, db.features.data_ptr(),
                                db.packed.data_ptr(), db.maxnorm.data_ptr(), db.n,
                                db.center.data_ptr() if db.center is not None else None, pair_q.data_ptr(),
                                pair_db.data_ptr(), P, int(mode), ws.data_ptr(), ws.numel(),
                                row_mask.data_ptr() if row_mask is not None else None, nn.data_ptr(),
                                d2.data_ptr() if want_d2 else None, stats.data_ptr(),
                                dbg.data_ptr() if want_debug else None, _lib.stream_ptr()), "orpcd_feature_match")
    launch_counter["kernels"] += 2 if mode == 0 else 1
    return nn, d2, stats, dbg


def fgr_batch(src_points, tgt_points, unit_src, unit_tgt, ij, ji, keys, division_factor=1.4, tuple_scale=0.9,
              max_correspondence_distance=0.5, iteration_number=100, decrease_mu=True, max_tuple_count=1000,
              tuple_test=True):
    """Cross check + tuple test + GNC solve per unit.  Returns (T [U,4,4] source->target, info int32 [U,4])."""
    t = _t()
    src_points, tgt_points = _f64(src_points), _f64(tgt_points)
    U = unit_src.shape[0]
    Ns, Nt = src_points.shape[1], tgt_points.shape[1]
    assert ij.dtype == t.int32 and ji.dtype == t.int32 and ij.shape[0] == U and ji.shape[0] == U
    assert ij.is_contiguous() and ji.is_contiguous()
    assert keys.dtype == t.int64 and keys.shape[0] == U and keys.is_cuda
    prm = FgrParams(U, Ns, Nt, ij.shape[1], ji.shape[1], int(iteration_number), int(bool(decrease_mu)),
                    int(max_tuple_count), int(bool(tuple_test)), float(division_factor), float(tuple_scale),
                    float(max_correspondence_distance))
    L = _lib.lib()
    need = L.orpcd_fgr_batch_workspace_bytes(C.byref(prm))
    if need == 0:
        raise _lib.OrpcdError("orpcd_fgr_batch: invalid parameters")
    ws = _lib.workspace(need)
    T = t.empty((U, 4, 4), dtype=t.float64, device="cuda")
    info = t.empty((U, 4), dtype=t.int32, device="cuda")
    check(L.orpcd_fgr_batch(C.byref(prm), src_points.data_ptr(), tgt_points.data_ptr(), unit_src.data_ptr(),
                            unit_tgt.data_ptr(), ij.data_ptr(), ji.data_ptr(),
                            keys.data_ptr(), ws.data_ptr(), ws.numel(), T.data_ptr(), info.data_ptr(),
                            _lib.stream_ptr()), "orpcd_fgr_batch")
    launch_counter["kernels"] += 3
    return T, info


def fgr_tuple_keys(src_points, tgt_points, unit_src, unit_tgt, seed: int):
    """Per-unit keys (int64 bit patterns of the uint64 keys) of the tuple-test streams, computed on the device."""
    t = _t()
    src_points, tgt_points = _f64(src_points), _f64(tgt_points)
    U = unit_src.shape[0]
    keys = t.empty(U, dtype=t.int64, device="cuda")
    check(_lib.lib().orpcd_fgr_tuple_keys(src_points.data_ptr(), src_points.shape[1], tgt_points.data_ptr(),
                                          tgt_points.shape[1], unit_src.data_ptr(), unit_tgt.data_ptr(), U,
                                          int(seed) & ((1 << 64) - 1), keys.data_ptr(), _lib.stream_ptr()),
          "orpcd_fgr_tuple_keys")
    launch_counter["kernels"] += 1
    return keys


def gicp_sweep_only(params: GicpParams, ws) -> None:
    check(_lib.lib().orpcd_gicp_sweep_only(C.byref(params), ws.data_ptr(), ws.numel(), _lib.stream_ptr()),
          "orpcd_gicp_sweep_only")
    launch_counter["kernels"] += 1


def ffma_burn(blocks: int, iters: int, packed: bool = False):
    t = _t()
    sink = t.empty(blocks * 256, dtype=t.float32, device="cuda")
    fn = _lib.lib().orpcd_ffma2_burn if packed else _lib.lib().orpcd_ffma_burn
    check(fn(blocks, iters, sink.data_ptr(), _lib.stream_ptr()), "orpcd_ffma_burn")
    launch_counter["kernels"] += 1
    return sink


def dfma_burn(blocks: int, iters: int):
    t = _t()
    sink = t.empty(blocks * 256, dtype=t.float64, device="cuda")
    check(_lib.lib().orpcd_dfma_burn(blocks, iters, sink.data_ptr(), _lib.stream_ptr()), "orpcd_dfma_burn")
    launch_counter["kernels"] += 1
    return sink
