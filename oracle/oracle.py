"""CPU ORACLE -- test infrastructure, not product code.

ctypes face of ``oracle/orpcd_oracle.c`` (the float64 restatement of the Open3D
0.18.0 algorithms the reference calls) plus a restatement of the reference's
pure-Python control flow on top of it.  PARITY UNPINNED for the arithmetic of the
Open3D entry points (the reference has no tests/golden vectors for this path and
Open3D cannot be installed here); the classes of THIS module are pinned bit for bit
to golden vectors of the reference's own classes (``tests/golden/make_golden*_ref.py``:
Aligner, GeneralizedICP, FastGlobalOptimizer, SOR, VoxelDownsampler run unmodified,
the Open3D-backed ones over a stub whose numerical entry points are this oracle's).
See the header of ``orpcd_oracle.c`` and DESIGN.md section 5.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this module.  Each function cites the reference file:line it follows
(paths relative to the reference repo root).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liborpcd_oracle.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_bp = C.POINTER(C.c_ubyte)


def build(force: bool = False) -> str:
    """Compile the C oracle (gcc) if the shared object is missing or stale."""
    src = os.path.join(_HERE, "orpcd_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _f64(a, cols=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if cols is not None:
        assert a.ndim == 2 and a.shape[1] == cols, a.shape
    return a


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_threads(n: int) -> None:
    lib().oracle_set_threads(C.c_int(int(n)))


# --------------------------------------------------------------------------
# Open3D primitives
# --------------------------------------------------------------------------
def knn(points: np.ndarray, queries: np.ndarray, k: int, method: str = "kdtree"):
    """KDTreeFlann.SearchKNN for every query (A.1): (idx (m,k) int32, d2 (m,k) f64),
    sorted by (d2, index); missing entries are -1 / inf."""
    points = _f64(points)
    queries = _f64(queries)
    n, dim = points.shape
    m = queries.shape[0]
    idx = np.empty((m, k), dtype=np.int32)
    d2 = np.empty((m, k), dtype=np.float64)
    rc = lib().oracle_knn(_d(points), C.c_int(n), C.c_int(dim), _d(queries), C.c_int(m), C.c_int(k),
                          C.c_int(0 if method == "kdtree" else 1), _i(idx), _d(d2))
    assert rc == 0
    return idx, d2


def sor(points: np.ndarray, nb_neighbors: int = 64, std_ratio: float = 2.0):
    """PointCloud.remove_statistical_outlier (A.2; sor.py:69-71).
    Returns (mask bool (n,), avg (n,), (mean, std, thr))."""
    points = _f64(points, 3)
    n = points.shape[0]
    mask = np.zeros(n, dtype=np.uint8)
    avg = np.empty(n, dtype=np.float64)
    stats = np.zeros(3, dtype=np.float64)
    rc = lib().oracle_sor(_d(points), C.c_int(n), C.c_int(nb_neighbors), C.c_double(std_ratio),
                          mask.ctypes.data_as(_bp), _d(avg), _d(stats))
    assert rc == 0
    return mask.astype(bool), avg, tuple(stats)


def normals(points: np.ndarray, knn_k: int = 20, radius: float = 0.0, gicp_eps: float = 1e-3):
    """EstimateNormals + GICP covariances (A.3).  radius<=0: KNN search; >0: Hybrid.
    Returns dict(cov (n,3,3), normal (n,3), gicp_cov (n,3,3), nn_idx (n,k), nn_cnt (n,))."""
    points = _f64(points, 3)
    n = points.shape[0]
    cov = np.empty((n, 9)); nrm = np.empty((n, 3)); gc = np.empty((n, 9))
    nn_idx = np.empty((n, knn_k), dtype=np.int32); nn_cnt = np.empty(n, dtype=np.int32)
    rc = lib().oracle_normals(_d(points), C.c_int(n), C.c_int(knn_k), C.c_double(radius),
                              C.c_double(gicp_eps), _d(cov), _d(nrm), _d(gc), _i(nn_idx), _i(nn_cnt))
    assert rc == 0
    return dict(cov=cov.reshape(n, 3, 3), normal=nrm, gicp_cov=gc.reshape(n, 3, 3),
                nn_idx=nn_idx, nn_cnt=nn_cnt)


def inv_sqrt_sym3(M: np.ndarray) -> np.ndarray:
    M = _f64(M).reshape(9).copy()
    W = np.empty(9)
    lib().oracle_inv_sqrt_sym3(_d(M), _d(W))
    return W.reshape(3, 3)


def ldlt_solve6(A: np.ndarray, b: np.ndarray) -> np.ndarray:
    A = _f64(A).reshape(36).copy(); b = _f64(b).reshape(6).copy()
    x = np.empty(6)
    lib().oracle_ldlt_solve6(_d(A), _d(b), _d(x))
    return x


def evaluate(source, target, max_dist, T):
    """registration::EvaluateRegistration (A.4): (fitness, rmse, n_corr, corr (ns,))."""
    source = _f64(source, 3); target = _f64(target, 3)
    T = _f64(T).reshape(16).copy()
    fit = C.c_double(); rmse = C.c_double(); nc = C.c_int()
    corr = np.empty(source.shape[0], dtype=np.int32)
    lib().oracle_evaluate(_d(source), C.c_int(source.shape[0]), _d(target), C.c_int(target.shape[0]),
                          C.c_double(max_dist), _d(T), C.byref(fit), C.byref(rmse), C.byref(nc), _i(corr))
    return fit.value, rmse.value, nc.value, corr


def gicp(source, target, max_dist=0.5, max_iter=100, eps=1e-3, src_cov=None, tgt_cov=None,
         want_trace=False):
    """registration_generalized_icp from identity (A.3-A.5; generalizedICP.py:59-70).
    Returns dict(T (4,4) Open3D column convention, fitness, rmse, iters, n_corr[, trace, xtrace])."""
    source = _f64(source, 3); target = _f64(target, 3)
    ns, nt = source.shape[0], target.shape[0]
    T = np.empty(16)
    fit = C.c_double(); rmse = C.c_double(); iters = C.c_int(); nc = C.c_int()
    trace = np.full((max_iter + 1, 2), np.nan) if want_trace else None
    xtrace = np.full((max_iter, 6), np.nan) if want_trace else None
    sc = _f64(src_cov).reshape(ns, 9) if src_cov is not None else None
    tc = _f64(tgt_cov).reshape(nt, 9) if tgt_cov is not None else None
    rc = lib().oracle_gicp(_d(source), C.c_int(ns), _d(target), C.c_int(nt), C.c_double(max_dist),
                           C.c_int(max_iter), C.c_double(eps),
                           _d(sc) if sc is not None else None, _d(tc) if tc is not None else None,
                           _d(T), C.byref(fit), C.byref(rmse), C.byref(iters), C.byref(nc),
                           _d(trace) if want_trace else None, _d(xtrace) if want_trace else None)
    assert rc == 0
    out = dict(T=T.reshape(4, 4), fitness=fit.value, rmse=rmse.value, iters=iters.value, n_corr=nc.value)
    if want_trace:
        out["trace"] = trace[: iters.value + 1]
        out["xtrace"] = xtrace[: iters.value]
    return out


def icp(source, target, max_dist=0.5, max_iter=200, init_T=None, mode="PointToPoint", target_normals=None):
    """registration_icp (A.4 loop with TransformationEstimationPointToPoint / PointToPlane;
    Aligner/Aligner.py:353-360).  init_T / returned T: Open3D column convention."""
    source = _f64(source, 3); target = _f64(target, 3)
    T = np.empty(16); fit = C.c_double(); rmse = C.c_double(); iters = C.c_int(); nc = C.c_int()
    it = _f64(init_T).reshape(16).copy() if init_T is not None else None
    tn = _f64(target_normals, 3) if target_normals is not None else None
    m = {"PointToPoint": 0, "PointToPlane": 1}[mode]
    rc = lib().oracle_icp(_d(source), C.c_int(source.shape[0]), _d(target), C.c_int(target.shape[0]),
                          C.c_double(max_dist), C.c_int(max_iter), _d(it) if it is not None else None, C.c_int(m),
                          _d(tn) if tn is not None else None, _d(T), C.byref(fit), C.byref(rmse), C.byref(iters),
                          C.byref(nc))
    if rc != 0:
        raise RuntimeError("TransformationEstimationPointToPlane and TransformationEstimationColoredICP "
                           "require pre-computed normal vectors for target PointCloud.")
    return dict(T=T.reshape(4, 4), fitness=fit.value, rmse=rmse.value, iters=iters.value, n_corr=nc.value)


def fpfh(points, normals_, knn_k: int = 20, radius: float = 0.1, want_spfh=False):
    """compute_fpfh_feature(cloud, Hybrid(radius, knn)) (A.7; fastGlobalOptimizer.py:130-143).
    Returns (n,33) float64 (Open3D's 33xN column-major data_ has the same memory order)."""
    points = _f64(points, 3); normals_ = _f64(normals_, 3)
    n = points.shape[0]
    out = np.empty((n, 33)); sp = np.empty((n, 33)) if want_spfh else None
    rc = lib().oracle_fpfh(_d(points), _d(normals_), C.c_int(n), C.c_int(knn_k), C.c_double(radius), _d(out),
                           _d(sp) if want_spfh else None)
    assert rc == 0
    return (out, sp) if want_spfh else out


def philox(counter: int, key: int):
    out = (C.c_uint32 * 4)()
    lib().oracle_philox(C.c_uint64(counter), C.c_uint64(key), out)
    return list(out)


_M64 = (1 << 64) - 1


def tuple_key(seed: int, source: np.ndarray, target: np.ndarray) -> int:
    """Key of the tuple-test random stream of one registration.  The reference never seeds
    Open3D's generator; this project derives the key from the call's own
    inputs (sizes and the bit patterns of the first/last point of each cloud), so the
    sequential reference control flow and a batched evaluation draw identical streams
    without any shared call counter."""
    h = (int(seed) * 0x9E3779B97F4A7C15 + 0x243F6A8885A308D3) & _M64
    words = [source.shape[0], target.shape[0]]
    for c in (source[0], source[-1], target[0], target[-1]):
        words.extend(int(v) for v in np.ascontiguousarray(c, dtype=np.float64).view(np.uint64))
    for w in words:
        h = ((h ^ (w & _M64)) * 0xBF58476D1CE4E5B9) & _M64
        h ^= h >> 31
    return h


def fgr(source, target, feat_s, feat_t, division_factor=1.4, tuple_scale=0.9, max_corr_dist=0.5,
        iteration_number=100, decrease_mu=True, max_tuple_count=1000, tuple_test=True, key=0, want_corres=False):
    """registration_fgr_based_on_feature_matching (A.8; fastGlobalOptimizer.py:158-174).
    Returns dict(T source->target (column convention), fitness, rmse, n_corr, n_initial, n_pairs, trials, swapped)."""
    source = _f64(source, 3); target = _f64(target, 3)
    feat_s = _f64(feat_s, 33); feat_t = _f64(feat_t, 33)
    T = np.empty(16); fit = C.c_double(); rmse = C.c_double(); nc = C.c_int()
    info = np.zeros(4, dtype=np.int32)
    cor = np.full((3 * max(max_tuple_count, 1) if tuple_test else max(feat_s.shape[0], 1), 2), -1, dtype=np.int32)
    rc = lib().oracle_fgr(_d(source), C.c_int(source.shape[0]), _d(target), C.c_int(target.shape[0]),
                          _d(feat_s), C.c_int(feat_s.shape[0]), _d(feat_t), C.c_int(feat_t.shape[0]),
                          C.c_double(division_factor), C.c_double(tuple_scale), C.c_double(max_corr_dist),
                          C.c_int(iteration_number), C.c_int(int(bool(decrease_mu))), C.c_int(max_tuple_count),
                          C.c_int(int(bool(tuple_test))), C.c_uint64(key), _d(T), C.byref(fit), C.byref(rmse),
                          C.byref(nc), _i(info), _i(cor))
    assert rc == 0
    out = dict(T=T.reshape(4, 4), fitness=fit.value, rmse=rmse.value, n_corr=nc.value, n_initial=int(info[0]),
               n_pairs=int(info[1]), trials=int(info[2]), swapped=bool(info[3]))
    if want_corres:
        out["corres"] = cor[: info[1]].copy()
    return out


# --------------------------------------------------------------------------
# The reference's Python layer, restated over the primitives above
# --------------------------------------------------------------------------
class RadiusScaler:
    """Preprocessor/Scalers/radiusScaler.py:18-30 (mean-centre, divide by max radius).
    scipy's cdist(euclidean) is sqrt(sum((a-b)**2)) in coordinate order."""

    is_scaler = True

    def __init__(self):
        self.mean = np.zeros((1, 3))
        self.scale = 1.0

    def process(self, cloud):
        center = np.mean(cloud, axis=0, keepdims=True)
        d = cloud - center
        radius = np.max(np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]))
        self.mean = center
        self.scale = radius
        return (cloud - center) / radius


class RandomDownsampler:
    """Preprocessor/Downsamplers/randomDownsampler.py:29-37 (global legacy numpy RNG)."""

    is_scaler = False

    def __init__(self, sample_size=4096, replace=False):
        self._sample_size = sample_size if sample_size > 0 else 4096
        self._replace = replace

    def process(self, cloud):
        return cloud[np.random.choice(cloud.shape[0], self._sample_size, replace=self._replace)]


class SOR:
    """Preprocessor/Outliers/sor.py:51-79 -> oracle_sor (A.2)."""

    is_scaler = False

    def __init__(self, nb_neighbours=64, std_ratio=2):
        self._nb_neighbours = nb_neighbours
        self._std_ratio = std_ratio

    def process(self, cloud):
        mask, _, _ = sor(np.asarray(cloud, dtype=np.float64), self._nb_neighbours, float(self._std_ratio))
        return np.ascontiguousarray(np.asarray(cloud, dtype=np.float64)[mask])


class Preprocessor:
    """Preprocessor/preprocessor.py:12-31 (auto-inserts RadiusScaler at index 0)."""

    def __init__(self, blocks: list):
        if not any(getattr(b, "is_scaler", False) for b in blocks):
            blocks.insert(0, RadiusScaler())
        self.preprocessor_blocks = blocks

    def preprocess(self, cloud):
        for b in self.preprocessor_blocks:
            cloud = b.process(cloud)
        return cloud


class GeneralizedICP:
    """Optimizer/generalizedICP.py:47-83: Open3D GICP from identity, rotation block
    transposed, ValueError when inlier_rmse == 0."""

    def __init__(self, max_correspondence_distance=0.5, max_iterations=100):
        self._max_correspondence_distance = max_correspondence_distance if max_correspondence_distance > 0 else 0.5
        self._max_iterations = max_iterations if max_iterations > 0 else 100
        self.calls = 0
        self.iterations = 0

    def optimize(self, source, target, **kwargs):
        r = gicp(source, target, self._max_correspondence_distance, self._max_iterations)
        self.calls += 1
        self.iterations += r["iters"]
        roto = np.copy(r["T"])
        roto[:3, :3] = roto[:3, :3].T
        if r["rmse"] == 0:
            raise ValueError("Optimization failed with loss = 0. Parameters are not well set.")
        return roto, r["rmse"]


def fps(points, first: int, k: int):
    """Farthest point sampling indices (farthestDownsampler.py:33-54 with the first index given)."""
    points = _f64(points, 3)
    out = np.empty(k, dtype=np.int32)
    rc = lib().oracle_fps(_d(points), C.c_int(points.shape[0]), C.c_int(int(first)), C.c_int(int(k)), _i(out))
    assert rc == 0
    return out


def voxel_downsample(points, voxel_size: float):
    """PointCloud.voxel_down_sample (voxelDownsampler.py:83,95), voxels in lexicographic index order.
    Returns (points (m,3), first point index of each voxel (m,))."""
    points = _f64(points, 3)
    out = np.empty_like(points); first = np.empty(points.shape[0], dtype=np.int32)
    lib().oracle_voxel_downsample.restype = C.c_int
    m = lib().oracle_voxel_downsample(_d(points), C.c_int(points.shape[0]), C.c_double(voxel_size), _d(out), _i(first))
    assert m >= 0
    return out[:m].copy(), first[:m].copy()


class FarthestDownsampler:
    """Preprocessor/Downsamplers/farthestDownsampler.py:10-57."""

    is_scaler = False

    def __init__(self, sample_size=4096):
        self._sample_size = sample_size if sample_size > 0 else 4096

    def process(self, cloud):
        first = np.random.randint(low=0, high=cloud.shape[0])
        self.last_indices = fps(cloud, first, self._sample_size).astype(np.int64)
        return cloud[self.last_indices]


class VoxelDownsampler:
    """Preprocessor/Downsamplers/voxelDownsampler.py:14-136 (compass search on the voxel size)."""

    is_scaler = False

    def __init__(self, target_points, base_voxel_size=0.01, min_voxel_size=0.0001, delta=0.01, eps=0.0005):
        if target_points <= 0:
            raise ValueError(f"target points cannot be 0 or less. Provided: {target_points}")
        self._target_points = target_points
        self._min_voxel_size = min_voxel_size if min_voxel_size > 0 else 0.0001
        self._base_voxel_size = base_voxel_size if base_voxel_size > 0 else 0.01
        self._delta = delta if delta > 0 else 0.01
        self._eps = eps if eps > 0 else 0.0005
        self.probes = []

    def _ds(self, cloud, size):
        pts, _ = voxel_downsample(cloud, size)
        self.probes.append((size, pts.shape[0]))
        return pts

    def compass_step(self, delta, cloud, current):
        new = current + delta
        if new <= self._min_voxel_size:
            new = self._min_voxel_size
        ds = self._ds(cloud, new)
        return ds, new, np.abs(self._target_points - ds.shape[0])

    def process(self, cloud):
        cur = self._base_voxel_size
        best = self._ds(cloud, cur)
        metric = np.abs(self._target_points - best.shape[0])
        while self._delta >= self._eps:
            ds, new, m = self.compass_step(self._delta, cloud, cur)
            if m < metric:
                cur, metric, best = new, m, ds
                continue
            ds, new, m = self.compass_step(-self._delta, cloud, cur)
            if m < metric:
                cur, metric, best = new, m, ds
                continue
            self._delta = self._delta / 2
        self.voxel_size = cur
        return best


class FastGlobalOptimizer:
    """Optimizer/fastGlobalOptimizer.py:23-190 restated.  ``feature_mode``:
    "reference_compat" reproduces the reference literally -- BOTH feature sets are the FPFH of the
    source cloud (fastGlobalOptimizer.py:137-138 passes source_copy twice);
    "fixed" computes the target's features from the target."""

    def __init__(self, division_factor=1.4, tuple_scale=0.9, maximum_correspondence_distance=0.5,
                 iteration_number=100, decrease_mu=True, normal_estimate_radius=0.1, normal_estimate_knn=20,
                 fpfh_radius=0.1, fpfh_knn=20, feature_mode="reference_compat", tuple_seed=0):
        self._division_factor = division_factor if division_factor > 0 else 1.4
        self._tuple_scale = tuple_scale if tuple_scale > 0 else 0.9
        self._maximum_correspondence_distance = (maximum_correspondence_distance
                                                 if maximum_correspondence_distance > 0 else 0.5)
        self._iteration_number = iteration_number if iteration_number > 0 else 100
        self._decrease_mu = decrease_mu
        self._normal_estimate_radius = normal_estimate_radius if normal_estimate_radius > 0 else 0.1
        self._normal_estimate_knn = normal_estimate_knn if normal_estimate_knn > 0 else 20
        self._fpfh_radius = fpfh_radius if fpfh_radius > 0 else 0.1
        self._fpfh_knn = fpfh_knn if fpfh_knn > 0 else 20
        self.feature_mode = feature_mode
        self.tuple_seed = tuple_seed
        self.calls = 0

    def get_fpfh_features(self, source, target):
        """fastGlobalOptimizer.py:109-144."""
        ns = normals(source, self._normal_estimate_knn, self._normal_estimate_radius)["normal"]
        fs = fpfh(source, ns, self._fpfh_knn, self._fpfh_radius)
        if self.feature_mode == "reference_compat":
            return fs, fs
        nt = normals(target, self._normal_estimate_knn, self._normal_estimate_radius)["normal"]
        return fs, fpfh(target, nt, self._fpfh_knn, self._fpfh_radius)

    def optimize(self, source, target, **kwargs):
        """fastGlobalOptimizer.py:146-190."""
        source = _f64(source, 3); target = _f64(target, 3)
        fs, ft = self.get_fpfh_features(source, target)
        r = fgr(source, target, fs, ft, self._division_factor, self._tuple_scale,
                self._maximum_correspondence_distance, self._iteration_number, self._decrease_mu,
                key=tuple_key(self.tuple_seed, source, target))
        self.calls += 1
        self.last = r
        roto = np.copy(r["T"])
        roto[:3, :3] = roto[:3, :3].T
        if r["n_corr"] == 0:
            raise Warning("Optimization failed with no correspondences. Parameters are not well set.")
        return roto, r["rmse"]


def initialize_rotation(deg: float, mu: float, std: float):
    """Aligner/Aligner.py:125-162: three scalar uniform draws then randn(3)."""
    t1 = np.random.uniform(low=-deg, high=deg)
    t2 = np.random.uniform(low=-deg, high=deg)
    t3 = np.random.uniform(low=-deg, high=deg)
    r1 = np.array([[1, 0, 0], [0, np.cos(t1), -np.sin(t1)], [0, np.sin(t1), np.cos(t1)]])
    r2 = np.array([[np.cos(t2), 0, np.sin(t2)], [0, 1, 0], [-np.sin(t2), 0, np.cos(t2)]])
    r3 = np.array([[np.cos(t3), -np.sin(t3), 0], [np.sin(t3), np.cos(t3), 0], [0, 0, 1]])
    rot = np.dot(r1, np.dot(r2, r3))
    trans = mu + np.random.randn(3) * std
    return rot, trans


class Aligner:
    """Aligner/Aligner.py:32-317 restated (no GUI, no logging, no refine step)."""

    def __init__(self, source_preprocessor, target_preprocessor, optimizer, attempts=30,
                 deg=np.pi / 2, mu=0.0, std=0.1, delta=0.2, max_iter=100, eps=0.05):
        self._source_preprocessor = source_preprocessor
        self._target_preprocessor = target_preprocessor
        self._optimizer = optimizer
        self._attempts = attempts if attempts > 0 else 30
        self._deg = deg if deg > 0 else np.pi / 2
        self._mu = mu if mu >= 0 else 0.0
        self._std = std if std > 0 else 0.1
        self._delta = delta if delta > 0 else 0.2
        self._max_iter = max_iter if max_iter > 0 else 100
        self._eps = eps if eps > 0 else 0.05
        self.decision_log: List[dict] = []
        self.transfromation = np.eye(4)
        self.scale_factors = np.ones((1, 3))

    def multistart_registration(self, source, target):
        """Aligner.py:164-204."""
        metric = np.inf
        best = np.eye(4)
        self.last_start_metrics = []
        for _ in range(self._attempts):
            rot, trans = initialize_rotation(self._deg, self._mu, self._std)
            src_init = np.dot(source.copy(), rot) + trans
            cur_T, cur_m = self._optimizer.optimize(src_init, target)
            self.last_start_metrics.append(cur_m)
            if cur_m < metric:
                metric = cur_m
                T = np.eye(4)
                T[:3, :3] = np.dot(rot, cur_T[:3, :3])
                T[:3, 3] = np.dot(trans, cur_T[:3, :3]).ravel() + cur_T[:3, 3]
                best = T
        return best, metric

    def compass_step(self, source, target, scale_factors, delta):
        """Aligner.py:206-226."""
        new_scale = scale_factors + delta
        target_scaled = target * new_scale
        rot, metric = self.multistart_registration(source, target_scaled)
        return new_scale, rot, metric

    def align(self, source, target, refine_registration=False):
        """Aligner.py:228-317 with refine_registration=False."""
        assert not refine_registration, "refinement (Open3D ICP) is outside the oracle's scope"
        source = self._source_preprocessor.preprocess(source)
        target = self._target_preprocessor.preprocess(target)
        self.source_pre, self.target_pre = source, target
        iteration = 0
        opt_scale = np.ones((1, 3))
        opt_T, opt_m = self.multistart_registration(source, target)
        self.decision_log.append(dict(scale=opt_scale.copy(), metric=opt_m, accepted=True, kind="init"))
        errors = [opt_m]
        directions = np.eye(3)
        while self._delta >= self._eps and iteration <= self._max_iter:
            iteration += 1
            for axis in range(3):
                step = self._delta * directions[:, axis]
                new_scale, new_T, new_m = self.compass_step(source, target, opt_scale, step)
                acc = bool(new_m <= opt_m)
                self.decision_log.append(dict(scale=new_scale.copy(), metric=new_m, accepted=acc, kind="+%d" % axis))
                if acc:
                    opt_m = new_m; opt_T = new_T; opt_scale += step; errors.append(new_m)
                    break
                step = -self._delta * directions[:, axis]
                new_scale, new_T, new_m = self.compass_step(source, target, opt_scale, step)
                acc = bool(new_m <= opt_m)
                self.decision_log.append(dict(scale=new_scale.copy(), metric=new_m, accepted=acc, kind="-%d" % axis))
                if acc:
                    opt_m = new_m; opt_T = new_T; opt_scale += step; errors.append(new_m)
                    break
            if new_m > opt_m:
                self._delta = self._delta / 2
        self.transfromation = opt_T
        self.scale_factors = opt_scale
        return opt_T, opt_m, opt_scale, errors

    def transfrom(self, source):
        """Aligner.py:367-394."""
        for b in self._source_preprocessor.preprocessor_blocks:
            if getattr(b, "is_scaler", False):
                s_mean, s_dist = b.mean, b.scale
        for b in self._target_preprocessor.preprocessor_blocks:
            if getattr(b, "is_scaler", False):
                t_mean, t_dist = b.mean, b.scale
        source = (source - s_mean) / s_dist
        source = np.dot(source, self.transfromation[:3, :3]) + self.transfromation[:3, 3]
        source = source / self.scale_factors
        return source * t_dist + t_mean


# --------------------------------------------------------------------------
# Fixtures
# --------------------------------------------------------------------------
def load_scan(name: str) -> np.ndarray:
    """Armadillo range scans (reference data/*.ply vertices, float32 -> float64 as
    Open3D's PLY reader does), stored as multi-scale-pointcloud-registration_b200/or_pcd_b200/data/armadillo_scans.npz (a data file; no product code is imported)."""
    p = os.path.join(os.path.dirname(_HERE), "multi-scale-pointcloud-registration_b200", "or_pcd_b200", "data",
                     "armadillo_scans.npz")
    with np.load(p) as z:
        return z[name].astype(np.float64)


def synthetic_cloud(n: int, rng: np.random.Generator) -> np.ndarray:
    """SURVEY.md 8(d) config-4 surface sampler."""
    u = rng.standard_normal((n, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    r = 1 + 0.3 * np.sin(3 * u[:, 0]) * np.cos(2 * u[:, 1]) + 0.2 * u[:, 2] ** 3
    return r[:, None] * u + rng.normal(0, 0.002, (n, 3))
