"""Fast Global Registration inner block on the B200
(reference: Optimizer/fastGlobalOptimizer.py:23-203, which wraps Open3D's estimate_normals,
compute_fpfh_feature and registration_fgr_based_on_feature_matching; algorithm SURVEY.md A.7-A.8).

``feature_mode``
  "reference_compat" (default) reproduces the reference literally: BOTH feature sets handed to
      Open3D are the FPFH of the SOURCE cloud (fastGlobalOptimizer.py:137-138 passes
      ``source_copy`` twice).  Pairs whose target index would exceed the target
      cloud -- an out-of-bounds read inside Open3D -- are dropped.
  "fixed" computes the target's features from the target cloud (what the code meant to do).

The reference never seeds Open3D's random generator, so its tuple test is not reproducible
(SURVEY.md B.5); here the stream is Philox4x32-10 keyed by ``tuple_key(tuple_seed, source, target)``."""
from typing import Optional, Sequence, Tuple

import numpy as np

from .. import _lib, ops
from ..utils.constants import (
    DECREASE_MU,
    DIVISION_FACTOR,
    FPFH_KNN,
    FPFH_RADIUS,
    ITERATION_NUMBER,
    MAXIMUM_CORRESPONDENCE_DISTANCE,
    NORMAL_ESTIMATE_KNN,
    NORMAL_ESTIMATE_RADIUS,
    TUPLE_SCALE,
)
from ..utils.logger_factory import LoggerFactory
from .iOptimizer import BatchResult, IOptimizer

_NO_CORR_MSG = ("Optimization failed with no correspondences. Parameters are not well set. "
                "Probably due to maximum_correspondence_distance set too low.")
_M64 = (1 << 64) - 1
MAXIMUM_TUPLE_COUNT = 1000   # FastGlobalRegistrationOption default, not exposed by the reference


def tuple_key(seed: int, source: np.ndarray, target: np.ndarray) -> int:
    """Key of one registration's tuple-test stream: a hash of the seed, the two cloud sizes and the
    bit patterns of the first and last point of each cloud, so that a sequential caller and a batched
    one draw the same stream for the same inputs with no shared counter."""
    h = (int(seed) * 0x9E3779B97F4A7C15 + 0x243F6A8885A308D3) & _M64
    words = [source.shape[0], target.shape[0]]
    for c in (source[0], source[-1], target[0], target[-1]):
        words.extend(int(v) for v in np.ascontiguousarray(c, dtype=np.float64).view(np.uint64))
    for w in words:
        h = ((h ^ (w & _M64)) * 0xBF58476D1CE4E5B9) & _M64
        h ^= h >> 31
    return h


def _positive(value, default, name, log):
    if value <= 0:
        log.warning(f"{name} cannot be 0 or less. Provided: {value}")
        return default
    return value


class FastGlobalOptimizer(IOptimizer):
    RERANK_KC = 48

    def __init__(
        self,
        division_factor: float = DIVISION_FACTOR,
        tuple_scale: float = TUPLE_SCALE,
        maximum_correspondence_distance: float = MAXIMUM_CORRESPONDENCE_DISTANCE,
        iteration_number: int = ITERATION_NUMBER,
        decrease_mu: bool = DECREASE_MU,
        normal_estimate_radius: float = NORMAL_ESTIMATE_RADIUS,
        normal_estimate_knn: int = NORMAL_ESTIMATE_KNN,
        fpfh_radius: float = FPFH_RADIUS,
        fpfh_knn: int = FPFH_KNN,
        feature_mode: str = "reference_compat",
        tuple_seed: int = 0,
    ):
        self._LOG = LoggerFactory.get_logger(log_name=self.__class__.__name__, log_on_file=False)
        log = self._LOG
        self._division_factor = _positive(division_factor, DIVISION_FACTOR, "division factor", log)
        self._tuple_scale = _positive(tuple_scale, TUPLE_SCALE, "tuple scale", log)
        self._maximum_correspondence_distance = _positive(
            maximum_correspondence_distance, MAXIMUM_CORRESPONDENCE_DISTANCE, "maximum correspondence distance", log)
        self._iteration_number = _positive(iteration_number, ITERATION_NUMBER, "iteration number", log)
        self._decrease_mu = decrease_mu
        self._normal_estimate_radius = _positive(normal_estimate_radius, NORMAL_ESTIMATE_RADIUS,
                                                 "normal estimate radius", log)
        self._normal_estimate_knn = _positive(normal_estimate_knn, NORMAL_ESTIMATE_KNN, "normal estimate knn", log)
        self._fpfh_radius = _positive(fpfh_radius, FPFH_RADIUS, "fpfh radius", log)
        self._fpfh_knn = _positive(fpfh_knn, FPFH_KNN, "fpfh knn", log)
        if feature_mode not in ("reference_compat", "fixed"):
            raise ValueError("feature_mode must be 'reference_compat' or 'fixed'")
        self.feature_mode = feature_mode
        self.tuple_seed = tuple_seed
        self.match_mode = 0            # 0: tcgen05 filter + exact re-score; 1: float64 scan only
        self._base_key = None
        self._base_lists = None
        self._ws = None
        self.last_info = None          # int array (U,4): cross-checked pairs, tuple pairs, trials, swapped
        self.last_exact_scans = None   # feature rows that needed the exact float64 scan
        self.last_rescored_chunks = None   # 32-row chunks re-scored in float64 by the certificate

    # ------------------------------------------------------------------ features
    def _neighbour_lists(self, d_clouds, k: int, rigid_base: Optional[np.ndarray]):
        """Exact k-NN lists (self query) of every cloud [B,N,3]; rigid images of one base cloud are
        re-ranked from the base cloud's candidate lists (same result, certified per point)."""
        t = _lib.torch()
        if rigid_base is not None and d_clouds.shape[1] > self.RERANK_KC and k < self.RERANK_KC:
            import hashlib

            base = np.ascontiguousarray(rigid_base, dtype=np.float64)
            key = (base.shape, hashlib.blake2b(base.tobytes(), digest_size=16).digest())
            if key != self._base_key:
                self._base_lists = ops.knn_candidates(_lib.to_device(base), self.RERANK_KC)
                self._base_key = key
            cand, bound, order = self._base_lists
            factor = t.full((d_clouds.shape[0],), 1.0 - 1e-9, dtype=t.float64, device="cuda")
            idx, d2, _ = ops.knn_rerank(cand, bound, d_clouds, factor, k, order=order)
            return idx, d2
        return ops.knn_exact(d_clouds, d_clouds, k)

    def _fpfh_sets(self, d_clouds, rigid_base=None):
        """get_fpfh_features for a batch of clouds (fastGlobalOptimizer.py:109-144): normals with
        Hybrid(normal radius, normal knn), then FPFH with Hybrid(fpfh radius, fpfh knn)."""
        kn, kf = int(self._normal_estimate_knn), int(self._fpfh_knn)
        idx, d2 = self._neighbour_lists(d_clouds, max(kn, kf), rigid_base)
        if kn == idx.shape[2]:
            n_idx, n_d2 = idx, d2
        else:
            n_idx, n_d2 = idx[:, :, :kn].contiguous(), d2[:, :, :kn].contiguous()
        normals = ops.normals_cov_from_knn(d_clouds, n_idx, n_d2, radius=self._normal_estimate_radius,
                                           want_normals=True, want_gicp_cov=False)["normal"]
        return ops.fpfh(d_clouds, normals, idx, d2, kf, self._fpfh_radius), normals

    def get_fpfh_features(self, source: np.ndarray, target: np.ndarray):
        """Host-array face of the reference's method: (source_fpfh, target_fpfh) as (N,33) arrays."""
        fs, _ = self._fpfh_sets(_lib.to_device(np.asarray(source, dtype=np.float64)[None]))
        fs = fs[0].cpu().numpy()
        if self.feature_mode == "reference_compat":
            return fs, fs
        ft, _ = self._fpfh_sets(_lib.to_device(np.asarray(target, dtype=np.float64)[None]))
        return fs, ft[0].cpu().numpy()

    # ------------------------------------------------------------------ reference API
    def optimize(self, source: np.ndarray, target: np.ndarray, **kwargs) -> Tuple[np.ndarray, float]:
        res = self.optimize_batch(np.asarray(source)[None], np.asarray(target)[None], [0], [0])
        if res.status[0] != 0:
            self.raise_failure()
        return res.transforms[0], float(res.rmse[0])

    def raise_failure(self):
        """Empty correspondence set (fastGlobalOptimizer.py:181-188)."""
        self._LOG.error(_NO_CORR_MSG)
        raise Warning(_NO_CORR_MSG)

    # ------------------------------------------------------------------ batched API
    accepts_device_sources = True   # start clouds built on the GPU are bit-identical to the host's (Aligner verifies)

    def optimize_batch(self, sources, targets, unit_src: Sequence[int], unit_tgt: Sequence[int],
                       rigid_base: Optional[np.ndarray] = None) -> BatchResult:
        t = _lib.torch()
        d_src = sources if hasattr(sources, "is_cuda") else _lib.to_device(np.asarray(sources))
        d_tgt = targets if hasattr(targets, "is_cuda") else _lib.to_device(np.asarray(targets))
        unit_src = np.asarray(unit_src, dtype=np.int32)
        unit_tgt = np.asarray(unit_tgt, dtype=np.int32)
        us, ut = t.as_tensor(unit_src).to("cuda"), t.as_tensor(unit_tgt).to("cuda")
        U = len(unit_src)

        fs, _ = self._fpfh_sets(d_src, rigid_base)
        pf_s = ops.feature_pack(fs)
        scans = 0
        if self.feature_mode == "reference_compat":
            # one table per source set, matched against itself; units index it
            sets = t.arange(d_src.shape[0], dtype=t.int32, device="cuda")
            nn, _, st, _ = ops.feature_match(pf_s, pf_s, sets, sets, mode=self.match_mode)
            ij = nn[us.long()].contiguous()
            ji = ij
            scans = st
        else:
            ft, _ = self._fpfh_sets(d_tgt)
            pf_t = ops.feature_pack(ft, center=pf_s.center)
            ij, _, st1, _ = ops.feature_match(pf_s, pf_t, us, ut, mode=self.match_mode)
            # the cross check reads the reverse table only at the images of the forward one (ji[ij[i]] == i):
            # the other target rows are not resolved
            needed = t.zeros((U, d_tgt.shape[1]), dtype=t.uint8, device="cuda")
            needed.scatter_(1, ij.clamp(min=0).long(), 1)
            ji, _, st2, _ = ops.feature_match(pf_t, pf_s, ut, us, mode=self.match_mode, row_mask=needed)
            scans = st1 + st2

        # tuple-test keys (tuple_key of every unit's two clouds), hashed on the device
        d_keys = ops.fgr_tuple_keys(d_src, d_tgt, us, ut, self.tuple_seed)

        T, info = ops.fgr_batch(d_src, d_tgt, us, ut, ij, ji, d_keys, self._division_factor, self._tuple_scale,
                                self._maximum_correspondence_distance, self._iteration_number, self._decrease_mu,
                                MAXIMUM_TUPLE_COUNT, True)
        stats, status, _, self._ws = ops.evaluate_batch(d_src, d_tgt, us, ut, T, self._maximum_correspondence_distance,
                                                       ws=self._ws)
        Th = T.cpu().numpy()
        stats = stats.cpu().numpy()
        self.last_info = info.cpu().numpy()
        self.last_exact_scans = int(scans[0].item()) if hasattr(scans, "cpu") else int(scans)
        self.last_rescored_chunks = int(scans[1].item()) if hasattr(scans, "cpu") else 0
        Th[:, :3, :3] = np.transpose(Th[:, :3, :3], (0, 2, 1)).copy()       # fastGlobalOptimizer.py:176-178
        # the reference fails iff the final correspondence set is EMPTY (fastGlobalOptimizer.py:181-188); an exact
        # alignment with rmse == 0 and correspondences is a success (unlike GICP's rmse == 0 rule)
        status = (stats[:, 3] == 0).astype(np.int32)
        return BatchResult(Th, stats[:, 1].copy(), stats[:, 0].copy(), self.last_info[:, 1].astype(np.int64),
                           status, [])

    def __repr__(self):
        return f"""{self.__class__.__name__}
            (division_factor={self._division_factor},
            tuple_scale={self._tuple_scale},
            maximum_correspondence_distance={self._maximum_correspondence_distance},
            iteration_number={self._iteration_number},
            decrease_mu={self._decrease_mu},
            normal_estimate_radius={self._normal_estimate_radius},
            normal_estimate_knn={self._normal_estimate_knn},
            fpfh_radius={self._fpfh_radius},
            fpfh_knn={self._fpfh_knn})
        """


