"""Generalized-ICP inner block on the B200
(reference: Optimizer/generalizedICP.py:15-89, which wraps Open3D's
registration_generalized_icp; algorithm SURVEY.md A.3-A.6)."""
from typing import Optional, Sequence, Tuple

import numpy as np

from .. import _lib, ops
from ..utils.constants import (
    GICP_COVARIANCE_KNN,
    GICP_EPSILON,
    ICP_RELATIVE_FITNESS,
    ICP_RELATIVE_RMSE,
    MAX_ITERATIONS,
    MAXIMUM_CORRESPONDENCE_DISTANCE,
)
from ..utils.logger_factory import LoggerFactory
from .iOptimizer import BatchResult, IOptimizer

_ZERO_LOSS_MSG = (
    "Optimization failed with loss = 0. Parameters are not well set. "
    "Probably due to maximum_correspondence_distance set too low."
)


class GeneralizedICP(IOptimizer):
    def __init__(
        self,
        max_correspondence_distance: float = MAXIMUM_CORRESPONDENCE_DISTANCE,
        max_iterations: int = MAX_ITERATIONS,
    ):
        self._LOG = LoggerFactory.get_logger(log_name=self.__class__.__name__, log_on_file=False)
        if max_correspondence_distance <= 0:
            self._LOG.warning(
                f"max correspondence distance cannot be 0 or less. Provided: {max_correspondence_distance}")
            max_correspondence_distance = MAXIMUM_CORRESPONDENCE_DISTANCE
        if max_iterations <= 0:
            self._LOG.warning(f"max iterations cannot be 0 or less. Provided: {max_iterations}")
            max_iterations = MAX_ITERATIONS
        self._max_correspondence_distance = max_correspondence_distance
        self._max_iterations = max_iterations
        self._ws = None
        self._base_key = None         # (shape, digest) of the cached base source cloud
        self._base_lists = None       # its candidate neighbour lists on the device
        # closed-form whitening of plane-covariance pairs (C ABI cov_mode 1); False runs the
        # generic 3x3 covariance path with a Jacobi eigen-solve per correspondence
        self.analytic_whitening = True
        self.profile = False          # set True to collect sweep timings in BatchResult.profile
        self.sweep_mode = 0           # tuning: 0 = persistent work-queue kernel, 5 = one fused launch per round, 4 = culled sweep + separate kernels, 1..3 = dense sweeps
        self.sweep_q = 0              # tuning: queries per thread (4 | 8), 0 = auto
        self.target_splits = 0        # tuning: target splits per sweep, 0 = auto
        self.sweep_tj = 0             # tuning: dense sweep targets per step; culled sweep warps per block (2->1, 4->2, 8->4)
        self.sweep_smem_pad = 0       # tuning: dense sweep occupancy pad; negative -(n+1): small patches below n
        self.half_patch_from_round = int(__import__("os").environ.get("ORPCD_HALF_FROM_ROUND", "0"))   # tuning: 0 = default
        self.last_profile = None

    # ------------------------------------------------------------------ reference API
    def optimize(self, source: np.ndarray, target: np.ndarray, **kwargs) -> Tuple[np.ndarray, float]:
        res = self.optimize_batch(np.asarray(source)[None], np.asarray(target)[None], [0], [0], poll_interval=8)
        if res.status[0] != 0:
            self.raise_failure()
        return res.transforms[0], float(res.rmse[0])

    def raise_failure(self):
        """inlier_rmse == 0 (generalizedICP.py:77-81)."""
        self._LOG.error(_ZERO_LOSS_MSG)
        raise ValueError(_ZERO_LOSS_MSG)

    # ------------------------------------------------------------------ batched API
    def optimize_batch(self, sources, targets, unit_src: Sequence[int], unit_tgt: Sequence[int],
                       poll_interval: int = 0, rigid_base: Optional[np.ndarray] = None) -> BatchResult:
        """Run ``len(unit_src)`` independent registrations concurrently.

        sources: (S,Ns,3) host float64 array (or CUDA tensor), the initial source clouds;
        targets: (C,Nt,3); unit u registers sources[unit_src[u]] onto targets[unit_tgt[u]].
        rigid_base: optional (Ns,3) cloud of which every source is a rigid image (the Aligner's
        random starts): their 20-NN graphs are then re-ranked from the base cloud's neighbour
        lists instead of searched from scratch -- same result, certified per point.
        Nothing raises for a unit whose final inlier RMSE is 0; it is flagged in ``status``."""
        t = _lib.torch()
        d_src = sources if hasattr(sources, "is_cuda") else _lib.to_device(np.asarray(sources))
        d_tgt = targets if hasattr(targets, "is_cuda") else _lib.to_device(np.asarray(targets))
        us = t.as_tensor(np.asarray(unit_src, dtype=np.int32)).to("cuda")
        ut = t.as_tensor(np.asarray(unit_tgt, dtype=np.int32)).to("cuda")
        src_cov, tgt_cov = self.covariances(d_src, d_tgt, rigid_base)
        d_res = self.optimize_batch_device(d_src, d_tgt, us, ut, poll_interval=poll_interval, src_cov=src_cov,
                                           tgt_cov=tgt_cov, rigid_copies=rigid_base is not None)
        T = d_res.T.cpu().numpy()
        stats = d_res.stats.cpu().numpy()
        status = d_res.status.cpu().numpy()
        T[:, :3, :3] = np.transpose(T[:, :3, :3], (0, 2, 1)).copy()   # generalizedICP.py:72-74
        self.last_profile = d_res.profile
        return BatchResult(T, stats[:, 1].copy(), stats[:, 0].copy(), stats[:, 2].astype(np.int64), status,
                           d_res.profile)

    def optimize_batch_records(self, sources, targets, unit_src: Sequence[int], unit_tgt: Sequence[int],
                               poll_interval: int = 0, rigid_base: Optional[np.ndarray] = None):
        """``optimize_batch`` with the results left on the device as one (U, 20) float64 record per unit --
        rmse, fitness, iterations, status, the 16 entries of the transform with its rotation block transposed
        (generalizedICP.py:72-74) -- the layout ShardedOptimizer all-gathers between the GPUs."""
        t = _lib.torch()
        d_src = sources if hasattr(sources, "is_cuda") else _lib.to_device(np.asarray(sources))
        d_tgt = targets if hasattr(targets, "is_cuda") else _lib.to_device(np.asarray(targets))
        us = t.as_tensor(np.asarray(unit_src, dtype=np.int32)).to("cuda")
        ut = t.as_tensor(np.asarray(unit_tgt, dtype=np.int32)).to("cuda")
        src_cov, tgt_cov = self.covariances(d_src, d_tgt, rigid_base)
        r = self.optimize_batch_device(d_src, d_tgt, us, ut, poll_interval=poll_interval, src_cov=src_cov,
                                       tgt_cov=tgt_cov, rigid_copies=rigid_base is not None)
        self.last_profile = r.profile
        U = us.shape[0]
        T = r.T.clone()
        T[:, :3, :3] = r.T[:, :3, :3].transpose(1, 2)
        rec = t.empty((U, 20), dtype=t.float64, device="cuda")
        rec[:, 0] = r.stats[:, 1]
        rec[:, 1] = r.stats[:, 0]
        rec[:, 2] = r.stats[:, 2]
        rec[:, 3] = r.status.to(t.float64)
        rec[:, 4:] = T.reshape(U, 16)
        return rec

    RERANK_KC = 48   # candidates kept per point of the base cloud (k = 20 are needed)
    accepts_device_sources = True   # the Aligner may hand over start clouds built on the GPU (CUDA tensors)

    def covariances(self, d_src, d_tgt, rigid_base: Optional[np.ndarray] = None):
        """GICP covariance data (plane axes or 3x3 covariances) of every source and target set.  The two sides are
        independent: the target side runs on a second CUDA stream next to the source side."""
        t = _lib.torch()
        ax = self.analytic_whitening
        key = "gicp_axis" if ax else "gicp_cov"
        main = t.cuda.current_stream()
        side = _lib.side_stream()
        side.wait_stream(main)
        with t.cuda.stream(side):
            tgt_cov = ops.normals_cov(d_tgt, knn=GICP_COVARIANCE_KNN, gicp_eps=GICP_EPSILON, want_gicp_cov=not ax,
                                      want_gicp_axis=ax)[key]
        if rigid_base is not None and d_src.shape[1] > self.RERANK_KC:
            src_cov = self._rigid_source_covariances(rigid_base, d_src)
        else:
            src_cov = ops.normals_cov(d_src, knn=GICP_COVARIANCE_KNN, gicp_eps=GICP_EPSILON, want_gicp_cov=not ax,
                                      want_gicp_axis=ax)[key]
        main.wait_stream(side)
        tgt_cov.record_stream(main)
        return src_cov, tgt_cov

    def _rigid_source_covariances(self, rigid_base: np.ndarray, d_src):
        """GICP covariances of rigid images of ``rigid_base`` via re-ranked neighbour lists."""
        import hashlib

        t = _lib.torch()
        base = np.ascontiguousarray(rigid_base, dtype=np.float64)
        key = (base.shape, hashlib.blake2b(base.tobytes(), digest_size=16).digest())
        if key != self._base_key:
            self._base_lists = ops.knn_candidates(_lib.to_device(base), self.RERANK_KC)
            self._base_key = key
        cand, bound, order = self._base_lists
        factor = t.full((d_src.shape[0],), 1.0 - 1e-9, dtype=t.float64, device="cuda")
        idx, d2, n_fixed = ops.knn_rerank(cand, bound, d_src, factor, GICP_COVARIANCE_KNN, order=order)
        self.last_rerank_fixed = n_fixed
        ax = self.analytic_whitening
        return ops.normals_cov_from_knn(d_src, idx, d2, gicp_eps=GICP_EPSILON, want_gicp_cov=not ax,
                                        want_gicp_axis=ax, order=order[0])["gicp_axis" if ax else "gicp_cov"]

    def optimize_batch_device(self, d_src, d_tgt, unit_src, unit_tgt, poll_interval: int = 0, want_trace=False,
                              src_cov=None, tgt_cov=None, rigid_copies: bool = False):
        """Device-resident variant: CUDA tensors in, ``ops.GicpBatchResult`` (CUDA tensors) out."""
        ax = self.analytic_whitening
        key = "gicp_axis" if ax else "gicp_cov"
        if src_cov is None:
            src_cov = ops.normals_cov(d_src, knn=GICP_COVARIANCE_KNN, gicp_eps=GICP_EPSILON, want_gicp_cov=not ax,
                                      want_gicp_axis=ax)[key]
        if tgt_cov is None:
            tgt_cov = ops.normals_cov(d_tgt, knn=GICP_COVARIANCE_KNN, gicp_eps=GICP_EPSILON, want_gicp_cov=not ax,
                                      want_gicp_axis=ax)[key]
        res = ops.gicp_batch(d_src, src_cov, d_tgt, tgt_cov, unit_src, unit_tgt,
                             max_correspondence_distance=self._max_correspondence_distance,
                             max_iterations=self._max_iterations, relative_fitness=ICP_RELATIVE_FITNESS,
                             relative_rmse=ICP_RELATIVE_RMSE, poll_interval=poll_interval, profile=self.profile,
                             want_trace=want_trace, ws=self._ws, gicp_epsilon=GICP_EPSILON,
                             sweep_mode=self.sweep_mode, sweep_q=self.sweep_q, target_splits=self.target_splits,
                             sweep_tj=self.sweep_tj, sweep_smem_pad=self.sweep_smem_pad,
                             src_rigid_copies=rigid_copies, half_patch_from_round=self.half_patch_from_round)
        self._ws = res.workspace
        return res

    def __repr__(self):
        return f"""{self.__class__.__name__}
            (max_correspondence_distance={self._max_correspondence_distance},
            max_iterations={self._max_iterations})
        """
