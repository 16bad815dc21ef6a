"""Outer block: multi-start pattern (compass) search over per-axis scale factors
(reference: Aligner/Aligner.py:32-408).

The reference evaluates one inner registration at a time.  Here every poll
candidate of a sweep (x+, x-, y+, y-, z+, z-) and every random start is one unit
of a single batched launch sequence on the GPU (``IOptimizer.optimize_batch``);
the accept logic then replays the reference's order and decisions exactly:
  * starts are drawn from the global legacy ``np.random`` stream in the order the
    sequential code would draw them, and the stream is rewound to the state the
    sequential code would have left after the first accepted candidate
    (SURVEY.md B.1);
  * a unit that would have raised in the reference raises here only if the
    sequential code would have reached it.
Optimizers without ``optimize_batch`` are driven sequentially, like the reference."""
import time
from typing import List, Optional, Tuple

import numpy as np

from ..Optimizer.iOptimizer import IOptimizer
from ..Preprocessor.preprocessor import Preprocessor
from ..Preprocessor.Scalers import BaseScaler
from ..utils.constants import (
    ALIGNER_DEG,
    ALIGNER_DELTA,
    ALIGNER_EPS,
    ALIGNER_MAX_ITER,
    ALIGNER_MU,
    ALIGNER_STD,
    MULTISTART_ATTEMPTS,
    REFINER_DISTANCE_THRESHOLD,
    REFINER_MAX_ITER,
)
from ..utils.logger_factory import LoggerFactory

_ZERO_LOSS_MSG = (
    "Optimization failed with loss = 0. Parameters are not well set. "
    "Probably due to maximum_correspondence_distance set too low."
)


def _checked(value, ok, default, name, log):
    if ok(value):
        return value
    log.warning(f"{name} cannot be 0 or less. Provided: {value}")
    return default


class Aligner:
    transfromation: np.ndarray = np.eye(4)   # (sic) attribute names follow the reference
    scale_factors: np.ndarray = np.ones((1, 3))

    def __init__(
        self,
        source_preprocessor: Preprocessor,
        target_preprocessor: Preprocessor,
        optimizer: IOptimizer,
        attempts: int = MULTISTART_ATTEMPTS,
        deg: float = ALIGNER_DEG,
        mu: float = ALIGNER_MU,
        std: float = ALIGNER_STD,
        delta: float = ALIGNER_DELTA,
        max_iter: int = ALIGNER_MAX_ITER,
        eps: float = ALIGNER_EPS,
        visualize_intermediate_steps: bool = False,
    ):
        self._LOG = LoggerFactory.get_logger(log_name=self.__class__.__name__, log_on_file=True)
        pos = lambda v: v > 0
        self._attempts = _checked(attempts, pos, MULTISTART_ATTEMPTS, "attempts", self._LOG)
        self._deg = _checked(deg, pos, ALIGNER_DEG, "deg", self._LOG)
        self._mu = _checked(mu, lambda v: v >= 0, ALIGNER_MU, "mu", self._LOG)
        self._std = _checked(std, pos, ALIGNER_STD, "std", self._LOG)
        self._delta = _checked(delta, pos, ALIGNER_DELTA, "delta", self._LOG)
        self._max_iter = _checked(max_iter, pos, ALIGNER_MAX_ITER, "max_iter", self._LOG)
        self._eps = _checked(eps, pos, ALIGNER_EPS, "eps", self._LOG)
        if visualize_intermediate_steps:
            self._LOG.warning("visualisation (Open3D GUI) is outside the B200 hot path; ignored")
        self._visualize_intermediate_steps = False
        self._source_preprocessor = source_preprocessor
        self._target_preprocessor = target_preprocessor
        self._optimizer = optimizer
        self.decision_log: List[dict] = []
        self.n_optimize_calls = 0          # inner registrations the sequential reference would have run
        self.n_units_evaluated = 0         # inner registrations actually run (speculation included)
        self.device_starts = True          # build the start clouds on the GPU when the optimizer allows it
        self._d_source, self._d_source_key = None, None
        self._d_target, self._d_target_key = None, None
        self._device_starts_verified = False

    # ------------------------------------------------------------------ starts
    def initialize_rotation(self) -> Tuple[np.ndarray, np.ndarray]:
        """One random start: three scalar ``uniform`` draws, then ``randn(3)`` --
        this order defines the RNG stream (reference: Aligner.py:125-162)."""
        a = np.random.uniform(low=-self._deg, high=self._deg)
        b = np.random.uniform(low=-self._deg, high=self._deg)
        g = np.random.uniform(low=-self._deg, high=self._deg)
        rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
        ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
        rz = np.array([[np.cos(g), -np.sin(g), 0], [np.sin(g), np.cos(g), 0], [0, 0, 1]])
        rotation = np.dot(rx, np.dot(ry, rz))
        translation = self._mu + np.random.randn(3) * self._std
        return rotation, translation

    def _device_starts(self) -> bool:
        """True when the optimizer takes start clouds built on the GPU: ``source @ R + t`` evaluated by
        orpcd_apply_rigid_batch with the arithmetic of numpy's BLAS (k-ordered fused multiply-adds), verified
        bit for bit against the host expression on the first batch -- a 1-ulp difference in a start cloud
        is enough to move a non-converged registration to another local minimum."""
        return self.device_starts and bool(getattr(self._optimizer, "accepts_device_sources", False))

    def _draw_starts(self, source: np.ndarray):
        """Draw ``attempts`` starts.  Host path: their initial source clouds are built with the reference's
        own expression (Aligner.py:183-185) so inputs are bit-identical; device path: clouds is None and
        ``_start_clouds`` builds the whole batch on the GPU."""
        rots, trans = [], []
        device = self._device_starts()
        clouds = None if device else np.empty((self._attempts,) + source.shape, dtype=np.float64)
        for n in range(self._attempts):
            r, t = self.initialize_rotation()
            rots.append(r)
            trans.append(t)
            if not device:
                clouds[n] = np.dot(source, r) + t
        return rots, trans, clouds

    def _start_clouds(self, source: np.ndarray, starts):
        """Initial source clouds of a list of ``_draw_starts`` results, one batch: host ndarray, or a CUDA
        tensor when the optimizer accepts device sources."""
        if not self._device_starts():
            return np.concatenate([s[2] for s in starts], axis=0)
        from .. import _lib, ops

        if self._d_source_key is not source:   # the array itself is kept, so the identity cannot be recycled
            self._d_source, self._d_source_key = _lib.to_device(source), source
        rots = np.stack([r for s in starts for r in s[0]])
        trans = np.stack([t for s in starts for t in s[1]])
        clouds = ops.apply_rigid_batch(self._d_source, _lib.to_device(rots), _lib.to_device(trans))
        if not self._device_starts_verified:
            k = len(rots) - 1
            if not np.array_equal(clouds[k].cpu().numpy(), np.dot(source, rots[k]) + trans[k]):
                self._LOG.warning("device start clouds differ from numpy's on this host: building them on the host")
                self.device_starts = False
                return np.stack([np.dot(source, r) + t for r, t in zip(rots, trans)])
            self._device_starts_verified = True
        return clouds

    def _batch_kwargs(self, source):
        """Tell optimizers that understand it that every start is a rigid image of ``source``."""
        import inspect

        try:
            if "rigid_base" in inspect.signature(self._optimizer.optimize_batch).parameters:
                return {"rigid_base": source}
        except (TypeError, ValueError):
            pass
        return {}

    @staticmethod
    def _compose(rot, trans, current):
        """Start pose followed by the optimizer's result (reference: Aligner.py:196-201)."""
        T = np.eye(4)
        T[:3, :3] = np.dot(rot, current[:3, :3])
        T[:3, 3] = np.dot(trans, current[:3, :3]).ravel() + current[:3, 3]
        return T

    def _best_of_starts(self, rots, trans, transforms, metrics, status):
        """Sequential strict-'<' selection over the starts (Aligner.py:178-202); raises where
        the reference's optimizer would have raised (generalizedICP.py:77-81)."""
        metric, best = np.inf, np.eye(4)
        for n in range(len(rots)):
            self.n_optimize_calls += 1
            if status[n] != 0:
                # what optimize() would have raised for this start (GICP: ValueError, FGR: Warning)
                raise_failure = getattr(self._optimizer, "raise_failure", None)
                if raise_failure is not None:
                    raise_failure()
                self._LOG.error(_ZERO_LOSS_MSG)
                raise ValueError(_ZERO_LOSS_MSG)
            if metrics[n] < metric:
                metric = metrics[n]
                best = self._compose(rots[n], trans[n], transforms[n])
        return best, metric

    # ------------------------------------------------------------------ reference API
    def multistart_registration(self, source: np.ndarray, target: np.ndarray) -> Tuple[np.ndarray, float]:
        """Best of ``attempts`` random starts (reference: Aligner.py:164-204)."""
        if not hasattr(self._optimizer, "optimize_batch"):
            return self._multistart_sequential(source, target)
        rots, trans, clouds = self._draw_starts(source)
        clouds = self._start_clouds(source, [(rots, trans, clouds)])
        res = self._optimizer.optimize_batch(clouds, np.asarray(target, dtype=np.float64)[None],
                                             np.arange(self._attempts), np.zeros(self._attempts, dtype=np.int32),
                                             **self._batch_kwargs(source))
        self.n_units_evaluated += self._attempts
        return self._best_of_starts(rots, trans, res.transforms, res.rmse, res.status)

    def _multistart_sequential(self, source, target):
        metric, best = np.inf, np.eye(4)
        for _ in range(self._attempts):
            rot, tr = self.initialize_rotation()
            cur_T, cur_m = self._optimizer.optimize(np.dot(np.copy(source), rot) + tr, target)
            self.n_optimize_calls += 1
            self.n_units_evaluated += 1
            if cur_m < metric:
                metric, best = cur_m, self._compose(rot, tr, cur_T)
        return best, metric

    def compass_step(self, source, target, scale_factors, delta):
        """One poll candidate (reference: Aligner.py:206-226)."""
        new_scale_factors = scale_factors + delta
        rotation, metric = self.multistart_registration(source, target * new_scale_factors)
        return new_scale_factors, rotation, metric

    def _scaled_targets(self, target: np.ndarray, scales: np.ndarray):
        """``target * scales[c]`` for every poll candidate (reference: Aligner.py:221-222).  One rounded float64
        multiply per coordinate either way; on the device (orpcd_apply_scale_batch) when the optimizer takes CUDA
        tensors, so a sweep uploads 144 bytes of scale factors instead of six scaled clouds."""
        if not self._device_starts():
            return np.stack([target * s for s in scales], axis=0)
        from .. import _lib, ops

        if self._d_target_key is not target:
            self._d_target, self._d_target_key = _lib.to_device(np.asarray(target, dtype=np.float64)), target
        return ops.apply_scale_batch(self._d_target, _lib.to_device(np.concatenate(scales, axis=0)))

    def _poll_sweep(self, source, target, scale, delta):
        """All six poll candidates of one sweep in a single batch (6 x attempts units).  Returns
        (steps, starts, states, res): the six scale steps in poll order, the starts drawn for each, the
        ``np.random`` state after each candidate's draws, and the optimizer's BatchResult; ``align`` replays the
        reference's sequential accept logic over them."""
        steps = []
        for axis in range(3):
            e = np.zeros(3)
            e[axis] = 1.0
            steps.append(delta * e)
            steps.append(-delta * e)
        starts, states = [], []
        for _ in steps:
            starts.append(self._draw_starts(source))
            states.append(np.random.get_state())
        A = self._attempts
        sources = self._start_clouds(source, starts)
        targets = self._scaled_targets(target, [scale + st for st in steps])
        unit_src = np.arange(6 * A, dtype=np.int32)
        unit_tgt = np.repeat(np.arange(6, dtype=np.int32), A)
        res = self._optimizer.optimize_batch(sources, targets, unit_src, unit_tgt, **self._batch_kwargs(source))
        self.n_units_evaluated += 6 * A
        return steps, starts, states, res

    def align(
        self,
        source: np.ndarray,
        target: np.ndarray,
        refine_registration: bool = True,
        icp_type: str = "PointToPlane",
    ) -> Tuple[np.ndarray, float, np.ndarray, List[float]]:
        """Two-block optimisation (reference: Aligner.py:228-317).  Returns
        (T 4x4 row-vector convention, inlier rmse, scale factors (1,3), errors)."""
        source = self._source_preprocessor.preprocess(source)
        target = self._target_preprocessor.preprocess(target)
        self.source_pre, self.target_pre = source, target
        iteration = 0
        optimal_scale_factors = np.ones((1, 3))
        start = time.time()
        optimal_transformation, optimal_metric = self.multistart_registration(source, target)
        self._LOG.info(f"Multi-start registration time: {time.time() - start}")
        self.decision_log.append(dict(scale=optimal_scale_factors.copy(), metric=optimal_metric, accepted=True,
                                      kind="init"))
        errors = [optimal_metric]
        batched = hasattr(self._optimizer, "optimize_batch")
        A = self._attempts
        while self._delta >= self._eps and iteration <= self._max_iter:
            iteration += 1
            self._LOG.info(f"Iteration number: {iteration} Current step: {self._delta}")
            if batched:
                steps, starts, states, res = self._poll_sweep(source, target, optimal_scale_factors, self._delta)
                new_metric = None
                for k, step in enumerate(steps):
                    sl = slice(k * A, (k + 1) * A)
                    new_rotation, new_metric = self._best_of_starts(
                        starts[k][0], starts[k][1], res.transforms[sl], res.rmse[sl], res.status[sl])
                    accepted = bool(new_metric <= optimal_metric)
                    self.decision_log.append(dict(scale=(optimal_scale_factors + step).copy(), metric=new_metric,
                                                  accepted=accepted, kind=("+" if k % 2 == 0 else "-") + str(k // 2)))
                    if accepted:
                        optimal_metric = new_metric
                        optimal_transformation = new_rotation
                        optimal_scale_factors += step
                        errors.append(new_metric)
                        np.random.set_state(states[k])   # later candidates were never drawn sequentially
                        break
            else:
                directions = np.eye(3)
                for axis in range(3):
                    for sign in (1.0, -1.0):
                        step = sign * self._delta * directions[:, axis]
                        _, new_rotation, new_metric = self.compass_step(source, target, optimal_scale_factors, step)
                        accepted = bool(new_metric <= optimal_metric)
                        self.decision_log.append(dict(scale=(optimal_scale_factors + step).copy(), metric=new_metric,
                                                      accepted=accepted, kind=("+" if sign > 0 else "-") + str(axis)))
                        if accepted:
                            optimal_metric = new_metric
                            optimal_transformation = new_rotation
                            optimal_scale_factors += step
                            errors.append(new_metric)
                            break
                    if accepted:
                        break
            if new_metric > optimal_metric:
                self._delta = self._delta / 2
                self._LOG.info(f"No improvement, decreasing step size to {self._delta}")

        if refine_registration:
            target = target * optimal_scale_factors
            optimal_transformation, optimal_metric = self.refine_registration(
                source=source, target=target, initial_transform=optimal_transformation, icp_type=icp_type)
            errors.append(optimal_metric)

        self.transfromation = optimal_transformation
        self.scale_factors = optimal_scale_factors
        return optimal_transformation, optimal_metric, optimal_scale_factors, errors

    def refine_registration(self, source, target, initial_transform, max_iteration: int = REFINER_MAX_ITER,
                            distance_threshold: float = REFINER_DISTANCE_THRESHOLD, icp_type: str = "PointToPoint",
                            target_normals: Optional[np.ndarray] = None):
        """ICP polish on the GPU (reference: Aligner.py:319-364, Open3D registration_icp).

        Mirrors the reference literally: ``initial_transform`` is handed over as Open3D's ``init``
        (column convention -- the reference passes its row-convention matrix here, SURVEY.md B.7, and
        so does ``align``), and the returned matrix is Open3D's ``result.transformation`` untouched.
        ``icp_type="PointToPlane"`` needs target normals; the reference never supplies them and Open3D
        raises, so without ``target_normals`` (an extension of this port: (M,3) array) the same
        RuntimeError is raised here."""
        from .. import _lib, ops

        if icp_type not in ("PointToPoint", "PointToPlane"):
            raise TypeError("icp_type must be 'PointToPoint' or 'PointToPlane'")
        if icp_type == "PointToPlane" and target_normals is None:
            raise RuntimeError("[Open3D Error] TransformationEstimationPointToPlane and "
                               "TransformationEstimationColoredICP require pre-computed normal vectors for target "
                               "PointCloud.")
        t = _lib.torch()
        d_src = _lib.to_device(np.asarray(source, dtype=np.float64)[None])
        d_tgt = _lib.to_device(np.asarray(target, dtype=np.float64)[None])
        zero = t.zeros(1, dtype=t.int32, device="cuda")
        init = _lib.to_device(np.asarray(initial_transform, dtype=np.float64).reshape(1, 4, 4))
        normals = None
        if icp_type == "PointToPlane":
            normals = _lib.to_device(np.asarray(target_normals, dtype=np.float64)[None])
        T, stats, _ = ops.icp_batch(d_src, d_tgt, zero, zero, init, normals, distance_threshold, max_iteration)
        return T[0].cpu().numpy(), float(stats[0, 1].item())

    def transfrom(self, source: np.ndarray) -> np.ndarray:
        """Apply the found registration to a raw cloud (reference: Aligner.py:367-394)."""
        for block in self._source_preprocessor.preprocessor_blocks:
            if isinstance(block, BaseScaler):
                source_mean, source_distance = block.mean, block.scale
        for block in self._target_preprocessor.preprocessor_blocks:
            if isinstance(block, BaseScaler):
                target_mean, target_distance = block.mean, block.scale
        source = (source - source_mean) / source_distance
        source = np.dot(source, self.transfromation[:3, :3]) + self.transfromation[:3, 3]
        source = source / self.scale_factors
        return source * target_distance + target_mean

    def __repr__(self):
        return f"""{self.__class__.__name__}
            (source_preprocessor={self._source_preprocessor},
            target_preprocessor={self._target_preprocessor},
            optimizer={self._optimizer},
            attempts={self._attempts},
            deg={self._deg},
            mu={self._mu},
            std={self._std},
            delta={self._delta},
            max_iter={self._max_iter},
            eps={self._eps})
        """
