"""Voxel-grid down-sampling to a target point count (reference:
Preprocessor/Downsamplers/voxelDownsampler.py:14-136).  The compass search over the voxel size is the
reference's, step for step (it mutates ``self._delta`` like the reference: an instance is single-use);
each probe is one voxel-grid averaging on the GPU (``orpcd_voxel_downsample``).  Open3D emits voxels in
hash-map order; here they come in lexicographic voxel-index order."""
import numpy as np

from ... import _lib, ops
from ...utils.constants import BASE_VOXEL_SIZE, MIN_VOXEL_SIZE, VOXEL_DELTA, VOXEL_EPS
from ...utils.logger_factory import LoggerFactory
from ..iProcessBlock import IProcessBlock


class VoxelDownsampler(IProcessBlock):
    def __init__(self, target_points: int, base_voxel_size: float = BASE_VOXEL_SIZE,
                 min_voxel_size: float = MIN_VOXEL_SIZE, delta: float = VOXEL_DELTA, eps: float = VOXEL_EPS):
        self._LOG = LoggerFactory.get_logger(log_name=self.__class__.__name__, log_on_file=False)
        if target_points <= 0:
            msg = f"target points cannot be 0 or less. Provided: {target_points}"
            self._LOG.error(msg)
            raise ValueError(msg)
        self._target_points = target_points

        def positive(v, default, name):
            if v <= 0:
                self._LOG.warning(f"{name} cannot be 0 or less. Provided: {v}. Using default value: {default}")
                return default
            return v

        self._min_voxel_size = positive(min_voxel_size, MIN_VOXEL_SIZE, "min voxel size")
        self._base_voxel_size = positive(base_voxel_size, BASE_VOXEL_SIZE, "base voxel size")
        self._delta = positive(delta, VOXEL_DELTA, "delta")
        self._eps = positive(eps, VOXEL_EPS, "eps")
        self._ws = None
        self.probes = []            # (voxel size, point count) of every probe, in order

    def _down_sample(self, d_cloud, voxel_size):
        pts, _, self._ws = ops.voxel_downsample(d_cloud, voxel_size, ws=self._ws)
        self.probes.append((voxel_size, pts.shape[0]))
        return pts

    def compass_step(self, delta: float, cloud, current_voxel_size: float):
        new_voxel_size = current_voxel_size + delta
        if new_voxel_size <= self._min_voxel_size:
            new_voxel_size = self._min_voxel_size
        cloud_ds = self._down_sample(cloud, new_voxel_size)
        metric = np.abs(self._target_points - cloud_ds.shape[0])
        return cloud_ds, new_voxel_size, metric

    def process(self, cloud: np.ndarray):
        d_cloud = _lib.to_device(np.ascontiguousarray(cloud, dtype=np.float64))
        current_voxel_size = self._base_voxel_size
        best_cloud = self._down_sample(d_cloud, current_voxel_size)
        metric = np.abs(self._target_points - best_cloud.shape[0])
        while self._delta >= self._eps:
            obtained_cloud, new_voxel_size, obtained_metric = self.compass_step(self._delta, d_cloud, current_voxel_size)
            if obtained_metric < metric:
                current_voxel_size, metric, best_cloud = new_voxel_size, obtained_metric, obtained_cloud
                continue
            obtained_cloud, new_voxel_size, obtained_metric = self.compass_step(-self._delta, d_cloud, current_voxel_size)
            if obtained_metric < metric:
                current_voxel_size, metric, best_cloud = new_voxel_size, obtained_metric, obtained_cloud
                continue
            self._delta = self._delta / 2
        self.voxel_size = current_voxel_size
        self._LOG.debug(f"Voxel size: {current_voxel_size}, Metric: {metric}")
        return best_cloud.cpu().numpy()

    def __repr__(self):
        return (f"{self.__class__.__name__}(target_points={self._target_points}, "
                f"base_voxel_size={self._base_voxel_size}, min_voxel_size={self._min_voxel_size}, "
                f"delta={self._delta}, eps={self._eps})")
