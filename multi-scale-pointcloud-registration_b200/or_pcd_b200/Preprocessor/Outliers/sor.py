"""Statistical outlier removal on the GPU (reference: Preprocessor/Outliers/sor.py:51-79,
Open3D PointCloud.remove_statistical_outlier)."""
import numpy as np

from ... import _lib, ops
from ...utils.constants import NB_NEIGHBOURS, STD_RATIO
from ...utils.logger_factory import LoggerFactory
from ..iProcessBlock import IProcessBlock


class SOR(IProcessBlock):
    def __init__(self, nb_neighbours: int = NB_NEIGHBOURS, std_ratio: float = STD_RATIO) -> None:
        super().__init__()
        self._LOG = LoggerFactory.get_logger(log_name=self.__class__.__name__, log_on_file=False)
        self._nb_neighbours = nb_neighbours
        self._std_ratio = std_ratio

    def process(self, cloud: np.ndarray) -> np.ndarray:
        cloud = np.ascontiguousarray(cloud, dtype=np.float64)
        if self._nb_neighbours < 1 or self._std_ratio <= 0:
            # Open3D logs an error and returns an empty selection for illegal arguments
            raise ValueError("SOR: nb_neighbours must be >= 1 and std_ratio > 0")
        if self._nb_neighbours > _lib.MAX_KNN:
            # the reference accepts any value (sor.py:24-37); the CUDA k-NN keeps its k-best sets in shared memory,
            # which bounds k at ORPCD_MAX_KNN = 512 (8x the reference's default of 64)
            raise ValueError(f"SOR: nb_neighbours > {_lib.MAX_KNN} exceeds the shared-memory k-NN lists of the CUDA kernel")
        d_cloud = _lib.to_device(cloud)
        mask, avg, stats = ops.sor_mask(d_cloud, self._nb_neighbours, float(self._std_ratio))
        self.last_mask = mask.cpu().numpy().astype(bool)
        inliers = cloud[self.last_mask]
        self._LOG.debug(f"SOR: {cloud.shape[0]} points in, {inliers.shape[0]} out")
        return inliers

    @property
    def nb_neighbours(self) -> int:
        return self._nb_neighbours

    @property
    def std_ratio(self) -> float:
        return self._std_ratio

    def __repr__(self):
        return f"{self.__class__.__name__}(nb_neighbours={self._nb_neighbours}, std_ratio={self._std_ratio})"
