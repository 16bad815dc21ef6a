"""Farthest point sampling on the GPU (reference: Preprocessor/Downsamplers/farthestDownsampler.py:10-57).

The first index comes from the same ``np.random.randint`` call as in the reference (it is part of the
global numpy stream); the k-1 sequential argmax steps run in one persistent CUDA block and reproduce
the reference's scipy ``cdist`` / ``np.argmax`` arithmetic bit for bit."""
import numpy as np

from ... import _lib, ops
from ...utils.constants import SAMPLE_SIZE
from ...utils.logger_factory import LoggerFactory
from ..iProcessBlock import IProcessBlock


class FarthestDownsampler(IProcessBlock):
    def __init__(self, sample_size: int = SAMPLE_SIZE):
        self._LOG = LoggerFactory.get_logger(log_name=self.__class__.__name__, log_on_file=False)
        if sample_size <= 0:
            self._LOG.warning(f"sample size cannot be 0 or less. Provided: {sample_size}. Using default value: "
                              f"{SAMPLE_SIZE}")
            sample_size = SAMPLE_SIZE
        self._sample_size = sample_size

    def process(self, cloud: np.ndarray):
        cloud = np.asarray(cloud)
        first = np.random.randint(low=0, high=cloud.shape[0])           # farthestDownsampler.py:35
        idx = ops.farthest_point_sample(_lib.to_device(np.ascontiguousarray(cloud, dtype=np.float64)), first,
                                        self._sample_size)
        self.last_indices = idx.cpu().numpy().astype(np.int64)
        return cloud[self.last_indices]

    def __repr__(self):
        return f"{self.__class__.__name__}(sample_size={self._sample_size})"
