"""Uniform random subsampling (reference: Preprocessor/Downsamplers/randomDownsampler.py:29-37).

The draw is part of the reference's single global legacy ``np.random`` stream
(SURVEY.md B.1), so the indices come from the same host call; only the gather
would be device work and it is a 120 KB copy."""
import numpy as np

from ...utils.constants import SAMPLE_SIZE
from ...utils.logger_factory import LoggerFactory
from ..iProcessBlock import IProcessBlock


class RandomDownsampler(IProcessBlock):
    def __init__(self, sample_size: int = SAMPLE_SIZE, replace: bool = False):
        self._LOG = LoggerFactory.get_logger(log_name=self.__class__.__name__, log_on_file=False)
        if sample_size <= 0:
            self._LOG.warning(f"sample size cannot be 0 or less. Provided: {sample_size}. Using {SAMPLE_SIZE}")
            sample_size = SAMPLE_SIZE
        self._sample_size = sample_size
        self._replace = replace

    def process(self, cloud: np.ndarray) -> np.ndarray:
        """``sample_size`` rows of ``cloud`` in the order of the permutation numpy drew (legacy global
        MT19937 stream: ``choice`` without replacement is ``permutation(N)[:k]``); the indices are kept in
        ``last_indices`` for the parity tests."""
        keep = np.random.choice(cloud.shape[0], self._sample_size, replace=self._replace)
        self.last_indices = keep
        return cloud[keep]

    def __repr__(self):
        return f"{self.__class__.__name__}(sample_size={self._sample_size}, replace={self._replace})"
