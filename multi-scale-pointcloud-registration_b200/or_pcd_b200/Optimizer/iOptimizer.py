"""Plugin ABC of the inner-block optimizer (reference: Optimizer/iOptimizer.py:7-21).

``optimize`` is the reference's contract.  B200 optimizers additionally offer
``optimize_batch`` so the Aligner can run every (scale candidate x random start)
registration of a poll round concurrently; optimizers without it still plug in
and are driven one call at a time."""
from abc import ABC, abstractmethod
from typing import Tuple

import numpy as np


class IOptimizer(ABC):
    @abstractmethod
    def optimize(self, source: np.ndarray, target: np.ndarray, **kwargs) -> Tuple[np.ndarray, float]:
        """(N,3) source, (M,3) target -> (4x4 roto-translation with the rotation block
        transposed for row-vector use, inlier RMSE)."""
