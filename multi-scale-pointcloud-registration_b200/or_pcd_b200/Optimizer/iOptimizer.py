"""Plugin ABC of the inner-block optimizer (reference: Optimizer/iOptimizer.py:7-21).

``optimize`` is the reference's contract.  B200 optimizers additionally offer
``optimize_batch`` so the Aligner can run every (scale candidate x random start)
registration of a poll round concurrently; optimizers without it still plug in
and are driven one call at a time."""
from abc import ABC, abstractmethod
from typing import NamedTuple, Tuple

import numpy as np


class BatchResult(NamedTuple):
    """What ``optimize_batch`` returns: one entry per unit (inner registration)."""

    transforms: np.ndarray   # (U,4,4) rotation block transposed (row-vector convention), as optimize()
    rmse: np.ndarray         # (U,)
    fitness: np.ndarray      # (U,)
    iterations: np.ndarray   # (U,) int
    status: np.ndarray       # (U,) 0 ok, 1 = optimize() would raise (GICP: inlier_rmse == 0; FGR: no correspondence)
    profile: list


class IOptimizer(ABC):
    @abstractmethod
    def optimize(self, source: np.ndarray, target: np.ndarray, **kwargs) -> Tuple[np.ndarray, float]:
        """(N,3) source, (M,3) target -> (4x4 roto-translation with the rotation block
        transposed for row-vector use, inlier RMSE)."""
