"""Sequential block runner (reference: Preprocessor/preprocessor.py:12-31).

One ``Preprocessor`` instance per cloud: the scaler inside it remembers that cloud's mean and scale.  The
Aligner runs ``source_preprocessor.preprocess(source)`` first and ``target_preprocessor.preprocess(target)``
second, which also fixes the order of the ``np.random`` draws of the down-samplers."""
from typing import List

import numpy as np

from ..utils.logger_factory import LoggerFactory
from .iProcessBlock import IProcessBlock
from .Scalers import BaseScaler, RadiusScaler


class Preprocessor:
    def __init__(self, preprocessor_blocks: List[IProcessBlock]):
        self._LOG = LoggerFactory.get_logger(self.__class__.__name__, log_on_file=False)
        # Like the reference (preprocessor.py:19-23) a RadiusScaler is put in front of the
        # CALLER'S list when it holds no scaler -- the list is mutated on purpose.
        if not any(isinstance(b, BaseScaler) for b in preprocessor_blocks):
            self._LOG.warning("no scaler among the preprocessor blocks: prepending RadiusScaler")
            preprocessor_blocks.insert(0, RadiusScaler())
        self.preprocessor_blocks = preprocessor_blocks

    def preprocess(self, cloud: np.ndarray) -> np.ndarray:
        """Run every block in list order on the (N,3) cloud; returns the last block's output."""
        for block in self.preprocessor_blocks:
            cloud = block.process(cloud)
        return cloud
