"""Plugin ABC of a preprocessing block (reference: Preprocessor/iProcessBlock.py:6-12)."""
from abc import ABC, abstractmethod

import numpy as np


class IProcessBlock(ABC):
    @abstractmethod
    def process(self, cloud: np.ndarray) -> np.ndarray:
        """(N,3) cloud in, (N',3) cloud out."""
