"""Plugin ABC of a preprocessing block (reference: Preprocessor/iProcessBlock.py:6-12).

Blocks are applied in list order by ``Preprocessor.preprocess``; each takes the (N,3) float64 cloud the
previous one returned.  The B200 blocks (``SOR``, the down-samplers) keep this host-array contract and
move the cloud to the GPU inside ``process``; they raise ``OrpcdError`` when the CUDA library or a CUDA
device is missing -- there is no CPU fallback."""
from abc import ABC, abstractmethod

import numpy as np


class IProcessBlock(ABC):
    @abstractmethod
    def process(self, cloud: np.ndarray) -> np.ndarray:
        """(N,3) cloud in, (N',3) cloud out."""
