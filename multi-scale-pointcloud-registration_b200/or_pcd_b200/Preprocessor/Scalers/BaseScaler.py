"""Base class of the scaling blocks (reference: Preprocessor/Scalers/BaseScaler.py:6-16).
``Aligner.transfrom`` finds the scaler of each pipeline by this type and reads
``.mean`` (1,3) and ``.scale``."""
import numpy as np

from ...utils.logger_factory import LoggerFactory
from ..iProcessBlock import IProcessBlock


class BaseScaler(IProcessBlock):
    mean: np.ndarray = np.zeros((1, 3))
    scale: float = 1.0

    def __init__(self):
        self._LOG = LoggerFactory.get_logger(log_name=self.__class__.__name__, log_on_file=False)
