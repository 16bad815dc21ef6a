from .iProcessBlock import IProcessBlock
from .preprocessor import Preprocessor
from .Downsamplers import RandomDownsampler
from .Outliers import SOR
from .Scalers import BaseScaler, RadiusScaler
