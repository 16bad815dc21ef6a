from .iProcessBlock import IProcessBlock
from .preprocessor import Preprocessor
from .Downsamplers import FarthestDownsampler, RandomDownsampler, VoxelDownsampler
from .Outliers import SOR
from .Scalers import BaseScaler, RadiusScaler
