from .iOptimizer import IOptimizer
from .generalizedICP import GeneralizedICP
