from .iOptimizer import BatchResult, IOptimizer
from .generalizedICP import GeneralizedICP
from .shardedOptimizer import ShardedOptimizer
