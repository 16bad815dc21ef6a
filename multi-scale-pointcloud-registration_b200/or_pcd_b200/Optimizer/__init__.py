from .iOptimizer import BatchResult, IOptimizer
from .generalizedICP import GeneralizedICP
from .fastGlobalOptimizer import FastGlobalOptimizer
from .shardedOptimizer import ShardedOptimizer
