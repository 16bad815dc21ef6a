"""Inner-registration plugins behind the C ABI (``csrc/liborpcd_b200.so``).

``IOptimizer``            the reference's plugin ABC (``optimize(source, target) -> (T, rmse)``)
``BatchResult``           what the batched extension ``optimize_batch`` returns, one entry per unit
``GeneralizedICP``        batched Generalized-ICP, one fused CUDA kernel per Gauss-Newton round
``FastGlobalOptimizer``   FPFH + tcgen05 feature matching + tuple test + graduated non-convexity
``ShardedOptimizer``      shards the units of a batch over the GPUs of one box (NCCL all-gather of results)
"""
from .iOptimizer import BatchResult, IOptimizer
from .generalizedICP import GeneralizedICP
from .fastGlobalOptimizer import FastGlobalOptimizer
from .shardedOptimizer import ShardedOptimizer
