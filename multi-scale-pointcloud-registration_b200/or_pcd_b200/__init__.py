"""or_pcd_b200 -- B200-native drop-in for the registration hot path of or_pcd.

Same plugin API as the reference package (``Preprocessor`` blocks, ``IOptimizer``
classes, ``Aligner.align``); the arithmetic runs in hand-written sm_100a CUDA
behind a C ABI (``include/orpcd_b200.h``).
"""
__version__ = "0.1.0"
