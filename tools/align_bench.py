"""Aligner.align() wall seconds on the Armadillo pairs (BASELINE configs 1 and 3), B200 path.
Usage: python tools/align_bench.py [back|onhead] [attempts] [gicp|fgr|fgr_fixed]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-scale-pointcloud-registration_b200"))


def main():
    import torch

    from or_pcd_b200 import ops
    from or_pcd_b200.Aligner.Aligner import Aligner
    from or_pcd_b200.data import load_sample_cloud
    from or_pcd_b200.Optimizer import FastGlobalOptimizer, GeneralizedICP
    from or_pcd_b200.Preprocessor import Preprocessor, RandomDownsampler, SOR

    pair = sys.argv[1] if len(sys.argv) > 1 else "back"
    attempts = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    kind = sys.argv[3] if len(sys.argv) > 3 else "gicp"   # gicp | fgr | fgr_fixed
    names = {"back": ("ArmadilloBack_330", "ArmadilloBack_0"),
             "onhead": ("ArmadilloOnHeadMultiple_150", "ArmadilloOnHeadMultiple_240")}[pair]
    src, tgt = load_sample_cloud(names[0]), load_sample_cloud(names[1])
    for rep in range(int(os.environ.get("REPS", "2"))):
        np.random.seed(0)
        al = Aligner(Preprocessor([RandomDownsampler(5000), SOR()]), Preprocessor([RandomDownsampler(5000), SOR()]),
                     GeneralizedICP() if kind == "gicp" else FastGlobalOptimizer(
                         feature_mode="fixed" if kind == "fgr_fixed" else "reference_compat"), attempts=attempts)
        k0 = ops.launch_counter["kernels"]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        T, m, scale, errors = al.align(src, tgt, refine_registration=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(json.dumps(dict(pair=pair, optimizer=kind, attempts=attempts, rep=rep, align_wall_s=dt, rmse=m, scale=scale.ravel().tolist(),
                              sequential_equivalent_calls=al.n_optimize_calls, units_evaluated=al.n_units_evaluated,
                              evals_per_s_sequential_equivalent=al.n_optimize_calls / dt,
                              evals_per_s_executed=al.n_units_evaluated / dt, errors=errors,
                              kernels=ops.launch_counter["kernels"] - k0)))


if __name__ == "__main__":
    main()
