"""Throughput of the batched FGR path on the bench workload (synthetic 5k-pt clouds, 6 scale candidates x 30
starts = 180 registrations per batch).  Usage: python tools/fgr_bench.py [reference_compat|fixed] [reps]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-scale-pointcloud-registration_b200"))
sys.path.insert(0, ROOT)


def main():
    import torch

    import bench
    from or_pcd_b200 import ops
    from or_pcd_b200.Optimizer import FastGlobalOptimizer

    mode = sys.argv[1] if len(sys.argv) > 1 else "fixed"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    src, sources, targets, unit_src, unit_tgt = bench.make_workload(1)
    opt = FastGlobalOptimizer(feature_mode=mode)
    for rep in range(reps):
        k0 = ops.launch_counter["kernels"]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = opt.optimize_batch(sources, targets, unit_src, unit_tgt, rigid_base=src)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(json.dumps(dict(mode=mode, rep=rep, units=len(unit_src), wall_s=dt, evals_per_s=len(unit_src) / dt,
                              kernels=ops.launch_counter["kernels"] - k0, exact_scan_rows=opt.last_exact_scans,
                              rows=int(2 * len(unit_src) * 5000 if mode == "fixed" else len(unit_src) * 5000),
                              rmse_min=float(res.rmse.min()), pairs_mean=float(opt.last_info[:, 1].mean()))))


if __name__ == "__main__":
    main()
