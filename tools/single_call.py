"""Latency of ONE GeneralizedICP.optimize() call (host arrays in, transform out) on a synthetic 5k-point pair:
the path the reference's sequential Aligner drives (Aligner.py:188-190).  usage: python tools/single_call.py [reps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-scale-pointcloud-registration_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import torch
    import oracle as O
    from or_pcd_b200.Optimizer import GeneralizedICP

    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    rng = np.random.default_rng(1234)
    src = O.RadiusScaler().process(O.synthetic_cloud(5000, rng))
    tgt = O.RadiusScaler().process(O.synthetic_cloud(5000, rng))
    np.random.seed(0)
    r, t = O.initialize_rotation(np.pi / 2, 0.0, 0.1)
    s = src @ r + t
    one = GeneralizedICP()
    for _ in range(3):
        one.optimize(s, tgt)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        T, rmse = one.optimize(s, tgt)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print("single optimize() ms: min %.3f median %.3f  rmse %.9g" % (min(ts), sorted(ts)[len(ts) // 2], rmse))


if __name__ == "__main__":
    main()
