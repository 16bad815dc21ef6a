"""Quick timing of one batched GICP poll round (not the contract bench)."""
import json
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
    from or_pcd_b200 import _lib, ops
    from or_pcd_b200.Optimizer import GeneralizedICP

    n_units = int(sys.argv[1]) if len(sys.argv) > 1 else 180
    rng = np.random.default_rng(1234)
    src = O.RadiusScaler().process(O.synthetic_cloud(5000, rng))
    tgt = O.RadiusScaler().process(O.synthetic_cloud(5000, rng))
    np.random.seed(0)
    starts = []
    for _ in range(n_units):
        r, t = O.initialize_rotation(np.pi / 2, 0.0, 0.1)
        starts.append(src @ r + t)
    sources = np.stack(starts)
    scales = np.array([[1.2, 1, 1], [0.8, 1, 1], [1, 1.2, 1], [1, 0.8, 1], [1, 1, 1.2], [1, 1, 0.8]])
    targets = np.stack([tgt * s for s in scales])
    opt = GeneralizedICP()
    opt.profile = True
    opt.sweep_mode = int(os.environ.get("SWEEP_MODE", "0"))
    opt.sweep_q = int(os.environ.get("SWEEP_Q", "0"))
    opt.target_splits = int(os.environ.get("SPLITS", "0"))
    opt.sweep_tj = int(os.environ.get("SWEEP_TJ", "0"))
    opt.sweep_smem_pad = int(os.environ.get("PAD", "0"))
    d_s, d_t = _lib.to_device(sources), _lib.to_device(targets)
    us = torch.arange(n_units, dtype=torch.int32, device="cuda")
    ut = (torch.arange(n_units, dtype=torch.int32, device="cuda") * 6 // n_units).to(torch.int32)
    for rep in range(int(os.environ.get("REPS", "3"))):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        opt._base_key = None
        sc, tc = opt.covariances(d_s, d_t, src if os.environ.get('RIGID', '1') == '1' else None)
        res = opt.optimize_batch_device(d_s, d_t, us, ut, src_cov=sc, tgt_cov=tc)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        p = res.profile
        st = res.stats.cpu().numpy()
        print(json.dumps(dict(rep=rep, units=n_units, wall_s=dt, evals_per_s=n_units / dt, launches=p[0],
                              sweep_ms=p[1], pairs_visited=p[2], pairs_brute=p[7] * 5000.0 * 5000.0,
                              visited_fraction=p[2] / max(p[7] * 25e6, 1), sweep_tflops_executed=8 * p[2] / max(p[1], 1e-9) / 1e9,
                              fixed=p[4], splits=p[5], unit_rounds=p[7], mean_iters=float(st[:, 2].mean()),
                              rmse_min=float(st[:, 1].min()))))


if __name__ == "__main__":
    main()
