"""Time the 1-NN sweep kernel in isolation for several tuning variants (Q, TJ, splits)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-scale-pointcloud-registration_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import torch
    import oracle as O
    from or_pcd_b200 import _lib, ops

    n_units = int(sys.argv[1]) if len(sys.argv) > 1 else 180
    n_pts = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
    rng = np.random.default_rng(1234)
    src = O.RadiusScaler().process(O.synthetic_cloud(n_pts, rng))
    tgt = O.RadiusScaler().process(O.synthetic_cloud(n_pts, rng))
    np.random.seed(0)
    sources = np.stack([src @ O.initialize_rotation(np.pi / 2, 0.0, 0.1)[0] for _ in range(n_units)])
    d_s, d_t = _lib.to_device(sources), _lib.to_device(tgt[None])
    us = torch.arange(n_units, dtype=torch.int32, device="cuda")
    ut = torch.zeros(n_units, dtype=torch.int32, device="cuda")
    ax_s = ops.normals_cov(d_s, want_gicp_cov=False, want_gicp_axis=True)["gicp_axis"]
    ax_t = ops.normals_cov(d_t, want_gicp_cov=False, want_gicp_axis=True)["gicp_axis"]
    variants = [(q, tj, 8, 0, m) for m in (1, 2, 3) for q in (4, 8) for tj in (2, 4, 8)]
    pairs = float(n_units) * n_pts * n_pts
    for q, tj, sp, pad, mode in variants:
        res = ops.gicp_batch(d_s, ax_s, d_t, ax_t, us, ut, max_iterations=0, sweep_q=q, sweep_tj=tj, target_splits=sp,
                             sweep_smem_pad=pad, sweep_mode=mode)
        for _ in range(3):
            ops.gicp_sweep_only(res.params, res.workspace)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            ops.gicp_sweep_only(res.params, res.workspace)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(json.dumps(dict(q=q, tj=tj, splits=sp, mode=mode, ms=ms, tpairs_per_s=pairs / ms / 1e9,
                              alg_tflops=8 * pairs / ms / 1e9)))


if __name__ == "__main__":
    main()
