"""BASELINE configs 4 and 5 at their full sizes: 50k points x 256 scale candidates, and 200k points x 1024
(start x candidate) units, GICP and FGR.  The oracle cannot finish these sizes in seconds, so the results are checked through
size-independent properties (tools/scale_check.py): a sample of the final correspondences against a plain
float64 exhaustive search (index-exact), fitness / rmse re-evaluated at the returned transform (1e-9 relative),
and the culled self k-NN against the brute-force kernel (bit-exact)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "scale_check.py"), *map(str, args)],
                         capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("n,scales,starts", [(50000, 256, 1), (200000, 32, 32)])
def test_gicp_configs_4_and_5_full_size(n, scales, starts):
    """Config 4: 256 scale candidates x 1 start at 50k points; config 5: 32 scales x 32 starts = 1024 units at 200k."""
    r = _run(n, scales, starts, "gicp", 1)
    assert r["units"] == scales * starts
    assert r["ok"] and r["failed_units"] == 0, r
    assert all(c["sample_mismatches"] == 0 and c["rel"] < 1e-9 for c in r["checks"])
    assert r["visited_fraction"] < 0.2          # the sweep culls: far fewer pairs than brute force
    print(json.dumps({k: r[k] for k in ("units", "n_points", "wall_s", "registrations_per_s", "mem_gb")}))


@pytest.mark.parametrize("n,scales,starts", [(50000, 4, 1), (200000, 32, 32)])
def test_fgr_full_size_clouds(n, scales, starts):
    """FastGlobalOptimizer(feature_mode="fixed") -- FPFH of every set, tcgen05 matching both ways for every unit, tuple
    test, GNC, evaluation -- at 50k points and at config 5's full 200k x 1024 units."""
    r = _run(n, scales, starts, "fgr", 1)
    assert r["units"] == scales * starts
    assert r["ok"] and r["failed_units"] == 0, r
    print(json.dumps({k: r[k] for k in ("units", "n_points", "wall_s", "registrations_per_s", "mem_gb")}))


def test_culled_knn_equals_brute_force_at_200k():
    import torch

    from or_pcd_b200 import _lib, ops

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import scale_check

    src, _, _ = scale_check.make_pair(200000, np.random.default_rng(5))
    d = _lib.to_device(src)
    idx, d2 = ops.knn_exact(d, d, 20)                     # self query: Hilbert-ordered culled kernel
    q = d[:4096].clone()                                  # a different tensor: brute-force kernel
    bidx, bd2 = ops.knn_exact(d, q, 20)
    assert torch.equal(idx[:4096], bidx) and torch.equal(d2[:4096], bd2)
