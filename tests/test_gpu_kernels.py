"""Parity of the CUDA kernels (through the C ABI) against the CPU oracle on the
same seeded inputs.  Index / mask work: bit-exact.  Floating point: tolerances
from BASELINE.json north_star, written beside each assert."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _dev(a):
    from or_pcd_b200 import _lib

    return _lib.to_device(a)


@pytest.mark.parametrize("k", [1, 20, 64])
def test_knn_bit_exact_armadillo(oracle, armadillo_pair, k):
    from or_pcd_b200 import ops

    s, t, _ = armadillo_pair
    idx, d2 = ops.knn_exact(_dev(t), _dev(s), k)
    oi, od = oracle.knn(t, s, k)
    assert np.array_equal(idx.cpu().numpy(), oi)
    assert np.array_equal(d2.cpu().numpy(), od)          # float64 distances bit-identical


def test_knn_ties_duplicates_and_small_clouds(oracle):
    from or_pcd_b200 import ops

    rng = np.random.default_rng(3)
    lattice = np.stack(np.meshgrid(np.arange(7.0), np.arange(6.0), np.arange(5.0), indexing="ij"), -1).reshape(-1, 3)
    dup = np.concatenate([lattice[:50], lattice[:50], rng.standard_normal((13, 3))])
    for pts in (lattice, dup, rng.standard_normal((5, 3))):
        for k in (1, 8, 20):
            idx, d2 = ops.knn_exact(_dev(pts), _dev(pts), k)
            oi, od = oracle.knn(pts, pts, k, method="brute")
            assert np.array_equal(idx.cpu().numpy(), oi), (len(pts), k)
            assert np.array_equal(d2.cpu().numpy(), od)


@pytest.mark.parametrize("k", [20, 48, 64])
def test_knn_self_query_culled_path(oracle, armadillo_pair, k):
    """Self query on clouds >= 2048 points runs the Hilbert-ordered, chunk-culled kernel: same lists as the
    oracle, on a real scan, on a lattice (every distance tied many times) and with duplicated points; batched
    launches of several clouds too."""
    from or_pcd_b200 import ops

    s, t, _ = armadillo_pair
    g = np.stack(np.meshgrid(np.arange(16.0), np.arange(16.0), np.arange(9.0), indexing="ij"), -1).reshape(-1, 3)
    rng = np.random.default_rng(11)
    dup = np.concatenate([s[:1500], s[:1500], s[200:300], rng.standard_normal((37, 3)) * 0.3])
    for pts in (s, g[rng.permutation(len(g))], dup):
        d = _dev(pts)
        idx, d2 = ops.knn_exact(d, d, k)
        oi, od = oracle.knn(pts, pts, k, method="brute")
        assert np.array_equal(idx.cpu().numpy(), oi), (len(pts), k)
        assert np.array_equal(d2.cpu().numpy(), od)
    if k == 20:
        batch = np.stack([t[:4900] * sc for sc in ([1, 1, 1], [1.2, 1, 1], [1, 0.8, 1])])
        d = _dev(batch)
        idx, d2 = ops.knn_exact(d, d, k)
        for b in range(3):
            oi, od = oracle.knn(batch[b], batch[b], k)
            assert np.array_equal(idx[b].cpu().numpy(), oi) and np.array_equal(d2[b].cpu().numpy(), od)


def test_knn_batched(oracle):
    from or_pcd_b200 import ops

    rng = np.random.default_rng(5)
    pts = rng.standard_normal((3, 700, 3))
    qs = rng.standard_normal((3, 333, 3))
    idx, d2 = ops.knn_exact(_dev(pts), _dev(qs), 7)
    for b in range(3):
        oi, od = oracle.knn(pts[b], qs[b], 7)
        assert np.array_equal(idx[b].cpu().numpy(), oi)
        assert np.array_equal(d2[b].cpu().numpy(), od)


@pytest.mark.parametrize("scan", ["ArmadilloBack_330", "ArmadilloBack_0", "ArmadilloOnHeadMultiple_150"])
def test_sor_mask_bit_exact(oracle, scan):
    from or_pcd_b200 import ops

    np.random.seed(1)
    cloud = oracle.RadiusScaler().process(oracle.load_scan(scan))
    cloud = cloud[np.random.choice(cloud.shape[0], 5000, replace=False)]
    mask, avg, stats = ops.sor_mask(_dev(cloud), 64, 2.0)
    om, oavg, ostats = oracle.sor(cloud, 64, 2.0)
    assert np.array_equal(mask.cpu().numpy().astype(bool), om)
    assert np.array_equal(avg.cpu().numpy(), oavg)        # sequential float64 sums: bit-identical
    st = stats.cpu().numpy()
    assert st[0] == ostats[0] and st[1] == ostats[1] and st[2] == ostats[2]
    assert int(st[3]) == int(om.sum())


@pytest.mark.parametrize("k", [65, 100, 200, 300, 512])
def test_sor_and_knn_beyond_64_neighbours(oracle, k):
    """The reference accepts any nb_neighbours (Preprocessor/Outliers/sor.py:24-37); the CUDA lists go to 512."""
    from or_pcd_b200 import ops
    from or_pcd_b200.Preprocessor import SOR

    rng = np.random.default_rng(k)
    cloud = oracle.RadiusScaler().process(oracle.synthetic_cloud(2600 if k < 300 else 1500, rng))
    cloud[100:110] = cloud[200:210]                       # duplicates: exact ties
    mask, avg, stats = ops.sor_mask(_dev(cloud), k, 2.0)
    om, oavg, ostats = oracle.sor(cloud, k, 2.0)
    assert np.array_equal(mask.cpu().numpy().astype(bool), om) and np.array_equal(avg.cpu().numpy(), oavg)
    assert np.array_equal(SOR(nb_neighbours=k).process(cloud), cloud[om])
    q = cloud[:333] + 1e-3
    idx, d2 = ops.knn_exact(_dev(cloud), _dev(q), k)      # foreign queries: the brute-force kernel
    oi, od = oracle.knn(cloud, q, k)
    assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(d2.cpu().numpy(), od)
    with pytest.raises(ValueError):
        SOR(nb_neighbours=513).process(cloud)


def test_normals_and_gicp_covariances(oracle, armadillo_pair):
    from or_pcd_b200 import ops

    s, t, _ = armadillo_pair
    for cloud, radius in ((s, 0.0), (t, 0.0), (s, 0.1)):
        out = ops.normals_cov(_dev(cloud), knn=20, radius=radius, gicp_eps=1e-3, want_normals=True, want_cov=True,
                              want_gicp_cov=True, want_nn=True, want_gicp_axis=True)
        ref = oracle.normals(cloud, 20, radius, 1e-3)
        assert np.array_equal(out["nn_idx"].cpu().numpy(), ref["nn_idx"])       # neighbour lists bit-exact
        cov = out["cov"].cpu().numpy()
        rc = ref["cov"]
        ref6 = np.stack([rc[:, 0, 0], rc[:, 0, 1], rc[:, 0, 2], rc[:, 1, 1], rc[:, 1, 2], rc[:, 2, 2]], 1)
        assert np.array_equal(cov, ref6)                                        # no-FMA float64: bit-exact
        # FastEigen3x3's acos / cos are the deterministic det_acos / det_cos on both sides: bit-exact
        assert np.array_equal(out["normal"].cpu().numpy(), ref["normal"])
        g = ref["gicp_cov"]
        g6 = np.stack([g[:, 0, 0], g[:, 0, 1], g[:, 0, 2], g[:, 1, 1], g[:, 1, 2], g[:, 2, 2]], 1)
        assert np.abs(out["gicp_cov"].cpu().numpy() - g6).max() < 1e-12
        # plane axis a: C = I - (1 - eps) a a^T reproduces the covariance (closed-form whitening input)
        a = out["gicp_axis"].cpu().numpy()
        Ca = np.eye(3)[None] - (1 - 1e-3) * a[:, :, None] * a[:, None, :]
        assert np.abs(Ca - g).max() < 1e-12


def test_apply_scale_batch_bit_exact():
    from or_pcd_b200 import ops

    rng = np.random.default_rng(0)
    pts = rng.standard_normal((1000, 3))
    scales = np.array([[1.0, 1.2, 1.0], [0.8, 1.0, 1.05], [1.0, 1.0, 1.0]])
    out = ops.apply_scale_batch(_dev(pts), _dev(scales)).cpu().numpy()
    for i, s in enumerate(scales):
        assert np.array_equal(out[i], pts * s[None, :])


def _rot(a, b, g):
    rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    rz = np.array([[np.cos(g), -np.sin(g), 0], [np.sin(g), np.cos(g), 0], [0, 0, 1]])
    return rx @ ry @ rz


def _rot_angle(Ra, Rb):
    c = (np.trace(Ra.T @ Rb) - 1) / 2
    return np.arccos(np.clip(c, -1, 1))


@pytest.mark.parametrize("analytic", [True, False])
def test_gicp_batch_matches_oracle_traces(oracle, armadillo_pair, analytic):
    """Several starts x two scale candidates in one batch vs the oracle one by one.
    Tolerances (north_star): rotation 1e-4 rad, translation 1e-5 x bbox diagonal,
    fitness/rmse 1e-5 relative; iteration counts must agree."""
    from or_pcd_b200 import _lib, ops
    from or_pcd_b200.Optimizer import GeneralizedICP

    s, t, _ = armadillo_pair
    rng = np.random.default_rng(11)
    starts = []
    for _ in range(5):
        a, b, g = rng.uniform(-np.pi / 2, np.pi / 2, 3)
        starts.append(s @ _rot(a, b, g) + rng.standard_normal(3) * 0.1)
    starts.append(s.copy())
    sources = np.stack(starts)
    targets = np.stack([t, t * np.array([[1.0, 1.2, 1.0]])])
    unit_src = np.array([0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 5], dtype=np.int32)
    unit_tgt = np.array([0] * 6 + [1] * 6, dtype=np.int32)
    opt = GeneralizedICP()
    opt.analytic_whitening = analytic
    t_ = _lib.torch()
    res = opt.optimize_batch_device(_lib.to_device(sources), _lib.to_device(targets),
                                    t_.as_tensor(unit_src).cuda(), t_.as_tensor(unit_tgt).cuda(), want_trace=True)
    T = res.T.cpu().numpy()
    stats = res.stats.cpu().numpy()
    trace = res.trace.cpu().numpy()
    diag = np.linalg.norm(t.max(0) - t.min(0))
    for u in range(len(unit_src)):
        ref = oracle.gicp(sources[unit_src[u]], targets[unit_tgt[u]], 0.5, 100, want_trace=True)
        assert int(stats[u, 2]) == ref["iters"], (u, stats[u], ref["iters"])
        assert int(stats[u, 3]) == ref["n_corr"]
        assert abs(stats[u, 0] - ref["fitness"]) <= 1e-5 * ref["fitness"]
        assert abs(stats[u, 1] - ref["rmse"]) <= 1e-5 * ref["rmse"]
        assert _rot_angle(T[u, :3, :3], ref["T"][:3, :3]) < 1e-4
        assert np.linalg.norm(T[u, :3, 3] - ref["T"][:3, 3]) < 1e-5 * diag
        n = ref["iters"] + 1
        assert np.allclose(trace[u, :n], ref["trace"], rtol=1e-7, atol=1e-12), u
        assert np.isnan(trace[u, n:]).all()


def test_gicp_optimize_plugin_and_error(oracle, armadillo_pair):
    from or_pcd_b200.Optimizer import GeneralizedICP

    s, t, _ = armadillo_pair
    T, rmse = GeneralizedICP().optimize(s, t)
    ref = oracle.GeneralizedICP().optimize(s, t)
    assert abs(rmse - ref[1]) <= 1e-5 * ref[1]
    assert _rot_angle(T[:3, :3], ref[0][:3, :3]) < 1e-4
    # no correspondence within the distance threshold -> inlier_rmse == 0 -> ValueError (generalizedICP.py:77-81)
    with pytest.raises(ValueError):
        GeneralizedICP(max_correspondence_distance=1e-3).optimize(s[:200] + 50.0, t[:300])
    with pytest.raises(ValueError):
        oracle.GeneralizedICP(max_correspondence_distance=1e-3).optimize(s[:200] + 50.0, t[:300])


def test_gicp_small_and_ragged_sizes(oracle):
    """Sizes that are not multiples of any tile (n_tgt < one chunk, n_src < one block)."""
    from or_pcd_b200.Optimizer import GeneralizedICP

    rng = np.random.default_rng(2)
    for ns, nt in ((37, 29), (130, 1025), (513, 64)):
        tgt = oracle.synthetic_cloud(nt, rng)
        src = oracle.synthetic_cloud(ns, rng) @ _rot(0.1, -0.05, 0.08) + 0.02
        T, rmse = GeneralizedICP().optimize(src, tgt)
        ref_T, ref_rmse = oracle.GeneralizedICP().optimize(src, tgt)
        assert abs(rmse - ref_rmse) <= 1e-5 * ref_rmse, (ns, nt)
        assert _rot_angle(T[:3, :3], ref_T[:3, :3]) < 1e-4


def test_knn_rerank_equals_exact_knn(oracle, armadillo_pair):
    """Re-ranked neighbour lists of rigidly moved / scaled copies == brute-force kNN on the copies,
    including the certified fallback (forced by a factor that certifies nothing)."""
    from or_pcd_b200 import _lib, ops

    s, t, _ = armadillo_pair
    rng = np.random.default_rng(4)
    moved = []
    for _ in range(3):
        a, b, g = rng.uniform(-np.pi / 2, np.pi / 2, 3)
        moved.append(s @ _rot(a, b, g) + rng.standard_normal(3) * 0.1)
    scale = np.array([1.2, 0.9, 1.0])
    moved.append(s * scale)
    moved = np.stack(moved)
    t_ = _lib.torch()
    cand, bound, order = ops.knn_candidates(_dev(s), 48)
    assert np.array_equal(np.sort(order[0].cpu().numpy()), np.arange(s.shape[0]))       # a permutation of the cloud
    for factors, expect_all_fixed, use_order in (([1 - 1e-9] * 3 + [0.81 * (1 - 1e-9)], False, True),
                                                 ([1 - 1e-9] * 3 + [0.81 * (1 - 1e-9)], False, False),
                                                 ([0.0] * 4, True, True)):
        # the processing order (the base cloud's Hilbert order: L1 locality of the gathers) never changes the result
        idx, d2, n_fixed = ops.knn_rerank(cand, bound, _dev(moved), _dev(np.array(factors)), 20,
                                          order=order if use_order else None)
        for b in range(moved.shape[0]):
            oi, od = oracle.knn(moved[b], moved[b], 20)
            assert np.array_equal(idx[b].cpu().numpy(), oi), b
            assert np.array_equal(d2[b].cpu().numpy(), od), b
        nf = int(n_fixed.cpu().numpy()[0])
        assert (nf == moved.shape[0] * s.shape[0]) if expect_all_fixed else (nf < 0.05 * s.shape[0] * moved.shape[0])


def test_optimize_batch_rigid_base_same_result(oracle, armadillo_pair):
    from or_pcd_b200.Optimizer import GeneralizedICP

    s, t, _ = armadillo_pair
    rng = np.random.default_rng(8)
    srcs = np.stack([s @ _rot(*rng.uniform(-1, 1, 3)) + rng.standard_normal(3) * 0.1 for _ in range(4)])
    opt = GeneralizedICP()
    a = opt.optimize_batch(srcs, t[None], [0, 1, 2, 3], [0, 0, 0, 0])
    b = opt.optimize_batch(srcs, t[None], [0, 1, 2, 3], [0, 0, 0, 0], rigid_base=s)
    # the neighbour lists re-ranked from the base cloud equal the searched ones (so do the covariances, bit for bit);
    # with rigid_base the source copies also share the base cloud's spatial order, so their 64-point patches hold other
    # points and the float64 Jacobian sums are taken in another order: same iterations, results equal to rounding
    assert np.array_equal(a.iterations, b.iterations)
    assert np.allclose(a.rmse, b.rmse, rtol=1e-9, atol=0) and np.allclose(a.transforms, b.transforms, rtol=0, atol=1e-9)


def test_fused_round_kernel_agrees_with_the_unfused_and_dense_paths(oracle, armadillo_pair):
    """Four implementations of the same rounds -- the persistent work-queue kernel (default: all rounds in one launch),
    one fused kernel launch per round (sweep_mode 5), culled sweep + separate kernels (sweep_mode 4), dense sweep +
    separate kernels (sweep_mode 1) -- find the same correspondences, so they differ only by the order of the float64
    sums: same iteration counts, rmse and transforms to 1e-9.  The persistent kernel and the launch-per-round kernel
    share every instruction of the round and sum in the same (patch) order: bit-identical."""
    from or_pcd_b200.Optimizer import GeneralizedICP

    s, t, _ = armadillo_pair
    np.random.seed(3)
    starts = [oracle.initialize_rotation(0.6, 0.0, 0.05) for _ in range(4)]
    sources = np.stack([s @ r + tr for r, tr in starts])
    targets = np.stack([t, t * np.array([1.0, 1.15, 1.0])])
    us, ut = [0, 1, 2, 3, 0, 1], [0, 0, 0, 0, 1, 1]
    res = []
    for mode in (0, 5, 4, 1):
        opt = GeneralizedICP()
        opt.sweep_mode = mode
        res.append(opt.optimize_batch(sources, targets, us, ut))
    assert np.array_equal(res[0].rmse, res[1].rmse) and np.array_equal(res[0].transforms, res[1].transforms)
    for other in res[1:]:
        assert np.array_equal(res[0].iterations, other.iterations)
        assert np.allclose(res[0].rmse, other.rmse, rtol=1e-9, atol=0)
        assert np.allclose(res[0].transforms, other.transforms, rtol=0, atol=1e-9)


def test_correspondence_search_degenerate_boxes_and_exact_ties():
    """The culled 1-NN sweep on geometry that degenerates its culling data: an axis-aligned planar lattice (every chunk's
    bounding box has zero extent in z, rows of the lattice give boxes of zero extent in two axes) queried from points
    exactly half-way between lattice nodes (two or four targets at EXACTLY the same float64 distance: the FP32 filter
    cannot decide, the box-based exact search must, and ties go to the lowest index).  Coordinates are multiples of
    2^-5, so every squared distance is exact in float64 and the reference answer is a plain numpy arg-min."""
    import torch

    from or_pcd_b200 import ops

    g = np.arange(48, dtype=np.float64) / 32.0
    tgt = np.stack([a.ravel() for a in np.meshgrid(g, g, [0.25], indexing="ij")], axis=1)          # 2304 targets, z = 0.25
    rng = np.random.default_rng(5)
    tgt = tgt[rng.permutation(len(tgt))]                     # index order unrelated to position
    base = np.stack([a.ravel() for a in np.meshgrid(g[:-1], g[:-1], [0.25], indexing="ij")], axis=1)
    sources = np.stack([base + np.array([1 / 64.0, 0.0, 0.0]),            # between two nodes along x: 2-way ties
                        base + np.array([1 / 64.0, 1 / 64.0, 1 / 32.0]),  # cell centres, off the plane: 4-way ties
                        base + np.array([0.0, 0.0, 1 / 16.0])])           # straight above a node: unique
    T = np.tile(np.eye(4), (3, 1, 1))
    stats, status, corr, _ = ops.evaluate_batch(_dev(sources), _dev(tgt[None]), torch.tensor([0, 1, 2], dtype=torch.int32, device="cuda"),
                                                torch.zeros(3, dtype=torch.int32, device="cuda"), _dev(T), 0.5, want_corr=True)
    corr = corr.cpu().numpy()
    for u in range(3):
        d2 = ((sources[u][:, None, 0] - tgt[None, :, 0]) ** 2 + (sources[u][:, None, 1] - tgt[None, :, 1]) ** 2) + \
             (sources[u][:, None, 2] - tgt[None, :, 2]) ** 2
        ref = d2.argmin(axis=1)                              # first minimum = lowest index
        assert np.array_equal(corr[u], ref), (u, int((corr[u] != ref).sum()))
        rmse = np.sqrt(d2.min(axis=1).mean())
        assert abs(float(stats[u, 1]) - rmse) <= 1e-12 * rmse
