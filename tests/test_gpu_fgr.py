"""Parity of the FGR path (FPFH, tensor-core feature matching, tuple test, GNC solve, evaluation)
against the CPU oracle on the same inputs.  Indices: bit-exact.  Floating point: tolerances from
BASELINE.json north_star written beside each assert."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _dev(a, dtype=np.float64):
    from or_pcd_b200 import _lib

    return _lib.to_device(a, dtype)


def _rot_angle(Ra, Rb):
    return np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1))


def _i32(a):
    import torch

    return torch.as_tensor(np.asarray(a, dtype=np.int32)).cuda()


def test_fpfh_matches_oracle(oracle, armadillo_pair):
    from or_pcd_b200 import ops

    s, t, _ = armadillo_pair
    for cloud in (s, t[:1777]):
        nrm = oracle.normals(cloud, 20, 0.1)["normal"]
        ref, ref_sp = oracle.fpfh(cloud, nrm, 20, 0.1, want_spfh=True)
        d = _dev(cloud[None])
        idx, d2 = ops.knn_exact(d, d, 20)
        out, sp = ops.fpfh(d, _dev(nrm[None]), idx, d2, 20, 0.1, want_spfh=True)
        out, sp = out[0].cpu().numpy(), sp[0].cpu().numpy()
        # same arithmetic on both sides, det_acos / det_atan2 included: SPFH and FPFH are bit-identical
        # (the device forms its 33 quotients per neighbour from one division, correctly rounded)
        assert np.array_equal(sp, ref_sp)
        assert np.array_equal(out, ref)
        # device normals (bit-identical to the oracle's) through the plugin class
        from or_pcd_b200.Optimizer import FastGlobalOptimizer

        fs, _ = FastGlobalOptimizer().get_fpfh_features(cloud, cloud)
        assert np.array_equal(fs, ref)


@pytest.mark.parametrize("mode", [0, 1])
def test_feature_match_exact_indices(oracle, mode):
    """Nearest neighbours in 33-D: indices and float64 distances equal to the oracle's kd-tree,
    with duplicates (exact ties -> lowest index), ragged sizes and several pairs per launch."""
    from or_pcd_b200 import ops

    rng = np.random.default_rng(7)
    nq, ndb = 700, 1000
    A = rng.gamma(2.0, 30.0, (2, nq, 33))
    B = rng.gamma(2.0, 30.0, (3, ndb, 33))
    B[0, 500:520] = B[0, 100:120]                  # exact duplicate rows
    A[0, :50] = B[0, 100:150]                      # queries that coincide with database rows
    B[2, :] = np.round(B[2, :] / 40.0) * 40.0      # coarse lattice: many exact ties
    pa = ops.feature_pack(_dev(A))
    pb = ops.feature_pack(_dev(B), center=pa.center)
    pq, pd = [0, 1, 0, 1], [0, 1, 2, 2]
    nn, d2, scans, dbg = ops.feature_match(pa, pb, _i32(pq), _i32(pd), mode=mode, want_d2=True,
                                           want_debug=(mode == 0))
    nn, d2 = nn.cpu().numpy(), d2.cpu().numpy()
    for p in range(4):
        oi, od = oracle.knn(B[pd[p]], A[pq[p]], 1)
        assert np.array_equal(nn[p], oi[:, 0]), (p, np.flatnonzero(nn[p] != oi[:, 0])[:10])
        assert np.array_equal(d2[p], od[:, 0])
        if mode == 0:
            v = dbg[p].cpu().numpy()[:, :ndb].astype(np.float64)
            mu = pa.center.cpu().numpy()                      # the filter runs on features centred on set 0's mean
            Ac, Bc = A[pq[p]] - mu, B[pd[p]] - mu
            ref = (Bc ** 2).sum(1)[None, :] - 2.0 * Ac @ Bc.T
            scale = np.sqrt((Ac ** 2).sum(1))[:, None] * np.sqrt((Bc ** 2).sum(1)).max()
            assert (np.abs(v - ref) / scale).max() < 2.0 ** -17          # 3xTF32 split: FP32-class accuracy
    if mode == 0:
        assert int(scans[0].item()) < 0.5 * 4 * nq                            # most rows certified by the filter
    else:
        assert int(scans[0].item()) == 4 * nq


def test_feature_match_on_real_fpfh(oracle, armadillo_pair):
    from or_pcd_b200 import ops

    s, t, _ = armadillo_pair
    fs = oracle.fpfh(s, oracle.normals(s, 20, 0.1)["normal"], 20, 0.1)
    ft = oracle.fpfh(t, oracle.normals(t, 20, 0.1)["normal"], 20, 0.1)
    pa = ops.feature_pack(_dev(fs[None]))
    pb = ops.feature_pack(_dev(ft[None]), center=pa.center)
    z = _i32([0])
    for q, db, fq, fdb in ((pa, pb, fs, ft), (pb, pa, ft, fs), (pa, pa, fs, fs)):
        nn, _, scans, _ = ops.feature_match(q, db, z, z)
        oi, _ = oracle.knn(fdb, fq, 1)
        assert np.array_equal(nn[0].cpu().numpy(), oi[:, 0])
        assert int(scans[0].item()) < 0.2 * fq.shape[0]


def test_evaluate_batch_matches_oracle(oracle, armadillo_pair):
    from or_pcd_b200 import ops

    s, t, _ = armadillo_pair
    rng = np.random.default_rng(3)
    Ts = []
    for k in range(3):
        a = rng.uniform(-0.3, 0.3)
        T = np.eye(4)
        T[:3, :3] = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
        T[:3, 3] = rng.standard_normal(3) * 0.05
        Ts.append(T)
    Ts.append(np.eye(4))
    Ts = np.stack(Ts)
    targets = np.stack([t, t * np.array([1.0, 1.2, 1.0])])
    us, ut = [0, 0, 0, 0], [0, 1, 0, 1]
    stats, status, corr, _ = ops.evaluate_batch(_dev(s[None]), _dev(targets), _i32(us), _i32(ut), _dev(Ts), 0.5,
                                                want_corr=True)
    stats, corr = stats.cpu().numpy(), corr.cpu().numpy()
    for u in range(4):
        fit, rmse, nc, oc = oracle.evaluate(s, targets[ut[u]], 0.5, Ts[u])
        assert np.array_equal(corr[u], oc) and int(stats[u, 3]) == nc
        assert abs(stats[u, 0] - fit) <= 1e-12 and abs(stats[u, 1] - rmse) <= 1e-12 * rmse
    # nothing within reach -> status 1 (the reference raises Warning)
    stats, status, _, _ = ops.evaluate_batch(_dev(s[None, :300] + 50.0), _dev(t[None, :200]), _i32([0]), _i32([0]),
                                             None, 0.5)
    assert int(status.cpu()[0]) == 1 and float(stats.cpu()[0, 1]) == 0.0


@pytest.mark.parametrize("mode", ["reference_compat", "fixed"])
def test_fgr_stages_match_oracle_given_same_features(oracle, armadillo_pair, mode):
    """Cross check, tuple test (shared Philox stream), GNC solve and final evaluation against the
    oracle when both sides start from the same feature tables."""
    from or_pcd_b200 import ops

    s, t, _ = armadillo_pair
    fs = oracle.fpfh(s, oracle.normals(s, 20, 0.1)["normal"], 20, 0.1)
    ft = fs if mode == "reference_compat" else oracle.fpfh(t, oracle.normals(t, 20, 0.1)["normal"], 20, 0.1)
    key = oracle.tuple_key(0, s, t)
    ref = oracle.fgr(s, t, fs, ft, key=key, want_corres=True)
    ij, _ = oracle.knn(ft, fs, 1)
    ji, _ = oracle.knn(fs, ft, 1)
    import torch

    keys = torch.as_tensor(np.array([key], dtype=np.uint64).view(np.int64)).cuda()
    T, info = ops.fgr_batch(_dev(s[None]), _dev(t[None]), _i32([0]), _i32([0]), _i32(ij[:, 0][None]),
                            _i32(ji[:, 0][None]), keys)
    info = info.cpu().numpy()[0]
    assert info[0] == ref["n_initial"] and info[1] == ref["n_pairs"] and info[2] == ref["trials"]
    assert bool(info[3]) == ref["swapped"]
    T = T.cpu().numpy()[0]
    diag = np.linalg.norm(t.max(0) - t.min(0))
    assert _rot_angle(T[:3, :3], ref["T"][:3, :3]) < 1e-4                   # north_star: 1e-4 rad
    assert np.linalg.norm(T[:3, 3] - ref["T"][:3, 3]) < 1e-5 * diag         # 1e-5 x bbox diagonal
    stats, status, _, _ = ops.evaluate_batch(_dev(s[None]), _dev(t[None]), _i32([0]), _i32([0]), _dev(T[None]), 0.5)
    stats = stats.cpu().numpy()[0]
    assert abs(stats[0] - ref["fitness"]) <= 1e-5 * ref["fitness"] and abs(stats[1] - ref["rmse"]) <= 1e-5 * ref["rmse"]


@pytest.mark.parametrize("mode", ["reference_compat", "fixed"])
def test_fast_global_optimizer_plugin(oracle, armadillo_pair, mode):
    from or_pcd_b200.Optimizer import FastGlobalOptimizer

    s, t, _ = armadillo_pair
    opt = FastGlobalOptimizer(feature_mode=mode)
    T, rmse = opt.optimize(s, t)
    ref = oracle.FastGlobalOptimizer(feature_mode=mode)
    rT, rrmse = ref.optimize(s, t)
    assert opt.last_info[0, 0] == ref.last["n_initial"] and opt.last_info[0, 1] == ref.last["n_pairs"]
    assert abs(rmse - rrmse) <= 1e-5 * rrmse
    assert _rot_angle(T[:3, :3], rT[:3, :3]) < 1e-4
    assert np.linalg.norm(T[:3, 3] - rT[:3, 3]) < 1e-5 * np.linalg.norm(t.max(0) - t.min(0))
    # swapped roles (target larger than source) and a batch of several units, one launch sequence
    res = opt.optimize_batch(np.stack([t[:4000], t[:4000] @ np.array([[0, -1.0, 0], [1, 0, 0], [0, 0, 1]])]),
                             np.stack([s, s * np.array([1.0, 1.2, 1.0])]), [0, 1, 0], [0, 0, 1])
    for u, (a, b) in enumerate(((0, 0), (1, 0), (0, 1))):
        srcs = [t[:4000], t[:4000] @ np.array([[0, -1.0, 0], [1, 0, 0], [0, 0, 1]])]
        tgts = [s, s * np.array([1.0, 1.2, 1.0])]
        rT, rrmse = ref.optimize(srcs[a], tgts[b])
        assert ref.last["swapped"]
        assert opt.last_info[u, 1] == ref.last["n_pairs"]
        assert abs(res.rmse[u] - rrmse) <= 1e-5 * rrmse, u
        assert _rot_angle(res.transforms[u][:3, :3], rT[:3, :3]) < 1e-4


def test_fast_global_optimizer_raises_warning_without_correspondences(oracle):
    from or_pcd_b200.Optimizer import FastGlobalOptimizer

    rng = np.random.default_rng(0)
    src = oracle.RadiusScaler().process(oracle.synthetic_cloud(600, rng))
    tgt = oracle.RadiusScaler().process(oracle.synthetic_cloud(500, rng)) * 1e-3 + 40.0
    with pytest.raises(Warning):
        FastGlobalOptimizer(maximum_correspondence_distance=1e-3).optimize(src, tgt)
    with pytest.raises(Warning):
        oracle.FastGlobalOptimizer(maximum_correspondence_distance=1e-3).optimize(src, tgt)


def test_aligner_with_fgr_replays_oracle_decisions(oracle):
    """Outer pattern search over the FGR inner block on 1500-point clouds, few starts: the decision
    sequence and the chosen scale triple equal the oracle's sequential run."""
    from or_pcd_b200.Aligner.Aligner import Aligner
    from or_pcd_b200.data import load_sample_cloud
    from or_pcd_b200.Optimizer import FastGlobalOptimizer
    from or_pcd_b200.Preprocessor import Preprocessor, RandomDownsampler, SOR

    src, tgt = load_sample_cloud("ArmadilloBack_330"), load_sample_cloud("ArmadilloBack_0")
    np.random.seed(2)
    ref = oracle.Aligner(oracle.Preprocessor([oracle.RandomDownsampler(1500), oracle.SOR()]),
                         oracle.Preprocessor([oracle.RandomDownsampler(1500), oracle.SOR()]),
                         oracle.FastGlobalOptimizer(feature_mode="fixed"), attempts=2, delta=0.2, eps=0.1)
    rT, rm, rscale, rerr = ref.align(src, tgt)
    np.random.seed(2)
    al = Aligner(Preprocessor([RandomDownsampler(1500), SOR()]), Preprocessor([RandomDownsampler(1500), SOR()]),
                 FastGlobalOptimizer(feature_mode="fixed"), attempts=2, delta=0.2, eps=0.1)
    T, m, scale, err = al.align(src, tgt, refine_registration=False)
    assert [d["accepted"] for d in al.decision_log] == [d["accepted"] for d in ref.decision_log]
    assert np.array_equal(scale, rscale)
    assert np.allclose(err, rerr, rtol=1e-5, atol=0) and abs(m - rm) <= 1e-5 * rm
    assert _rot_angle(T[:3, :3], rT[:3, :3]) < 1e-4
