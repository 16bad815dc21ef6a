"""End-to-end: Aligner.align() with the B200 GeneralizedICP / SOR blocks on the Armadillo
pair of BASELINE config 1 against the committed oracle run (tests/golden/golden_v1.npz:
870 sequential inner registrations of the restated reference).  Tolerances from
BASELINE.json north_star: chosen scale triple identical; rmse 1e-5 relative; rotation
1e-4 rad; translation 1e-5 x bbox diagonal; masks bit-exact."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")


def _rot_angle(Ra, Rb):
    return np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1))


@pytest.fixture(scope="module")
def golden():
    with np.load(GOLDEN) as z:
        return {k: z[k] for k in z.files}


def test_preprocessing_blocks_match_golden(golden):
    from or_pcd_b200.data import load_sample_cloud
    from or_pcd_b200.Preprocessor import Preprocessor, RandomDownsampler, SOR

    for tag, names in (("back", ("ArmadilloBack_330", "ArmadilloBack_0")),
                       ("onhead", ("ArmadilloOnHeadMultiple_150", "ArmadilloOnHeadMultiple_240"))):
        np.random.seed(0)
        for name in names:
            blocks = [RandomDownsampler(5000), SOR()]
            out = Preprocessor(blocks).preprocess(load_sample_cloud(name))
            assert np.array_equal(np.packbits(blocks[2].last_mask), golden[f"{tag}_{name}_sor_mask"])   # bit-exact
            assert out.shape[0] == int(golden[f"{tag}_{name}_n_after"])
            assert np.isclose(blocks[0].scale, golden[f"{tag}_{name}_scale"], rtol=1e-14)
        assert np.array_equal(np.random.get_state()[1][:8], golden[f"{tag}_rng_after_pre"])


def test_align_config1_reproduces_the_oracle_run(golden):
    from or_pcd_b200.Aligner.Aligner import Aligner
    from or_pcd_b200.data import load_sample_cloud
    from or_pcd_b200.Optimizer import GeneralizedICP
    from or_pcd_b200.Preprocessor import Preprocessor, RandomDownsampler, SOR

    np.random.seed(0)
    al = Aligner(Preprocessor([RandomDownsampler(5000), SOR()]), Preprocessor([RandomDownsampler(5000), SOR()]),
                 GeneralizedICP())
    T, m, scale, errors = al.align(load_sample_cloud("ArmadilloBack_330"), load_sample_cloud("ArmadilloBack_0"),
                                   refine_registration=False)
    # decisions: the same candidates, in the same order, with the same verdicts
    acc = np.array([d["accepted"] for d in al.decision_log])
    assert np.array_equal(acc, golden["align_back_log_accepted"])
    assert np.array_equal(np.array([d["scale"].ravel() for d in al.decision_log]), golden["align_back_log_scale"])
    gm = golden["align_back_log_metric"]
    lm = np.array([d["metric"] for d in al.decision_log])
    assert np.abs(lm - gm).max() <= 1e-5 * gm.min()                       # rmse 1e-5 relative
    assert np.array_equal(scale, golden["align_back_scale"])              # chosen scale triple identical
    assert abs(m - float(golden["align_back_metric"])) <= 1e-5 * m
    assert np.allclose(errors, golden["align_back_errors"], rtol=1e-5, atol=0)
    diag = np.linalg.norm(al.target_pre.max(0) - al.target_pre.min(0))
    assert _rot_angle(T[:3, :3], golden["align_back_T"][:3, :3]) < 1e-4
    assert np.linalg.norm(T[:3, 3] - golden["align_back_T"][:3, 3]) < 1e-5 * diag
    # sequential-equivalent work and RNG position
    assert al.n_optimize_calls == int(golden["align_back_calls"]) == 870
    assert np.array_equal(np.random.get_state()[1][:8], golden["align_back_rng_after"])
    # record how tight the accept margins were (decision parity is only meaningful if they
    # exceed the run-to-run agreement)
    print("max |rmse_gpu - rmse_oracle| over the decision log:", np.abs(lm - gm).max())


def _golden(name):
    with np.load(os.path.join(os.path.dirname(GOLDEN), name)) as z:
        return {k: z[k] for k in z.files}


def _check_against_run(al, T, m, scale, errors, g, calls):
    acc = np.array([d["accepted"] for d in al.decision_log])
    assert np.array_equal(acc, g["log_accepted"])
    assert np.array_equal(np.array([d["scale"].ravel() for d in al.decision_log]), g["log_scale"])
    lm = np.array([d["metric"] for d in al.decision_log])
    assert np.abs(lm - g["log_metric"]).max() <= 1e-5 * g["log_metric"].min()      # rmse 1e-5 relative
    assert np.array_equal(scale, g["scale"])                                      # chosen scale triple identical
    assert abs(m - float(g["metric"])) <= 1e-5 * m
    assert np.allclose(errors, g["errors"], rtol=1e-5, atol=0)
    diag = np.linalg.norm(al.target_pre.max(0) - al.target_pre.min(0))
    assert _rot_angle(T[:3, :3], g["T"][:3, :3]) < 1e-4
    assert np.linalg.norm(T[:3, 3] - g["T"][:3, 3]) < 1e-5 * diag
    assert al.n_optimize_calls == int(g["calls"]) == calls
    assert np.array_equal(np.random.get_state()[1][:8], g["rng_after"])


def test_align_config3_onhead_64_starts():
    """BASELINE config 3: OnHead pair, attempts=64 (2496 sequential-equivalent registrations)."""
    from or_pcd_b200.Aligner.Aligner import Aligner
    from or_pcd_b200.data import load_sample_cloud
    from or_pcd_b200.Optimizer import GeneralizedICP
    from or_pcd_b200.Preprocessor import Preprocessor, RandomDownsampler, SOR

    np.random.seed(0)
    al = Aligner(Preprocessor([RandomDownsampler(5000), SOR()]), Preprocessor([RandomDownsampler(5000), SOR()]),
                 GeneralizedICP(), attempts=64)
    T, m, scale, errors = al.align(load_sample_cloud("ArmadilloOnHeadMultiple_150"),
                                   load_sample_cloud("ArmadilloOnHeadMultiple_240"), refine_registration=False)
    _check_against_run(al, T, m, scale, errors, _golden("golden_onhead64.npz"), 2496)


def test_align_config2_fgr_reference_compat():
    """BASELINE config 2: Back pair, FastGlobalOptimizer in the reference's literal feature mode."""
    from or_pcd_b200.Aligner.Aligner import Aligner
    from or_pcd_b200.data import load_sample_cloud
    from or_pcd_b200.Optimizer import FastGlobalOptimizer
    from or_pcd_b200.Preprocessor import Preprocessor, RandomDownsampler, SOR

    np.random.seed(0)
    al = Aligner(Preprocessor([RandomDownsampler(5000), SOR()]), Preprocessor([RandomDownsampler(5000), SOR()]),
                 FastGlobalOptimizer(feature_mode="reference_compat"))
    T, m, scale, errors = al.align(load_sample_cloud("ArmadilloBack_330"), load_sample_cloud("ArmadilloBack_0"),
                                   refine_registration=False)
    g = _golden("golden_fgr_back.npz")
    _check_against_run(al, T, m, scale, errors, g, int(g["calls"]))


def test_fgr_single_calls_match_golden(golden, armadillo_pair):
    from or_pcd_b200.Optimizer import FastGlobalOptimizer

    s, t, _ = armadillo_pair
    for mode in ("reference_compat", "fixed"):
        opt = FastGlobalOptimizer(feature_mode=mode)
        T, rmse = opt.optimize(s, t)
        gs = golden[f"back_fgr_{mode}_stats"]          # fitness, rmse, n_corr, n_initial, n_pairs, trials
        assert opt.last_info[0, 0] == int(gs[3]) and opt.last_info[0, 1] == int(gs[4]) and opt.last_info[0, 2] == int(gs[5])
        assert abs(rmse - gs[1]) <= 1e-5 * gs[1]
        assert _rot_angle(T[:3, :3], golden[f"back_fgr_{mode}_T"][:3, :3]) < 1e-4
    fs, _ = FastGlobalOptimizer().get_fpfh_features(s, t)
    assert np.allclose(fs[:64], golden["back_src_fpfh_head"], rtol=0, atol=1e-9)
    assert np.allclose(fs.sum(0), golden["back_src_fpfh_colsum"], rtol=1e-12)


def test_refine_registration_icp(oracle, armadillo_pair):
    """Aligner.refine_registration: point-to-point and point-to-plane ICP against
    the oracle's restatement of Open3D registration_icp, literal conventions of the reference."""
    from or_pcd_b200.Aligner.Aligner import Aligner
    from or_pcd_b200.Optimizer import GeneralizedICP
    from or_pcd_b200.Preprocessor import Preprocessor, RadiusScaler

    s, t, _ = armadillo_pair
    al = Aligner(Preprocessor([RadiusScaler()]), Preprocessor([RadiusScaler()]), GeneralizedICP())
    a = 0.3
    init = np.eye(4)
    init[:3, :3] = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    init[:3, 3] = [0.02, -0.01, 0.03]
    diag = np.linalg.norm(t.max(0) - t.min(0))
    for init_T in (np.eye(4), init):
        T, rmse = al.refine_registration(s, t, init_T, icp_type="PointToPoint")
        ref = oracle.icp(s, t, 0.5, 200, init_T, "PointToPoint")
        assert abs(rmse - ref["rmse"]) <= 1e-5 * ref["rmse"]                       # north_star tolerances
        assert _rot_angle(T[:3, :3], ref["T"][:3, :3]) < 1e-4
        assert np.linalg.norm(T[:3, 3] - ref["T"][:3, 3]) < 1e-5 * diag
        tn = oracle.normals(t, 20, 0.0)["normal"]
        T, rmse = al.refine_registration(s, t, init_T, icp_type="PointToPlane", target_normals=tn)
        ref = oracle.icp(s, t, 0.5, 200, init_T, "PointToPlane", tn)
        assert abs(rmse - ref["rmse"]) <= 1e-5 * ref["rmse"]
        assert _rot_angle(T[:3, :3], ref["T"][:3, :3]) < 1e-4
        assert np.linalg.norm(T[:3, 3] - ref["T"][:3, 3]) < 1e-5 * diag
    # the reference's default call (PointToPlane, no target normals) raises inside Open3D
    with pytest.raises(RuntimeError):
        al.refine_registration(s, t, np.eye(4), icp_type="PointToPlane")
    with pytest.raises(RuntimeError):
        oracle.icp(s, t, mode="PointToPlane")


def test_device_start_clouds_equal_numpy_bit_for_bit(armadillo_pair):
    """orpcd_apply_rigid_batch reproduces np.dot(source, R) + t (Aligner.py:183-185) exactly: the Aligner builds
    its start clouds on the GPU only under that condition."""
    from or_pcd_b200 import _lib, ops
    from or_pcd_b200.Aligner.Aligner import Aligner
    from or_pcd_b200.Optimizer import GeneralizedICP
    from or_pcd_b200.Preprocessor import Preprocessor

    s, _, _ = armadillo_pair
    np.random.seed(5)
    al = Aligner(Preprocessor([]), Preprocessor([]), GeneralizedICP(), attempts=40)
    rots, trans = zip(*[al.initialize_rotation() for _ in range(40)])
    dev = ops.apply_rigid_batch(_lib.to_device(s), _lib.to_device(np.stack(rots)), _lib.to_device(np.stack(trans)))
    host = np.stack([np.dot(s, r) + t for r, t in zip(rots, trans)])
    assert np.array_equal(dev.cpu().numpy(), host)


def test_single_calls_equal_the_references_own_wrapper_classes():
    """GeneralizedICP.optimize / FastGlobalOptimizer.optimize through the C ABI against the golden vectors of the
    reference's OWN wrapper classes (tests/golden/make_golden_wrappers_ref.py: generalizedICP.py / fastGlobalOptimizer.py
    run unmodified over an oracle-backed open3d stub): the rotation block comes out transposed like the reference's, the
    constructor arguments mean the same, and the failure rules raise the same exceptions.  north_star tolerances."""
    from or_pcd_b200.Optimizer import FastGlobalOptimizer, GeneralizedICP

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_wrappers_ref.npz"))
    s, t = g["source"], g["target"]
    diag = np.linalg.norm(t.max(0) - t.min(0))
    for name, kw in (("gicp_default", {}), ("gicp_params", dict(max_correspondence_distance=0.3, max_iterations=7)),
                     ("gicp_nonpositive", dict(max_correspondence_distance=-1.0, max_iterations=0))):
        T, rmse = GeneralizedICP(**kw).optimize(s, t)
        assert abs(rmse - float(g[f"{name}_rmse"])) <= 1e-5 * rmse, name
        assert _rot_angle(T[:3, :3], g[f"{name}_T"][:3, :3]) < 1e-4, name
        assert np.linalg.norm(T[:3, 3] - g[f"{name}_T"][:3, 3]) < 1e-5 * diag, name
    with pytest.raises(ValueError):
        GeneralizedICP(max_correspondence_distance=1e-3).optimize(s[:200] + 50.0, t[:300])
    for name, kw in (("fgr_default", {}), ("fgr_params", dict(tuple_scale=0.8, iteration_number=40, decrease_mu=False,
                                                              maximum_correspondence_distance=0.3))):
        T, rmse = FastGlobalOptimizer(**kw).optimize(s, t)
        assert abs(rmse - float(g[f"{name}_rmse"])) <= 1e-5 * rmse, name
        assert _rot_angle(T[:3, :3], g[f"{name}_T"][:3, :3]) < 1e-4, name
        assert np.linalg.norm(T[:3, 3] - g[f"{name}_T"][:3, 3]) < 1e-5 * diag, name
    with pytest.raises(Warning):
        FastGlobalOptimizer(maximum_correspondence_distance=1e-4).optimize(s, t + 30.0)


def test_refine_registration_equals_the_references_own_method():
    """Aligner.refine_registration against the reference's OWN method (golden_aligner_ref.npz: Aligner.py:319-364 run
    unmodified over an open3d stub whose registration_icp is the oracle's): the initial transform is handed over
    unchanged, the defaults are distance 0.5 / 200 iterations, the result comes back untransposed, and the default
    "PointToPlane" without target normals raises."""
    import inspect

    from or_pcd_b200.Aligner.Aligner import Aligner
    from or_pcd_b200.Optimizer import GeneralizedICP
    from or_pcd_b200.Preprocessor import Preprocessor, RadiusScaler

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_aligner_ref.npz"))
    assert int(g["refine_init_passed_unchanged"]) == 1 and int(g["refine_point_to_plane_without_normals_raises"]) == 1
    al = Aligner(Preprocessor([RadiusScaler()]), Preprocessor([RadiusScaler()]), GeneralizedICP())
    sig = inspect.signature(al.refine_registration).parameters
    assert sig["max_iteration"].default == int(g["refine_max_iteration"])
    assert sig["distance_threshold"].default == float(g["refine_max_dist"])
    s, t, init = g["refine_source"], g["refine_target"], g["refine_init"]
    T, rmse = al.refine_registration(s, t, init, icp_type="PointToPoint")
    diag = np.linalg.norm(t.max(0) - t.min(0))
    assert abs(rmse - float(g["refine_rmse"])) <= 1e-5 * rmse
    assert _rot_angle(T[:3, :3], g["refine_T"][:3, :3]) < 1e-4
    assert np.linalg.norm(T[:3, 3] - g["refine_T"][:3, 3]) < 1e-5 * diag
    with pytest.raises(RuntimeError):
        al.refine_registration(s, t, init, icp_type="PointToPlane")
