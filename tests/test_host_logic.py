"""Host-side logic of the plugin layer, no GPU needed: the preprocessing blocks that stay
in numpy, and the Aligner's speculative batch + replay of the reference's sequential
poll order, accept decisions, RNG consumption and exception behaviour
(reference: Aligner/Aligner.py:164-317)."""
import os

import numpy as np
import pytest

from fakes import FakeBatched, FakeSequential


def _cloud(seed, n=600):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, 3)) * np.array([1.0, 0.7, 0.4]) + rng.standard_normal(3)


def test_radius_scaler_matches_oracle_and_definition(oracle):
    from or_pcd_b200.Preprocessor import RadiusScaler

    c = _cloud(0, 3000)
    a, b = RadiusScaler(), oracle.RadiusScaler()
    out = a.process(c)
    assert np.array_equal(out, b.process(c))
    assert np.array_equal(a.mean, b.mean) and a.scale == b.scale and a.mean.shape == (1, 3)
    assert abs(np.linalg.norm(out, axis=1).max() - 1.0) < 1e-12          # radiusScaler.py:25-30
    assert np.abs(out.mean(0)).max() < 1e-12


def test_random_downsampler_consumes_the_global_stream(oracle):
    from or_pcd_b200.Preprocessor import RandomDownsampler

    c = _cloud(1, 2000)
    np.random.seed(5)
    out = RandomDownsampler(300).process(c)
    after = np.random.get_state()[1].copy()
    np.random.seed(5)
    keep = np.random.choice(2000, 300, replace=False)                     # randomDownsampler.py:35-37
    assert np.array_equal(out, c[keep])
    assert np.array_equal(after, np.random.get_state()[1])
    # non-positive size falls back to the default with a log line, not an exception
    assert RandomDownsampler(0)._sample_size == 4096


def test_preprocessor_inserts_radius_scaler_in_callers_list():
    from or_pcd_b200.Preprocessor import BaseScaler, Preprocessor, RadiusScaler, RandomDownsampler

    blocks = [RandomDownsampler(100)]
    p = Preprocessor(blocks)
    assert isinstance(blocks[0], RadiusScaler) and p.preprocessor_blocks is blocks   # preprocessor.py:19-24
    blocks2 = [RandomDownsampler(100), RadiusScaler()]
    Preprocessor(blocks2)
    assert len(blocks2) == 2 and isinstance(blocks2[1], BaseScaler)
    np.random.seed(0)
    assert p.preprocess(_cloud(2)).shape == (100, 3)


def _run_pair(oracle, fail_above=np.inf, attempts=5, seed=3):
    """oracle.Aligner over the sequential fake vs the B200 Aligner over the batched fake."""
    from or_pcd_b200.Aligner.Aligner import Aligner
    from or_pcd_b200.Preprocessor import Preprocessor, RandomDownsampler

    src = _cloud(10, 900)
    tgt = (_cloud(10, 900) * np.array([1.0, 1.31, 0.9]))[::-1].copy()
    outs = []
    for kind in ("seq", "batch"):
        np.random.seed(seed)
        if kind == "seq":
            al = oracle.Aligner(oracle.Preprocessor([oracle.RandomDownsampler(400)]),
                                oracle.Preprocessor([oracle.RandomDownsampler(400)]),
                                FakeSequential(fail_above), attempts=attempts)
        else:
            al = Aligner(Preprocessor([RandomDownsampler(400)]), Preprocessor([RandomDownsampler(400)]),
                         FakeBatched(fail_above), attempts=attempts)
        err = None
        try:
            res = al.align(src, tgt, refine_registration=False)
        except ValueError as e:
            res, err = None, e
        outs.append((al, res, err, np.random.get_state()[1].copy(), np.random.get_state()[2]))
    return outs


def test_aligner_replays_reference_decisions_and_rng(oracle):
    (a, ra, ea, sa, pa), (b, rb, eb, sb, pb) = _run_pair(oracle)
    assert ea is None and eb is None
    assert len(a.decision_log) == len(b.decision_log) > 7
    for x, y in zip(a.decision_log, b.decision_log):
        assert x["accepted"] == y["accepted"] and x["kind"] == y["kind"]
        assert np.array_equal(x["scale"], y["scale"]) and x["metric"] == y["metric"]
    assert np.array_equal(ra[0], rb[0]) and ra[1] == rb[1] and np.array_equal(ra[2], rb[2]) and ra[3] == rb[3]
    assert rb[2].shape == (1, 3)
    # the global numpy stream is left exactly where the sequential code leaves it
    assert np.array_equal(sa, sb) and pa == pb
    # the sequential code ran fewer registrations than the speculative batch evaluated
    assert b.n_optimize_calls == a._optimizer.calls
    assert b.n_units_evaluated >= b.n_optimize_calls
    assert any(d["accepted"] for d in b.decision_log[1:]), "test workload should accept at least one move"


def test_aligner_exception_replay(oracle):
    """A failing unit raises iff the sequential code would have reached it."""
    # threshold chosen so that only enlarged-z candidates fail
    probe = _run_pair(oracle)[0][0]
    z = probe.target_pre.std(0)[2]
    for fail_above in (z * 1.1, z * 0.5):
        (a, ra, ea, sa, pa), (b, rb, eb, sb, pb) = _run_pair(oracle, fail_above=fail_above)
        assert (ea is None) == (eb is None)
        if ea is None:
            assert np.array_equal(ra[2], rb[2]) and ra[1] == rb[1]
            assert np.array_equal(sa, sb) and pa == pb
        else:
            assert b.n_optimize_calls == a._optimizer.calls


def test_aligner_sequential_plugin_still_works(oracle):
    """An optimizer with only optimize() (the reference's contract) is driven one call at a time."""
    from or_pcd_b200.Aligner.Aligner import Aligner
    from or_pcd_b200.Preprocessor import Preprocessor, RandomDownsampler

    src, tgt = _cloud(4, 500), _cloud(5, 500)
    np.random.seed(1)
    ref = oracle.Aligner(oracle.Preprocessor([oracle.RandomDownsampler(200)]),
                         oracle.Preprocessor([oracle.RandomDownsampler(200)]), FakeSequential(), attempts=3)
    r0 = ref.align(src, tgt)
    np.random.seed(1)
    al = Aligner(Preprocessor([RandomDownsampler(200)]), Preprocessor([RandomDownsampler(200)]), FakeSequential(),
                 attempts=3)
    r1 = al.align(src, tgt, refine_registration=False)
    assert np.array_equal(r0[0], r1[0]) and r0[1] == r1[1] and np.array_equal(r0[2], r1[2]) and r0[3] == r1[3]


def test_transfrom_identity(oracle):
    """Aligner.transfrom (Aligner.py:367-394): with T = I and unit scale factors the call maps
    the source's normalised frame onto the target's raw frame."""
    from or_pcd_b200.Aligner.Aligner import Aligner
    from or_pcd_b200.Preprocessor import Preprocessor, RadiusScaler

    src, tgt = _cloud(6, 300), _cloud(7, 300)
    ps, pt = Preprocessor([RadiusScaler()]), Preprocessor([RadiusScaler()])
    al = Aligner(ps, pt, FakeSequential(), attempts=1)
    ps.preprocess(src); pt.preprocess(tgt)
    al.transfromation, al.scale_factors = np.eye(4), np.ones((1, 3))
    out = al.transfrom(src)
    s, t = ps.preprocessor_blocks[0], pt.preprocessor_blocks[0]
    assert np.allclose(out, (src - s.mean) / s.scale * t.scale + t.mean, rtol=0, atol=1e-14)


def test_aligner_argument_defaults():
    from or_pcd_b200.Aligner.Aligner import Aligner
    from or_pcd_b200.Preprocessor import Preprocessor, RadiusScaler

    al = Aligner(Preprocessor([RadiusScaler()]), Preprocessor([RadiusScaler()]), FakeSequential(), attempts=0,
                 deg=-1, std=0, delta=0, max_iter=0, eps=-2)
    assert (al._attempts, al._delta, al._eps, al._max_iter, al._std) == (30, 0.2, 0.05, 100, 0.1)   # constants.py:71-77
    assert "attempts=30" in repr(al)
    # the reference's default refinement (PointToPlane without target normals) raises inside Open3D
    with pytest.raises(RuntimeError):
        al.refine_registration(np.zeros((4, 3)), np.zeros((4, 3)), np.eye(4), icp_type="PointToPlane")


def test_cloud_io_ply_and_npz(tmp_path):
    """On-disk formats either side of the path: big/little-endian binary and ascii
    PLY vertices (float32 widened to float64 like Open3D's reader) and the driver's .npz layout."""
    from or_pcd_b200.data import load_cloud, load_sample_cloud, read_ply_vertices, sample_cloud_names

    rng = np.random.default_rng(0)
    v = rng.standard_normal((57, 3)).astype(np.float32)
    for fmt, order in (("binary_big_endian", ">"), ("binary_little_endian", "<")):
        hdr = (f"ply\nformat {fmt} 1.0\nobj_info test\nelement vertex 57\nproperty float x\nproperty float y\n"
               "property float z\nelement range_grid 2\nproperty list uchar int vertex_indices\nend_header\n")
        path = tmp_path / f"{fmt}.ply"
        path.write_bytes(hdr.encode() + v.astype(order + "f4").tobytes() + b"\x01\x00\x00\x00\x05\x00")
        out = read_ply_vertices(str(path))
        assert out.dtype == np.float64 and np.array_equal(out, v.astype(np.float64))
    path = tmp_path / "ascii.ply"
    path.write_text("ply\nformat ascii 1.0\nelement vertex 3\nproperty float z\nproperty float x\nproperty float y\n"
                    "end_header\n3 1 2\n6 4 5\n9 7 8\n")
    assert np.array_equal(load_cloud(str(path)), [[1, 2, 3], [4, 5, 6], [7, 8, 9]])
    np.savez(tmp_path / "c.npz", points=np.hstack([v, np.ones((57, 2), np.float32)]))
    assert np.array_equal(load_cloud(str(tmp_path / "c.npz")), v.astype(np.float64))
    assert "ArmadilloBack_330" in sample_cloud_names() and load_sample_cloud("ArmadilloBack_330").shape == (29116, 3)
    ref = "/root/reference/data/ArmadilloBack_330.ply"
    if os.path.exists(ref):      # authoring container only: the bundled scans equal the reference's PLY files
        assert np.array_equal(read_ply_vertices(ref), load_sample_cloud("ArmadilloBack_330"))
    with pytest.raises(FileNotFoundError):
        load_sample_cloud("nope")
    # the reference's signature: load_sample_cloud(name="ArmadilloBack_180") (data/__init__.py:5)
    assert load_sample_cloud().shape == (20535, 3) and np.array_equal(load_sample_cloud(name="ArmadilloBack_180"), load_sample_cloud())
    # object arrays are never unpickled (an untrusted .npz must not execute code)
    np.savez(tmp_path / "evil.npz", points=np.array([{"a": 1}, None], dtype=object))
    with pytest.raises(ValueError):
        load_cloud(str(tmp_path / "evil.npz"))
    np.savez(tmp_path / "nopoints.npz", other=v)
    with pytest.raises(ValueError):
        load_cloud(str(tmp_path / "nopoints.npz"))


# ---------------------------------------------------------------------------------------------------------------------
# Against the REFERENCE'S OWN Aligner class (tests/golden/make_golden_aligner_ref.py ran the unmodified
# /root/reference/src/or_pcd/Aligner/Aligner.py with open3d stubbed out and the fake optimizer of tests/fakes.py)
# ---------------------------------------------------------------------------------------------------------------------
_REF_CASES = ["plain", "more_starts", "fails_late", "fails_at_once"]


def _run_against_reference_golden(g, name, make_aligner, optimizer):
    seed, attempts, size = int(g[f"{name}_seed"]), int(g[f"{name}_attempts"]), int(g[f"{name}_size"])
    np.random.seed(seed)
    al = make_aligner(size, optimizer, attempts)
    raised, res = 0, None
    try:
        res = al.align(g["source"], g["target"], refine_registration=False)
    except ValueError:
        raised = 1
    st = np.random.get_state()
    assert raised == int(g[f"{name}_raised"]), name
    if not raised:
        T, metric, scale, errors = res
        assert np.array_equal(np.asarray(T), g[f"{name}_T"]), name
        assert metric == float(g[f"{name}_metric"]), name
        assert np.array_equal(np.asarray(scale), g[f"{name}_scale"]) and np.asarray(scale).shape == (1, 3), name
        assert np.array_equal(np.asarray(errors), g[f"{name}_errors"]), name
        # the global numpy stream is left exactly where the reference leaves it
        assert np.array_equal(st[1], g[f"{name}_rng_keys"]) and st[2] == int(g[f"{name}_rng_pos"]), name
    return al


@pytest.mark.parametrize("name", _REF_CASES)
def test_oracle_aligner_equals_the_references_own_class(oracle, name):
    """oracle.Aligner (the restatement every GPU test is judged against) takes the decisions of the reference's class:
    transformation, metric, scale factors, error list bit for bit, the same number of optimizer calls, the same
    exception behaviour and the same position of the global numpy stream."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_aligner_ref.npz"))
    opt = FakeSequential(float(g[f"{name}_fail_above"]))
    _run_against_reference_golden(
        g, name, lambda size, o, attempts: oracle.Aligner(oracle.Preprocessor([oracle.RandomDownsampler(size)]),
                                                          oracle.Preprocessor([oracle.RandomDownsampler(size)]), o,
                                                          attempts=attempts), opt)
    assert opt.calls == int(g[f"{name}_calls"])


@pytest.mark.parametrize("name", _REF_CASES)
@pytest.mark.parametrize("batched", [False, True])
def test_b200_aligner_equals_the_references_own_class(name, batched):
    """The B200 Aligner -- driven one call at a time by a reference-style optimizer, or speculatively in batches -- against
    the same golden vectors of the reference's class.  The batched path evaluates more units than the sequential code
    reaches; the calls the sequential code WOULD have made are what is compared."""
    from or_pcd_b200.Aligner.Aligner import Aligner
    from or_pcd_b200.Preprocessor import Preprocessor, RandomDownsampler

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_aligner_ref.npz"))
    opt = (FakeBatched if batched else FakeSequential)(float(g[f"{name}_fail_above"]))
    al = _run_against_reference_golden(
        g, name, lambda size, o, attempts: Aligner(Preprocessor([RandomDownsampler(size)]),
                                                   Preprocessor([RandomDownsampler(size)]), o, attempts=attempts), opt)
    if batched:
        if not int(g[f"{name}_raised"]):      # (after a failure the sequential-equivalent count is the failing call's)
            assert al.n_optimize_calls == int(g[f"{name}_calls"])
    else:
        assert opt.calls == int(g[f"{name}_calls"])
    if name == "plain":     # Aligner.transfrom (Aligner.py:367-394) with this run's result, against the reference's class
        assert np.array_equal(al.transfrom(g["source"]), g["plain_transformed_source"])


def test_b200_optimizer_constructors_equal_the_references_own_classes():
    """Constructor semantics of the plugin classes against what the reference's own classes handed to Open3D
    (golden_wrappers_ref.npz): defaults, explicit values, and non-positive values replaced by the defaults."""
    from or_pcd_b200.Optimizer import FastGlobalOptimizer, GeneralizedICP

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_wrappers_ref.npz"))
    for name, kw in (("gicp_default", {}), ("gicp_params", dict(max_correspondence_distance=0.3, max_iterations=7)),
                     ("gicp_nonpositive", dict(max_correspondence_distance=-1.0, max_iterations=0))):
        opt = GeneralizedICP(**kw)
        assert opt._max_correspondence_distance == float(g[f"{name}_max_dist"]), name
        assert opt._max_iterations == int(g[f"{name}_max_iteration"]), name
    for name, kw in (("fgr_default", {}), ("fgr_params", dict(tuple_scale=0.8, iteration_number=40, decrease_mu=False,
                                                              maximum_correspondence_distance=0.3))):
        opt = FastGlobalOptimizer(**kw)
        o = g[f"{name}_option"]
        assert (opt._division_factor, opt._tuple_scale, opt._maximum_correspondence_distance, opt._iteration_number,
                float(opt._decrease_mu)) == (o[0], o[1], o[2], int(o[3]), o[4]), name
        assert (opt._normal_estimate_radius, opt._normal_estimate_knn) == (float(g[f"{name}_normal_radius"]),
                                                                           int(g[f"{name}_normal_knn"]))
        assert (opt._fpfh_radius, opt._fpfh_knn) == (float(g[f"{name}_fpfh_radius"]), int(g[f"{name}_fpfh_knn"]))
        assert opt.feature_mode == "reference_compat"     # the default reproduces the reference: both FPFH sets from the source
        assert list(g[f"{name}_fpfh_clouds"]) == ["source", "source"]
