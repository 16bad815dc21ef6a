"""Golden vectors of the REFERENCE'S OWN `Aligner` (multi-start + compass search + accept logic + RNG consumption,
/root/reference/src/or_pcd/Aligner/Aligner.py:125-317), run in the authoring container.

The reference's control flow is plain Python / numpy; only its inner optimizers need Open3D.  So the class is imported
from /root/reference/src by path with `open3d`, the Visualizer and the file logger stubbed out (none of them is touched
by `align(..., refine_registration=False)`), and driven with the deterministic stand-in optimizer of tests/fakes.py.
What it returns -- transformation, metric, scale factors, error list, number of optimizer calls, the state of the
global numpy stream afterwards, and whether / when it raised -- is committed as tests/golden/golden_aligner_ref.npz;
tests/test_host_logic.py checks oracle.Aligner (the restatement) and the batched B200 Aligner against it.

    python tests/golden/make_golden_aligner_ref.py        # needs /root/reference; not run on the GPU box
"""
import importlib.util
import logging
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from fakes import FakeSequential  # noqa: E402

REF = "/root/reference/src/or_pcd"


ICP_CALLS = []      # what the reference's refine_registration handed to registration_icp


def make_open3d_stub():
    """`align(refine_registration=False)` and `transfrom` never touch Open3D; `refine_registration` calls
    registration_icp, which the stub answers with the oracle's restatement (oracle.icp) and records."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
    import oracle as O

    o3d = types.ModuleType("open3d")
    pipelines, registration = types.ModuleType("open3d.pipelines"), types.ModuleType("open3d.pipelines.registration")

    class TransformationEstimationPointToPoint:
        mode = "PointToPoint"

    class TransformationEstimationPointToPlane:
        mode = "PointToPlane"

    class ICPConvergenceCriteria:
        def __init__(self, relative_fitness=1e-6, relative_rmse=1e-6, max_iteration=30):
            self.relative_fitness, self.relative_rmse, self.max_iteration = relative_fitness, relative_rmse, max_iteration

    class Result:
        pass

    def registration_icp(source, target, max_correspondence_distance, init, estimation_method, criteria):
        ICP_CALLS.append(dict(max_dist=max_correspondence_distance, init=np.array(init), mode=estimation_method.mode,
                              max_iteration=criteria.max_iteration))
        r = O.icp(source, target, max_correspondence_distance, criteria.max_iteration, init, estimation_method.mode)
        out = Result()
        out.transformation, out.inlier_rmse, out.fitness = r["T"], r["rmse"], r["fitness"]
        return out

    registration.TransformationEstimationPointToPoint = TransformationEstimationPointToPoint
    registration.TransformationEstimationPointToPlane = TransformationEstimationPointToPlane
    registration.ICPConvergenceCriteria = ICPConvergenceCriteria
    registration.registration_icp = registration_icp
    pipelines.registration = registration
    o3d.pipelines = pipelines
    return o3d


def load_reference_aligner():
    """The reference's Aligner, Preprocessor, RadiusScaler and RandomDownsampler classes (unmodified source files)."""
    for name in ["or_pcd", "or_pcd.utils", "or_pcd.Preprocessor", "or_pcd.Preprocessor.Scalers",
                 "or_pcd.Preprocessor.Downsamplers", "or_pcd.Optimizer", "or_pcd.Aligner", "or_pcd.Visualizer"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["open3d"] = make_open3d_stub()          # only refine_registration reaches it (oracle-backed ICP)
    lf = types.ModuleType("or_pcd.utils.logger_factory")

    class LoggerFactory:
        @staticmethod
        def get_logger(log_name=None, log_on_file=False, **kw):          # no log files next to the golden vectors
            lg = logging.getLogger("ref." + str(log_name))
            lg.setLevel(logging.CRITICAL)
            return lg

    lf.LoggerFactory = LoggerFactory
    sys.modules["or_pcd.utils.logger_factory"] = lf
    vis = types.ModuleType("or_pcd.Visualizer.Visualizer")
    vis.visualize_point_clouds = vis.draw_registration_result = lambda *a, **k: None
    sys.modules["or_pcd.Visualizer.Visualizer"] = vis
    cc = types.ModuleType("or_pcd.utils.create_cloud")
    cc.create_cloud = lambda points: points
    sys.modules["or_pcd.utils.create_cloud"] = cc

    def load(mod, rel):
        spec = importlib.util.spec_from_file_location(mod, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[mod] = m
        spec.loader.exec_module(m)
        return m

    load("or_pcd.utils.constants", "utils/constants.py")
    load("or_pcd.Preprocessor.iProcessBlock", "Preprocessor/iProcessBlock.py")
    load("or_pcd.Preprocessor.Scalers.BaseScaler", "Preprocessor/Scalers/BaseScaler.py")
    rs = load("or_pcd.Preprocessor.Scalers.radiusScaler", "Preprocessor/Scalers/radiusScaler.py")
    sys.modules["or_pcd.Preprocessor.Scalers"].RadiusScaler = rs.RadiusScaler       # `from ...Scalers import RadiusScaler`
    rd = load("or_pcd.Preprocessor.Downsamplers.randomDownsampler", "Preprocessor/Downsamplers/randomDownsampler.py")
    pp = load("or_pcd.Preprocessor.preprocessor", "Preprocessor/preprocessor.py")
    load("or_pcd.Optimizer.iOptimizer", "Optimizer/iOptimizer.py")
    al = load("or_pcd.Aligner.Aligner", "Aligner/Aligner.py")
    return al.Aligner, pp.Preprocessor, rs.RadiusScaler, rd.RandomDownsampler


def cloud(seed, n):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, 3)) * np.array([1.0, 0.7, 0.4]) + rng.standard_normal(3)


# (name, seed of the numpy stream, attempts, sample size, fail_above as a multiple of the target's z spread or None)
CASES = [("plain", 3, 5, 400, None), ("more_starts", 11, 9, 250, None), ("fails_late", 3, 5, 400, 1.1),
         ("fails_at_once", 3, 5, 400, 0.5)]


def main():
    import contextlib
    import io

    Aligner, Preprocessor, RadiusScaler, RandomDownsampler = load_reference_aligner()
    src = cloud(10, 900)
    tgt = (cloud(10, 900) * np.array([1.0, 1.31, 0.9]))[::-1].copy()
    out = {"source": src, "target": tgt, "cases": np.array([c[0] for c in CASES])}
    for name, seed, attempts, size, fail in CASES:
        fail_above = np.inf
        if fail is not None:          # the z spread of the preprocessed target of this very run (same stream position)
            np.random.seed(seed)
            Preprocessor([RandomDownsampler(size)]).preprocess(src)
            z = Preprocessor([RandomDownsampler(size)]).preprocess(tgt).std(0)[2]
            fail_above = z * fail
        np.random.seed(seed)
        opt = FakeSequential(fail_above)
        with contextlib.redirect_stdout(io.StringIO()):       # the reference prints "EPS ..." in its constructor
            al = Aligner(Preprocessor([RandomDownsampler(size)]), Preprocessor([RandomDownsampler(size)]), opt,
                         attempts=attempts)
        raised = 0
        T, metric, scale, errors = np.full((4, 4), np.nan), np.nan, np.full((1, 3), np.nan), []
        try:
            T, metric, scale, errors = al.align(src, tgt, refine_registration=False)
        except ValueError:
            raised = 1
        st = np.random.get_state()
        out.update({f"{name}_seed": seed, f"{name}_attempts": attempts, f"{name}_size": size,
                    f"{name}_fail_above": fail_above, f"{name}_raised": raised, f"{name}_T": np.asarray(T),
                    f"{name}_metric": metric, f"{name}_scale": np.asarray(scale), f"{name}_errors": np.asarray(errors),
                    f"{name}_calls": opt.calls, f"{name}_rng_keys": st[1].copy(), f"{name}_rng_pos": st[2]})
        print(name, "raised" if raised else f"metric {metric:.6g} scale {np.asarray(scale).ravel()} "
              f"errors {len(errors)}", "optimizer calls", opt.calls)
        if name == "plain":
            # Aligner.transfrom (Aligner.py:367-394) of the raw source with the result of this run, and
            # Aligner.refine_registration (Aligner.py:319-364) called directly: what reaches registration_icp and what comes back
            out["plain_transformed_source"] = al.transfrom(src)
            s2 = RadiusScaler().process(cloud(4, 500))
            t2 = RadiusScaler().process(cloud(4, 520)[::-1].copy() * np.array([1.0, 1.05, 1.0]))
            a = 0.2
            init = np.eye(4)
            init[:3, :3] = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
            init[:3, 3] = [0.02, -0.01, 0.03]
            T2, rmse2 = al.refine_registration(s2, t2, init, icp_type="PointToPoint")
            call = ICP_CALLS[-1]
            out.update({"refine_source": s2, "refine_target": t2, "refine_init": init, "refine_T": np.asarray(T2),
                        "refine_rmse": rmse2, "refine_init_passed_unchanged": int(np.array_equal(call["init"], init)),
                        "refine_max_dist": call["max_dist"], "refine_max_iteration": call["max_iteration"]})
            raised_plane = 0
            try:      # the default icp_type of align(refine_registration=True): no target normals -> Open3D raises
                al.refine_registration(s2, t2, init, icp_type="PointToPlane")
            except RuntimeError:
                raised_plane = 1
            out["refine_point_to_plane_without_normals_raises"] = raised_plane
            print("refine: rmse", rmse2, "max_dist", call["max_dist"], "max_iteration", call["max_iteration"],
                  "init unchanged", out["refine_init_passed_unchanged"], "PointToPlane raises", raised_plane)
    np.savez_compressed(os.path.join(HERE, "golden_aligner_ref.npz"), **out)


if __name__ == "__main__":
    main()
