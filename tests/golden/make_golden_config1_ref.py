"""BASELINE config 1 END TO END with the REFERENCE'S OWN classes: `Aligner`, `Preprocessor`, `RandomDownsampler`, `SOR`,
`RadiusScaler` and `GeneralizedICP` of /root/reference/src/or_pcd, all unmodified, exactly as the reference's README
wires them --

    np.random.seed(0)
    aligner = Aligner(Preprocessor([RandomDownsampler(5000), SOR()]), Preprocessor([RandomDownsampler(5000), SOR()]),
                      GeneralizedICP())
    T, rmse, scale, errors = aligner.align(ArmadilloBack_330, ArmadilloBack_0, refine_registration=False)

-- over an `open3d` stub in which only the two numerical entry points this pipeline reaches
(`remove_statistical_outlier`, `registration_generalized_icp`) are answered by the oracle.  870 registrations, a few
minutes on 8 cores.  The result must equal tests/golden/golden_v1.npz (the run of oracle.Aligner that every GPU test of
`Aligner.align()` is held to) bit for bit: transformation, rmse, the chosen scale (1, 1.15, 1), the error list, the
number of optimizer calls and the state of the numpy stream.  Output: tests/golden/golden_config1_ref.npz.

    python tests/golden/make_golden_config1_ref.py [1|2|3]   # needs /root/reference; not run on the GPU box

Config 2 (the same pair through `FastGlobalOptimizer`, the stub also answering `estimate_normals`,
`compute_fpfh_feature` and `registration_fgr_based_on_feature_matching`) and config 3 (the OnHead pair, 64 starts) are
checked the same way against golden_fgr_back.npz / golden_onhead64.npz and written to golden_config{2,3}_ref.npz.
"""
import contextlib
import importlib.util
import io
import logging
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402

REF = "/root/reference/src/or_pcd"
COUNT = {"gicp": 0, "sor": 0}


def make_open3d_stub():
    o3d = types.ModuleType("open3d")
    geometry, utility = types.ModuleType("open3d.geometry"), types.ModuleType("open3d.utility")
    pipelines, registration = types.ModuleType("open3d.pipelines"), types.ModuleType("open3d.pipelines.registration")

    class PointCloud:
        def __init__(self):
            self.points, self.normals = None, None

        def estimate_normals(self, search_param):
            self.normals = O.normals(self.points, search_param.max_nn, search_param.radius)["normal"]

        def remove_statistical_outlier(self, nb_neighbors, std_ratio):
            COUNT["sor"] += 1
            mask, _, _ = O.sor(self.points, nb_neighbors, std_ratio)
            ind = np.nonzero(mask)[0]
            out = PointCloud()
            out.points = self.points[ind]
            return out, list(ind)

    class Result:
        pass

    class ICPConvergenceCriteria:
        def __init__(self, relative_fitness=1e-6, relative_rmse=1e-6, max_iteration=30):
            self.max_iteration = max_iteration

    class TransformationEstimationForGeneralizedICP:
        pass

    class KDTreeSearchParamHybrid:
        def __init__(self, radius, max_nn):
            self.radius, self.max_nn = radius, max_nn

    class Feature:
        def __init__(self, data):
            self.data = data

    class FastGlobalRegistrationOption:
        def __init__(self, division_factor=1.4, use_absolute_scale=False, decrease_mu=False,
                     maximum_correspondence_distance=0.025, iteration_number=64, tuple_scale=0.95,
                     maximum_tuple_count=1000, tuple_test=True):
            self.__dict__.update(division_factor=division_factor, decrease_mu=decrease_mu,
                                 maximum_correspondence_distance=maximum_correspondence_distance,
                                 iteration_number=iteration_number, tuple_scale=tuple_scale,
                                 maximum_tuple_count=maximum_tuple_count, tuple_test=tuple_test)

    def compute_fpfh_feature(cloud, search_param):
        return Feature(O.fpfh(cloud.points, cloud.normals, search_param.max_nn, search_param.radius))

    def registration_fgr_based_on_feature_matching(source, target, source_feature, target_feature, option):
        COUNT["fgr"] = COUNT.get("fgr", 0) + 1
        r = O.fgr(source.points, target.points, source_feature.data, target_feature.data, option.division_factor,
                  option.tuple_scale, option.maximum_correspondence_distance, option.iteration_number, option.decrease_mu,
                  option.maximum_tuple_count, option.tuple_test, key=O.tuple_key(0, source.points, target.points))
        out = Result()
        out.transformation, out.inlier_rmse, out.fitness = r["T"], r["rmse"], r["fitness"]
        out.correspondence_set = np.zeros((r["n_corr"], 2), dtype=np.int32)
        return out

    def registration_generalized_icp(source, target, max_correspondence_distance, init, estimation_method, criteria):
        COUNT["gicp"] += 1
        assert np.array_equal(init, np.eye(4))
        r = O.gicp(source.points, target.points, max_correspondence_distance, criteria.max_iteration)
        out = Result()
        out.transformation, out.inlier_rmse, out.fitness = r["T"], r["rmse"], r["fitness"]
        return out

    geometry.PointCloud, geometry.KDTreeSearchParamHybrid = PointCloud, KDTreeSearchParamHybrid
    registration.Feature, registration.RegistrationResult = Feature, Result
    registration.FastGlobalRegistrationOption = FastGlobalRegistrationOption
    registration.compute_fpfh_feature = compute_fpfh_feature
    registration.registration_fgr_based_on_feature_matching = registration_fgr_based_on_feature_matching
    utility.Vector3dVector = lambda a: np.array(a, dtype=np.float64)
    registration.ICPConvergenceCriteria = ICPConvergenceCriteria
    registration.TransformationEstimationForGeneralizedICP = TransformationEstimationForGeneralizedICP
    registration.registration_generalized_icp = registration_generalized_icp
    pipelines.registration = registration
    o3d.geometry, o3d.utility, o3d.pipelines = geometry, utility, pipelines
    return o3d


def load_reference():
    for name in ["or_pcd", "or_pcd.utils", "or_pcd.Preprocessor", "or_pcd.Preprocessor.Scalers",
                 "or_pcd.Preprocessor.Downsamplers", "or_pcd.Preprocessor.Outliers", "or_pcd.Optimizer", "or_pcd.Aligner",
                 "or_pcd.Visualizer"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["open3d"] = make_open3d_stub()
    lf = types.ModuleType("or_pcd.utils.logger_factory")

    class LoggerFactory:
        @staticmethod
        def get_logger(log_name=None, log_on_file=False, **kw):
            lg = logging.getLogger("ref." + str(log_name))
            lg.setLevel(logging.CRITICAL + 1)
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
    sys.modules["or_pcd.Preprocessor.Scalers"].RadiusScaler = rs.RadiusScaler
    rd = load("or_pcd.Preprocessor.Downsamplers.randomDownsampler", "Preprocessor/Downsamplers/randomDownsampler.py")
    so = load("or_pcd.Preprocessor.Outliers.sor", "Preprocessor/Outliers/sor.py")
    pp = load("or_pcd.Preprocessor.preprocessor", "Preprocessor/preprocessor.py")
    load("or_pcd.Optimizer.iOptimizer", "Optimizer/iOptimizer.py")
    gi = load("or_pcd.Optimizer.generalizedICP", "Optimizer/generalizedICP.py")
    fg = load("or_pcd.Optimizer.fastGlobalOptimizer", "Optimizer/fastGlobalOptimizer.py")
    al = load("or_pcd.Aligner.Aligner", "Aligner/Aligner.py")
    return al.Aligner, pp.Preprocessor, rd.RandomDownsampler, so.SOR, gi.GeneralizedICP, fg.FastGlobalOptimizer


# config -> (source scan, target scan, optimizer, attempts, golden file of the oracle.Aligner run, key prefix, output)
CONFIGS = {
    "1": ("ArmadilloBack_330", "ArmadilloBack_0", "gicp", 30, "golden_v1.npz", "align_back_", "golden_config1_ref.npz"),
    "2": ("ArmadilloBack_330", "ArmadilloBack_0", "fgr", 30, "golden_fgr_back.npz", "", "golden_config2_ref.npz"),
    "3": ("ArmadilloOnHeadMultiple_150", "ArmadilloOnHeadMultiple_240", "gicp", 64, "golden_onhead64.npz", "",
          "golden_config3_ref.npz"),
}


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "1"
    s_name, t_name, which, attempts, gold_file, pre, out_file = CONFIGS[cfg]
    Aligner, Preprocessor, RandomDownsampler, SOR, GeneralizedICP, FastGlobalOptimizer = load_reference()
    src, tgt = O.load_scan(s_name), O.load_scan(t_name)
    np.random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):           # the reference prints "EPS ..." in its constructor
        opt = GeneralizedICP() if which == "gicp" else FastGlobalOptimizer()
        al = Aligner(Preprocessor([RandomDownsampler(5000), SOR()]), Preprocessor([RandomDownsampler(5000), SOR()]),
                     opt, attempts=attempts, visualize_intermediate_steps=False)
    t0 = time.perf_counter()
    T, rmse, scale, errors = al.align(src, tgt, refine_registration=False)
    wall = time.perf_counter() - t0
    rng_after = np.random.get_state()[1][:8].copy()
    calls = COUNT["gicp"] if which == "gicp" else COUNT.get("fgr", 0)
    print(f"config {cfg}: reference classes over the stub: {calls} registrations, {COUNT['sor']} SOR calls, {wall:.1f} s, "
          f"scale {np.asarray(scale).ravel()}, rmse {rmse!r}")
    g = np.load(os.path.join(HERE, gold_file))
    key = (lambda k: pre + k) if pre else (lambda k: {"metric": "metric"}.get(k, k))
    same = dict(T=np.array_equal(T, g[key("T")]), rmse=rmse == float(g[key("metric")]),
                scale=np.array_equal(scale, g[key("scale")]),
                errors=np.array_equal(np.asarray(errors), g[key("errors")]),
                calls=calls == int(g[key("calls")]), rng=np.array_equal(rng_after, g[key("rng_after")]))
    print(f"equal to {gold_file} (the oracle.Aligner run):", same)
    assert all(same.values()), same
    np.savez_compressed(os.path.join(HERE, out_file), T=np.asarray(T), rmse=np.float64(rmse),
                        scale=np.asarray(scale), errors=np.asarray(errors), calls=np.int64(calls),
                        sor_calls=np.int64(COUNT["sor"]), rng_after=rng_after, wall_s=np.float64(wall))


if __name__ == "__main__":
    main()
