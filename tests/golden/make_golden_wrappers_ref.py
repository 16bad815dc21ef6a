"""Golden vectors of the REFERENCE'S OWN optimizer wrapper classes -- `GeneralizedICP.optimize`
(/root/reference/src/or_pcd/Optimizer/generalizedICP.py:47-83) and `FastGlobalOptimizer.get_fpfh_features / optimize`
(fastGlobalOptimizer.py:109-190) -- run in the authoring container over an `open3d` STUB whose numerical entry points
are the oracle's (oracle/oracle.py).

What this pins is everything the wrappers do around the Open3D calls, with the reference's own code executing it: which
arguments reach `registration_generalized_icp` (identity start, the correspondence distance, the iteration cap), the
transpose of the rotation block on the way out, the failure rules (`inlier_rmse == 0` -> ValueError, empty
correspondence set -> Warning), the search parameters of the normal and FPFH estimation, the option object of the FGR
call -- and the reference's feature bug: both `compute_fpfh_feature` calls receive the SOURCE cloud.  The
arithmetic under the stub is the oracle's, so this does not pin the oracle's arithmetic to Open3D (nothing offline can);
it pins the oracle's restated WRAPPERS (`oracle.GeneralizedICP`, `oracle.FastGlobalOptimizer`) and the plugin
semantics of the B200 classes to the reference's own classes.  Output: tests/golden/golden_wrappers_ref.npz.

    python tests/golden/make_golden_wrappers_ref.py        # needs /root/reference; not run on the GPU box
"""
import copy
import importlib.util
import logging
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402

REF = "/root/reference/src/or_pcd"
TRACE = []          # (call name, details) in the order the reference's code made them


def make_open3d_stub():
    o3d = types.ModuleType("open3d")
    geometry, utility = types.ModuleType("open3d.geometry"), types.ModuleType("open3d.utility")
    pipelines, registration = types.ModuleType("open3d.pipelines"), types.ModuleType("open3d.pipelines.registration")

    class PointCloud:
        def __init__(self):
            self.points, self.normals, self.tag = None, None, None

        def estimate_normals(self, search_param):
            TRACE.append(("estimate_normals", dict(tag=self.tag, radius=search_param.radius, max_nn=search_param.max_nn)))
            self.normals = O.normals(self.points, search_param.max_nn, search_param.radius)["normal"]

    class KDTreeSearchParamHybrid:
        def __init__(self, radius, max_nn):
            self.radius, self.max_nn = radius, max_nn

    class Feature:
        def __init__(self, data, tag):
            self.data, self.tag = data, tag

    class Result:
        def __init__(self, T, fitness, rmse, n_corr):
            self.transformation, self.fitness, self.inlier_rmse = T, fitness, rmse
            self.correspondence_set = np.zeros((n_corr, 2), dtype=np.int32)

    class ICPConvergenceCriteria:
        def __init__(self, relative_fitness=1e-6, relative_rmse=1e-6, max_iteration=30):
            self.relative_fitness, self.relative_rmse, self.max_iteration = relative_fitness, relative_rmse, max_iteration

    class TransformationEstimationForGeneralizedICP:
        epsilon = 1e-3

    class FastGlobalRegistrationOption:
        def __init__(self, division_factor=1.4, use_absolute_scale=False, decrease_mu=False,
                     maximum_correspondence_distance=0.025, iteration_number=64, tuple_scale=0.95,
                     maximum_tuple_count=1000, tuple_test=True):
            self.__dict__.update(division_factor=division_factor, use_absolute_scale=use_absolute_scale,
                                 decrease_mu=decrease_mu, maximum_correspondence_distance=maximum_correspondence_distance,
                                 iteration_number=iteration_number, tuple_scale=tuple_scale,
                                 maximum_tuple_count=maximum_tuple_count, tuple_test=tuple_test)

    def vector3d(a):
        return np.array(a, dtype=np.float64)

    def registration_generalized_icp(source, target, max_correspondence_distance, init, estimation_method, criteria):
        TRACE.append(("registration_generalized_icp", dict(max_dist=max_correspondence_distance, init=np.array(init),
                                                            max_iteration=criteria.max_iteration,
                                                            relative_fitness=criteria.relative_fitness,
                                                            relative_rmse=criteria.relative_rmse)))
        assert np.array_equal(init, np.eye(4))
        r = O.gicp(source.points, target.points, max_correspondence_distance, criteria.max_iteration)
        return Result(r["T"], r["fitness"], r["rmse"], r["n_corr"])

    def compute_fpfh_feature(cloud, search_param):
        TRACE.append(("compute_fpfh_feature", dict(tag=cloud.tag, radius=search_param.radius, max_nn=search_param.max_nn)))
        return Feature(O.fpfh(cloud.points, cloud.normals, search_param.max_nn, search_param.radius), cloud.tag)

    def registration_fgr_based_on_feature_matching(source, target, source_feature, target_feature, option):
        TRACE.append(("registration_fgr", dict(source_feature=source_feature.tag, target_feature=target_feature.tag,
                                               **option.__dict__)))
        r = O.fgr(source.points, target.points, source_feature.data, target_feature.data, option.division_factor,
                  option.tuple_scale, option.maximum_correspondence_distance, option.iteration_number, option.decrease_mu,
                  option.maximum_tuple_count, option.tuple_test, key=O.tuple_key(0, source.points, target.points))
        return Result(r["T"], r["fitness"], r["rmse"], r["n_corr"])

    geometry.PointCloud, geometry.KDTreeSearchParamHybrid = PointCloud, KDTreeSearchParamHybrid
    utility.Vector3dVector = vector3d
    registration.Feature, registration.RegistrationResult = Feature, Result
    registration.ICPConvergenceCriteria = ICPConvergenceCriteria
    registration.TransformationEstimationForGeneralizedICP = TransformationEstimationForGeneralizedICP
    registration.FastGlobalRegistrationOption = FastGlobalRegistrationOption
    registration.registration_generalized_icp = registration_generalized_icp
    registration.compute_fpfh_feature = compute_fpfh_feature
    registration.registration_fgr_based_on_feature_matching = registration_fgr_based_on_feature_matching
    pipelines.registration = registration
    o3d.geometry, o3d.utility, o3d.pipelines = geometry, utility, pipelines
    return o3d


def load_reference_optimizers():
    for name in ["or_pcd", "or_pcd.utils", "or_pcd.Optimizer"]:
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

    def load(mod, rel):
        spec = importlib.util.spec_from_file_location(mod, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[mod] = m
        spec.loader.exec_module(m)
        return m

    load("or_pcd.utils.constants", "utils/constants.py")
    load("or_pcd.Optimizer.iOptimizer", "Optimizer/iOptimizer.py")
    g = load("or_pcd.Optimizer.generalizedICP", "Optimizer/generalizedICP.py")
    f = load("or_pcd.Optimizer.fastGlobalOptimizer", "Optimizer/fastGlobalOptimizer.py")
    return g.GeneralizedICP, f.FastGlobalOptimizer


def tagged_pointcloud_patch():
    """The reference builds its clouds itself; the stub tags them in creation order (0 = source, 1 = target of a call)
    and a deep copy keeps the tag, so the trace can say WHICH cloud reached each Open3D call."""
    o3d = sys.modules["open3d"]
    counter = {"n": 0}
    base = o3d.geometry.PointCloud

    class Tagged(base):
        def __init__(self):
            super().__init__()
            self.tag = ("source", "target")[counter["n"] % 2]
            counter["n"] += 1

    o3d.geometry.PointCloud = Tagged
    return counter


def main():
    GeneralizedICP, FastGlobalOptimizer = load_reference_optimizers()
    counter = tagged_pointcloud_patch()
    z = np.load(os.path.join(ROOT, "multi-scale-pointcloud-registration_b200", "or_pcd_b200", "data", "armadillo_scans.npz"))
    rng = np.random.default_rng(0)
    s = O.RadiusScaler().process(z["ArmadilloBack_330"][rng.choice(len(z["ArmadilloBack_330"]), 700, replace=False)])
    t = O.RadiusScaler().process(z["ArmadilloBack_0"][rng.choice(len(z["ArmadilloBack_0"]), 650, replace=False)])
    out = {"source": s, "target": t}

    # --- GeneralizedICP: defaults, explicit parameters, non-positive parameters (fall back to the defaults), failure
    for name, kw in (("gicp_default", {}), ("gicp_params", dict(max_correspondence_distance=0.3, max_iterations=7)),
                     ("gicp_nonpositive", dict(max_correspondence_distance=-1.0, max_iterations=0))):
        TRACE.clear(); counter["n"] = 0
        T, rmse = GeneralizedICP(**kw).optimize(s, t)
        call = [d for n, d in TRACE if n == "registration_generalized_icp"][0]
        out.update({f"{name}_T": T, f"{name}_rmse": rmse, f"{name}_max_dist": call["max_dist"],
                    f"{name}_max_iteration": call["max_iteration"]})
        print(name, "rmse", rmse, "max_dist", call["max_dist"], "max_iteration", call["max_iteration"])
    raised = 0
    try:
        GeneralizedICP(max_correspondence_distance=1e-3).optimize(s[:200] + 50.0, t[:300])
    except ValueError:
        raised = 1
    out["gicp_far_apart_raises_valueerror"] = raised

    # --- FastGlobalOptimizer
    for name, kw in (("fgr_default", {}), ("fgr_params", dict(tuple_scale=0.8, iteration_number=40, decrease_mu=False,
                                                              maximum_correspondence_distance=0.3))):
        TRACE.clear(); counter["n"] = 0
        T, rmse = FastGlobalOptimizer(**kw).optimize(s, t)
        fp = [d for n, d in TRACE if n == "compute_fpfh_feature"]
        nr = [d for n, d in TRACE if n == "estimate_normals"]
        fg = [d for n, d in TRACE if n == "registration_fgr"][0]
        out.update({f"{name}_T": T, f"{name}_rmse": rmse,
                    f"{name}_fpfh_clouds": np.array([d["tag"] for d in fp]),          # the feature bug: source, source
                    f"{name}_normal_clouds": np.array([d["tag"] for d in nr]),
                    f"{name}_normal_radius": nr[0]["radius"], f"{name}_normal_knn": nr[0]["max_nn"],
                    f"{name}_fpfh_radius": fp[0]["radius"], f"{name}_fpfh_knn": fp[0]["max_nn"],
                    f"{name}_option": np.array([fg["division_factor"], fg["tuple_scale"],
                                                fg["maximum_correspondence_distance"], fg["iteration_number"],
                                                float(fg["decrease_mu"]), fg["maximum_tuple_count"],
                                                float(fg["tuple_test"]), float(fg["use_absolute_scale"])])})
        print(name, "rmse", rmse, "fpfh computed on", [d["tag"] for d in fp], "option", out[f"{name}_option"])
    raised = 0
    try:
        FastGlobalOptimizer(maximum_correspondence_distance=1e-4).optimize(s, t + 30.0)
    except Warning:
        raised = 1
    out["fgr_far_apart_raises_warning"] = raised
    print("failure rules:", out["gicp_far_apart_raises_valueerror"], out["fgr_far_apart_raises_warning"])
    np.savez_compressed(os.path.join(HERE, "golden_wrappers_ref.npz"), **out)


if __name__ == "__main__":
    main()
