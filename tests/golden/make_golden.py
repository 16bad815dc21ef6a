"""Generate the committed golden vectors under tests/golden/.

Run in the authoring container (needs /root/reference for the pure-numpy reference blocks):
    python tests/golden/make_golden.py
Sources of truth:
  * RadiusScaler / RandomDownsampler: the REFERENCE's own classes, imported from
    /root/reference/src (they do not import open3d);
  * everything that Open3D computes in the reference: the CPU oracle (oracle/), a
    restatement -- "parity unpinned", see oracle/orpcd_oracle.c.
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402

REF = "/root/reference/src/or_pcd"


def _load_reference_blocks():
    """Import the reference's RadiusScaler and RandomDownsampler without triggering `import open3d`
    (package __init__ files pull in Open3D-backed siblings, so the modules are loaded by path)."""
    import logging

    pkgs = ["or_pcd", "or_pcd.utils", "or_pcd.Preprocessor", "or_pcd.Preprocessor.Scalers",
            "or_pcd.Preprocessor.Downsamplers"]
    for name in pkgs:
        sys.modules.setdefault(name, types.ModuleType(name))
    lf = types.ModuleType("or_pcd.utils.logger_factory")

    class LoggerFactory:
        @staticmethod
        def get_logger(log_name, log_on_file=False):
            return logging.getLogger("ref." + log_name)

    lf.LoggerFactory = LoggerFactory
    sys.modules["or_pcd.utils.logger_factory"] = lf

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
    rd = load("or_pcd.Preprocessor.Downsamplers.randomDownsampler", "Preprocessor/Downsamplers/randomDownsampler.py")
    return rs.RadiusScaler, rd.RandomDownsampler


def main():
    RefRadiusScaler, RefRandomDownsampler = _load_reference_blocks()
    out = {}
    # ---- reference numpy blocks on the config-1 / config-3 scans, seed 0, README pipeline order
    for tag, (a, b) in {"back": ("ArmadilloBack_330", "ArmadilloBack_0"),
                        "onhead": ("ArmadilloOnHeadMultiple_150", "ArmadilloOnHeadMultiple_240")}.items():
        np.random.seed(0)
        pre = []
        for name in (a, b):
            raw = O.load_scan(name)
            sc = RefRadiusScaler()
            scaled = sc.process(raw)
            ds = RefRandomDownsampler(5000)
            state_before = np.random.get_state()[1][:4].copy()
            sampled = ds.process(scaled)
            # recover the drawn indices by replaying the same call
            mask, avg, stats = O.sor(sampled, 64, 2.0)
            out[f"{tag}_{name}_mean"] = sc.mean
            out[f"{tag}_{name}_scale"] = np.float64(sc.scale)
            out[f"{tag}_{name}_sampled_head"] = sampled[:8]
            out[f"{tag}_{name}_sampled_sum"] = sampled.sum(0)
            out[f"{tag}_{name}_sor_mask"] = np.packbits(mask)
            out[f"{tag}_{name}_sor_stats"] = np.array(stats)
            out[f"{tag}_{name}_n_after"] = np.int64(mask.sum())
            pre.append(np.ascontiguousarray(sampled[mask]))
        out[f"{tag}_rng_after_pre"] = np.random.get_state()[1][:8].copy()
        s, t = pre
        # ---- oracle GICP: identity start and three random starts at scale (1,1,1)
        res = O.gicp(s, t, 0.5, 100, want_trace=True)
        out[f"{tag}_gicp_identity_T"] = res["T"]
        out[f"{tag}_gicp_identity_stats"] = np.array([res["fitness"], res["rmse"], res["iters"], res["n_corr"]])
        out[f"{tag}_gicp_identity_trace"] = res["trace"]
        Ts, st = [], []
        for _ in range(3):
            r, tr = O.initialize_rotation(np.pi / 2, 0.0, 0.1)
            rr = O.gicp(np.dot(s, r) + tr, t, 0.5, 100)
            Ts.append(rr["T"]); st.append([rr["fitness"], rr["rmse"], rr["iters"], rr["n_corr"]])
        out[f"{tag}_gicp_starts_T"] = np.array(Ts)
        out[f"{tag}_gicp_starts_stats"] = np.array(st)
        nn = O.normals(s, 20, 0.0, 1e-3)
        out[f"{tag}_src_nn20_head"] = nn["nn_idx"][:64]
        out[f"{tag}_src_normals_head"] = nn["normal"][:64]
        # ---- oracle FGR (config 2): FPFH of the source, one registration per feature mode
        fs = O.fpfh(s, O.normals(s, 20, 0.1)["normal"], 20, 0.1)
        out[f"{tag}_src_fpfh_head"] = fs[:64]
        out[f"{tag}_src_fpfh_colsum"] = fs.sum(0)
        for mode in ("reference_compat", "fixed"):
            fo = O.FastGlobalOptimizer(feature_mode=mode)
            T, rm = fo.optimize(s, t)
            out[f"{tag}_fgr_{mode}_T"] = T
            out[f"{tag}_fgr_{mode}_stats"] = np.array([fo.last["fitness"], rm, fo.last["n_corr"], fo.last["n_initial"],
                                                        fo.last["n_pairs"], fo.last["trials"]])

    # ---- full align() decision log of config 1 (README pipeline, GICP, attempts=30, seed 0)
    np.random.seed(0)
    al = O.Aligner(O.Preprocessor([O.RandomDownsampler(5000), O.SOR()]),
                   O.Preprocessor([O.RandomDownsampler(5000), O.SOR()]), O.GeneralizedICP())
    T, m, scale, errors = al.align(O.load_scan("ArmadilloBack_330"), O.load_scan("ArmadilloBack_0"))
    out["align_back_T"] = T
    out["align_back_metric"] = np.float64(m)
    out["align_back_scale"] = scale
    out["align_back_errors"] = np.array(errors)
    out["align_back_log_scale"] = np.array([d["scale"].ravel() for d in al.decision_log])
    out["align_back_log_metric"] = np.array([d["metric"] for d in al.decision_log])
    out["align_back_log_accepted"] = np.array([d["accepted"] for d in al.decision_log])
    out["align_back_calls"] = np.int64(al._optimizer.calls)
    out["align_back_iterations"] = np.int64(al._optimizer.iterations)
    out["align_back_rng_after"] = np.random.get_state()[1][:8].copy()
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    print("wrote golden_v1.npz:", len(out), "arrays; align calls", al._optimizer.calls, "scale", scale, "rmse", m)


if __name__ == "__main__":
    main()
