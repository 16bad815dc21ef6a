"""Golden vectors of the REFERENCE'S OWN preprocessing blocks that wrap Open3D -- `SOR.process`
(/root/reference/src/or_pcd/Preprocessor/Outliers/sor.py:51-79) and `VoxelDownsampler.process` with its compass search
over the voxel size (Preprocessor/Downsamplers/voxelDownsampler.py:78-126) -- run in the authoring container over an
`open3d` stub whose two numerical entry points (`remove_statistical_outlier`, `voxel_down_sample`) are the oracle's.

Pinned with the reference's own code executing: the arguments that reach Open3D (nb_neighbors = 64, std_ratio = 2 by
default), that `SOR.process` returns the inlier POINTS in their original order, and the whole probe sequence of the
voxel-size search (every voxel size it tries, in order, the accept rule `<`, the halving of delta, the clamp at
min_voxel_size) together with the cloud it finally returns.  The arithmetic under the stub is the oracle's.
Output: tests/golden/golden_blocks_ref.npz.

    python tests/golden/make_golden_blocks_ref.py        # needs /root/reference; not run on the GPU box
"""
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
CALLS = []


def make_open3d_stub():
    o3d = types.ModuleType("open3d")
    geometry, utility = types.ModuleType("open3d.geometry"), types.ModuleType("open3d.utility")

    class PointCloud:
        def __init__(self):
            self.points = None

        def remove_statistical_outlier(self, nb_neighbors, std_ratio):
            CALLS.append(("remove_statistical_outlier", nb_neighbors, std_ratio))
            mask, _, _ = O.sor(self.points, nb_neighbors, std_ratio)
            ind = np.nonzero(mask)[0]
            out = PointCloud()
            out.points = self.points[ind]
            return out, list(ind)

        def voxel_down_sample(self, voxel_size):
            pts, _ = O.voxel_downsample(self.points, voxel_size)
            CALLS.append(("voxel_down_sample", voxel_size, pts.shape[0]))
            out = PointCloud()
            out.points = pts
            return out

    geometry.PointCloud = PointCloud
    utility.Vector3dVector = lambda a: np.array(a, dtype=np.float64)
    o3d.geometry, o3d.utility = geometry, utility
    return o3d


def load_reference_blocks():
    for name in ["or_pcd", "or_pcd.utils", "or_pcd.Preprocessor", "or_pcd.Preprocessor.Outliers",
                 "or_pcd.Preprocessor.Downsamplers"]:
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
    load("or_pcd.Preprocessor.iProcessBlock", "Preprocessor/iProcessBlock.py")
    s = load("or_pcd.Preprocessor.Outliers.sor", "Preprocessor/Outliers/sor.py")
    v = load("or_pcd.Preprocessor.Downsamplers.voxelDownsampler", "Preprocessor/Downsamplers/voxelDownsampler.py")
    return s.SOR, v.VoxelDownsampler


def main():
    SOR, VoxelDownsampler = load_reference_blocks()
    out = {}
    for name, n, seed in (("ArmadilloBack_30", 3000, 4), ("ArmadilloOnface_270", 2500, 9)):
        scan = O.load_scan(name)
        cloud = O.RadiusScaler().process(scan[np.random.default_rng(seed).choice(len(scan), n, replace=False)])
        out[f"{name}_cloud"] = cloud
        for tag, kw in (("default", {}), ("k16", dict(nb_neighbours=16, std_ratio=1.0))):
            CALLS.clear()
            kept = SOR(**kw).process(cloud)
            call = CALLS[0]
            out[f"sor_{name}_{tag}_kept"] = kept
            out[f"sor_{name}_{tag}_args"] = np.array([call[1], call[2]], dtype=np.float64)
            print("SOR", name, tag, "args", call[1:], "kept", kept.shape[0], "of", n)
        for tag, target, kw in (("t800", 800, {}), ("t300_coarse", 300, dict(base_voxel_size=0.05, delta=0.02, eps=0.001)),
                                ("t2900_clamped", 2900, dict(base_voxel_size=0.002, min_voxel_size=0.0015, delta=0.004))):
            CALLS.clear()
            pts = VoxelDownsampler(target, **kw).process(cloud)
            probes = np.array([(c[1], c[2]) for c in CALLS if c[0] == "voxel_down_sample"])
            out[f"voxel_{name}_{tag}_probes"] = probes              # (voxel size, points) of every probe, in order
            out[f"voxel_{name}_{tag}_points"] = pts
            print("Voxel", name, tag, "probes", len(probes), "final points", pts.shape[0], "last size", probes[-1, 0])
    np.savez_compressed(os.path.join(HERE, "golden_blocks_ref.npz"), **out)


if __name__ == "__main__":
    main()
