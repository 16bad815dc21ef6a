"""Golden vectors for the down-sampling blocks.  FarthestDownsampler: produced by the
REFERENCE's own class (pure numpy/scipy, imported from /root/reference/src by path).  VoxelDownsampler:
the oracle (Open3D's voxel_down_sample is not installable here).
    python tests/golden/make_golden_downsamplers.py  ->  tests/golden/golden_downsamplers.npz"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import make_golden as MG  # noqa: E402
import oracle as O  # noqa: E402


def main():
    MG._load_reference_blocks()
    spec = importlib.util.spec_from_file_location("or_pcd.Preprocessor.Downsamplers.farthestDownsampler",
                                                  os.path.join(MG.REF, "Preprocessor/Downsamplers/farthestDownsampler.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = {}
    for name, k, seed in (("ArmadilloBack_30", 700, 4), ("ArmadilloOnface_270", 1200, 9)):
        cloud = O.RadiusScaler().process(O.load_scan(name))
        np.random.seed(seed)
        ref = mod.FarthestDownsampler(k).process(cloud)          # the reference itself
        np.random.seed(seed)
        first = np.random.randint(low=0, high=cloud.shape[0])
        idx = O.fps(cloud, first, k)
        assert np.array_equal(cloud[idx], ref)
        out[f"fps_{name}_indices"] = idx
        out[f"fps_{name}_rng_after"] = np.random.get_state()[1][:4].copy()
        v = O.VoxelDownsampler(2000)
        pts = v.process(cloud)
        out[f"voxel_{name}_probes"] = np.array(v.probes)
        out[f"voxel_{name}_size"] = np.float64(v.voxel_size)
        out[f"voxel_{name}_count"] = np.int64(pts.shape[0])
        out[f"voxel_{name}_sum"] = pts.sum(0)
        out[f"voxel_{name}_head"] = pts[:16]
    np.savez_compressed(os.path.join(HERE, "golden_downsamplers.npz"), **out)
    print("wrote golden_downsamplers.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
