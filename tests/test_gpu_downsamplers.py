"""Down-sampling blocks on the GPU against the oracle and the committed
fixtures (the FPS fixture was produced by the reference's own FarthestDownsampler class).  Index work
is bit-exact; voxel averages are bit-exact too (sums in index order, float64)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_downsamplers.npz")


@pytest.fixture(scope="module")
def gold():
    with np.load(GOLD) as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize("name,k,seed", [("ArmadilloBack_30", 700, 4), ("ArmadilloOnface_270", 1200, 9)])
def test_farthest_downsampler_equals_reference(oracle, gold, name, k, seed):
    from or_pcd_b200.Preprocessor import FarthestDownsampler

    cloud = oracle.RadiusScaler().process(oracle.load_scan(name))
    np.random.seed(seed)
    blk = FarthestDownsampler(k)
    out = blk.process(cloud)
    assert np.array_equal(blk.last_indices, gold[f"fps_{name}_indices"])          # the reference's own indices
    assert np.array_equal(out, cloud[gold[f"fps_{name}_indices"]])
    assert np.array_equal(np.random.get_state()[1][:4], gold[f"fps_{name}_rng_after"])


def test_farthest_point_sample_large_and_ties(oracle):
    from or_pcd_b200 import _lib, ops

    rng = np.random.default_rng(1)
    big = rng.standard_normal((20000, 3))                                         # > 16384: workspace path
    lattice = np.stack(np.meshgrid(np.arange(9.0), np.arange(8.0), np.arange(7.0), indexing="ij"), -1).reshape(-1, 3)
    for pts, first, k in ((big, 77, 300), (lattice, 0, 200), (rng.standard_normal((5, 3)), 3, 5)):
        idx = ops.farthest_point_sample(_lib.to_device(pts), first, k).cpu().numpy()
        assert np.array_equal(idx, oracle.fps(pts, first, k))                     # equal distances -> first index


@pytest.mark.parametrize("size", [0.2, 0.05, 0.01, 0.0031])
def test_voxel_downsample_bit_exact(oracle, size):
    from or_pcd_b200 import _lib, ops

    cloud = oracle.RadiusScaler().process(oracle.load_scan("ArmadilloBack_60"))
    pts, first, _ = ops.voxel_downsample(_lib.to_device(cloud), size)
    rp, rf = oracle.voxel_downsample(cloud, size)
    assert pts.shape[0] == rp.shape[0]
    assert np.array_equal(first.cpu().numpy(), rf)
    assert np.array_equal(pts.cpu().numpy(), rp)


@pytest.mark.parametrize("name", ["ArmadilloBack_30", "ArmadilloOnface_270"])
def test_voxel_downsampler_search_matches_golden(oracle, gold, name):
    from or_pcd_b200.Preprocessor import VoxelDownsampler

    cloud = oracle.RadiusScaler().process(oracle.load_scan(name))
    blk = VoxelDownsampler(2000)
    out = blk.process(cloud)
    assert np.array_equal(np.array(blk.probes), gold[f"voxel_{name}_probes"])     # same probes, same counts
    assert blk.voxel_size == float(gold[f"voxel_{name}_size"]) and out.shape[0] == int(gold[f"voxel_{name}_count"])
    assert np.array_equal(out[:16], gold[f"voxel_{name}_head"])
    assert np.allclose(out.sum(0), gold[f"voxel_{name}_sum"], rtol=1e-13)
    with pytest.raises(ValueError):
        VoxelDownsampler(0)


def test_voxel_grid_limits_are_told_apart():
    """Open3D's own failure (voxel_size * INT_MAX < extent) and this implementation's sort-word limit raise different
    messages: the second is not an Open3D error."""
    from or_pcd_b200 import _lib, ops

    rng = np.random.default_rng(0)
    cloud = rng.uniform(-50, 50, (4000, 3))
    with pytest.raises(_lib.OrpcdError, match="voxel_size is too small"):
        ops.voxel_downsample(_lib.to_device(cloud), 1e-9)
    with pytest.raises(_lib.OrpcdError, match="64-bit sort key"):
        ops.voxel_downsample(_lib.to_device(cloud), 1e-4)
    out, first, _ = ops.voxel_downsample(_lib.to_device(cloud), 5.0)
    assert 0 < out.shape[0] <= 4000



def test_blocks_equal_the_references_own_classes():
    """The B200 SOR and VoxelDownsampler blocks through the C ABI against golden vectors of the reference's OWN classes
    (tests/golden/make_golden_blocks_ref.py): inlier points bit for bit; the voxel-size search probe for probe and the
    returned voxel averages bit for bit."""
    import os

    from or_pcd_b200.Preprocessor import SOR, VoxelDownsampler

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_blocks_ref.npz"))
    for name in ("ArmadilloBack_30", "ArmadilloOnface_270"):
        cloud = g[f"{name}_cloud"]
        for tag, kw in (("default", {}), ("k16", dict(nb_neighbours=16, std_ratio=1.0))):
            assert np.array_equal(SOR(**kw).process(cloud), g[f"sor_{name}_{tag}_kept"]), (name, tag)
        for tag, target, kw in (("t800", 800, {}), ("t300_coarse", 300, dict(base_voxel_size=0.05, delta=0.02, eps=0.001)),
                                ("t2900_clamped", 2900, dict(base_voxel_size=0.002, min_voxel_size=0.0015, delta=0.004))):
            blk = VoxelDownsampler(target, **kw)
            pts = blk.process(cloud)
            assert np.array_equal(np.array(blk.probes, dtype=np.float64), g[f"voxel_{name}_{tag}_probes"]), (name, tag)
            assert np.array_equal(pts, g[f"voxel_{name}_{tag}_points"]), (name, tag)
