"""The C-ABI shared library loads on a machine without a GPU and exports every entry point
that include/orpcd_b200.h declares (no compute calls here: those are the -m gpu tests)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "orpcd_b200.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(orpcd_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = _declared_functions()
    for must in ("orpcd_knn_exact", "orpcd_sor_mask", "orpcd_normals_cov", "orpcd_gicp_batch",
                 "orpcd_gicp_batch_workspace_bytes", "orpcd_apply_scale_batch"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from or_pcd_b200 import _lib

    L = _lib.lib()
    missing = [n for n in _declared_functions() if not hasattr(L, n)]
    assert not missing, missing
    assert L.orpcd_compiled_arch() == 100          # sm_100a only
    assert L.orpcd_version() >= 100
    assert L.orpcd_error_string(0) is not None


def test_argument_validation_needs_no_gpu():
    """Invalid arguments are rejected before any CUDA call."""
    from or_pcd_b200 import _lib

    L = _lib.lib()
    p = _lib.GicpParams()                            # all zeros: invalid
    assert L.orpcd_gicp_batch_workspace_bytes(ctypes.byref(p)) == 0
    p = _lib.GicpParams(n_units=180, n_src=5000, n_tgt=5000, n_src_sets=180, n_tgt_sets=6, max_iterations=100,
                        cov_mode=1, max_correspondence_distance=0.5, relative_fitness=1e-6, relative_rmse=1e-6,
                        gicp_epsilon=1e-3)
    need = L.orpcd_gicp_batch_workspace_bytes(ctypes.byref(p))
    assert 50e6 < need < 2e9                         # sized for HBM, not for smem
    assert L.orpcd_knn_exact(None, 10, 30, None, 10, 30, 1, 0, None, None, None, 0, None) < 0      # k = 0
    assert L.orpcd_knn_exact(None, 10, 30, None, 10, 30, 1, 513, None, None, None, 0, None) < 0     # k > ORPCD_MAX_KNN


def test_product_path_refuses_to_run_without_cuda():
    """No CPU fallback: on a box without a CUDA device the B200 blocks raise."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    import numpy as np

    from or_pcd_b200 import _lib
    from or_pcd_b200.Optimizer import GeneralizedICP
    from or_pcd_b200.Preprocessor import SOR

    with pytest.raises(_lib.OrpcdError):
        GeneralizedICP().optimize(np.zeros((10, 3)), np.zeros((10, 3)))
    with pytest.raises(_lib.OrpcdError):
        SOR().process(np.zeros((100, 3)))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "multi-scale-pointcloud-registration_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in src and "liborpcd_oracle" not in src, os.path.join(dirpath, f)
