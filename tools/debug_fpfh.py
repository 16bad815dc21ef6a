import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-scale-pointcloud-registration_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O
from or_pcd_b200 import _lib, ops
np.random.seed(0)
s = O.Preprocessor([O.RandomDownsampler(5000), O.SOR()]).preprocess(O.load_scan("ArmadilloBack_330"))
t = O.Preprocessor([O.RandomDownsampler(5000), O.SOR()]).preprocess(O.load_scan("ArmadilloBack_0"))
for name, cloud in (("s", s), ("t", t)):
    on = O.normals(cloud, 20, 0.1)
    nrm = on["normal"]
    ref, ref_sp = O.fpfh(cloud, nrm, 20, 0.1, want_spfh=True)
    d = _lib.to_device(cloud[None])
    idx, d2 = ops.knn_exact(d, d, 20)
    gn = ops.normals_cov_from_knn(d, idx, d2, radius=0.1, want_normals=True, want_gicp_cov=False)["normal"]
    gnh = gn[0].cpu().numpy()
    print(name, "normal max diff", np.abs(gnh - nrm).max(), "rows>1e-9:", (np.abs(gnh - nrm).max(1) > 1e-9).sum())
    for label, nn in (("oracle normals", _lib.to_device(nrm[None])), ("device normals", gn)):
        out, sp = ops.fpfh(d, nn, idx, d2, 20, 0.1, want_spfh=True)
        out, sp = out[0].cpu().numpy(), sp[0].cpu().numpy()
        bad_sp = np.flatnonzero(np.abs(sp - ref_sp).max(1) > 1e-9)
        bad = np.flatnonzero(np.abs(out - ref).max(1) > 1e-7)
        print(name, label, "spfh rows differing:", len(bad_sp), "fpfh rows differing:", len(bad), "max fpfh diff", np.abs(out - ref).max(),
              "max fpfh diff on rows without spfh-neighbour issue", np.abs(out - ref)[np.setdiff1d(np.arange(len(ref)), bad)].max())
        for r in bad_sp[:4]:
            j = np.flatnonzero(np.abs(sp[r] - ref_sp[r]) > 1e-9)
            print("   row", r, "bins", j, "gpu", sp[r][j], "ref", ref_sp[r][j], "cnt", on["nn_cnt"][r])
