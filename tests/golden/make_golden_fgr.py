"""Golden run of BASELINE config 2 (Back pair, FastGlobalOptimizer, attempts=30) from the CPU oracle in
the reference's literal feature mode.  About 900 sequential FGR registrations (~1 s each on 8 cores).
Writes tests/golden/golden_fgr_back.npz.   python tests/golden/make_golden_fgr.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import oracle as O  # noqa: E402


def main():
    np.random.seed(0)
    al = O.Aligner(O.Preprocessor([O.RandomDownsampler(5000), O.SOR()]),
                   O.Preprocessor([O.RandomDownsampler(5000), O.SOR()]),
                   O.FastGlobalOptimizer(feature_mode="reference_compat"))
    T, m, scale, errors = al.align(O.load_scan("ArmadilloBack_330"), O.load_scan("ArmadilloBack_0"))
    out = dict(T=T, metric=np.float64(m), scale=scale, errors=np.array(errors),
               log_scale=np.array([d["scale"].ravel() for d in al.decision_log]),
               log_metric=np.array([d["metric"] for d in al.decision_log]),
               log_accepted=np.array([d["accepted"] for d in al.decision_log]),
               calls=np.int64(al._optimizer.calls), rng_after=np.random.get_state()[1][:8].copy())
    np.savez_compressed(os.path.join(HERE, "golden_fgr_back.npz"), **out)
    print("calls", al._optimizer.calls, "scale", scale, "rmse", m)


if __name__ == "__main__":
    main()
