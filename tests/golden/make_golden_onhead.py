"""Golden run of BASELINE config 3 (OnHead pair, attempts=64) from the CPU oracle: ~2500 sequential inner
registrations, about ten minutes on 8 cores.  Writes tests/golden/golden_onhead64.npz.
Run in the authoring container:  python tests/golden/make_golden_onhead.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import oracle as O  # noqa: E402


def main():
    np.random.seed(0)
    al = O.Aligner(O.Preprocessor([O.RandomDownsampler(5000), O.SOR()]),
                   O.Preprocessor([O.RandomDownsampler(5000), O.SOR()]), O.GeneralizedICP(), attempts=64)
    T, m, scale, errors = al.align(O.load_scan("ArmadilloOnHeadMultiple_150"), O.load_scan("ArmadilloOnHeadMultiple_240"))
    out = dict(T=T, metric=np.float64(m), scale=scale, errors=np.array(errors),
               log_scale=np.array([d["scale"].ravel() for d in al.decision_log]),
               log_metric=np.array([d["metric"] for d in al.decision_log]),
               log_accepted=np.array([d["accepted"] for d in al.decision_log]),
               calls=np.int64(al._optimizer.calls), iterations=np.int64(al._optimizer.iterations),
               rng_after=np.random.get_state()[1][:8].copy())
    np.savez_compressed(os.path.join(HERE, "golden_onhead64.npz"), **out)
    print("calls", al._optimizer.calls, "scale", scale, "rmse", m)


if __name__ == "__main__":
    main()
