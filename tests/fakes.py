"""Deterministic stand-in optimizers for the host-logic tests (no GPU, no Open3D):
the metric is a cheap smooth function of the two clouds so that the Aligner's
accept / poll / RNG-replay logic can be compared between the sequential reference
control flow (oracle.Aligner) and the batched B200 Aligner."""
import numpy as np


def fake_metric(source, target):
    ms, mt = source.mean(0), target.mean(0)
    ss, st = source.std(0), target.std(0)
    return float(np.abs(ss - st).sum() + 0.05 * np.linalg.norm(ms - mt))


def fake_transform(source, target):
    T = np.eye(4)
    T[:3, 3] = target.mean(0) - source.mean(0)
    return T


class FakeSequential:
    """Reference-style optimizer: optimize() only.  Raises ValueError like
    generalizedICP.py:77-81 when the target's z-extent exceeds ``fail_above``."""

    def __init__(self, fail_above=np.inf):
        self.fail_above = fail_above
        self.calls = 0

    def optimize(self, source, target, **kw):
        self.calls += 1
        if target.std(0)[2] > self.fail_above:
            raise ValueError("loss = 0")
        return fake_transform(source, target), fake_metric(source, target)


class FakeBatched(FakeSequential):
    """Same arithmetic with the optimize_batch face of the B200 optimizers."""

    def optimize_batch(self, sources, targets, unit_src, unit_tgt, rigid_base=None, **kw):
        from or_pcd_b200.Optimizer.iOptimizer import BatchResult

        sources, targets = np.asarray(sources), np.asarray(targets)
        n = len(unit_src)
        T = np.empty((n, 4, 4)); rmse = np.empty(n); status = np.zeros(n, dtype=np.int32)
        for u in range(n):
            s, t = sources[unit_src[u]], targets[unit_tgt[u]]
            self.calls += 1
            T[u] = fake_transform(s, t)
            rmse[u] = fake_metric(s, t)
            if t.std(0)[2] > self.fail_above:
                status[u] = 1
                rmse[u] = 0.0
        return BatchResult(T, rmse, np.ones(n), np.zeros(n, dtype=np.int64), status, [])
