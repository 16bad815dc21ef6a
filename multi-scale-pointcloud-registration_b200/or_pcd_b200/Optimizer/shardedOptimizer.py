"""Candidate sharding across the GPUs of one box.

The units of a poll round (scale candidate x random start) are independent inner
registrations, so they shard with no data-path collective: every rank (one process
per GPU, ``torch.distributed``) derives the same unit list from the same numpy
stream, runs the units ``u = rank (mod world)`` on its own GPU and contributes one
fixed-size record per unit -- {rmse, fitness, iterations, status, T[16]} = 20
float64 = 160 bytes -- to a single all-gather (NCCL over NVLink on the GPU box,
gloo in the CPU tests).  Every rank then holds the full result and replays the
Aligner's accept logic identically; no decision broadcast is needed.

The reference has no counterpart (single process; Aligner/Aligner.py:164-204 runs
the starts one after another)."""
import inspect
from typing import Optional, Sequence, Tuple

import numpy as np

from .iOptimizer import BatchResult, IOptimizer

RECORD_F64 = 20   # rmse, fitness, iterations, status, 16 transform entries


def unit_shard(n_units: int, rank: int, world: int) -> np.ndarray:
    """Indices of the units rank ``rank`` runs: round-robin, so every rank gets the same
    mix of scale candidates (units are ordered candidate-major) and of easy/hard starts."""
    return np.arange(rank, n_units, world, dtype=np.int64)


def pack_records(res: BatchResult, n_slots: int) -> np.ndarray:
    """(n_slots, 20) float64 records; unused slots are NaN-filled with status -1."""
    n = len(res.rmse)
    rec = np.full((n_slots, RECORD_F64), np.nan, dtype=np.float64)
    rec[:, 3] = -1.0
    rec[:n, 0] = res.rmse
    rec[:n, 1] = res.fitness
    rec[:n, 2] = res.iterations
    rec[:n, 3] = res.status
    rec[:n, 4:] = np.asarray(res.transforms, dtype=np.float64).reshape(n, 16)
    return rec


def unpack_records(all_rec: np.ndarray, n_units: int, world: int) -> BatchResult:
    """Inverse of the round-robin shard: all_rec is (world, n_slots, 20)."""
    T = np.empty((n_units, 4, 4), dtype=np.float64)
    rmse = np.empty(n_units, dtype=np.float64)
    fit = np.empty(n_units, dtype=np.float64)
    iters = np.empty(n_units, dtype=np.int64)
    status = np.empty(n_units, dtype=np.int32)
    for r in range(world):
        mine = unit_shard(n_units, r, world)
        rec = all_rec[r, : len(mine)]
        if (rec[:, 3] < 0).any():
            raise RuntimeError(f"rank {r} contributed fewer records than its shard holds")
        rmse[mine] = rec[:, 0]
        fit[mine] = rec[:, 1]
        iters[mine] = rec[:, 2].astype(np.int64)
        status[mine] = rec[:, 3].astype(np.int32)
        T[mine] = rec[:, 4:].reshape(-1, 4, 4)
    return BatchResult(T, rmse, fit, iters, status, [])


class ShardedOptimizer(IOptimizer):
    """Wraps a batched optimizer; ``optimize_batch`` runs this rank's share of the units and
    all-gathers the records.  With no process group (or world size 1) it is a pass-through."""

    def __init__(self, inner: IOptimizer, group=None):
        if not hasattr(inner, "optimize_batch"):
            raise TypeError("ShardedOptimizer needs an optimizer with optimize_batch")
        self._inner = inner
        self._group = group
        self._takes_rigid_base = "rigid_base" in inspect.signature(inner.optimize_batch).parameters
        # start clouds built on the GPU by the Aligner (CUDA tensors) are passed through when the inner optimizer takes them
        self.accepts_device_sources = bool(getattr(inner, "accepts_device_sources", False))
        self.last_local_units = 0

    def _dist(self):
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist, dist.get_rank(self._group), dist.get_world_size(self._group)
        return None, 0, 1

    def optimize(self, source: np.ndarray, target: np.ndarray, **kwargs) -> Tuple[np.ndarray, float]:
        return self._inner.optimize(source, target, **kwargs)

    def optimize_batch(self, sources, targets, unit_src: Sequence[int], unit_tgt: Sequence[int],
                       rigid_base: Optional[np.ndarray] = None, **kwargs) -> BatchResult:
        dist, rank, world = self._dist()
        if self._takes_rigid_base:
            kwargs["rigid_base"] = rigid_base
        if world == 1:
            res = self._inner.optimize_batch(sources, targets, unit_src, unit_tgt, **kwargs)
            self.last_local_units = len(res.rmse)
            return res
        import torch

        unit_src = np.asarray(unit_src)
        unit_tgt = np.asarray(unit_tgt)
        n_units = len(unit_src)
        mine = unit_shard(n_units, rank, world)
        # only the source / target sets this rank's units touch travel to its GPU
        src_sets, local_src = np.unique(unit_src[mine], return_inverse=True)
        tgt_sets, local_tgt = np.unique(unit_tgt[mine], return_inverse=True)
        n_slots = (n_units + world - 1) // world
        if len(mine):
            def pick(sets, idx):
                if hasattr(sets, "is_cuda"):
                    return sets.index_select(0, torch.as_tensor(idx, dtype=torch.int64, device=sets.device))
                return np.asarray(sets)[idx]

            res = self._inner.optimize_batch(pick(sources, src_sets), pick(targets, tgt_sets),
                                             local_src.astype(np.int32), local_tgt.astype(np.int32), **kwargs)
            rec = pack_records(res, n_slots)
        else:
            rec = pack_records(BatchResult(np.empty((0, 4, 4)), *(np.empty(0),) * 4, []), n_slots)
        self.last_local_units = len(mine)
        backend = dist.get_backend(self._group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        mine_t = torch.from_numpy(rec).to(dev)
        out = torch.empty((world * n_slots, RECORD_F64), dtype=torch.float64, device=dev)   # rank-major concat
        dist.all_gather_into_tensor(out, mine_t, group=self._group)
        return unpack_records(out.cpu().numpy().reshape(world, n_slots, RECORD_F64), n_units, world)

    def raise_failure(self):
        raise_failure = getattr(self._inner, "raise_failure", None)
        if raise_failure is not None:
            raise_failure()
        raise ValueError("inner registration failed")

    def __repr__(self):
        return f"{self.__class__.__name__}({self._inner!r})"
