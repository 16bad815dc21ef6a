"""Candidate sharding across the GPUs of one box.

The units of a poll round (scale candidate x random start) are independent inner
registrations, so they shard with no data-path collective: every rank (one process
per GPU, ``torch.distributed``) derives the same unit list from the same numpy
stream, runs one block of it on its own GPU and contributes one fixed-size record
per unit -- {rmse, fitness, iterations, status, T[16]} = 20 float64 = 160 bytes --
to a single all-gather (NCCL over NVLink on the GPU box, gloo in the CPU tests).
Every rank then holds the full result and replays the Aligner's accept logic
identically; no decision broadcast is needed.

Blocks follow the TARGET sets: the units are ordered by target set (stable) and cut
into ``world`` contiguous runs of equal weight, so a rank prepares (k-NN graph,
covariances, Hilbert order, chunk spheres) only the few target sets its own units
touch -- with round-robin dealing every rank would repeat that work for every set.
With an optimizer that offers ``optimize_batch_records`` (GeneralizedICP) the records
are packed on the device and go GPU -> all-gather -> one D2H copy of the result.

The reference has no counterpart (single process; Aligner/Aligner.py:164-204 runs
the starts one after another)."""
import inspect
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .iOptimizer import BatchResult, IOptimizer

RECORD_F64 = 20   # rmse, fitness, iterations, status, 16 transform entries


def shard_plan(unit_tgt: Sequence[int], world: int, weights: Optional[Sequence[float]] = None) -> List[np.ndarray]:
    """Unit indices of every rank.  Units are ordered by target set (stable: ties keep the caller's order) and
    the ordered list is cut into ``world`` contiguous runs of (nearly) equal total weight -- equal counts when
    ``weights`` is None -- so each rank touches at most ``ceil(sets / world) + 1`` target sets."""
    unit_tgt = np.asarray(unit_tgt)
    n = len(unit_tgt)
    order = np.argsort(unit_tgt, kind="stable").astype(np.int64)
    if weights is None:
        cuts = [(n * r) // world for r in range(world + 1)]
    else:
        w = np.asarray(weights, dtype=np.float64)[order]
        if (w < 0).any() or not np.isfinite(w).all():
            raise ValueError("shard weights must be finite and non-negative")
        c = np.concatenate([[0.0], np.cumsum(w)])
        total = c[-1]
        if total <= 0:
            cuts = [(n * r) // world for r in range(world + 1)]
        else:   # unit i goes to the rank whose weight interval holds the midpoint of its own
            mid = 0.5 * (c[:-1] + c[1:]) / total * world
            owner = np.minimum(mid.astype(np.int64), world - 1)
            cuts = [int(np.searchsorted(owner, r, side="left")) for r in range(world)] + [n]
    return [order[cuts[r]:cuts[r + 1]] for r in range(world)]


def unit_shard(unit_tgt: Sequence[int], rank: int, world: int, weights=None) -> np.ndarray:
    """Indices of the units rank ``rank`` runs (see shard_plan)."""
    return shard_plan(unit_tgt, world, weights)[rank]


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


def unpack_records(all_rec: np.ndarray, plan: List[np.ndarray]) -> BatchResult:
    """Inverse of the shard: all_rec is (world, n_slots, 20), ``plan`` the shard_plan every rank derived."""
    n_units = int(sum(len(p) for p in plan))
    T = np.empty((n_units, 4, 4), dtype=np.float64)
    rmse = np.empty(n_units, dtype=np.float64)
    fit = np.empty(n_units, dtype=np.float64)
    iters = np.empty(n_units, dtype=np.int64)
    status = np.empty(n_units, dtype=np.int32)
    for r, mine in enumerate(plan):
        rec = all_rec[r, : len(mine)]
        if (rec[:, 3] < 0).any() or np.isnan(rec[:, 3]).any():
            raise RuntimeError(f"rank {r} contributed fewer records than its shard holds")
        rmse[mine] = rec[:, 0]
        fit[mine] = rec[:, 1]
        iters[mine] = rec[:, 2].astype(np.int64)
        status[mine] = rec[:, 3].astype(np.int32)
        T[mine] = rec[:, 4:].reshape(-1, 4, 4)
    return BatchResult(T, rmse, fit, iters, status, [])


class ShardedOptimizer(IOptimizer):
    """Wraps a batched optimizer; ``optimize_batch`` runs this rank's share of the units and
    all-gathers the records.  With no process group (or world size 1) it is a pass-through.

    ``unit_weights`` (optional, set per call through ``optimize_batch(..., unit_weights=w)``): a difficulty proxy per
    unit -- e.g. the iteration counts of the same candidates in the previous poll sweep -- used to cut blocks of
    equal expected work instead of equal unit counts."""

    def __init__(self, inner: IOptimizer, group=None):
        if not hasattr(inner, "optimize_batch"):
            raise TypeError("ShardedOptimizer needs an optimizer with optimize_batch")
        self._inner = inner
        self._group = group
        self._takes_rigid_base = "rigid_base" in inspect.signature(inner.optimize_batch).parameters
        # start clouds built on the GPU by the Aligner (CUDA tensors) are passed through when the inner optimizer takes them
        self.accepts_device_sources = bool(getattr(inner, "accepts_device_sources", False))
        self.last_local_units = 0
        self.last_local_target_sets = 0
        self.last_timing = {}     # {"compute_ms", "gather_ms"} of the last sharded call (CUDA events / perf_counter)
        self.last_iterations = None   # iterations per unit of the last call: a difficulty proxy for the next one

    def _dist(self):
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist, dist.get_rank(self._group), dist.get_world_size(self._group)
        return None, 0, 1

    def optimize(self, source: np.ndarray, target: np.ndarray, **kwargs) -> Tuple[np.ndarray, float]:
        return self._inner.optimize(source, target, **kwargs)

    def optimize_batch(self, sources, targets, unit_src: Sequence[int], unit_tgt: Sequence[int],
                       rigid_base: Optional[np.ndarray] = None, unit_weights=None, **kwargs) -> BatchResult:
        dist, rank, world = self._dist()
        if self._takes_rigid_base:
            kwargs["rigid_base"] = rigid_base
        if world == 1:
            res = self._inner.optimize_batch(sources, targets, unit_src, unit_tgt, **kwargs)
            self.last_local_units = len(res.rmse)
            self.last_local_target_sets = len(np.unique(np.asarray(unit_tgt)))
            self.last_iterations = res.iterations
            return res
        import time

        import torch

        unit_src = np.asarray(unit_src)
        unit_tgt = np.asarray(unit_tgt)
        plan = shard_plan(unit_tgt, world, unit_weights)
        mine = plan[rank]
        n_slots = max(len(p) for p in plan)
        backend = dist.get_backend(self._group)
        on_gpu = backend == "nccl"
        dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
        # only the source / target sets this rank's units touch travel to (or stay on) its GPU
        src_sets, local_src = np.unique(unit_src[mine], return_inverse=True)
        tgt_sets, local_tgt = np.unique(unit_tgt[mine], return_inverse=True)
        self.last_local_units, self.last_local_target_sets = len(mine), len(tgt_sets)

        def pick(sets, idx):
            if len(idx) == sets.shape[0]:
                return sets
            if len(idx) and int(idx[-1]) - int(idx[0]) + 1 == len(idx):
                # a contiguous run (the usual case: blocks follow the unit order): a view, so a pinned host array
                # stays pinned for its H2D copy and a CUDA tensor is not copied at all
                return sets[int(idx[0]):int(idx[-1]) + 1]
            if hasattr(sets, "is_cuda"):
                return sets.index_select(0, torch.as_tensor(idx, dtype=torch.int64, device=sets.device))
            return np.asarray(sets)[idx]

        t0 = time.perf_counter()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if on_gpu else None
        if ev:
            ev[0].record()
        rec_t = None
        if len(mine):
            args = (pick(sources, src_sets), pick(targets, tgt_sets), local_src.astype(np.int32), local_tgt.astype(np.int32))
            records = getattr(self._inner, "optimize_batch_records", None) if on_gpu else None
            if records is not None:      # records packed on the device: nothing comes back to the host before the gather
                rec_local = records(*args, **kwargs)
                rec_t = torch.full((n_slots, RECORD_F64), float("nan"), dtype=torch.float64, device=dev)
                rec_t[:, 3] = -1.0
                rec_t[: len(mine)] = rec_local
            else:
                rec = pack_records(self._inner.optimize_batch(*args, **kwargs), n_slots)
        else:
            rec = pack_records(BatchResult(np.empty((0, 4, 4)), *(np.empty(0),) * 4, []), n_slots)
        if rec_t is None:
            rec_t = torch.from_numpy(rec).to(dev)
        t1 = time.perf_counter()
        if ev:
            ev[1].record()
        out = torch.empty((world * n_slots, RECORD_F64), dtype=torch.float64, device=dev)   # rank-major concat
        dist.all_gather_into_tensor(out, rec_t, group=self._group)
        if ev:
            ev[2].record()
        host = out.cpu().numpy().reshape(world, n_slots, RECORD_F64)    # the one D2H copy (synchronises)
        t2 = time.perf_counter()
        if ev:
            self.last_timing = {"compute_ms": ev[0].elapsed_time(ev[1]), "gather_ms": ev[1].elapsed_time(ev[2])}
        else:
            self.last_timing = {"compute_ms": 1e3 * (t1 - t0), "gather_ms": 1e3 * (t2 - t1)}
        res = unpack_records(host, plan)
        self.last_iterations = res.iterations
        return res

    def raise_failure(self):
        raise_failure = getattr(self._inner, "raise_failure", None)
        if raise_failure is not None:
            raise_failure()
        raise ValueError("inner registration failed")

    def __repr__(self):
        return f"{self.__class__.__name__}({self._inner!r})"
