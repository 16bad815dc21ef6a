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


def target_set_weights(unit_tgt: Sequence[int], iterations: Sequence[float], base: float = 4.0) -> np.ndarray:
    """Per-unit weights from a previous evaluation of the same layout: the mean iteration count of the unit's target
    set (a scale candidate keeps its difficulty from one poll sweep to the next; a random start does not), plus a
    constant for the per-unit set-up and the expensive first rounds."""
    unit_tgt = np.asarray(unit_tgt)
    it = np.asarray(iterations, dtype=np.float64)
    w = np.empty(len(unit_tgt), dtype=np.float64)
    for t in np.unique(unit_tgt):
        m = unit_tgt == t
        w[m] = it[m].mean() + base
    return w


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
        self._last_unit_tgt = None
        # no weights given and the same (unit -> target set) layout as the last call -- consecutive poll sweeps of the
        # Aligner -- : weigh every unit by the mean iteration count of its target set (scale candidate) in that call.
        # Results never depend on the partition, only the time does.
        self.auto_weights = True
        # second-order correction of that proxy by the measured rank times of the previous call (see
        # _auto_unit_weights).  Off by default: on the 5k-point batches a rank's time is dominated by the luck of its
        # random starts, the correction chases that noise from call to call (measured: 12.8 -> 14.9 ms per step on
        # 4 GPUs, every new partition also re-allocating workspaces); on 200k-point batches it gains 2-4 %.
        self.time_feedback = False
        self._cost_scale = None          # per-target-set correction of the iteration-count proxy (measured rank times)
        self._last_plan = self._last_weights = None
        self._last_auto = False
        self.last_rank_compute_ms = None

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
        auto = False
        if (unit_weights is None and self.auto_weights and self.last_iterations is not None
                and self._last_unit_tgt is not None and np.array_equal(self._last_unit_tgt, unit_tgt)):
            unit_weights = self._auto_unit_weights(unit_tgt, world)
            auto = True
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
                rec_t = torch.full((n_slots + 1, RECORD_F64), float("nan"), dtype=torch.float64, device=dev)
                rec_t[:, 3] = -1.0
                rec_t[: len(mine)] = rec_local
            else:
                rec = pack_records(self._inner.optimize_batch(*args, **kwargs), n_slots + 1)
        else:
            rec = pack_records(BatchResult(np.empty((0, 4, 4)), *(np.empty(0),) * 4, []), n_slots + 1)
        if ev:
            ev[1].record()
            ev[1].synchronize()          # this rank's compute time travels with its records (last row)
            mine_ms = ev[0].elapsed_time(ev[1])
        else:
            mine_ms = 1e3 * (time.perf_counter() - t0)
        if rec_t is None:
            rec[n_slots, 0] = mine_ms
            rec_t = torch.from_numpy(rec).to(dev)
        else:
            rec_t[n_slots, 0] = mine_ms
        t1 = time.perf_counter()
        if ev:
            ev[1].record()
        out = torch.empty((world * (n_slots + 1), RECORD_F64), dtype=torch.float64, device=dev)   # rank-major concat
        dist.all_gather_into_tensor(out, rec_t, group=self._group)
        if ev:
            ev[2].record()
        host = out.cpu().numpy().reshape(world, n_slots + 1, RECORD_F64)    # the one D2H copy (synchronises)
        t2 = time.perf_counter()
        if ev:
            self.last_timing = {"compute_ms": mine_ms, "gather_ms": ev[1].elapsed_time(ev[2])}
        else:
            self.last_timing = {"compute_ms": 1e3 * (t1 - t0), "gather_ms": 1e3 * (t2 - t1)}
        res = unpack_records(host[:, :n_slots], plan)
        self.last_iterations, self._last_unit_tgt = res.iterations, unit_tgt.copy()
        # what every rank measured for its block (the extra record of each rank): feedback for the next partition
        self._last_plan = plan
        self._last_weights = None if unit_weights is None else np.asarray(unit_weights, dtype=np.float64).copy()
        self._last_auto = auto
        self.last_rank_compute_ms = host[:, n_slots, 0].copy()
        return res

    def _auto_unit_weights(self, unit_tgt, world):
        """Cost per unit for the next partition of a repeated layout.  First guess: the mean iteration count of the
        unit's target set in the previous call (target_set_weights).  Iterations are not the whole cost (a poorly
        converging candidate also visits more chunks per round), so once the ranks' measured compute times of a call
        partitioned with these weights are known, every target set's cost is corrected by the measured / predicted
        ratio of the ranks its units ran on."""
        w = target_set_weights(unit_tgt, self.last_iterations)
        if self._cost_scale is None or not self._last_auto or self._last_weights is None:
            self._cost_scale = {}
        if (self.time_feedback and self._last_auto and self._last_weights is not None
                and self.last_rank_compute_ms is not None):
            t = np.asarray(self.last_rank_compute_ms, dtype=np.float64)
            pred = np.array([self._last_weights[p].sum() if len(p) else 0.0 for p in self._last_plan])
            ok = (pred > 0) & (t > 0)
            if ok.sum() == world:
                ratio = (t / t.mean()) / (pred / pred.mean())            # > 1: the rank's units cost more than predicted
                for ts in np.unique(unit_tgt):
                    num = den = 0.0
                    for r, p in enumerate(self._last_plan):
                        n = float((unit_tgt[p] == ts).sum())
                        num += n * ratio[r]
                        den += n
                    # damped (square root of the measured / predicted ratio, bounded): moving one cut changes two ranks
                    # at once, and part of a rank's time is the luck of its random starts, not its target sets' cost
                    step = min(max((num / den) if den else 1.0, 0.8), 1.25) ** 0.5
                    self._cost_scale[int(ts)] = min(max(self._cost_scale.get(int(ts), 1.0) * step, 0.5), 2.0)
        for ts, sc in self._cost_scale.items():
            w[unit_tgt == ts] *= sc
        return w

    def raise_failure(self):
        raise_failure = getattr(self._inner, "raise_failure", None)
        if raise_failure is not None:
            raise_failure()
        raise ValueError("inner registration failed")

    def __repr__(self):
        return f"{self.__class__.__name__}({self._inner!r})"
