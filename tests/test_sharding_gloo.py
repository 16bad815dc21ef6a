"""N>1 path on CPU: two processes over the gloo backend shard the units of a poll round
in contiguous target-set blocks, all-gather the 160-byte records, and every rank replays the same decisions
(SURVEY.md 8(e)).  The inner optimizer is the deterministic fake (no GPU)."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "multi-scale-pointcloud-registration_b200"))
    import torch.distributed as dist

    from fakes import FakeBatched
    from or_pcd_b200.Aligner.Aligner import Aligner
    from or_pcd_b200.Optimizer import ShardedOptimizer
    from or_pcd_b200.Preprocessor import Preprocessor, RandomDownsampler

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(10)
        src = rng.standard_normal((900, 3)) * np.array([1.0, 0.7, 0.4])
        tgt = src[::-1] * np.array([1.0, 1.31, 0.9])
        # raw batch: 7 units over 3 sources x 2 targets (odd count -> ragged shards)
        inner = FakeBatched()
        opt = ShardedOptimizer(inner)
        sources = np.stack([src, src + 0.1, src * 1.1])
        targets = np.stack([tgt, tgt * 1.2])
        us, ut = [0, 1, 2, 0, 1, 2, 0], [0, 0, 0, 1, 1, 1, 1]
        res = opt.optimize_batch(sources, targets, us, ut)
        local_calls = inner.calls
        np.random.seed(3)
        al = Aligner(Preprocessor([RandomDownsampler(400)]), Preprocessor([RandomDownsampler(400)]),
                     ShardedOptimizer(FakeBatched()), attempts=5)
        T, m, scale, errors = al.align(src, tgt, refine_registration=False)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), T=res.transforms, rmse=res.rmse, status=res.status,
                 iters=res.iterations, calls=local_calls, aT=T, am=m, ascale=scale, aerr=np.array(errors),
                 rng=np.random.get_state()[1], local_units=al._optimizer.last_local_units)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_matches_single_process(tmp_path):
    import torch.multiprocessing as mp

    sys.path.insert(0, HERE)
    from fakes import FakeBatched
    from or_pcd_b200.Aligner.Aligner import Aligner
    from or_pcd_b200.Optimizer import ShardedOptimizer
    from or_pcd_b200.Optimizer.shardedOptimizer import unit_shard
    from or_pcd_b200.Preprocessor import Preprocessor, RandomDownsampler

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]

    # single-process truth (ShardedOptimizer is a pass-through without a process group)
    rng = np.random.default_rng(10)
    src = rng.standard_normal((900, 3)) * np.array([1.0, 0.7, 0.4])
    tgt = src[::-1] * np.array([1.0, 1.31, 0.9])
    sources = np.stack([src, src + 0.1, src * 1.1]); targets = np.stack([tgt, tgt * 1.2])
    us, ut = [0, 1, 2, 0, 1, 2, 0], [0, 0, 0, 1, 1, 1, 1]
    truth = ShardedOptimizer(FakeBatched()).optimize_batch(sources, targets, us, ut)
    np.random.seed(3)
    al = Aligner(Preprocessor([RandomDownsampler(400)]), Preprocessor([RandomDownsampler(400)]),
                 FakeBatched(), attempts=5)
    T, m, scale, errors = al.align(src, tgt, refine_registration=False)

    for k in range(world):
        assert np.array_equal(r[k]["T"], truth.transforms) and np.array_equal(r[k]["rmse"], truth.rmse)
        assert np.array_equal(r[k]["status"], truth.status) and np.array_equal(r[k]["iters"], truth.iterations)
        assert int(r[k]["calls"]) == len(unit_shard(ut, k, world))          # each rank ran only its share
        assert np.array_equal(r[k]["aT"], T) and float(r[k]["am"]) == m
        assert np.array_equal(r[k]["ascale"], scale) and np.array_equal(r[k]["aerr"], np.array(errors))
        assert np.array_equal(r[k]["rng"], np.random.get_state()[1])        # every rank replays the same stream
        assert 1 <= int(r[k]["local_units"]) <= 29                          # 6 candidates x 5 starts over 2 ranks, blocks weighted
    assert sum(int(r[k]["local_units"]) for k in range(world)) == 30       # by the previous sweep's iterations per candidate


def test_shard_plan_is_a_partition_in_target_blocks():
    from or_pcd_b200.Optimizer.shardedOptimizer import pack_records, shard_plan, unit_shard, unpack_records
    from or_pcd_b200.Optimizer.iOptimizer import BatchResult

    rng = np.random.default_rng(1)
    cases = [(np.repeat(np.arange(6), 30), 8), (np.array([0, 0, 0, 1, 1, 1, 1]), 2), (np.array([0, 1, 0]), 4),
             (np.repeat(np.arange(32), 32), 8), (rng.integers(0, 5, 97), 3), (np.tile(np.arange(32), 32), 4)]
    for ut, w in cases:
        plan = shard_plan(ut, w)
        assert np.array_equal(np.sort(np.concatenate(plan)), np.arange(len(ut)))          # a partition
        assert max(len(p) for p in plan) - min(len(p) for p in plan) <= 1                   # equal counts
        n_sets = len(np.unique(ut))
        for r in range(w):                                                                  # few target sets per rank
            assert len(np.unique(ut[plan[r]])) <= -(-n_sets // w) + 1
            assert np.array_equal(plan[r], unit_shard(ut, r, w))
        ordered = ut[np.concatenate(plan)]
        assert (np.diff(ordered) >= 0).all()                                                # contiguous in target order
    # weights: blocks of equal expected work instead of equal counts
    ut = np.repeat(np.arange(4), 8)
    wts = np.where(ut < 2, 3.0, 1.0)           # the first two target sets are three times as hard
    plan = shard_plan(ut, 2, wts)
    assert np.array_equal(np.sort(np.concatenate(plan)), np.arange(32))
    loads = [wts[p].sum() for p in plan]
    assert abs(loads[0] - loads[1]) <= 3.0 and len(plan[0]) < len(plan[1])
    with pytest.raises(ValueError):
        shard_plan(ut, 2, -wts)
    # the automatic weights of consecutive calls: per-target-set mean iterations of the previous call
    from or_pcd_b200.Optimizer.shardedOptimizer import target_set_weights
    w = target_set_weights(np.array([0, 0, 1, 1]), [10, 20, 50, 70])
    assert np.array_equal(w, [19.0, 19.0, 64.0, 64.0])
    # records survive the round trip, ragged shards included
    n, w = 7, 2
    ut = np.array([0, 0, 0, 1, 1, 1, 1])
    plan = shard_plan(ut, w)
    full = BatchResult(rng.standard_normal((n, 4, 4)), rng.random(n), rng.random(n), np.arange(n), np.arange(n) % 2, [])
    slots = max(len(p) for p in plan)
    recs = np.stack([pack_records(BatchResult(*(np.asarray(f)[plan[r]] for f in full[:5]), []), slots)
                     for r in range(w)])
    back = unpack_records(recs, plan)
    for a, b in zip(full[:5], back[:5]):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    bad = recs.copy()
    bad[0, 2, 3] = -1.0                                # rank 0 left one of its three slots unfilled
    with pytest.raises(RuntimeError):
        unpack_records(bad, plan)
