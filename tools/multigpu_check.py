"""N-GPU check of the product's sharded path (launch with torchrun, one rank per GPU, NCCL):
  1. a raw batch through ShardedOptimizer(GeneralizedICP()) equals, bit for bit on every rank, the same batch run
     whole on the rank's own GPU (units are independent; records are all-gathered);
  2. Aligner.align() of BASELINE config 1 with the sharded optimizer takes the same decisions as the committed
     oracle run on every rank (tests/golden/golden_v1.npz).
Prints one JSON line per rank; exits non-zero on any mismatch."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-scale-pointcloud-registration_b200"))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import bench
    from or_pcd_b200.Aligner.Aligner import Aligner
    from or_pcd_b200.data import load_sample_cloud
    from or_pcd_b200.Optimizer import GeneralizedICP, ShardedOptimizer
    from or_pcd_b200.Preprocessor import Preprocessor, RandomDownsampler, SOR

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = dict(rank=rank, world=world)
    try:
        # ---- 1. raw batch: 5 target sets x 7 starts (35 units: ragged shards for every world size > 1)
        rng = np.random.default_rng(3)
        src = bench.radius_scale(bench.synthetic_cloud(3000, rng))
        tgt = bench.radius_scale(bench.synthetic_cloud(2900, rng))
        rs = np.random.RandomState(5)
        sources = np.stack([src @ r + t for r, t in (bench.start_pose(rs, 0.6, 0.05) for _ in range(7))])
        targets = np.stack([tgt * s for s in ([1, 1, 1], [1.1, 1, 1], [1, 0.9, 1], [1, 1, 1.2], [0.95, 1.05, 1])])
        us = np.tile(np.arange(7, dtype=np.int32), 5)
        ut = np.repeat(np.arange(5, dtype=np.int32), 7)
        whole = GeneralizedICP().optimize_batch(sources, targets, us, ut, rigid_base=src)
        sh = ShardedOptimizer(GeneralizedICP())
        part = sh.optimize_batch(sources, targets, us, ut, rigid_base=src)
        same = all(np.array_equal(np.asarray(a), np.asarray(b)) for a, b in zip(whole[:5], part[:5]))
        out.update(raw_batch_identical=bool(same), local_units=sh.last_local_units,
                   local_target_sets=sh.last_local_target_sets, timing=sh.last_timing)
        # ---- 2. align() of config 1, decisions against the committed oracle run
        gold = np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))
        a, b = load_sample_cloud("ArmadilloBack_330"), load_sample_cloud("ArmadilloBack_0")
        np.random.seed(0)
        al = Aligner(Preprocessor([RandomDownsampler(5000), SOR()]), Preprocessor([RandomDownsampler(5000), SOR()]),
                     ShardedOptimizer(GeneralizedICP()))
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        T, m, scale, errors = al.align(a, b, refine_registration=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        acc = np.array([d["accepted"] for d in al.decision_log])
        ok = bool(np.array_equal(scale, gold["align_back_scale"]) and np.array_equal(acc, gold["align_back_log_accepted"])
                  and abs(m - float(gold["align_back_metric"])) <= 1e-5 * m
                  and np.array_equal(np.random.get_state()[1][:8], gold["align_back_rng_after"]))
        # every rank holds the same answer: compare a digest across ranks
        dig = torch.tensor(np.concatenate([T.ravel(), scale.ravel(), [m]]), dtype=torch.float64, device="cuda")
        gathered = [torch.empty_like(dig) for _ in range(world)]
        dist.all_gather(gathered, dig)
        ranks_agree = all(torch.equal(g, gathered[0]) for g in gathered)
        out.update(align_matches_oracle_run=ok, ranks_agree=bool(ranks_agree), align_wall_s=dt, rmse=float(m),
                   scale=scale.ravel().tolist(), sequential_equivalent_calls=al.n_optimize_calls)
        print(json.dumps(out), flush=True)
        if not (same and ok and ranks_agree):
            sys.exit(1)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
