"""Aligner.align() with the units of every poll round sharded over the GPUs of one box
(torchrun --nproc-per-node N tools/multigpu_align.py [back|onhead] [attempts]).  Every rank must end
with the same decisions as the committed oracle run."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-scale-pointcloud-registration_b200"))


def main():
    import torch
    import torch.distributed as dist

    from or_pcd_b200.Aligner.Aligner import Aligner
    from or_pcd_b200.data import load_sample_cloud
    from or_pcd_b200.Optimizer import GeneralizedICP, ShardedOptimizer
    from or_pcd_b200.Preprocessor import Preprocessor, RandomDownsampler, SOR

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pair = sys.argv[1] if len(sys.argv) > 1 else "back"
    attempts = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    names = {"back": ("ArmadilloBack_330", "ArmadilloBack_0"),
             "onhead": ("ArmadilloOnHeadMultiple_150", "ArmadilloOnHeadMultiple_240")}[pair]
    gold = np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz" if pair == "back" else "golden_onhead64.npz"))
    g_scale = gold["align_back_scale"] if pair == "back" else gold["scale"]
    g_metric = float(gold["align_back_metric"] if pair == "back" else gold["metric"])
    src, tgt = load_sample_cloud(names[0]), load_sample_cloud(names[1])
    for rep in range(3):
        np.random.seed(0)
        al = Aligner(Preprocessor([RandomDownsampler(5000), SOR()]), Preprocessor([RandomDownsampler(5000), SOR()]),
                     ShardedOptimizer(GeneralizedICP()), attempts=attempts)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        T, m, scale, errors = al.align(src, tgt, refine_registration=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        ok = bool(np.array_equal(scale, g_scale) and abs(m - g_metric) <= 1e-5 * g_metric)
        print(json.dumps(dict(rank=rank, world=world, pair=pair, rep=rep, align_wall_s=dt, rmse=m,
                              scale=scale.ravel().tolist(), matches_oracle_run=ok,
                              sequential_equivalent_calls=al.n_optimize_calls)), flush=True)
        assert ok
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
