#!/usr/bin/env python
"""bench.py -- GICP candidate evals/s on synthetic 5k-point clouds (BASELINE.json metric).

One STEP = one compass poll round of the Aligner's pattern search recast as a batch:
6 scale candidates x 30 random starts = 180 independent inner registrations
(Generalized-ICP to Open3D's stopping rule, max 100 iterations) of a 5000-point source
onto 5000-point targets, all advanced concurrently on one GPU.  With N GPUs every rank
runs its own 180 units (30 different starts per candidate: more GPUs = more multi-starts,
SURVEY.md 8(e)) and a tiny all-gather collects the (rmse, transform) records.

  value  : inner registrations per second, inputs resident in HBM, CUDA-event timed.
  e2e    : same through the plugin API (GeneralizedICP.optimize_batch) with HOST buffers:
           H2D of the step's clouds and D2H of its results inside the timed region.
  roofline: the dominant kernel (gicp_round_kernel: one launch per Gauss-Newton round = pending update +
           chunk-culled FP32 1-NN sweep + certificate + exact float64 re-rank + Jacobian sums + solve) --
           8 FLOP per (query,target) pair the kernel actually evaluated (chunks it culls are not counted;
           the brute-force equivalent is reported beside it) over the CUDA-event time spent in that kernel
           during the timed steps, against the FFMA peak measured in this run.  The dense (un-culled)
           sweep kernel, which is purely FMA-pipe work, is timed beside it as `dense_sweep_reference`.
  cpu_baseline: the restated CPU path (oracle/, C + OpenMP, KD-tree) on the host cores.

`--impl reference` times that CPU path alone (Open3D, the reference's arithmetic, is an
un-vendored dependency that cannot be installed offline).
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "multi-scale-pointcloud-registration_b200"))

N_POINTS = 5000
ATTEMPTS = 30
SCALES = np.array([[1.2, 1, 1], [0.8, 1, 1], [1, 1.2, 1], [1, 0.8, 1], [1, 1, 1.2], [1, 1, 0.8]], dtype=np.float64)
WORKLOAD = "synthetic 5k-pt pair, GICP, one compass poll round = 6 scale candidates x 30 starts = 180 inner registrations/step/GPU"


# ----------------------------------------------------------------------------- synthetic workload
def synthetic_cloud(n, rng):
    """SURVEY.md 8(d) surface sampler."""
    u = rng.standard_normal((n, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    r = 1 + 0.3 * np.sin(3 * u[:, 0]) * np.cos(2 * u[:, 1]) + 0.2 * u[:, 2] ** 3
    return r[:, None] * u + rng.normal(0, 0.002, (n, 3))


def radius_scale(c):
    c = c - c.mean(0, keepdims=True)
    return c / np.sqrt((c * c).sum(1)).max()


def start_pose(rs, deg=np.pi / 2, std=0.1):
    a, b, g = rs.uniform(-deg, deg), rs.uniform(-deg, deg), rs.uniform(-deg, deg)
    rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    rz = np.array([[np.cos(g), -np.sin(g), 0], [np.sin(g), np.cos(g), 0], [0, 0, 1]])
    return rx @ ry @ rz, rs.standard_normal(3) * std


def make_workload(rank, pinned=None):
    """Same clouds on every rank (seed 1234); the random starts differ per rank."""
    rng = np.random.default_rng(1234)
    src = radius_scale(synthetic_cloud(N_POINTS, rng))
    tgt = radius_scale(synthetic_cloud(N_POINTS, rng) @ start_pose(np.random.RandomState(7), 0.3, 0.05)[0])
    rs = np.random.RandomState(1000 + rank)
    n_units = len(SCALES) * ATTEMPTS
    sources = pinned((n_units, N_POINTS, 3)) if pinned else np.empty((n_units, N_POINTS, 3))
    for u in range(n_units):
        r, t = start_pose(rs)
        sources[u] = src @ r + t
    targets = pinned((len(SCALES), N_POINTS, 3)) if pinned else np.empty((len(SCALES), N_POINTS, 3))
    for i, s in enumerate(SCALES):
        targets[i] = tgt * s[None, :]
    unit_src = np.arange(n_units, dtype=np.int32)
    unit_tgt = np.repeat(np.arange(len(SCALES), dtype=np.int32), ATTEMPTS)
    return src, sources, targets, unit_src, unit_tgt


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """`nvidia-smi -lms 100` streamed for the duration of the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.samples = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is None:
            return
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 6:
                self.samples.append(f)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        mhz = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(mhz)}


# ----------------------------------------------------------------------------- CPU path
def cpu_registrations(src_clouds, tgt_clouds, unit_src, unit_tgt, units):
    """Restated CPU path (oracle, C + OpenMP): `units` inner registrations, wall seconds."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O

    O.build()
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1 to its workers)
    O.set_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    t0 = time.perf_counter()
    iters = 0
    for u in units:
        r = O.gicp(src_clouds[unit_src[u]], tgt_clouds[unit_tgt[u]], 0.5, 100)
        iters += r["iters"]
    return time.perf_counter() - t0, iters, O.num_threads()


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores.
    Open3D cannot be installed here, so this is the restated C/OpenMP port (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _, sources, targets, unit_src, unit_tgt = make_workload(0)
    per_step = 24          # bounded sample: 24 of the step's 180 registrations (4 per scale candidate)
    units = [c * ATTEMPTS + k for k in range(per_step // len(SCALES)) for c in range(len(SCALES))]
    for _ in range(args.warmup):
        cpu_registrations(sources, targets, unit_src, unit_tgt, units[:6])
    total_s, cores = 0.0, 1
    for _ in range(args.steps):
        s, _, cores = cpu_registrations(sources, targets, unit_src, unit_tgt, units)
        total_s += s
    value = per_step * args.steps / total_s
    line = {"impl": "reference", "metric": "GICP candidate evals/s (5k-pt clouds)", "value": value,
            "unit": "registrations/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "restated CPU path (oracle/, C + OpenMP, KD-tree), not Open3D"},
            "cpu_baseline": {"value": value, "unit": "registrations/s", "cores": cores, "kind": "port",
                             "sample": f"{per_step} of the step's 180 registrations per step (4 starts per scale candidate)"},
            "e2e": {"value": value, "unit": "registrations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU path
def run_b200(args):
    import torch
    import torch.distributed as dist

    from or_pcd_b200 import _lib, ops
    from or_pcd_b200.Optimizer import GeneralizedICP

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.lib()

    src, sources, targets, unit_src, unit_tgt = make_workload(rank, pinned=_lib.pinned_empty)
    n_units = len(unit_src)
    opt = GeneralizedICP()
    opt.profile = True
    d_src, d_tgt = _lib.to_device(sources), _lib.to_device(targets)
    us, ut = torch.as_tensor(unit_src).cuda(), torch.as_tensor(unit_tgt).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2
    gather_buf = torch.empty((world, n_units, 20), dtype=torch.float64, device="cuda") if world > 1 else None

    def device_step():
        """Hot path with inputs resident in HBM: source covariances, target covariances, batched GICP."""
        opt._base_key = None   # no neighbour lists carried over from the previous step: every step does all its work
        sc, tc = opt.covariances(d_src, d_tgt, src)
        res = opt.optimize_batch_device(d_src, d_tgt, us, ut, src_cov=sc, tgt_cov=tc)
        if world > 1:   # (rmse, fitness, iters, status, T) record per unit, one small all-gather
            rec = torch.cat([res.stats, res.status.to(torch.float64)[:, None], res.T.reshape(n_units, 16)], 1)[:, :20]
            dist.all_gather_into_tensor(gather_buf, rec.contiguous())
        return res

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()

    # ---- timed: K device-resident steps, L2 flushed between them
    k0 = ops.launch_counter["kernels"]
    sweep_ms, pairs, fixed, unit_rounds = 0.0, 0.0, 0.0, 0.0   # pairs = (query,target) pairs evaluated by the sweep
    dev_ms = 0.0
    with ClockSampler(local) as clk:
        for _ in range(args.steps):
            flush.fill_(1)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = device_step()
            e1.record()
            barrier()
            dev_ms += e0.elapsed_time(e1)
            p = res.profile
            sweep_ms += p[1]; pairs += p[2]; fixed += p[4]; unit_rounds += p[7]
            sweep_launches = p[6]
    launches = ops.launch_counter["kernels"] - k0
    clocks = clk.summary()
    mean_iters = float(res.stats[:, 2].mean().item())

    # ---- e2e: host buffers in, host results out, through the plugin API
    opt.profile = False
    for _ in range(2):
        opt.optimize_batch(sources, targets, unit_src, unit_tgt, rigid_base=src)
    barrier()
    e2e_s = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        barrier()
        opt._base_key = None
        t0 = time.perf_counter()
        out = opt.optimize_batch(sources, targets, unit_src, unit_tgt, rigid_base=src)
        torch.cuda.synchronize()
        e2e_s += time.perf_counter() - t0
    h2d = sources.nbytes + targets.nbytes + unit_src.nbytes + unit_tgt.nbytes
    d2h = out.transforms.nbytes + 4 * 8 * n_units + 4 * n_units

    # ---- second half of the metric: Aligner.align() wall seconds on BASELINE config 1 (rank 0, N = 1)
    align_info = None
    fgr_info = None
    if world == 1:
        align_info = align_config1()
        fgr_info = fgr_throughput(src, sources, targets, unit_src, unit_tgt)

    # ---- max over ranks
    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        # FFMA peak of this box, this run (roofline denominator)
        def peak(packed):
            best = 0.0
            blocks, iters = 148 * 16, 4096
            ops.ffma_burn(blocks, 64, packed)
            torch.cuda.synchronize()
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); ops.ffma_burn(blocks, iters, packed); b.record()
                torch.cuda.synchronize()
                best = max(best, blocks * 256.0 * iters * 16 * 8 * (4 if packed else 2) / (a.elapsed_time(b) * 1e-3) / 1e12)
            return best

        ffma_peak = max(peak(False), peak(True))
        achieved = 8.0 * pairs / (sweep_ms * 1e-3) / 1e12

        # the brute-force sweep kernel alone (sweep_mode 1: every (query, target) pair), same clouds
        def dense_sweep():
            ax_s = ops.normals_cov(d_src, want_gicp_cov=False, want_gicp_axis=True)["gicp_axis"]
            ax_t = ops.normals_cov(d_tgt, want_gicp_cov=False, want_gicp_axis=True)["gicp_axis"]
            r0 = ops.gicp_batch(d_src, ax_s, d_tgt, ax_t, us, ut, max_iterations=0, sweep_mode=1)
            for _ in range(2):
                ops.gicp_sweep_only(r0.params, r0.workspace)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                ops.gicp_sweep_only(r0.params, r0.workspace)
            b.record()
            torch.cuda.synchronize()
            tf = 8.0 * n_units * N_POINTS * N_POINTS * 5 / (a.elapsed_time(b) * 1e-3) / 1e12
            return {"kernel": "nn_sweep_kernel (dense, not on the default path)", "achieved": tf, "unit": "TFLOP/s",
                    "frac": tf / ffma_peak, "us_per_launch": a.elapsed_time(b) * 1e3 / 5}

        dense_ref = dense_sweep()
        try:
            hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            hbm_src = "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            hbm_peak, hbm_src = 6549.8, "fallback: MEASURED_PEAKS.json of this pool (6549.8 GB/s)"
        # CPU baseline on this box's host cores: a bounded sample of the same step (N = 1 only)
        cpu_units = [c * ATTEMPTS + k for k in range(2) for c in range(len(SCALES))]
        if world == 1:
            cpu_s, _, cores = cpu_registrations(sources, targets, unit_src, unit_tgt, cpu_units)
            cpu_baseline = {"value": len(cpu_units) / cpu_s, "unit": "registrations/s", "cores": cores, "kind": "port",
                            "sample": f"{len(cpu_units)} of the step's 180 registrations (2 starts per scale candidate), "
                                      "restated CPU path (oracle/, C + OpenMP, KD-tree), not Open3D"}
        else:
            cpu_baseline = None
        value = world * n_units * args.steps / (dev_ms * 1e-3)
        line = {
            "metric": "GICP candidate evals/s (5k-pt clouds)", "value": value, "unit": "registrations/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_src": N_POINTS, "n_tgt": N_POINTS, "units_per_gpu": n_units,
                       "max_iterations": 100, "mean_iterations": mean_iters, "l2": "256 MB flush between timed steps",
                       "filter_dtype": "f32 (certified, chunk-culled) + f64 exact re-rank", "parallelism": f"units x{world}"},
            "e2e": {"value": world * n_units * args.steps / (e2e_ms * 1e-3), "unit": "registrations/s",
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": ffma_peak, "unit": "TFLOP/s",
                         "frac": achieved / ffma_peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch with all 180 units active, from the
                         # ncu --set full capture in profiles/r1_05 (48.7 MB + 22.7 MB; the working set is L2-resident)
                         "traffic": 71.4e6, "traffic_unit": "bytes/launch (ncu, profiles/r1_05_gicp_round_ncu_full_details.csv)",
                         "kernel": "gicp_round_kernel", "avg_launch_us": 1e3 * sweep_ms / max(sweep_launches * args.steps, 1),
                         "share_of_step": sweep_ms / dev_ms,
                         "note": "8 FLOP per (query,target) pair x pairs the kernel evaluated / CUDA-event time inside the "
                                 "kernel; peak = FFMA/FFMA2 micro-benchmark measured in this run (MEASURED_PEAKS.json has "
                                 "no FP32 figure); brute force would evaluate pairs_brute_force_per_step.  The kernel "
                                 "fuses the whole round (float64 update, certificate, exact re-rank, Jacobian sums, "
                                 "6x6 solve) around the FP32 sweep, so its time is not FMA-pipe time alone: see "
                                 "profiles/ for the per-phase split",
                         "dense_sweep_reference": dense_ref,
                         # the same kernel against the HBM roofline: per source point and round it reads point, plane
                         # axis, previous match and that match's target point (76 B), writes point, axis and match
                         # (52 B), gathers the new match's point and axis (48 B) and streams pairs/32 chunks of 512 B per
                         # 64-point patch -- all of it L2-resident, which is why DRAM traffic is a quarter of it
                         "hbm_view": {"algorithmic_bytes_per_step": unit_rounds * N_POINTS * 176.0 / args.steps + pairs / args.steps / 64.0 * 16.0,
                                      "achieved_GBps": (unit_rounds * N_POINTS * 176.0 + pairs / 64.0 * 16.0) / (sweep_ms * 1e-3) / 1e9,
                                      "peak_GBps": hbm_peak, "peak_source": hbm_src,
                                      "frac": (unit_rounds * N_POINTS * 176.0 + pairs / 64.0 * 16.0) / (sweep_ms * 1e-3) / 1e9 / hbm_peak},
                         "pairs_per_step": pairs / args.steps,
                         "pairs_brute_force_per_step": unit_rounds * N_POINTS * N_POINTS / args.steps,
                         "culled_fraction": 1.0 - pairs / max(unit_rounds * N_POINTS * N_POINTS, 1),
                         "brute_force_equivalent_tflops": 8.0 * unit_rounds * N_POINTS * N_POINTS / (sweep_ms * 1e-3) / 1e12,
                         "exact_fallback_fraction": fixed / max(unit_rounds * N_POINTS, 1)},
        }
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        if align_info is not None:
            line["align"] = align_info
        if fgr_info is not None:
            line["fgr"] = fgr_info
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def fgr_throughput(src, sources, targets, unit_src, unit_tgt):
    """The same 180-unit batch through FastGlobalOptimizer.optimize_batch (host buffers in, host results out):
    FPFH of every source copy and target set, tcgen05 feature matching both ways, tuple test, GNC solve,
    evaluation.  Secondary figure (BASELINE config 2's optimizer on the bench clouds)."""
    import torch

    from or_pcd_b200.Optimizer import FastGlobalOptimizer

    out = {}
    for mode in ("fixed", "reference_compat"):
        opt = FastGlobalOptimizer(feature_mode=mode)
        best = None
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            opt.optimize_batch(sources, targets, unit_src, unit_tgt, rigid_base=src)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        out[mode] = {"registrations_per_s": len(unit_src) / best, "wall_s": best,
                     "feature_rows_exact_scanned": opt.last_exact_scans}
    out["note"] = "FastGlobalOptimizer.optimize_batch on the bench batch, host buffers, best of 3"
    return out


def align_config1():
    """Aligner.align() end to end (preprocessing included) on the Armadillo pair of BASELINE config 1, checked
    against the committed oracle run (tests/golden/golden_v1.npz)."""
    import torch

    from or_pcd_b200.Aligner.Aligner import Aligner
    from or_pcd_b200.data import load_sample_cloud
    from or_pcd_b200.Optimizer import GeneralizedICP
    from or_pcd_b200.Preprocessor import Preprocessor, RandomDownsampler, SOR

    src, tgt = load_sample_cloud("ArmadilloBack_330"), load_sample_cloud("ArmadilloBack_0")
    gold = np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))
    best = None
    for _ in range(2):
        np.random.seed(0)
        al = Aligner(Preprocessor([RandomDownsampler(5000), SOR()]), Preprocessor([RandomDownsampler(5000), SOR()]),
                     GeneralizedICP())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        T, m, scale, errors = al.align(src, tgt, refine_registration=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    ok = bool(np.array_equal(scale, gold["align_back_scale"]) and abs(m - float(gold["align_back_metric"])) <= 1e-5 * m)
    return {"config": "ArmadilloBack_330 -> ArmadilloBack_0, Preprocessor[RandomDownsampler(5000), SOR()], GeneralizedICP, "
                      "attempts=30, seed 0", "wall_s": best, "sequential_equivalent_registrations": al.n_optimize_calls,
            "registrations_executed": al.n_units_evaluated, "scale": scale.ravel().tolist(), "rmse": m,
            "matches_oracle_run": ok, "oracle_wall_s_8_cores_authoring_box": 178.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
