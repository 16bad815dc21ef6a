#!/usr/bin/env python
"""bench.py -- GICP candidate evals/s on synthetic 5k-point clouds (BASELINE.json metric).

One STEP = one compass poll round of the Aligner's pattern search recast as a batch: 6 scale candidates x 30 random
starts = 180 independent inner registrations PER GPU (Generalized-ICP to Open3D's stopping rule, max 100 iterations)
of a 5000-point source onto 5000-point targets.  With N GPUs the step is 6 candidates x 30 N starts (weak scaling:
more GPUs = more multi-starts); it goes through the product's own multi-GPU path,
``ShardedOptimizer(GeneralizedICP()).optimize_batch`` -- contiguous target-set blocks per rank, records packed on
the device, one small NCCL all-gather, every rank ends with every result.

  value  : inner registrations per second, inputs resident in HBM, CUDA-event timed, profiling OFF.
  e2e    : the same call with HOST buffers: H2D of the step's clouds and D2H of its results inside the timed region.
           Also carries the second half of BASELINE's metric (Aligner.align() wall seconds on config 1, checked against
           the committed oracle run), FastGlobalOptimizer throughput on the same batch, and a STRONG-scaling record
           (BASELINE config 5: 1024 units of 200k points, the same batch at every N).
  roofline: the dominant kernel, gicp_persistent_kernel -- ONE launch per step: every warp takes (unit, patch) items
           from a work queue; an item is a fused Gauss-Newton round of 64 source points (pending update, chunk-culled
           FP32 1-NN sweep, certificate, exact float64 re-rank, Jacobian sums on the float64 tensor cores); the warp that
           closes a unit's round solves the 6x6 system and enqueues the unit's next round.  Measured in a SEPARATE pass
           (profile mode: two CUDA events around the launch on its stream + device counters read back) so that the
           timed `value` region carries no instrumentation.  `traffic` is read from the committed ncu raw page.
  cpu_baseline: the restated CPU path (oracle/, C + OpenMP, KD-tree) on the host cores -- or, when Open3D is importable,
           the unmodified reference.

`--impl reference` times the reference's CPU implementation alone: the unmodified `or_pcd` classes over Open3D when
`import open3d` works (baseline/_ref or /root/reference/src on sys.path; kind "reference"), else the restated C/OpenMP
port (kind "port": Open3D is an un-vendored wheel that cannot be installed offline).
"""
import argparse
import contextlib
import csv
import glob
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
UNITS_PER_GPU = len(SCALES) * ATTEMPTS
WORKLOAD = ("synthetic 5k-pt pair, GICP, one compass poll round = 6 scale candidates x 30 starts = 180 inner "
            "registrations/step/GPU")
METRIC = "GICP candidate evals/s (5k-pt clouds)"


def base_config(world):
    """The `config` object of both arms (identical keys and values)."""
    return {"workload": WORKLOAD, "n_src": N_POINTS, "n_tgt": N_POINTS, "units_per_gpu": UNITS_PER_GPU,
            "units_per_step": UNITS_PER_GPU * world, "max_iterations": 100, "max_correspondence_distance": 0.5,
            "l2": "256 MB flush between timed steps", "parallelism": f"units x{world} (target-set blocks per rank)"}


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


def make_workload(world=1, pinned=None):
    """The step's batch, identical on every rank: clouds from seed 1234, 6 candidates x (30 * world) starts drawn from
    RandomState(1000) in candidate-major order (unit u = candidate * starts + start, its own source cloud)."""
    rng = np.random.default_rng(1234)
    src = radius_scale(synthetic_cloud(N_POINTS, rng))
    tgt = radius_scale(synthetic_cloud(N_POINTS, rng) @ start_pose(np.random.RandomState(7), 0.3, 0.05)[0])
    rs = np.random.RandomState(1000)
    starts = ATTEMPTS * world
    n_units = len(SCALES) * starts
    sources = pinned((n_units, N_POINTS, 3)) if pinned else np.empty((n_units, N_POINTS, 3))
    for u in range(n_units):
        r, t = start_pose(rs)
        sources[u] = src @ r + t
    targets = pinned((len(SCALES), N_POINTS, 3)) if pinned else np.empty((len(SCALES), N_POINTS, 3))
    for i, s in enumerate(SCALES):
        targets[i] = tgt * s[None, :]
    unit_src = np.arange(n_units, dtype=np.int32)
    unit_tgt = np.repeat(np.arange(len(SCALES), dtype=np.int32), starts)
    return src, sources, targets, unit_src, unit_tgt


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons sampled every 50 ms for the duration of the timed region (B200_PROFILING.md clocks
    line).  Read in-process through NVML (pynvml: two light queries per sample from a background thread); a streamed
    `nvidia-smi -lms 100` is the fallback.  (The nvidia-smi process measurably disturbed what it watched: its NVML
    start-up and its polls take driver locks, and 8-GPU runs lost steps of 12.9 instead of 10.8 ms to it.)"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.samples = index, None, []
        self.thread, self.stop, self.source = None, None, None

    def _nvml_loop(self, nv, handle, max_mhz):
        bits = [nv.nvmlClocksThrottleReasonHwSlowdown, nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                nv.nvmlClocksThrottleReasonSwThermalSlowdown, nv.nvmlClocksThrottleReasonSwPowerCap]
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
                r = int(get_reasons(handle))
                self.samples.append([str(mhz), str(max_mhz)] + ["Active" if r & b else "Not Active" for b in bits])
            except Exception:
                pass
            self.stop.wait(0.05)

    def __enter__(self):
        try:
            import threading

            import pynvml as nv

            nv.nvmlInit()
            handle = None
            try:      # the device THIS process computes on, whatever CUDA_VISIBLE_DEVICES says: by UUID
                import torch

                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                handle = nv.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
            except Exception:
                handle = None
            if handle is None:
                vis = [x.strip() for x in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if x.strip()]
                if self.index < len(vis) and vis[self.index].startswith(("GPU-", "MIG-")):
                    handle = nv.nvmlDeviceGetHandleByUUID(vis[self.index])
                else:
                    phys = int(vis[self.index]) if self.index < len(vis) and vis[self.index].isdigit() else self.index
                    handle = nv.nvmlDeviceGetHandleByIndex(phys)
            max_mhz = nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM)
            self.stop = threading.Event()
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle, max_mhz), daemon=True)
            self.thread.start()
            self.source = "nvml"
            return self
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.thread is not None:
            self.stop.set()
            self.thread.join(timeout=2)
            return
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
                "samples": len(mhz), "source": self.source}


# ----------------------------------------------------------------------------- CPU paths
def _host_threads():
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1 to its workers)
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def port_registrations(src_clouds, tgt_clouds, unit_src, unit_tgt, units):
    """Restated CPU path (oracle, C + OpenMP): `units` inner registrations, wall seconds."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O

    O.build()
    O.set_threads(_host_threads())
    t0 = time.perf_counter()
    iters = 0
    for u in units:
        r = O.gicp(src_clouds[unit_src[u]], tgt_clouds[unit_tgt[u]], 0.5, 100)
        iters += r["iters"]
    return time.perf_counter() - t0, iters, O.num_threads()


def open3d_reference():
    """The unmodified reference's GeneralizedICP class (over Open3D), or None with the reason it is unavailable."""
    for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference/src"):
        if os.path.isdir(os.path.join(p, "or_pcd")) and p not in sys.path:
            sys.path.insert(0, p)
    try:
        import open3d  # noqa: F401
    except Exception as e:      # not installable offline (un-vendored wheel, no network)
        return None, f"import open3d failed: {type(e).__name__}: {e}"
    try:
        from or_pcd.Optimizer.generalizedICP import GeneralizedICP as RefGICP
    except Exception as e:
        return None, f"import or_pcd failed: {type(e).__name__}: {e}"
    return RefGICP, None


def reference_registrations(RefGICP, src_clouds, tgt_clouds, unit_src, unit_tgt, units):
    """The reference's own public API and stock code path, one optimize() call per unit like Aligner.py:188-190."""
    opt = RefGICP()
    os.environ.setdefault("OMP_NUM_THREADS", str(_host_threads()))
    t0 = time.perf_counter()
    for u in units:
        try:
            opt.optimize(np.array(src_clouds[unit_src[u]]), np.array(tgt_clouds[unit_tgt[u]]))
        except ValueError:      # inlier_rmse == 0 (generalizedICP.py:77-81): still one registration's worth of work
            pass
    return time.perf_counter() - t0, 0, _host_threads()


def cpu_sample(sources, targets, unit_src, unit_tgt, per_candidate, steps=1):
    """`per_candidate` starts of each of the 6 scale candidates of the N = 1 step, on the host cores."""
    starts = len(unit_src) // len(SCALES)
    units = [c * starts + k for k in range(per_candidate) for c in range(len(SCALES))]
    RefGICP, why = open3d_reference()
    total, cores = 0.0, 1
    for _ in range(steps):
        if RefGICP is not None:
            s, _, cores = reference_registrations(RefGICP, sources, targets, unit_src, unit_tgt, units)
        else:
            s, _, cores = port_registrations(sources, targets, unit_src, unit_tgt, units)
        total += s
    kind = "reference" if RefGICP is not None else "port"
    what = ("unmodified or_pcd.GeneralizedICP over Open3D" if RefGICP is not None else
            "restated CPU path (oracle/, C + OpenMP, KD-tree), not Open3D (" + why + ")")
    return {"value": len(units) * steps / total, "unit": "registrations/s", "cores": cores, "kind": kind,
            "sample": f"{len(units)} of the step's {UNITS_PER_GPU} registrations per step ({per_candidate} starts per scale "
                      f"candidate), {what}"}, total


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _, sources, targets, unit_src, unit_tgt = make_workload(1)
    if args.warmup > 0:
        cpu_sample(sources, targets, unit_src, unit_tgt, 1, 1)
    cb, _ = cpu_sample(sources, targets, unit_src, unit_tgt, 4, max(args.steps, 1))
    value = cb["value"]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "registrations/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup,
            # a 180-registration step at the measured rate (each timed step ran a bounded sample of it, see cpu_baseline)
            "ms_per_step": 1e3 * UNITS_PER_GPU / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": base_config(max(args.gpus, 1)),
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": "registrations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU path
def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the newest committed ncu
    raw-page export (profiles/*gicp_persistent*raw.csv); None when there is none."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*gicp_persistent*raw*.csv")))
    if not files:
        return None, None
    try:
        rows = list(csv.reader(open(files[-1])))
        h, units = rows[0], rows[1]
        tot = 0.0
        for r in rows[2:]:
            if len(r) < len(h) or "gicp_persistent_kernel" not in " ".join(r[:12]):
                continue
            for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                i = h.index(name)
                mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
                tot += float(r[i]) * mult
            return tot, os.path.relpath(files[-1], ROOT)
    except Exception:
        pass
    return None, None


def run_b200(args):
    import torch
    import torch.distributed as dist

    from or_pcd_b200 import _lib, ops
    from or_pcd_b200.Optimizer import GeneralizedICP, ShardedOptimizer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.lib()

    src, sources, targets, unit_src, unit_tgt = make_workload(world, pinned=_lib.pinned_empty)
    n_units = len(unit_src)
    inner = GeneralizedICP()
    opt = ShardedOptimizer(inner)                 # the product's multi-GPU path (a pass-through at N = 1)
    d_src, d_tgt = _lib.to_device(sources), _lib.to_device(targets)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(s, t):
        inner._base_key = None   # no neighbour lists carried over from the previous step: every step does all its work
        return opt.optimize_batch(s, t, unit_src, unit_tgt, rigid_base=src)

    # The clock sampler is started BEFORE the warm-up so that its start-up -- NVML initialisation takes driver locks on
    # every GPU of the box; eight `nvidia-smi` processes starting inside the timed region cost one rank a 70 ms stall
    # in one step of an 8-GPU run -- happens outside the timed region, and only rank 0 samples (its own GPU): the samples
    # cover the warm-up and the timed steps, all under load.
    warmup = max(args.warmup, 3)
    with (ClockSampler(local) if rank == 0 else contextlib.nullcontext()) as clk:
        time.sleep(0.3 if rank == 0 else 0.0)
        for _ in range(warmup):
            step(d_src, d_tgt)
        barrier()

        # ---- timed: K steps, inputs resident in HBM, L2 flushed between them, no profiling
        k0 = ops.launch_counter["kernels"]
        dev_ms = gather_ms = compute_ms = 0.0
        for _ in range(args.steps):
            flush.fill_(1)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = step(d_src, d_tgt)
            e1.record()
            barrier()
            dev_ms += e0.elapsed_time(e1)
            gather_ms += opt.last_timing.get("gather_ms", 0.0)
            compute_ms += opt.last_timing.get("compute_ms", 0.0)
    launches = ops.launch_counter["kernels"] - k0
    clocks = clk.summary() if clk is not None else None
    mean_iters = float(np.mean(res.iterations))

    # ---- e2e: host buffers in, host results out, through the same public call
    for _ in range(2):
        step(sources, targets)
    barrier()
    e2e_s = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        barrier()
        t0 = time.perf_counter()
        out = step(sources, targets)
        torch.cuda.synchronize()
        e2e_s += time.perf_counter() - t0
    if world > 1:      # bytes per rank: only the source / target sets of the rank's own block travel
        lu, lt = opt.last_local_units, opt.last_local_target_sets
        h2d = lu * N_POINTS * 24 + lt * N_POINTS * 24 + 8 * lu
        d2h = 160 * n_units                          # the all-gathered records, one D2H copy
    else:
        h2d = sources.nbytes + targets.nbytes + unit_src.nbytes + unit_tgt.nbytes
        d2h = out.transforms.nbytes + 4 * 8 * n_units + 4 * n_units

    # ---- N = 1: roofline pass, CPU baseline, Aligner.align(), FGR (before the 200k-point batch heats the GPU up)
    extras = None
    if world == 1:
        extras = single_gpu_extras(inner, src, sources, targets, unit_src, unit_tgt, d_src, d_tgt, dev_ms / args.steps)

    # ---- strong scaling: BASELINE config 5 (1024 units of 200k points), the SAME batch at every N
    strong = strong_scaling(opt, world, barrier) if not args.no_strong else None

    # ---- max over ranks; per-rank numbers for the attribution of any scaling loss
    t = torch.tensor([dev_ms, e2e_s * 1e3, gather_ms, compute_ms], dtype=torch.float64, device="cuda")
    per_rank = None
    if world > 1:
        allr = torch.empty((world, 4), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allr, t)
        per_rank = allr.cpu().numpy()
        dev_ms, e2e_ms = float(per_rank[:, 0].max()), float(per_rank[:, 1].max())
    else:
        dev_ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        line = {
            "metric": METRIC, "value": n_units * args.steps / (dev_ms * 1e-3), "unit": "registrations/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": base_config(world),
            "e2e": {"value": n_units * args.steps / (e2e_ms * 1e-3), "unit": "registrations/s",
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "note": "bytes per rank; with N > 1 every rank uploads only the source / target sets of its own block"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        # everything beyond the contract's keys lives inside `e2e` / `roofline` (top-level extras are dropped by the driver)
        line["e2e"]["mean_iterations"] = mean_iters
        if per_rank is not None:
            line["e2e"]["multi_gpu"] = {
                "path": "ShardedOptimizer(GeneralizedICP()).optimize_batch over NCCL",
                "per_rank_ms_per_step": [float(x) / args.steps for x in per_rank[:, 0]],
                "per_rank_compute_ms_per_step": [float(x) / args.steps for x in per_rank[:, 3]],
                "per_rank_allgather_ms_per_step": [float(x) / args.steps for x in per_rank[:, 2]],
                "record_bytes_all_gathered_per_step": 160 * n_units,
                "note": "every rank runs 180 units of the same batch definition (its own block of starts): the spread of "
                        "per_rank_compute is workload variance, per_rank_allgather is the collective including the wait "
                        "for the slowest rank"}
        if strong is not None:
            line["e2e"]["strong_scaling"] = strong
        if extras is not None:
            line["e2e"].update(extras.pop("e2e_extra"))
            line.update(extras)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def strong_scaling(opt, world, barrier):
    """BASELINE config 5: 32 scale candidates x 32 starts = 1024 registrations of a 200k-point pair, sharded over the N
    GPUs by the product path; the batch is the same at every N, so value(N) / value(1) is strong-scaling speed-up."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import scale_check

    from or_pcd_b200 import _lib

    n, n_scales, n_starts = 200000, 32, 32
    src, tgt, _ = scale_check.make_pair(n, np.random.default_rng(1234))
    scales = scale_check.latin_hypercube(n_scales, np.random.default_rng(7))
    rs = np.random.RandomState(11)
    sources = np.empty((n_starts, n, 3))
    for k in range(n_starts):
        r, t = start_pose(rs, 0.3, 0.05)
        sources[k] = src @ r + t
    d_src = _lib.to_device(sources)
    d_tgt = _lib.to_device(np.stack([tgt * s[None, :] for s in scales]))
    unit_src = np.tile(np.arange(n_starts, dtype=np.int32), n_scales)
    unit_tgt = np.repeat(np.arange(n_scales, dtype=np.int32), n_starts)
    inner = opt._inner
    best, weights, walls = None, None, []
    opt.time_feedback = True       # 200k-point units: correct the iteration proxy by the measured rank times
    for rep in range(6 if world > 1 else 2):
        # rep 0 cuts blocks of equal unit counts; rep 1 cuts blocks of equal expected work: ShardedOptimizer weighs every
        # unit by the mean iteration count of its target set (scale candidate) in the previous call of the same layout
        # -- what the Aligner has from the previous poll sweep; reps 2.. correct that proxy by the measured rank times of
        # the call before (damped: a square-root step per call, so the partition settles over a few calls of the layout)
        inner._base_key = None
        barrier()
        t0 = time.perf_counter()
        res = opt.optimize_batch(d_src, d_tgt, unit_src, unit_tgt, rigid_base=src, unit_weights=weights)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        walls.append(dt)
        weights = None      # (ShardedOptimizer weighs by the per-target-set mean iterations of the previous call itself)
        best = dt if best is None else min(best, dt)
    tw = torch.tensor([best, opt.last_timing.get("compute_ms", 0.0) * 1e-3], dtype=torch.float64, device="cuda")
    per_rank_compute = None
    if world > 1:
        import torch.distributed as dist

        allr = torch.empty((world, 2), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allr, tw)
        allr = allr.cpu().numpy()
        wall = float(allr[:, 0].max())
        per_rank_compute = [float(x) for x in allr[:, 1]]      # of the last (weighted) repetition
    else:
        wall = float(tw[0])
    opt.time_feedback = False
    del d_src, d_tgt
    inner._ws = None
    torch.cuda.empty_cache()
    return {"workload": "BASELINE config 5: synthetic 200k-pt pair, 32 scale candidates x 32 starts = 1024 GICP "
                        "registrations, inputs resident in HBM, same batch at every N",
            "units": 1024, "wall_s": wall, "value": 1024 / wall, "unit": "registrations/s", "scaling": "strong",
            "failed_units": int((res.status != 0).sum()), "local_units": opt.last_local_units,
            "local_target_sets": opt.last_local_target_sets, "mean_iterations": float(np.mean(res.iterations)),
            "wall_s_per_repetition": walls, "per_rank_compute_s_weighted": per_rank_compute,
            "sharding": "contiguous target-set blocks; repetition 0 equal counts, 1 weighted by per-target-set mean iterations of the previous call, 2..5 corrected (damped) by the measured rank times of the call before; wall_s is the best repetition"}


def single_gpu_extras(inner, src, sources, targets, unit_src, unit_tgt, d_src, d_tgt, ms_per_step):
    """Roofline of the dominant kernel (separate profiling pass), CPU baseline, Aligner.align() wall time, FGR."""
    import torch

    from or_pcd_b200 import _lib, ops

    n_units = len(unit_src)
    us, ut = torch.as_tensor(unit_src).cuda(), torch.as_tensor(unit_tgt).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    # ---- profiling pass: same step, profile mode (two events around the one launch, counters read back)
    inner.profile = True
    kernel_ms = pairs = fixed = unit_rounds = 0.0
    n_prof = 5
    for _ in range(n_prof):
        flush.fill_(1)
        torch.cuda.synchronize()
        inner._base_key = None
        sc, tc = inner.covariances(d_src, d_tgt, src)
        p = inner.optimize_batch_device(d_src, d_tgt, us, ut, src_cov=sc, tgt_cov=tc).profile
        kernel_ms += p[1]; pairs += p[2]; fixed += p[4]; unit_rounds += p[7]
        kernel_launches = p[6]
    inner.profile = False

    def peak(packed):          # FFMA / FFMA2 peak of this box, this run (MEASURED_PEAKS.json has no FP32 figure)
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
    achieved = 8.0 * pairs / (kernel_ms * 1e-3) / 1e12

    def dense_sweep():         # the brute-force sweep kernel alone (sweep_mode 1: every (query, target) pair), same clouds
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

    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        hbm_src = "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        hbm_peak, hbm_src = 6549.8, "fallback: MEASURED_PEAKS.json of this pool (6549.8 GB/s)"
    traffic, traffic_file = ncu_traffic()
    # algorithmic bytes of the fused round (192 B gathered per correspondence; here per source point and
    # round: read point, plane axis, previous match and that match's target point (76 B), write point, axis and match
    # (52 B), gather the new match's point and axis (48 B); plus the chunk stream, 512 B per 32 targets x 64 queries)
    alg_bytes = unit_rounds * N_POINTS * 176.0 + pairs / 64.0 * 16.0
    roofline = {
        "bound": "fp32", "achieved": achieved, "peak": ffma_peak, "unit": "TFLOP/s", "frac": achieved / ffma_peak,
        "traffic": traffic, "traffic_unit": "bytes/launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full)",
        "traffic_source": traffic_file, "kernel": "gicp_persistent_kernel",
        "avg_launch_us": 1e3 * kernel_ms / max(kernel_launches * n_prof, 1), "launches_per_step": kernel_launches,
        "share_of_step": (kernel_ms / n_prof) / ms_per_step,
        "note": "achieved = 8 FLOP per (query, target) pair x pairs the kernel evaluated (its own counter) / CUDA-event time "
                "of the launch, profiling pass outside the timed region; peak = FFMA / FFMA2 micro-benchmark of this run. "
                "The kernel fuses the whole registration (float64 update, certificate, exact re-rank, Jacobian sums on the "
                "float64 tensor cores, 6x6 solves, work queue) around the FP32 sweep and culls 89 % of the brute-force pairs, "
                "so the FMA pipe is not what bounds it: ncu shows the shared-memory / L1 data pipe (target-chunk broadcasts) "
                "as the busiest unit; see profiles/ and DESIGN.md 4.1.  brute_force_equivalent_tflops counts the pairs a "
                "brute-force search (SURVEY 8(d)'s per-unit figure) would evaluate in the same time.",
        "dense_sweep_reference": dense_sweep(),
        "hbm_view": {"algorithmic_bytes_per_launch": alg_bytes / n_prof,
                     "achieved_GBps": alg_bytes / (kernel_ms * 1e-3) / 1e9, "peak_GBps": hbm_peak, "peak_source": hbm_src,
                     "frac": alg_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak},
        "pairs_per_step": pairs / n_prof, "pairs_brute_force_per_step": unit_rounds * N_POINTS * N_POINTS / n_prof,
        "culled_fraction": 1.0 - pairs / max(unit_rounds * N_POINTS * N_POINTS, 1),
        "brute_force_equivalent_tflops": 8.0 * unit_rounds * N_POINTS * N_POINTS / (kernel_ms * 1e-3) / 1e12,
        "exact_fallback_fraction": fixed / max(unit_rounds * N_POINTS, 1),
        "unit_rounds_per_step": unit_rounds / n_prof,
    }
    cpu_baseline, _ = cpu_sample(sources, targets, unit_src, unit_tgt, 2)
    extras = {"roofline": roofline, "cpu_baseline": cpu_baseline}
    align = align_config1()
    fgr = fgr_throughput(src, sources, targets, unit_src, unit_tgt)
    extras["e2e_extra"] = {"align": align, "fgr": fgr}
    return extras


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
    for _ in range(3):
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
    # single optimize() call: the path the reference's own sequential Aligner would drive (Aligner.py:188-190)
    one = GeneralizedICP()
    s1, t1 = al.source_pre, al.target_pre
    one.optimize(s1, t1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        one.optimize(s1, t1)
    torch.cuda.synchronize()
    single_ms = (time.perf_counter() - t0) / 5 * 1e3
    return {"config": "ArmadilloBack_330 -> ArmadilloBack_0, Preprocessor[RandomDownsampler(5000), SOR()], GeneralizedICP, "
                      "attempts=30, seed 0", "wall_s": best, "sequential_equivalent_registrations": al.n_optimize_calls,
            "registrations_executed": al.n_units_evaluated, "scale": scale.ravel().tolist(), "rmse": m,
            "matches_oracle_run": ok, "oracle_wall_s_8_cores_authoring_box": 178.0,
            "single_optimize_call_ms": single_ms}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-strong", action="store_true", help="skip the config-5 strong-scaling record")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
