"""The profile helpers run on the committed ncu exports (no GPU needed)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_launch_summary_reads_the_committed_launch_list():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"),
                          os.path.join(ROOT, "profiles", "r1_05_gicp_step_launches.csv")],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "gicp_round_kernel" in out.stdout and "Sum " in out.stdout
    first = [l for l in out.stdout.splitlines() if l.startswith("| gicp_round_kernel")][0]
    share = float(first.split("|")[4].strip().rstrip("%"))
    assert share > 50.0          # the fused round kernel dominates a GICP step


def test_bench_line_of_the_round_has_the_contract_keys():
    import json

    d = json.loads(open(os.path.join(ROOT, "profiles", "r1_05_bench_line.json")).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert d["gpu_launches"] > 0 and d["config"]["workload"]


def test_round2_bench_lines_and_traffic_lookup():
    """The committed bench lines of the round's final state (1, 2, 4 and 8 GPUs) carry the contract keys, and bench.py
    finds the dominant kernel's DRAM traffic in the newest committed ncu raw page."""
    import json

    sys.path.insert(0, ROOT)
    import bench

    for name, n in (("r2_02_bench_line.json", 1), ("r2_02_bench_line_2gpu.json", 2), ("r2_02_bench_line_4gpu.json", 4),
                    ("r2_02_bench_line_8gpu.json", 8)):
        d = json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])
        assert d["n_gpus"] == n and d["value"] > 0 and d["gpu_launches"] > 0
        assert d["clocks"]["reasons"] == [] and d["clocks"]["sm_mhz"] > 0
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
        if n == 1:
            assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
            assert d["e2e"]["align"]["matches_oracle_run"] is True
        else:
            assert len(d["e2e"]["multi_gpu"]["per_rank_compute_ms_per_step"]) == n
            assert d["e2e"]["strong_scaling"]["failed_units"] == 0
    traffic, source = bench.ncu_traffic()
    assert source.endswith("r2_02_gicp_persistent_ncu_raw.csv") and 1e8 < traffic < 2e9


def test_clock_sampler_degrades_without_a_gpu():
    """No NVML device and no nvidia-smi (this container): the sampler reports "unsampled" instead of raising."""
    sys.path.insert(0, ROOT)
    import bench

    with bench.ClockSampler(0) as s:
        pass
    out = s.summary()
    assert out["reasons"] == ["unsampled"] or out["sm_mhz"] is not None
