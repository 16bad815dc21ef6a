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
