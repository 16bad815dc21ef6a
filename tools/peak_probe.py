"""Measure the FP32 FMA-pipe peak with scalar FFMA and packed FFMA2 (CUDA events)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-scale-pointcloud-registration_b200"))


def measure(packed: bool, blocks: int = 148 * 16, iters: int = 4096, reps: int = 5):
    import torch
    from or_pcd_b200 import ops

    best = 0.0
    for _ in range(3):
        ops.ffma_burn(blocks, 64, packed)
    torch.cuda.synchronize()
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sink = ops.ffma_burn(blocks, iters, packed)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        flops = blocks * 256.0 * iters * 16 * 8 * (4 if packed else 2)
        best = max(best, flops / (ms * 1e-3) / 1e12)
    return best


if __name__ == "__main__":
    out = {"ffma_tflops": measure(False), "ffma2_tflops": measure(True)}
    print(json.dumps(out))
