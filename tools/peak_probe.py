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


def measure_dfma(blocks: int = 148 * 16, iters: int = 512, reps: int = 5):
    import torch
    from or_pcd_b200 import ops

    best = 0.0
    ops.dfma_burn(blocks, 16)
    torch.cuda.synchronize()
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sink = ops.dfma_burn(blocks, iters)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, blocks * 256.0 * iters * 16 * 8 * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def measure_mix(alu: int, blocks: int = 148 * 16, iters: int = 2048, reps: int = 3):
    """ms for the FFMA2 stream with `alu` FMNMX3 per FFMA2 interleaved (same FFMA2 count each time)."""
    import torch
    from or_pcd_b200 import _lib

    sink = torch.empty(blocks * 256, dtype=torch.float32, device="cuda")
    best = 1e30
    for _ in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(_lib.lib().orpcd_mix_burn(blocks, iters, alu, sink.data_ptr(), _lib.stream_ptr()), "mix")
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


if __name__ == "__main__":
    out = {"ffma_tflops": measure(False), "ffma2_tflops": measure(True), "dfma_tflops": measure_dfma(),
           "mix_ms_alu0": measure_mix(0), "mix_ms_alu1": measure_mix(1), "mix_ms_alu2": measure_mix(2)}
    print(json.dumps(out))
