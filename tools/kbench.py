#!/usr/bin/env python3
"""tools/kbench.py — times the secondary kernels only (bench.py's `extras`: G.711, conversion, general-ratio
resample), device resident, CUDA events.  Development aid for kernel work; bench.py stays the contract."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from soundkit_b200 import capi
    ctx = capi.Context(0)
    ctx.use_torch_stream()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    only = sys.argv[1:] or None
    res = bench.extras(ctx, torch, peaks, only=only)
    res["ffma_peak_tflops"] = ctx.ffma_peak_tflops()
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
