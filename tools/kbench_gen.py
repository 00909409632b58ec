#!/usr/bin/env python3
"""tools/kbench_gen.py [streams ...] — times the general-ratio resampler (44.1 -> 48 kHz and 44.1 -> 16 kHz, stereo f32,
one-shot) for a few batch sizes, device resident, CUDA events.  Development aid (ncu target); bench.py stays the contract."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from soundkit_b200 import capi
    ctx = capi.Context(0)
    ctx.use_torch_stream()
    sizes = [int(a) for a in sys.argv[1:]] or [16, 64, 256]
    reps = int(os.environ.get("KB_REPS", "3"))
    ctx.set_resample_mode(os.environ.get("KB_FAST", "0") == "1")
    out = {}
    for in_rate, out_rate in ((44100, 48000), (44100, 16000)):
        for S in sizes:
            frames = 2646000 if S <= 64 else 661500
            n_out = ctx.batch_plan(in_rate, out_rate, frames, capi.SEM_ONESHOT)
            x = torch.rand((S, frames * 2), device="cuda") - 0.5
            o = torch.empty((S, 2, n_out), dtype=torch.float32, device="cuda")
            d = capi.BatchDesc(x.data_ptr(), x.stride(0) * 4, S, frames, 32, 1, 0, 2, capi.MODE_A, in_rate, out_rate, capi.SEM_ONESHOT,
                               capi.MIX_NONE, o.data_ptr(), 2 * n_out, n_out)
            ctx.normalize_batch_dev(d); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e9
            for _ in range(reps):
                e0.record(); ctx.normalize_batch_dev(d); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) * 1e-3)
            out[f"{in_rate}->{out_rate} S={S}"] = {"TFLOP/s": round(1027.0 * S * 2 * n_out / best / 1e12, 2), "ms": round(best * 1e3, 3)}
            del x, o
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
