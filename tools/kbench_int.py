#!/usr/bin/env python3
"""tools/kbench_int.py — times the integer-step fused chain (s24le stereo -> f32 mono, streaming semantics) for every integer
ratio the library instantiates, device resident, CUDA events.  Development aid; bench.py stays the contract."""
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
    out = {}
    only = os.environ.get("KB_ONLY")  # e.g. "48000-8000": one ratio (ncu target)
    ch, bits = int(os.environ.get("KB_CH", "2")), int(os.environ.get("KB_BITS", "24"))  # source shape (default: the headline's s24 stereo)
    mix = capi.MIX_SCALE_THEN_SUM if ch > 1 else capi.MIX_NONE
    for in_rate, out_rate in ((48000, 24000), (48000, 16000), (48000, 12000), (48000, 8000), (96000, 8000), (32000, 16000), (44100, 22050)):
        if only and only != f"{in_rate}-{out_rate}":
            continue
        S, frames = 16, in_rate * 120
        n_out = ctx.batch_plan(in_rate, out_rate, frames, capi.SEM_STREAMING)
        fb = ch * bits // 8
        stride = (frames * fb + 15) // 16 * 16
        raw = torch.randint(0, 256, (S, stride), dtype=torch.uint8, device="cuda")
        o = torch.empty((S, n_out), dtype=torch.float32, device="cuda")
        d = capi.BatchDesc(raw.data_ptr(), stride, S, frames, bits, capi.FMT_SIGNED, 0, ch, capi.MODE_B, in_rate, out_rate, capi.SEM_STREAMING,
                           mix, o.data_ptr(), n_out, 0)
        ctx.normalize_batch_dev(d); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(5):
            e0.record(); ctx.normalize_batch_dev(d); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3)
        out[f"{in_rate}->{out_rate}"] = {"TFLOP/s": round(512.0 * ch * S * n_out / best / 1e12, 2), "G frames/s": round(S * n_out / best / 1e9, 2),
                                         "GB/s": round((S * frames * fb + S * n_out * 4) / best / 1e9, 1), "ms": round(best * 1e3, 3)}
        del raw, o
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
