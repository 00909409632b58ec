#!/usr/bin/env python3
"""tools/kbench_n48.py — times the 48 kHz stereo normalizer (sk_normalize48k_batch_dev) over sample formats and source rates,
device resident, CUDA events; GB/s = algorithmic bytes (source bytes read + 8 B written per output frame).  Development aid."""
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
    for bits, fmt, ch, rate in ((16, 0, 2, 44100), (16, 0, 1, 22050), (16, 0, 2, 48000), (16, 0, 2, 96000), (24, 0, 2, 44100), (32, 0, 2, 44100),
                                (32, 1, 2, 44100), (32, 1, 2, 96000), (16, 0, 2, 8000)):
        S, seconds = 32, 60
        frames = rate * seconds
        fb = bits // 8 * ch
        stride = (frames * fb + 15) // 16 * 16
        raw = torch.randint(0, 256, (S, stride), dtype=torch.uint8, device="cuda")
        if fmt == 1:
            raw = (torch.rand((S, stride // 4), device="cuda") * 2 - 1).view(torch.uint8)
        n_out = ctx.normalize48k_out_frames(frames, rate)
        plane = (n_out + 3) // 4 * 4
        o = torch.empty((S, 2, plane), dtype=torch.float32, device="cuda")
        run = lambda: ctx.normalize48k_batch_dev(raw, stride, S, frames, bits, fmt, False, ch, rate, o, 2 * plane, plane)
        run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(5):
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3)
        nbytes = S * (frames * fb + n_out * 8)
        out[f"{'f32' if fmt else 's%d' % bits} {ch}ch {rate}"] = {"GB/s": round(nbytes / best / 1e9, 1), "frac_of_6539": round(nbytes / best / 6538.9e9, 3),
                                                                    "G frames/s": round(S * n_out / best / 1e9, 1)}
        del raw, o
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
