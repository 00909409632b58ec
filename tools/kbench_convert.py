#!/usr/bin/env python3
"""tools/kbench_convert.py — times sk_pcm_to_planar_f32_dev (interleaved PCM -> planar f32, decoder convention) over channel counts and
sample formats, device resident, CUDA events; GB/s = bytes read + bytes written.  Development aid."""
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
    for ch in (1, 2, 3, 6, 8):
        for bits, fmt in ((16, 0), (24, 0), (32, 1)):
            S = 16
            frames = 2880000 * 2 // ch // 16 * 16
            raw = torch.randint(0, 256, (S, frames * ch * bits // 8), dtype=torch.uint8, device="cuda")
            o = torch.empty((S, ch, frames), dtype=torch.float32, device="cuda")
            run = lambda: ctx.pcm_to_planar_f32_dev(raw, raw.stride(0), S, frames, bits, fmt, False, ch, capi.MODE_B, o, ch * frames)
            run(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e9
            for _ in range(5):
                e0.record(); run(); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) * 1e-3)
            nbytes = S * frames * ch * (bits // 8 + 4)
            out[f"{'f32' if fmt else 's%d' % bits} {ch}ch"] = {"GB/s": round(nbytes / best / 1e9, 1), "frac_of_6539": round(nbytes / best / 6538.9e9, 3)}
            del raw, o
    # the other direction: planar f32 -> interleaved PCM bytes (f32_channels_to_bytes)
    for ch in (2, 3, 6, 8):
        for bits, fmt in ((16, 0), (24, 0), (32, 1)):
            frames = 2880000 * 16 // ch // 16 * 16
            x = torch.rand((ch, frames), device="cuda") * 2 - 1
            o = torch.empty((frames * ch * bits // 8,), dtype=torch.uint8, device="cuda")
            run = lambda: ctx.planar_f32_to_pcm_dev(x, frames, ch, bits, fmt, o)
            run(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e9
            for _ in range(5):
                e0.record(); run(); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) * 1e-3)
            nbytes = frames * ch * (bits // 8 + 4)
            out[f"planar f32 {ch}ch -> {'f32' if fmt else 's%d' % bits}"] = {"GB/s": round(nbytes / best / 1e9, 1), "frac_of_6539": round(nbytes / best / 6538.9e9, 3)}
            del x, o
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
