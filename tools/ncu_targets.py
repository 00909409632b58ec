#!/usr/bin/env python3
"""tools/ncu_targets.py — launches every HBM-bound / codec kernel of the library exactly once on a > L2 sized input, so that one
`ncu --set full` run of this script yields one capture per kernel (profiles/*_ncu_full_hbm_kernels.csv).  Not a benchmark."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from soundkit_b200 import capi
    ctx = capi.Context(0)
    ctx.use_torch_stream()
    n = 1 << 29
    codes = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda")
    pcm = torch.empty((n,), dtype=torch.int16, device="cuda")
    ctx.g711_decode_dev(capi.LAW_MULAW, codes, n, capi.KIND_I16, pcm)
    ctx.g711_decode_dev(capi.LAW_ALAW, codes, n, capi.KIND_I16, pcm)
    ctx.g711_encode_dev(capi.LAW_MULAW, capi.KIND_I16, pcm, n, codes)
    ctx.g711_encode_dev(capi.LAW_ALAW, capi.KIND_I16, pcm, n, codes)
    f = torch.empty((n // 2,), dtype=torch.float32, device="cuda")
    ctx.g711_decode_dev(capi.LAW_MULAW, codes, n // 2, capi.KIND_F32, f)
    del f, pcm
    S, frames = 32, 2880000
    out = torch.empty((S, 2, frames), dtype=torch.float32, device="cuda")
    for bits, fmt in ((16, 0), (24, 0), (32, 1)):
        raw = torch.randint(0, 256, (S, frames * 2 * bits // 8), dtype=torch.uint8, device="cuda")
        ctx.pcm_to_planar_f32_dev(raw, raw.stride(0), S, frames, bits, fmt, False, 2, capi.MODE_B, out, 2 * frames)
        del raw
    del out
    m = 1 << 28
    raw = torch.randint(0, 256, (m * 4,), dtype=torch.uint8, device="cuda")
    o = torch.empty((m * 4,), dtype=torch.uint8, device="cuda")
    for name, ib in (("s24le_to_i32", 3), ("s16be_to_i16", 2), ("s32le_to_f32", 4), ("f32le_to_i16", 4)):
        ctx.flat_convert_dev(name, raw, m * ib, o)
    fr = m // 2
    planar = torch.rand((2, fr), device="cuda") * 2 - 1
    for bits, fmt in ((16, 0), (24, 0), (32, 1)):
        ctx.planar_f32_to_pcm_dev(planar, fr, 2, bits, fmt, o)
    del raw, o, planar
    S, frames, rate = 32, 2646000, 44100
    raw = torch.randint(0, 256, (S, frames * 4), dtype=torch.uint8, device="cuda")
    n_out = ctx.normalize48k_out_frames(frames, rate)
    plane = (n_out + 3) // 4 * 4
    o = torch.empty((S, 2, plane), dtype=torch.float32, device="cuda")
    ctx.normalize48k_batch_dev(raw, raw.stride(0), S, frames, 16, 0, False, 2, rate, o, 2 * plane, plane)
    del raw, o
    S, k = 65536, 8000
    c6 = torch.randint(0, 256, (S, k // 2), dtype=torch.uint8, device="cuda")
    p6 = torch.empty((S, k), dtype=torch.int16, device="cuda")
    ctx.g726_decode_batch_dev(4, 0, c6, k // 2, S, k // 2, capi.KIND_I16, p6, k)
    ctx.g726_encode_batch_dev(4, 0, capi.KIND_I16, p6, k, S, k, c6, k // 2)
    ctx.g722_decode_batch_dev(c6, k // 2, S, k // 2, capi.KIND_I16, p6, k)
    ctx.g722_encode_batch_dev(capi.KIND_I16, p6, k, S, k, c6, k // 2)
    # multi-channel conversion both ways (shared-memory tile kernels) and the mono chain (conversion pass + planar TMA FIR)
    S, ch, frames = 8, 6, 960000
    raw = torch.randint(0, 256, (S, frames * ch * 2), dtype=torch.uint8, device="cuda")
    o = torch.empty((S, ch, frames), dtype=torch.float32, device="cuda")
    ctx.pcm_to_planar_f32_dev(raw, raw.stride(0), S, frames, 16, 0, False, ch, capi.MODE_B, o, ch * frames)
    b = torch.empty((frames * ch * 2,), dtype=torch.uint8, device="cuda")
    ctx.planar_f32_to_pcm_dev(o[0], frames, ch, 16, 0, b)
    del raw, o, b
    S, frames = 16, 5760000
    raw = torch.randint(0, 256, (S, frames * 2), dtype=torch.uint8, device="cuda")
    n_out = ctx.batch_plan(48000, 16000, frames, capi.SEM_STREAMING)
    o = torch.empty((S, n_out), dtype=torch.float32, device="cuda")
    ctx.normalize_batch_dev(capi.BatchDesc(raw.data_ptr(), raw.stride(0), S, frames, 16, capi.FMT_SIGNED, 0, 1, capi.MODE_B, 48000, 16000, capi.SEM_STREAMING,
                                           capi.MIX_NONE, o.data_ptr(), n_out, 0))
    del raw, o
    torch.cuda.synchronize()
    print("launched", ctx.launches, "kernels")


if __name__ == "__main__":
    main()
