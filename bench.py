#!/usr/bin/env python3
"""bench.py — headline benchmark of the normalization hot path (driver contract in the task brief).

Workload (config.workload): one WAVE of BASELINE.json configs[4] — S streams x 600 s of s24le stereo
48 kHz normalised to f32 mono 16 kHz (decoder convention convert -> windowed-sinc resample per channel
-> mono downmix) by ONE fused sm_100a kernel per step.  configs[4] as a whole (32 768 streams, 6.9 TB)
does not fit a GPU; it is processed in such waves, and at N GPUs every rank runs its own wave with no
collective (weak scaling, independent streams).

  value     output Msamples/s (mono f32 frames), inputs resident in HBM, CUDA-event timed
  e2e       same metric through the host-buffer C-ABI call (sk_normalize_batch): pinned host input,
            H2D + kernel + D2H inside the timed region; beside it the machine's own ceiling for the
            same bytes (concurrent pinned H2D + D2H cudaMemcpyAsync, all ranks at once)
  roofline  FP32 CUDA-core roofline of the dominant kernel (resample_int_tma_kernel<3, 3>), peak = FFMA
            microbenchmark measured in the same run (nominal figure beside it); HBM figures alongside
  parity    outside the timed region: whole streams of this very wave (device path and e2e path) are
            compared bit for bit with the CPU oracle; a mismatch fails the run
  cpu_baseline  the oracle (C restatement of the reference's Rust path, rubato AVX2+FMA order) timed on
            the box's host cores on a bounded sample

`--impl reference` times that CPU path only (rank 0) on the SAME wave (same streams, same lengths).
No NCCL anywhere: ranks meet on a gloo barrier for timing only.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IN_RATE, OUT_RATE = 48000, 16000
BITS, CHANNELS = 24, 2
FLOPS_PER_FRAME = 1024.0  # 2 channels x 256 taps x 2 flop (P = 1: integer t_ratio)
BYTES_PER_FRAME = 22.0    # 3 input frames x 6 B + 4 B out
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # 74.4: SMs x FP32 lanes x 2 flop x max SM clock
METRIC = "output Msamples/s, 48k->16k sinc resample"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=32, help="streams per GPU per step (one wave)")
    ap.add_argument("--seconds", type=float, default=600.0, help="stream length in seconds")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary kernel measurements")
    ap.add_argument("--extras", default=None, help="comma list of extras groups to run (g711,convert,bytes,flat,g726,g722,norm48k,resample,latency,stream)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    return ap.parse_args()


def config_dict(args, frames, n_out, world):
    return {"workload": f"BASELINE configs[4] wave: {args.streams} streams/GPU x {args.seconds:g} s s24le stereo 48 kHz -> "
                        f"f32 mono 16 kHz (fused convert + sinc resample + downmix, streaming semantics)",
            "streams_per_gpu": args.streams, "streams_total": args.streams * world, "frames_per_stream": frames, "out_frames_per_stream": n_out,
            "in_rate": IN_RATE, "out_rate": OUT_RATE, "bits": BITS, "channels": CHANNELS,
            "l2": "inputs larger than L2 (each stream's input alone is %.0f MB)" % (frames * 6 / 1e6),
            "wave_size": "32 streams = 5.5 GB in + 1.2 GB out per GPU; the persistent kernel's rate does not depend on the wave size "
                         "(tiles are drawn by ticket), and the e2e leg needs the same wave in pinned HOST memory on every rank, which is "
                         "what bounds it (8 ranks x 6.8 GB pinned), not the 180 GB of HBM",
            "parallelism": "streams sharded over GPUs, no collective (gloo barrier for timing only)"}


# ------------------------------------------------------------------------------------------- CPU leg
def cpu_wave(threads, n_streams, seconds, seed=0):
    """the oracle chain over `n_streams` whole streams on `threads` host threads (streams dealt round robin, one call per
    stream); returns (mono frames produced, wall seconds).  Synthetic input: one random stream buffer shared by all streams
    (the cost is data independent)."""
    import numpy as np
    from oracle import oracle as O
    O.build()
    frames = int(seconds * IN_RATE)
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 256, frames * CHANNELS * 3, dtype=np.uint8)
    O.normalize_chain_mono(raw[: 4800 * 6], BITS, O.SIGNED, False, CHANNELS, IN_RATE, OUT_RATE)  # warm the table code
    threads = max(1, min(threads, n_streams))
    outs = [0] * threads

    def work(t):
        n = 0
        for _ in range(t, n_streams, threads):
            n += O.normalize_chain_mono(raw, BITS, O.SIGNED, False, CHANNELS, IN_RATE, OUT_RATE).size
        outs[t] = n

    th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for x in th:
        x.start()
    for x in th:
        x.join()
    return sum(outs), time.perf_counter() - t0


def avx_label():
    from oracle import oracle as O
    return "AVX2+FMA" if O.lib().sko_resampler_uses_avx() else "scalar fmaf"


def cpu_baseline(seconds=30.0, target_wall=8.0):
    cores = os.cpu_count() or 1
    frames, wall = cpu_wave(cores, cores, seconds)
    reps = max(1, min(16, int(target_wall / max(wall, 1e-3))))
    if reps > 1:
        frames, wall = cpu_wave(cores, cores * reps, seconds)
    return {"value": frames / wall / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port",
            "sample": f"{cores} threads x {reps} streams x {seconds:g} s s24le stereo 48k -> f32 mono 16k, oracle C restatement "
                      f"(rubato order, {avx_label()}, both interpolation products)",
            "wall_s": wall}


def run_reference(args):
    """the reference's CPU path (its C restatement: the Rust crate cannot be built here) over the SAME wave our arm runs:
    streams_per_gpu x n_gpus whole streams per step on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(1, args.gpus)
    frames = int(args.seconds * IN_RATE)
    cores = os.cpu_count() or 1
    n_streams = args.streams * world
    for _ in range(max(0, args.warmup)):
        cpu_wave(cores, min(cores, n_streams), min(args.seconds, 30.0))  # warm caches / clocks on a short sample
    total_frames, total_wall = 0, 0.0
    for _ in range(max(1, args.steps)):
        f, w = cpu_wave(cores, n_streams, args.seconds)
        total_frames += f
        total_wall += w
    n_out = total_frames // (n_streams * max(1, args.steps))
    val = total_frames / total_wall / 1e6
    sample = (f"each step = the whole wave: {n_streams} streams x {args.seconds:g} s on {min(cores, n_streams)} host threads, one stream per call "
              f"(the Rust reference cannot be built here: no cargo/rustc; this is its C restatement, {avx_label()}, both interpolation "
              f"products as rubato executes them); warm-up steps run a 30 s sample")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Msamples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_wall / max(1, args.steps) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, frames, n_out, world),
            "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        for ln in out.splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------- GPU leg
def extras(ctx, torch, peaks, only=None):
    """secondary kernels (BASELINE configs 1-3 and the SURVEY 8(f) rows), device resident, CUDA-event timed, each > L2 sized.
    Every entry carries the best and the median of its repetitions.  `only`: optional list of group names."""
    import numpy as np
    from soundkit_b200 import capi
    res = {}

    def want(group):
        return only is None or group in only

    def timed(fn, reps=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(reps):
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        return min(ts), statistics.median(ts)

    hbm = peaks.get("hbm_gbs", 6650.0)

    def hbm_entry(nbytes, t, samples):
        best, med = t
        return {"GB/s": nbytes / best / 1e9, "frac_of_measured_hbm": nbytes / best / 1e9 / hbm, "Msamples/s": samples / best / 1e6,
                "median": {"GB/s": nbytes / med / 1e9, "frac_of_measured_hbm": nbytes / med / 1e9 / hbm, "Msamples/s": samples / med / 1e6}}

    def flop_entry(flops, t, samples):
        best, med = t
        return {"TFLOP/s": flops / best / 1e12, "Msamples/s": samples / best / 1e6,
                "median": {"TFLOP/s": flops / med / 1e12, "Msamples/s": samples / med / 1e6}}

    if want("g711"):
        n = 1 << 30  # 1 Gi samples of G.711 (BASELINE configs[2] is 15.7 Gi samples; one wave)
        codes = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda")
        pcm = torch.empty((n,), dtype=torch.int16, device="cuda")
        for law, lname in ((capi.LAW_MULAW, "ulaw"), (capi.LAW_ALAW, "alaw")):
            t = timed(lambda: ctx.g711_decode_dev(law, codes, n, capi.KIND_I16, pcm))
            res[f"g711_decode_{lname}_i16"] = hbm_entry(3 * n, t, n)
            t = timed(lambda: ctx.g711_encode_dev(law, capi.KIND_I16, pcm, n, codes))
            res[f"g711_encode_{lname}_i16"] = hbm_entry(3 * n, t, n)
        del pcm
        m = n // 2
        f = torch.empty((m,), dtype=torch.float32, device="cuda")
        t = timed(lambda: ctx.g711_decode_dev(capi.LAW_MULAW, codes, m, capi.KIND_F32, f))
        res["g711_decode_ulaw_f32"] = hbm_entry(5 * m, t, m)
        del f
        i32 = torch.randint(-2 ** 31, 2 ** 31 - 1, (m,), dtype=torch.int32, device="cuda")
        t = timed(lambda: ctx.g711_encode_dev(capi.LAW_ALAW, capi.KIND_I32, i32, m, codes))
        res["g711_encode_alaw_i32"] = hbm_entry(5 * m, t, m)
        del codes, i32
    if want("convert"):
        # conversion: 64 streams x 60 s stereo 48 kHz -> planar f32 (BASELINE configs[1] wave)
        S, frames = 64, 2880000
        out = torch.empty((S, 2, frames), dtype=torch.float32, device="cuda")
        for bits, fmt, name in ((16, 0, "s16le"), (24, 0, "s24le"), (32, 1, "f32le")):
            raw = torch.randint(0, 256, (S, frames * 2 * bits // 8), dtype=torch.uint8, device="cuda")
            t = timed(lambda: ctx.pcm_to_planar_f32_dev(raw, raw.stride(0), S, frames, bits, fmt, False, 2, capi.MODE_B, out, 2 * frames))
            nb = S * frames * 2 * (bits // 8 + 4)
            res[f"convert_{name}_stereo_to_planar_f32"] = hbm_entry(nb, t, S * frames * 2)
            del raw
        del out
    if want("bytes"):
        # the headline chain with the byte epilogue: s24le stereo 48 kHz -> s16le mono 16 kHz (apply_output_options for ASR), 8 streams x 600 s
        S, frames = 8, 28800000
        n_out = ctx.batch_plan(48000, 16000, frames, capi.SEM_STREAMING)
        stride = (frames * 6 + 15) // 16 * 16
        raw = torch.randint(0, 256, (S, stride), dtype=torch.uint8, device="cuda")
        ostride = (n_out * 2 + 15) // 16 * 16
        o = torch.empty((S, ostride), dtype=torch.uint8, device="cuda")
        d = capi.BatchDesc(raw.data_ptr(), stride, S, frames, 24, capi.FMT_SIGNED, 0, 2, capi.MODE_B, 48000, 16000, capi.SEM_STREAMING,
                           capi.MIX_SCALE_THEN_SUM, 0, 0, 0)
        t = timed(lambda: ctx.normalize_batch_pcm_dev(d, 16, 0, o, ostride), reps=3)
        res["chain_s24_stereo_48k_to_s16_mono_16k_bytes"] = flop_entry(1024.0 * S * n_out, t, S * n_out)
        del raw, o
        # a mono source through the same chain (s16le mono 48 kHz -> f32 mono 16 kHz, 16 streams x 600 s): conversion to planar rows + the
        # planar TMA FIR over pairs of streams
        S = 16
        stride = (frames * 2 + 15) // 16 * 16
        raw = torch.randint(0, 256, (S, stride), dtype=torch.uint8, device="cuda")
        o = torch.empty((S, n_out), dtype=torch.float32, device="cuda")
        d = capi.BatchDesc(raw.data_ptr(), stride, S, frames, 16, capi.FMT_SIGNED, 0, 1, capi.MODE_B, 48000, 16000, capi.SEM_STREAMING,
                           capi.MIX_NONE, o.data_ptr(), n_out, 0)
        t = timed(lambda: ctx.normalize_batch_dev(d), reps=3)
        res["chain_s16_mono_48k_to_f32_mono_16k"] = flop_entry(512.0 * S * n_out, t, S * n_out)
        del raw, o
    if want("flat"):
        # a few of the remaining R1 entry points (soundkit/src/audio_bytes.rs flat functions, f32_channels_to_bytes)
        n = 1 << 28
        raw = torch.randint(0, 256, (n * 4,), dtype=torch.uint8, device="cuda")
        o = torch.empty((n * 4,), dtype=torch.uint8, device="cuda")
        for name, ib, ob in (("s24le_to_i32", 3, 4), ("s16be_to_i16", 2, 2), ("s32le_to_f32", 4, 4), ("f32le_to_i16", 4, 2)):
            t = timed(lambda: ctx.flat_convert_dev(name, raw, n * ib, o))
            res[f"flat_{name}"] = hbm_entry(n * (ib + ob), t, n)
        frames = n // 2
        planar = torch.rand((2, frames), device="cuda") * 2 - 1
        for bits, fmt, nm in ((16, 0, "s16"), (24, 0, "s24"), (32, 1, "f32")):
            t = timed(lambda: ctx.planar_f32_to_pcm_dev(planar, frames, 2, bits, fmt, o))
            res[f"planar_f32_stereo_to_{nm}_bytes"] = hbm_entry(2 * frames * (4 + bits // 8), t, 2 * frames)
        del raw, o, planar
    if want("g726"):
        # G.726-32 decode/encode, 65 536 telephony streams x 4 s (one thread per stream; latency bound)
        S, n = 65536, 32000
        codes = torch.randint(0, 256, (S, n // 2), dtype=torch.uint8, device="cuda")
        pcm = torch.empty((S, n), dtype=torch.int16, device="cuda")
        b, m = timed(lambda: ctx.g726_decode_batch_dev(4, 0, codes, n // 2, S, n // 2, capi.KIND_I16, pcm, n), reps=3)
        res["g726_32k_decode_i16"] = {"Msamples/s": S * n / b / 1e6, "GB/s": S * n * 2.5 / b / 1e9, "streams": S, "median": {"Msamples/s": S * n / m / 1e6}}
        b, m = timed(lambda: ctx.g726_encode_batch_dev(4, 0, capi.KIND_I16, pcm, n, S, n, codes, n // 2), reps=3)
        res["g726_32k_encode_i16"] = {"Msamples/s": S * n / b / 1e6, "GB/s": S * n * 2.5 / b / 1e9, "streams": S, "median": {"Msamples/s": S * n / m / 1e6}}
        del codes, pcm
    if want("g722"):
        # G.722 64 kbit/s decode/encode, 65 536 wideband telephony streams x 1 s (one thread per stream; latency bound)
        S, nb = 65536, 8000
        codes = torch.randint(0, 256, (S, nb), dtype=torch.uint8, device="cuda")
        pcm = torch.empty((S, 2 * nb), dtype=torch.int16, device="cuda")
        b, m = timed(lambda: ctx.g722_decode_batch_dev(codes, nb, S, nb, capi.KIND_I16, pcm, 2 * nb), reps=3)
        res["g722_64k_decode_i16"] = {"Msamples/s": S * 2 * nb / b / 1e6, "GB/s": S * nb * 5 / b / 1e9, "streams": S, "median": {"Msamples/s": S * 2 * nb / m / 1e6}}
        b, m = timed(lambda: ctx.g722_encode_batch_dev(capi.KIND_I16, pcm, 2 * nb, S, 2 * nb, codes, nb), reps=3)
        res["g722_64k_encode_i16"] = {"Msamples/s": S * 2 * nb / b / 1e6, "GB/s": S * nb * 5 / b / 1e9, "streams": S, "median": {"Msamples/s": S * 2 * nb / m / 1e6}}
        del codes, pcm
    if want("norm48k"):
        # StreamingStereo48kNormalizer, 64 whole streams x 60 s s16le stereo 44.1 kHz -> f32 stereo 48 kHz
        S, frames, rate = 64, 2646000, 44100
        raw = torch.randint(0, 256, (S, frames * 4), dtype=torch.uint8, device="cuda")
        n_out = ctx.normalize48k_out_frames(frames, rate)
        plane = (n_out + 3) // 4 * 4
        o = torch.empty((S, 2, plane), dtype=torch.float32, device="cuda")
        t = timed(lambda: ctx.normalize48k_batch_dev(raw, raw.stride(0), S, frames, 16, 0, False, 2, rate, o, 2 * plane, plane))
        res["normalizer48k_s16le_stereo_44k1"] = hbm_entry(S * (frames * 4 + n_out * 8), t, S * n_out * 2)
        del raw, o
    if want("resample"):
        # non-integer ratios, stereo f32 (BASELINE configs[3] wave: 44.1k -> 48k, one-shot and chunk-4096 streaming; plus the
        # common downsampling case 44.1k -> 16k).  P = 2: both interpolation products, 1027 flop per output sample.
        for in_rate, out_rate, sem, tag, fast in ((44100, 48000, capi.SEM_ONESHOT, "", False), (44100, 48000, capi.SEM_STREAMING, "_streaming", False),
                                                  (44100, 16000, capi.SEM_ONESHOT, "", False),
                                                  # the opt-in tolerance mode (sk_ctx_set_resample_mode: one product against the interpolated row per
                                                  # output; <= 1e-5 / >= 120 dB against the bit-exact kernel, asserted in the tests).  TFLOP/s is quoted
                                                  # on the same ALGORITHMIC 1027 flop per sample: the kernel executes half of them.
                                                  (44100, 48000, capi.SEM_ONESHOT, "_fast_mode", True), (44100, 16000, capi.SEM_ONESHOT, "_fast_mode", True)):
            ctx.set_resample_mode(fast)
            S, frames = 64, 2646000  # 64 streams x 2 channels = 128 lanes (configs[3] is 4096 streams: lane columns, no period structure needed)
            n_out = ctx.batch_plan(in_rate, out_rate, frames, sem)
            x = (torch.rand((S, frames * 2), device="cuda") - 0.5)
            o = torch.empty((S, 2, n_out), dtype=torch.float32, device="cuda")
            d = capi.BatchDesc(x.data_ptr(), x.stride(0) * 4, S, frames, 32, 1, 0, 2, capi.MODE_A, in_rate, out_rate, sem,
                               capi.MIX_NONE, o.data_ptr(), 2 * n_out, n_out)
            f0 = ctx.resample_fallback_outputs()
            t = timed(lambda: ctx.normalize_batch_dev(d), reps=3)
            fb = (ctx.resample_fallback_outputs() - f0) / 4.0 / (S * 2 * n_out)
            e = flop_entry(1027.0 * S * 2 * n_out, t, S * 2 * n_out)
            e["fallback_share"] = fb
            e["arithmetic"] = "tolerance mode (1e-5 / 120 dB), algorithmic flops" if fast else "bit-exact (rubato's order)"
            res[f"resample_{in_rate}_to_{out_rate}_stereo_f32{tag}"] = e
            del x, o
        ctx.set_resample_mode(False)
    if want("stream") and hasattr(ctx, "resampler_batch"):
        # batched stateful streaming (sk_resampler_batch_*): N live streams, one 4096-frame chunk each per call, state on the device
        for in_rate, out_rate in ((48000, 16000), (44100, 48000)):
            S, ch, chunks = 4096, 2, 8
            rb = ctx.resampler_batch(in_rate, out_rate, ch, S)
            x = torch.rand((S, ch, 4096), device="cuda") - 0.5
            o = torch.empty((S, ch, rb.out_max), dtype=torch.float32, device="cuda")
            def run():
                for _ in range(chunks):
                    rb.process_dev(x, o)
            t = timed(run, reps=3)
            n_per = 4096.0 * out_rate / in_rate
            P = 1 if in_rate % out_rate == 0 else 2
            res[f"stream_batch_{in_rate}_to_{out_rate}_{S}x{ch}ch_chunk4096"] = flop_entry((512.0 * P + 3) * S * ch * n_per * chunks, t, S * ch * n_per * chunks)
            if hasattr(rb, "process_in_place_dev"):
                # the same pushes with the producer writing in place (sk_resampler_batch_input_dev): no append copy.  The rows keep
                # whatever the last copying push left there; the FIR's cost does not depend on the data
                def run_in_place():
                    for _ in range(chunks):
                        rb.process_in_place_dev(4096, o)
                t = timed(run_in_place, reps=3)
                res[f"stream_batch_{in_rate}_to_{out_rate}_{S}x{ch}ch_chunk4096_in_place"] = flop_entry((512.0 * P + 3) * S * ch * n_per * chunks, t, S * ch * n_per * chunks)
            rb.close()
            del x, o
    if want("latency"):
        res["latency"] = latency_table(ctx, np)
    return res


def latency_table(ctx, np):
    """small-call latency of the host-pointer drop-in API (no CPU fallback: every call is a PCIe round trip) with the CPU port
    beside it: median microseconds per call."""
    from oracle import oracle as O
    from soundkit_b200 import capi
    O.build()
    rng = np.random.default_rng(5)
    out = {}

    def med_us(fn, reps):
        fn(); fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        return statistics.median(ts) * 1e6

    for n, reps in ((160, 200), (4096, 200), (4 << 20, 20)):
        codes = rng.integers(0, 256, n, dtype=np.uint8)
        out[f"g711_decode_ulaw_{n}B"] = {"gpu_us": med_us(lambda: ctx.g711_decode(capi.LAW_MULAW, codes), reps),
                                         "cpu_port_us": med_us(lambda: O.g711_decode(capi.LAW_MULAW, codes), reps)}
    # one 4096-frame stereo chunk through the stateful resampler (StreamingResampler::process granularity)
    x = rng.uniform(-0.5, 0.5, (2, 4096)).astype(np.float32)
    r = ctx.resampler(48000, 16000, 2, 4096)
    ro = O.Resampler(48000, 16000, 4096, 2)
    out["resampler_chunk_4096x2ch_48k_16k"] = {"gpu_us": med_us(lambda: r.process(x), 100), "cpu_port_us": med_us(lambda: ro.process(x), 100)}
    r.close()
    # one 20 ms block (882 frames of s16 stereo 44.1 kHz) through the streaming 48 kHz normalizer (StreamingStereo48kNormalizer::push)
    blk = rng.integers(0, 256, 882 * 4, dtype=np.uint8)
    gn, on = ctx.normalizer48k(), O.Normalizer48k()
    out["normalizer48k_push_882_frames_s16_stereo_44k1"] = {"gpu_us": med_us(lambda: gn.push(blk, 16, 0, False, 2, 44100), 100),
                                                            "cpu_port_us": med_us(lambda: on.push(blk, 16, 0, False, 2, 44100), 100)}
    gn.close()
    # a 1 s utterance (s16le stereo 48 kHz -> f32 mono 16 kHz) through the same call: the short-clip regime of downsample_audio
    clip = rng.integers(-20000, 20000, 48000 * 2).astype(np.int16).view(np.uint8)
    out["downsample_audio_1s_s16_stereo"] = {
        "gpu_us": med_us(lambda: ctx.normalize_stream(clip, 16, 0, False, 2, capi.MODE_A, 48000, 16000, capi.SEM_ONESHOT, capi.MIX_SUM_THEN_SCALE), 100),
        "cpu_port_us": med_us(lambda: O.mixdown_to_mono_f32(O.downsample_audio(clip, 16, 0, 2, 48000, 16000, threads=1)), 20)}
    # BASELINE configs[0] (C1): one 60 s s16le stereo 48 kHz stream -> f32 mono 16 kHz via downsample_audio, single thread
    frames = 2_880_000
    pcm = rng.integers(-20000, 20000, frames * 2).astype(np.int16).view(np.uint8)
    g = med_us(lambda: ctx.normalize_stream(pcm, 16, 0, False, 2, capi.MODE_A, 48000, 16000, capi.SEM_ONESHOT, capi.MIX_SUM_THEN_SCALE), 5)
    c = med_us(lambda: O.mixdown_to_mono_f32(O.downsample_audio(pcm, 16, 0, 2, 48000, 16000, threads=1)), 3)
    out["c1_downsample_audio_60s_s16_stereo"] = {"gpu_us": g, "cpu_port_us": c, "gpu_Msamples/s": 959956 / g, "cpu_port_1thread_Msamples/s": 959956 / c,
                                                  "note": "host buffers in pageable memory both ways (what a Rust Vec gives); CPU port executes both "
                                                          "interpolation products like rubato"}
    return out


def aggregate_extras(per_rank, hbm):
    """sum the per-rank rates of the extras (every rank ran the same kernels on its own GPU at the same time)"""
    if len(per_rank) == 1:
        return per_rank[0]
    agg = {}
    rate_keys = ("GB/s", "Msamples/s", "TFLOP/s")
    for name, e0 in per_rank[0].items():
        if not isinstance(e0, dict) or name == "latency":
            agg[name] = e0
            continue
        e = dict(e0)
        for k in rate_keys:
            if k in e0:
                e[k] = sum(r.get(name, {}).get(k, 0.0) for r in per_rank)
        if "frac_of_measured_hbm" in e0:
            e["frac_of_measured_hbm"] = e["GB/s"] / (hbm * len(per_rank))
        if "median" in e0:
            e["median"] = {k: sum(r.get(name, {}).get("median", {}).get(k, 0.0) for r in per_rank) for k in e0["median"] if k in rate_keys}
        e["per_gpu_min"] = {k: min(r.get(name, {}).get(k, 0.0) for r in per_rank) for k in rate_keys if k in e0}
        agg[name] = e
    return agg


def copy_ceiling(torch, h_in, h_out, d_in, d_out, reps, barrier):
    """what the machine gives for the e2e step's bytes: the same pinned buffers copied H2D and D2H concurrently on two
    streams, nothing else running; all ranks start together.  Returns seconds per repetition (wall clock, synchronised)."""
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def once():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    once(); torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def run_ours(args):
    import numpy as np
    import torch
    from soundkit_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")  # timing barrier + max over ranks only: the data path has no collective, NCCL is not used
    from soundkit_b200 import shard
    numa_node = shard.bind_to_gpu_numa_node(local) if world > 1 else None  # before any pinned allocation
    ctx = capi.Context(local)
    ctx.use_torch_stream()

    # weak scaling: the job is a batch of streams_per_gpu x world streams; rank g owns the contiguous range
    # [g*S/G, (g+1)*S/G) of it and nothing crosses NVLink on the data path
    lo, hi = shard.stream_range(args.streams * world, world, rank)
    S, frames = hi - lo, int(args.seconds * IN_RATE)
    n_out = ctx.batch_plan(IN_RATE, OUT_RATE, frames, capi.SEM_STREAMING)
    stride = frames * CHANNELS * 3
    stride = (stride + 15) // 16 * 16
    g = torch.Generator(device="cuda"); g.manual_seed(shard.stream_seed(0x5EED0000, 4, lo))
    raw = torch.randint(0, 256, (S, stride), dtype=torch.uint8, device="cuda", generator=g)
    n_pad = (n_out + 3) // 4 * 4
    out = torch.empty((S, n_pad), dtype=torch.float32, device="cuda")
    desc = capi.BatchDesc(raw.data_ptr(), stride, S, frames, BITS, capi.FMT_SIGNED, 0, CHANNELS, capi.MODE_B, IN_RATE, OUT_RATE,
                          capi.SEM_STREAMING, capi.MIX_SCALE_THEN_SUM, out.data_ptr(), n_pad, 0)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing
    for _ in range(max(args.warmup, 3)):
        ctx.normalize_batch_dev(desc)
    barrier()
    sampler = ClockSampler(local)
    time.sleep(0.25)
    l0 = ctx.launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for k in range(args.steps):
        ctx.normalize_batch_dev(desc)
        ev[k + 1].record()
    barrier()
    launches = ctx.launches - l0
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    clocks = sampler.stop()
    total_ms_max = shard.max_over_ranks(total_ms, dist)
    frames_all = float(S) * n_out * world * args.steps
    value = frames_all / (total_ms_max * 1e-3) / 1e6

    # ---- end to end through the host-buffer C-ABI call
    h_in = torch.empty((S, stride), dtype=torch.uint8, pin_memory=True)
    h_in.copy_(raw)
    h_out = torch.empty((S, n_pad), dtype=torch.float32, pin_memory=True)
    hdesc = capi.BatchDesc(h_in.data_ptr(), stride, S, frames, BITS, capi.FMT_SIGNED, 0, CHANNELS, capi.MODE_B, IN_RATE, OUT_RATE,
                           capi.SEM_STREAMING, capi.MIX_SCALE_THEN_SUM, h_out.data_ptr(), n_pad, 0)
    e2e_steps = max(2, min(args.steps, 5))
    ctx.normalize_batch_host(hdesc)  # warm-up (allocates the slab pipeline)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.normalize_batch_host(hdesc)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_s = shard.max_over_ranks(e2e_s, dist)
    e2e_val = float(S) * n_out * world * e2e_steps / e2e_s / 1e6
    same = bool(torch.equal(h_out[:, :n_out], out[:, :n_out].cpu()))

    # ---- parity, outside every timed region: whole streams of THIS wave against the CPU oracle (first and last of the rank),
    # both the device-resident result and what the host-buffer call returned
    from oracle import oracle as O
    O.build()
    checked, exact = [], True
    for s in sorted({0, S - 1}):
        want = O.normalize_chain_mono(h_in[s, : frames * 6].numpy(), BITS, O.SIGNED, False, CHANNELS, IN_RATE, OUT_RATE)
        ok = want.size == n_out and np.array_equal(want.view(np.uint32), out[s, :n_out].cpu().numpy().view(np.uint32)) \
            and np.array_equal(want.view(np.uint32), h_out[s, :n_out].numpy().view(np.uint32))
        checked.append(lo + s)
        exact = exact and bool(ok)
    parity = {"oracle": "oracle/ C restatement (bit-for-bit)", "streams_checked": checked, "frames_each": n_out, "bit_exact": exact,
              "e2e_equals_device_path": same}
    if not exact or not same:
        print(json.dumps({"error": "parity failure: the wave's output differs from the oracle", "parity": parity}), flush=True)
        raise SystemExit(3)

    # the machine's ceiling for the same bytes (concurrent H2D + D2H of the pinned buffers, all ranks at once)
    ceil_s = copy_ceiling(torch, h_in, h_out, raw, out, 3, barrier)
    ceil_s = shard.max_over_ranks(ceil_s, dist)
    ceil_val = float(S) * n_out * world / ceil_s / 1e6

    only = [x for x in args.extras.split(",") if x] if args.extras else None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    ext = None
    if not args.no_extras:
        try:
            del raw, out, h_in, h_out
            torch.cuda.empty_cache()
            barrier()
            groups = only if only is not None else ["g711", "convert", "bytes", "flat", "g726", "g722", "norm48k", "resample", "stream", "latency"]
            mine = extras(ctx, torch, peaks, [x for x in groups if rank == 0 or x != "latency"])
        except Exception as e:  # extras must never cost the headline line
            mine = {"error": repr(e)}
        if dist is not None:
            allx = [None] * world
            dist.all_gather_object(allx, mine)
            ext = aggregate_extras(allx, hbm_peak) if rank == 0 else None
        else:
            ext = mine

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    ffma_peak = ctx.ffma_peak_tflops()
    launch_ms = statistics.mean(per_launch_ms)
    kernels_per_step = max(1, launches // max(1, args.steps))
    ach = FLOPS_PER_FRAME * S * n_out / (launch_ms * 1e-3) / 1e12
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (per launch, same workload only)
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        w = tr["workload"]
        if (w["streams_per_gpu"], w["frames_per_stream"], w["bits"], w["channels"]) == (S, frames, BITS, CHANNELS):
            traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]  # bytes per launch; algorithmic: BYTES_PER_FRAME * S * n_out
            traffic_src = tr["source"]
    except Exception:
        pass
    roofline = {"bound": "fp32", "kernel": "resample_int_tma_kernel<3, 3> (fused s24->f32 + 256-tap FIR x2ch + downmix)",
                "achieved": ach, "peak": ffma_peak, "unit": "TFLOP/s", "frac": ach / ffma_peak,
                "peak_source": "FFMA / FFMA2 register microbenchmark measured in this run (sk_ctx_measure_ffma_peak); MEASURED_PEAKS.json has no FP32 "
                               "CUDA-core entry",
                "peak_nominal": FP32_NOMINAL_TFLOPS, "frac_of_nominal": ach / FP32_NOMINAL_TFLOPS,
                "algorithmic_flops_per_launch": FLOPS_PER_FRAME * S * n_out, "launch_ms": launch_ms, "launch_ms_median": statistics.median(per_launch_ms),
                "kernels_per_step": kernels_per_step, "traffic": traffic, "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                "traffic_source": traffic_src, "algorithmic_bytes_per_launch": BYTES_PER_FRAME * S * n_out,
                "hbm": {"achieved_GBs": BYTES_PER_FRAME * S * n_out / (launch_ms * 1e-3) / 1e9, "peak_GBs": hbm_peak,
                        "frac": BYTES_PER_FRAME * S * n_out / (launch_ms * 1e-3) / 1e9 / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"}}
    line = {"metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (uniform random s24 samples, full range)",
            "config": config_dict(args, frames, n_out, world), "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "Msamples/s", "h2d_bytes_per_step": S * frames * CHANNELS * 3,
                    "d2h_bytes_per_step": S * n_out * 4, "steps": e2e_steps, "api": "sk_normalize_batch (host pointers, pinned; 3-deep slab pipeline)",
                    "numa_node": numa_node, "copy_ceiling": {"value": ceil_val, "unit": "Msamples/s",
                                                             "how": "the step's H2D and D2H bytes as two concurrent cudaMemcpyAsync on the same pinned "
                                                                    "buffers, no kernel, all ranks at once",
                                                             "GB/s_h2d_plus_d2h_all_ranks": world * (S * frames * 6 + S * n_out * 4) / ceil_s / 1e9},
                    "frac_of_copy_ceiling": e2e_val / ceil_val},
            "parity": parity, "gpu_launches": launches, "roofline": roofline}
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline()
    if ext is not None:
        line["extras"] = ext
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
