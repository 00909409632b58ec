"""Parity at BASELINE.json stream lengths: every config is compared with the CPU oracle bit for bit at its REAL stream
length, so the steady state of the persistent kernels is covered (a CTA of the fused 48 -> 16 kHz kernel walks tens of
1536-output tiles here: mbarrier phase flips, raw-buffer refills and the 4-entry geometry ring for k >= 2), not just the
first two tiles per CTA that short inputs reach.

  C1  60 s s16le stereo 48 kHz -> f32 mono 16 kHz, one-shot (soundkit/src/audio_pipeline.rs:438-493 + mixdown :105-131)
  C3  G.711 mu/A-law, 1024 streams x 240 000 samples in one call (soundkit-g711/src/lib.rs:36-66)
  C4  60 s f32 stereo 44.1 -> 48 kHz, one-shot and chunk-4096 streaming (soundkit-decoder/src/lib.rs:1970-2059)
  C5  600 s s24le stereo 48 kHz -> f32 mono 16 kHz (and s16 bytes), streaming (soundkit-decoder/src/lib.rs:3324-3456)
"""
import concurrent.futures as cf

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from soundkit_b200 import capi  # noqa: E402


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _sines_s16_stereo(frames, rate, seed):
    """SURVEY 8(d) C1 input: 220 Hz + 1 kHz + 6.5 kHz (0.3 / 0.2 / 0.1 FS), per-channel phase, +-1 LSB dither"""
    rng = np.random.default_rng(seed)
    t = np.arange(frames, dtype=np.float64) / rate
    out = np.empty((frames, 2), np.int16)
    for c in range(2):
        ph = rng.uniform(0, 2 * np.pi, 3)
        x = 0.3 * np.sin(2 * np.pi * 220 * t + ph[0]) + 0.2 * np.sin(2 * np.pi * 1000 * t + ph[1]) + 0.1 * np.sin(2 * np.pi * 6500 * t + ph[2])
        out[:, c] = np.clip(np.rint(x * 32767 + rng.uniform(-1, 1, frames)), -32768, 32767).astype(np.int16)
    return out.reshape(-1)


def test_c1_oneshot_60s_s16_stereo_to_mono(ctx, oracle):
    """BASELINE configs[0]: the reference's own CPU-runnable case, 2 880 000 frames -> 959 956 frames, 625 tiles"""
    frames = 2_880_000
    pcm = _sines_s16_stereo(frames, 48000, 0x5EED0000 + 1)
    data = pcm.view(np.uint8)
    got = ctx.normalize_stream(data, 16, capi.FMT_SIGNED, False, 2, capi.MODE_A, 48000, 16000, capi.SEM_ONESHOT, capi.MIX_SUM_THEN_SCALE)
    planar = oracle.downsample_audio(data, 16, 0, 2, 48000, 16000, threads=4)
    want = oracle.mixdown_to_mono_f32(planar)
    assert got.shape == (1, 959_956) and want.shape == (959_956,)
    assert np.array_equal(_bits(got[0]), _bits(want))
    # and the planar form downsample_audio itself returns
    got2 = ctx.normalize_stream(data, 16, capi.FMT_SIGNED, False, 2, capi.MODE_A, 48000, 16000, capi.SEM_ONESHOT, capi.MIX_NONE)
    assert np.array_equal(_bits(got2), _bits(planar))


def test_c3_g711_1024_streams_x_30s_one_call(ctx, oracle):
    import torch
    n = 1024 * 240_000
    g = torch.Generator(device="cuda"); g.manual_seed(0x5EED0000 + 2 * (1 << 20))
    codes = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda", generator=g)
    pcm = torch.randint(-32768, 32768, (n,), dtype=torch.int16, device="cuda", generator=g)
    # one stream enumerates every code / every sample value
    codes[:256] = torch.arange(256, dtype=torch.uint8, device="cuda")
    pcm[:65536] = torch.arange(-32768, 32768, dtype=torch.int32, device="cuda").to(torch.int16)
    h_codes, h_pcm = codes.cpu().numpy(), pcm.cpu().numpy()
    c = capi.Context(0)
    c.use_torch_stream()
    try:
        out_pcm = torch.empty((n,), dtype=torch.int16, device="cuda")
        out_codes = torch.empty((n,), dtype=torch.uint8, device="cuda")
        for law in (capi.LAW_MULAW, capi.LAW_ALAW):
            c.g711_decode_dev(law, codes, n, capi.KIND_I16, out_pcm)
            c.g711_encode_dev(law, capi.KIND_I16, pcm, n, out_codes)
            torch.cuda.synchronize()
            assert np.array_equal(out_pcm.cpu().numpy(), oracle.g711_decode(law, h_codes))
            assert np.array_equal(out_codes.cpu().numpy(), oracle.g711_encode(law, h_pcm))
    finally:
        c.close()


@pytest.mark.parametrize("semantics", ["oneshot", "streaming"])
def test_c4_60s_f32_stereo_44k1_to_48k(ctx, oracle, semantics):
    """BASELINE configs[3] at its stream length: white noise in [-0.5, 0.5] on one stream, 10 kHz / 18 kHz sines on another
    (index drift shows up as phase error there).  One-shot: 2 879 859 frames per channel."""
    import torch
    frames = 2_646_000
    rng = np.random.default_rng(0x5EED0000 + 3 * (1 << 20))
    t = np.arange(frames, dtype=np.float64) / 44100.0
    noise = rng.uniform(-0.5, 0.5, (frames, 2)).astype(np.float32)
    tones = np.stack([0.9 * np.sin(2 * np.pi * 10000 * t), 0.9 * np.sin(2 * np.pi * 18000 * t + 0.3)], axis=1).astype(np.float32)
    streams = [noise.reshape(-1), tones.reshape(-1)]
    sem = capi.SEM_ONESHOT if semantics == "oneshot" else capi.SEM_STREAMING
    c = capi.Context(0)
    c.use_torch_stream()
    try:
        n_out = c.batch_plan(44100, 48000, frames, sem)
        if semantics == "oneshot":
            assert n_out == 2_879_859
        x = torch.from_numpy(np.stack(streams)).cuda()
        o = torch.empty((2, 2, n_out), dtype=torch.float32, device="cuda")
        d = capi.BatchDesc(x.data_ptr(), x.stride(0) * 4, 2, frames, 32, capi.FMT_FLOAT, 0, 2, capi.MODE_A, 44100, 48000, sem, capi.MIX_NONE,
                           o.data_ptr(), 2 * n_out, n_out)
        before = c.resample_fallback_outputs()
        c.normalize_batch_dev(d)
        torch.cuda.synchronize()
        fallback = c.resample_fallback_outputs() - before
        got = o.cpu().numpy()
    finally:
        c.close()

    def want_of(s):
        if semantics == "oneshot":
            return oracle.downsample_audio(streams[s].view(np.uint8), 32, 1, 2, 44100, 48000, threads=4)
        st = oracle.Stream(44100, 48000, 2)
        planar = np.ascontiguousarray(streams[s].reshape(-1, 2).T)
        a = st.process(planar)
        return np.concatenate([a, st.flush()], axis=1)

    for s in range(2):
        want = want_of(s)
        assert want.shape == (2, n_out)
        assert np.array_equal(_bits(got[s]), _bits(want)), f"stream {s} differs"
    # outputs that left the tiled fast path (non-uniform rows where the f64 walk lands within rounding of an integer,
    # subindex 255, ragged ends) still took the exact per-output arithmetic; they must stay a small minority
    total = 2 * 2 * n_out
    assert 0 <= fallback <= total // 20, (fallback, total)
    print(f"[c4 {semantics}] fallback outputs {fallback} of {total} ({100.0 * fallback / total:.3f} %)")


def _c5_inputs(S, frames):
    import torch
    stride = (frames * 6 + 15) // 16 * 16
    g = torch.Generator(device="cuda"); g.manual_seed(0x5EED0000 + 4 * (1 << 20))
    raw = torch.randint(0, 256, (S, stride), dtype=torch.uint8, device="cuda", generator=g)
    return raw, stride


def test_c5_wave_4_streams_x_600s_bit_exact(oracle):
    """BASELINE configs[4] wave at the real stream length: 4 x 28 800 000 frames s24le stereo -> 4 x 9 599 956 mono frames =
    25 000 tiles of the persistent TMA kernel, ~84 per CTA.  Every stream is compared in full, f32 and s16-byte epilogue."""
    import torch
    S, frames = 4, 28_800_000
    c = capi.Context(0)
    c.use_torch_stream()
    try:
        raw, stride = _c5_inputs(S, frames)
        n_out = c.batch_plan(48000, 16000, frames, capi.SEM_STREAMING)
        assert n_out == 9_599_956
        n_pad = (n_out + 3) // 4 * 4
        out = torch.zeros((S, n_pad), dtype=torch.float32, device="cuda")
        d = capi.BatchDesc(raw.data_ptr(), stride, S, frames, 24, capi.FMT_SIGNED, 0, 2, capi.MODE_B, 48000, 16000, capi.SEM_STREAMING,
                           capi.MIX_SCALE_THEN_SUM, out.data_ptr(), n_pad, 0)
        c.normalize_batch_dev(d)
        ostride = (n_out * 2 + 15) // 16 * 16
        ob = torch.zeros((S, ostride), dtype=torch.uint8, device="cuda")
        c.normalize_batch_pcm_dev(d, 16, capi.FMT_SIGNED, ob, ostride)
        torch.cuda.synchronize()
        got = out.cpu().numpy()[:, :n_out]
        got_b = ob.cpu().numpy()[:, : n_out * 2]
        host = raw.cpu().numpy()
    finally:
        c.close()

    def one(s):
        want = oracle.normalize_chain_mono(host[s, : frames * 6], 24, 0, False, 2, 48000, 16000)
        return want, oracle.f32_channels_to_bytes(want[None, :], 16, 0)

    with cf.ThreadPoolExecutor(4) as ex:
        wants = list(ex.map(one, range(S)))
    for s, (want, want_b) in enumerate(wants):
        assert want.shape == (n_out,)
        assert np.array_equal(_bits(got[s]), _bits(want)), f"stream {s} f32 differs"
        assert np.array_equal(got_b[s], np.frombuffer(want_b, np.uint8)), f"stream {s} s16 bytes differ"


def test_c5_host_batch_pipeline_matches_oracle(oracle):
    """sk_normalize_batch (host pointers, 3-deep slab pipeline: what bench.py's e2e number goes through) on 5 x 120 s streams:
    more slabs than pipeline buffers, so slab buffers are reused while earlier copies are still in flight."""
    import torch
    S, frames = 5, 5_760_000
    c = capi.Context(0)
    try:
        stride = frames * 6
        rng = np.random.default_rng(77)
        h_in = torch.from_numpy(rng.integers(0, 256, (S, stride), dtype=np.uint8)).pin_memory()
        n_out = c.batch_plan(48000, 16000, frames, capi.SEM_STREAMING)
        n_pad = (n_out + 3) // 4 * 4
        h_out = torch.zeros((S, n_pad), dtype=torch.float32).pin_memory()
        d = capi.BatchDesc(h_in.data_ptr(), stride, S, frames, 24, capi.FMT_SIGNED, 0, 2, capi.MODE_B, 48000, 16000, capi.SEM_STREAMING,
                           capi.MIX_SCALE_THEN_SUM, h_out.data_ptr(), n_pad, 0)
        c.set_batch_slab_bytes(40 << 20)  # one 34.6 MB stream per slab: 5 slabs through 3 buffers
        c.normalize_batch_host(d)
        got = h_out.numpy()[:, :n_out]
        host = h_in.numpy()
        for s in range(S):
            want = oracle.normalize_chain_mono(host[s], 24, 0, False, 2, 48000, 16000)
            assert np.array_equal(_bits(got[s]), _bits(want)), f"stream {s} differs"
    finally:
        c.close()


@pytest.mark.parametrize("rates", [(44100, 48000), (44100, 16000), (48000, 44100), (8000, 16000), (32000, 48000), (22050, 16000)])
@pytest.mark.parametrize("semantics", ["oneshot", "streaming"])
def test_general_ratio_many_lanes_bit_exact(oracle, rates, semantics):
    """>= 64 (stream, channel) lanes in one call: the packed 4 x 2 kernel runs with lane columns only (no period structure, no
    fallback lanes).  41 stereo streams = 82 lanes: one full block of 64 and one ragged block of 18."""
    import torch
    i, o = rates
    S, frames = 41, 30011
    rng = np.random.default_rng(i * 7 + o)
    x = rng.uniform(-0.5, 0.5, (S, frames * 2)).astype(np.float32)
    sem = capi.SEM_ONESHOT if semantics == "oneshot" else capi.SEM_STREAMING
    c = capi.Context(0)
    c.use_torch_stream()
    try:
        n_out = c.batch_plan(i, o, frames, sem)
        xd = torch.from_numpy(x).cuda()
        od = torch.zeros((S, 2, n_out), dtype=torch.float32, device="cuda")
        d = capi.BatchDesc(xd.data_ptr(), xd.stride(0) * 4, S, frames, 32, capi.FMT_FLOAT, 0, 2, capi.MODE_A, i, o, sem, capi.MIX_NONE,
                           od.data_ptr(), 2 * n_out, n_out)
        f0 = c.resample_fallback_outputs()
        c.normalize_batch_dev(d)
        torch.cuda.synchronize()
        fb = c.resample_fallback_outputs() - f0
        got = od.cpu().numpy()
    finally:
        c.close()
    for s in (0, 1, 31, 32, 40):
        if semantics == "oneshot":
            want = oracle.downsample_audio(x[s].view(np.uint8), 32, 1, 2, i, o)
        else:
            st = oracle.Stream(i, o, 2)
            want = np.concatenate([st.process(np.ascontiguousarray(x[s].reshape(-1, 2).T)), st.flush()], axis=1)
        assert want.shape == (2, n_out)
        assert np.array_equal(_bits(got[s]), _bits(want)), f"stream {s} differs"
    print(f"[many lanes {i}->{o} {semantics}] fallback outputs {fb} of {S * 2 * n_out}")


def _rubato_cases():
    import glob
    import os
    from conftest import GOLDEN
    return sorted(glob.glob(os.path.join(GOLDEN, "rubato", "*.out.f32")))


@pytest.mark.skipif(not _rubato_cases(), reason="no rubato 0.14.1 dump under tests/golden/rubato (needs a Rust toolchain: see its README)")
@pytest.mark.parametrize("path", _rubato_cases() or ["none"])
def test_cuda_kernels_against_real_rubato_dump(ctx, path):
    """the CUDA path against real rubato 0.14.1 output (rust/rubato-dump): north-star contract 1e-5 / 120 dB; reports bit-equality"""
    import json
    meta = json.load(open(path.replace(".out.f32", ".json")))
    ch = meta["channels"]
    x = np.fromfile(path.replace(".out.f32", ".in.f32"), "<f4").reshape(ch, -1)
    want = np.fromfile(path, "<f4").reshape(ch, -1)
    inter = np.ascontiguousarray(x.T).reshape(-1)
    sem = capi.SEM_ONESHOT if meta["semantics"] == "oneshot" else capi.SEM_STREAMING
    got = ctx.normalize_stream(inter.view(np.uint8), 32, capi.FMT_FLOAT, False, ch, capi.MODE_A, meta["in_rate"], meta["out_rate"], sem, capi.MIX_NONE)
    assert got.shape == want.shape
    err = np.max(np.abs(got.astype(np.float64) - want))
    noise = np.sum((got.astype(np.float64) - want) ** 2)
    snr = np.inf if noise == 0 else 10 * np.log10(np.sum(want.astype(np.float64) ** 2) / noise)
    assert err <= 1e-5 and snr >= 120.0, (err, snr)
    print("bit-identical to rubato:", bool(np.array_equal(_bits(got), _bits(want))))


@pytest.mark.parametrize("rates", [(44100, 48000), (44100, 16000), (48000, 44100), (8000, 16000), (32000, 48000), (22050, 16000)])
@pytest.mark.parametrize("semantics", ["oneshot", "streaming"])
def test_fast_resample_mode_stays_inside_the_north_star_contract(oracle, rates, semantics):
    """sk_ctx_set_resample_mode(SK_RESAMPLE_FAST): one product against the interpolated table row per output instead of rubato's
    two — opt-in, lane columns only.  Gate: max-abs <= 1e-5 of full scale and SNR >= 120 dB against the bit-exact kernel AND against
    the CPU oracle, full-scale noise and near-Nyquist tones; batches below 64 lanes must stay bit-exact even with the mode on."""
    import torch
    i, o = rates
    S, frames = 40, 50021
    rng = np.random.default_rng(i + o)
    t = np.arange(frames) / i
    x = rng.uniform(-1.0, 1.0, (S, frames, 2)).astype(np.float32)
    x[1, :, 0] = 0.99 * np.sin(2 * np.pi * 0.47 * min(i, o) * t)
    x[1, :, 1] = 0.99 * np.sin(2 * np.pi * 997.0 * t)
    x = x.reshape(S, -1)
    sem = capi.SEM_ONESHOT if semantics == "oneshot" else capi.SEM_STREAMING
    c = capi.Context(0)
    c.use_torch_stream()
    try:
        n_out = c.batch_plan(i, o, frames, sem)
        xd = torch.from_numpy(x).cuda()
        res = {}
        for fast in (False, True):
            c.set_resample_mode(fast)
            od = torch.zeros((S, 2, n_out), dtype=torch.float32, device="cuda")
            d = capi.BatchDesc(xd.data_ptr(), xd.stride(0) * 4, S, frames, 32, capi.FMT_FLOAT, 0, 2, capi.MODE_A, i, o, sem, capi.MIX_NONE,
                               od.data_ptr(), 2 * n_out, n_out)
            c.normalize_batch_dev(d)
            torch.cuda.synchronize()
            res[fast] = od.cpu().numpy()
        # fewer than 64 lanes: the fast kernel does not apply, the result is the bit-exact one
        od = torch.zeros((3, 2, n_out), dtype=torch.float32, device="cuda")
        d = capi.BatchDesc(xd.data_ptr(), xd.stride(0) * 4, 3, frames, 32, capi.FMT_FLOAT, 0, 2, capi.MODE_A, i, o, sem, capi.MIX_NONE,
                           od.data_ptr(), 2 * n_out, n_out)
        c.normalize_batch_dev(d)
        torch.cuda.synchronize()
        small = od.cpu().numpy()
    finally:
        c.close()
    exact, fast = res[False].astype(np.float64), res[True].astype(np.float64)
    assert np.array_equal(_bits(small), _bits(res[False][:3]))
    err = np.max(np.abs(exact - fast))
    snr = 10 * np.log10(np.sum(exact ** 2) / max(np.sum((exact - fast) ** 2), 1e-300))
    assert err <= 1e-5 and snr >= 120.0, (err, snr)
    for s in (0, 1, 39):
        if semantics == "oneshot":
            want = oracle.downsample_audio(x[s].view(np.uint8), 32, 1, 2, i, o)
        else:
            st = oracle.Stream(i, o, 2)
            want = np.concatenate([st.process(np.ascontiguousarray(x[s].reshape(-1, 2).T)), st.flush()], axis=1)
        assert np.array_equal(_bits(res[False][s]), _bits(want))
        w = want.astype(np.float64)
        assert np.max(np.abs(w - fast[s])) <= 1e-5
        assert 10 * np.log10(np.sum(w ** 2) / max(np.sum((w - fast[s]) ** 2), 1e-300)) >= 120.0
    print(f"[fast mode {i}->{o} {semantics}] max-abs {err:.2e}, SNR {snr:.1f} dB vs the bit-exact kernel")
