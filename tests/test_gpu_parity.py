"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs and against the reference's golden fixtures.  Integer / byte work must be
bit-exact; the resampler is compared bit-for-bit too (the kernels keep rubato's accumulation order),
with the north-star tolerance (max-abs 1e-5 FS, SNR >= 120 dB) asserted as the contract."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden

pytestmark = pytest.mark.gpu

from soundkit_b200 import capi  # noqa: E402

RNG = np.random.default_rng(0x5EED0000)


def s24_bytes(vals, big_endian=False):
    v = np.asarray(vals, np.int64) & 0xFFFFFF
    cols = [(v >> 16) & 0xFF, (v >> 8) & 0xFF, v & 0xFF] if big_endian else [v & 0xFF, (v >> 8) & 0xFF, (v >> 16) & 0xFF]
    return np.stack(cols, axis=1).astype(np.uint8).reshape(-1)


def snr_db(ref, got):
    ref = ref.astype(np.float64); got = got.astype(np.float64)
    noise = np.sum((ref - got) ** 2)
    return np.inf if noise == 0 else 10 * np.log10(np.sum(ref ** 2) / noise)


def assert_resample_close(want, got):
    assert want.shape == got.shape
    if want.size == 0:
        return
    assert np.max(np.abs(want.astype(np.float64) - got.astype(np.float64))) <= 1e-5   # north-star tolerance
    assert snr_db(want, got) >= 120.0
    # and in fact identical: same table, same index walk, same accumulation order
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))


# ---------------------------------------------------------------------------------------------- G.711
@pytest.mark.parametrize("law", [capi.LAW_MULAW, capi.LAW_ALAW])
def test_g711_all_codes_and_all_samples(ctx, oracle, law):
    codes = np.arange(256, dtype=np.uint8)
    for kind, name in ((capi.KIND_I16, "i16"), (capi.KIND_I32, "i32"), (capi.KIND_F32, "f32")):
        got = ctx.g711_decode(law, codes, kind)
        want = oracle.g711_decode(law, codes, name)
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8))
    samples = np.arange(-32768, 32768).astype(np.int16)
    assert np.array_equal(ctx.g711_encode(law, samples), oracle.g711_encode(law, samples))
    s32 = RNG.integers(-2 ** 31, 2 ** 31, 70001, dtype=np.int64).astype(np.int32)
    assert np.array_equal(ctx.g711_encode(law, s32), oracle.g711_encode(law, s32))


@pytest.mark.parametrize("law,stem", [(capi.LAW_MULAW, "g711_ulaw.ulaw"), (capi.LAW_ALAW, "g711_alaw.alaw")])
def test_g711_golden_fixture(ctx, law, stem):
    raw = golden(stem)
    want = np.frombuffer(golden(stem.split(".")[0] + ".decoded.s16le"), "<i2")
    assert np.array_equal(ctx.g711_decode(law, raw), want)
    # chunked like the reference test (137-byte chunks, soundkit-g711/src/lib.rs:384): chunking cannot matter
    parts = [ctx.g711_decode(law, raw[i:i + 137]) for i in range(0, 137 * 5, 137)]
    assert np.array_equal(np.concatenate(parts), want[:137 * 5])


@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 4099, 1 << 20])
def test_g711_ragged_sizes(ctx, oracle, n):
    codes = RNG.integers(0, 256, n, dtype=np.uint8)
    pcm = RNG.integers(-32768, 32768, n).astype(np.int16)
    for law in (capi.LAW_MULAW, capi.LAW_ALAW):
        assert np.array_equal(ctx.g711_decode(law, codes), oracle.g711_decode(law, codes))
        assert np.array_equal(ctx.g711_encode(law, pcm), oracle.g711_encode(law, pcm))


def test_g711_idempotent_requantisation_full_size(ctx):
    """size-independent property at a large size: dec(enc(dec(c))) == dec(c) (soundkit-g711/src/lib.rs:417-434)"""
    n = 1 << 24
    codes = RNG.integers(0, 256, n, dtype=np.uint8)
    for law in (capi.LAW_MULAW, capi.LAW_ALAW):
        pcm = ctx.g711_decode(law, codes)
        again = ctx.g711_decode(law, ctx.g711_encode(law, pcm))
        assert np.array_equal(pcm, again)


# ---------------------------------------------------------------------------------------------- flat ops
def _flat_input(name, n):
    if name.startswith(("i16", "s16")):
        return RNG.integers(0, 256, 2 * n, dtype=np.uint8)
    if name.startswith("s24"):
        return RNG.integers(0, 256, 3 * n, dtype=np.uint8)
    if name.startswith("s32"):
        return RNG.integers(0, 256, 4 * n, dtype=np.uint8)
    # f32 sources: [-1.25, 1.25] plus NaN, +-Inf, -0.0 and exact +-1
    x = RNG.uniform(-1.25, 1.25, n).astype(np.float32)
    x[::97] = np.nan; x[1::89] = np.inf; x[2::83] = -np.inf; x[3::79] = -0.0; x[4::73] = 1.0; x[5::71] = -1.0
    b = x.view(np.uint8).reshape(-1, 4)
    if "be" in name.split("_")[0]:
        b = b[:, ::-1]
    return np.ascontiguousarray(b).reshape(-1)


@pytest.mark.parametrize("name", capi.FLAT_OPS)
@pytest.mark.parametrize("n", [0, 3, 16, 1000, 65537])
def test_flat_ops_bit_exact(ctx, oracle, name, n):
    data = _flat_input(name, n)
    got, want = ctx.flat_convert(name, data), oracle.flat(name, data)
    assert got.dtype == want.dtype and got.shape == want.shape
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8))


def test_flat_partial_trailing_sample_is_dropped(ctx, oracle):
    data = RNG.integers(0, 256, 3 * 100 + 2, dtype=np.uint8)
    assert np.array_equal(ctx.flat_convert("s24le_to_i32", data), oracle.flat("s24le_to_i32", data))


def test_endian_swap_matches_reference_hashes(ctx):
    """CAF/MOV PCM payloads: BE->LE canonicalisation must hash to what the reference asserts
    (soundkit-audio-demux/src/caf.rs:777-842, src/lib.rs:5568-5690)."""
    man = json.load(open(os.path.join(GOLDEN, "MANIFEST.json")))["endian_payloads"]
    for name, info in man.items():
        raw = golden(name)
        kind, be = info["kind"], info["big_endian"]
        if kind == "s16":
            le = ctx.flat_convert("s16be_to_i16" if be else "s16le_to_i16", raw).astype("<i2").tobytes()
        elif kind == "s24":
            # through the planar path and back: s24 -> f32 (mode B) is exact, f32 -> s24 is not the inverse;
            # use the i32 flat op for LE and narrow/compare for BE via 16-bit
            if be:
                v = ctx.flat_convert("s24be_to_i16", raw)
                want = ctx.flat_convert("s24le_to_i16", golden("pcm-s24le.mov.bin"))
                assert np.array_equal(v, want)
                continue
            v = ctx.flat_convert("s24le_to_i32", raw)
            le = s24_bytes(v).tobytes()
        elif kind == "s32":
            le = ctx.flat_convert("s32be_to_i32" if be else "s32le_to_i32", raw).astype("<i4").tobytes()
        else:
            pl = ctx.pcm_to_planar_f32(raw, 32, capi.FMT_FLOAT, be, 2, capi.MODE_B)
            le = np.ascontiguousarray(pl.T).astype("<f4").tobytes()
        assert hashlib.sha256(le).hexdigest() == info["le_sha256"], name


# ---------------------------------------------------------------------------------------------- planar
@pytest.mark.parametrize("bits,fmt", [(16, 0), (24, 0), (32, 0), (32, 1)])
@pytest.mark.parametrize("channels", [1, 2, 3, 4, 6, 8])  # 3 .. 16 channels: the shared-memory tile kernel (4 / 8: its swizzled form)
@pytest.mark.parametrize("mode", [capi.MODE_A, capi.MODE_B])
@pytest.mark.parametrize("be", [False, True])
def test_pcm_to_planar_f32(ctx, oracle, bits, fmt, channels, mode, be):
    for frames in (0, 1, 7, 64, 4099, 4100, 66000):  # 4100 / 66000: vector kernels (frames % 4 == 0) + scalar tail
        if fmt == 1:
            data = _flat_input("f32be" if be else "f32le", frames * channels)
        else:
            data = RNG.integers(0, 256, frames * channels * bits // 8, dtype=np.uint8)
        got = ctx.pcm_to_planar_f32(data, bits, fmt, be, channels, mode)
        want = oracle.pcm_to_planar_f32(data, bits, fmt, be, channels, mode)
        assert got.shape == want.shape
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_pcm_to_planar_mode_b_rejects_partial_frame(ctx):
    with pytest.raises(capi.SoundkitError, match="partial frame"):
        ctx.pcm_to_planar_f32(np.zeros(7, np.uint8), 16, 0, False, 2, capi.MODE_B)


@pytest.mark.parametrize("kind", ["i16", "s24", "f32"])
@pytest.mark.parametrize("channels", [1, 2, 5, 8, 16, 17])
def test_deinterleave_typed(ctx, oracle, kind, channels):
    bps = {"i16": 2, "s24": 3, "f32": 4}[kind]
    for frames in (0, 3, 8, 1001, 40000):
        data = RNG.integers(0, 256, frames * channels * bps + (1 if frames else 0), dtype=np.uint8)  # + partial frame
        got, want = ctx.deinterleave(kind, data, channels), oracle.deinterleave(kind, data, channels)
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8))


def test_interleave_i16_and_stereo_to_mono(ctx, oracle):
    p = RNG.integers(-32768, 32768, (3, 1001)).astype(np.int16)
    assert np.array_equal(ctx.interleave_i16(p), oracle.interleave_vecs_i16(p))
    s = RNG.integers(-32768, 32768, 2 * 5003).astype(np.int16)
    assert np.array_equal(ctx.stereo_to_mono_i16(s, False), oracle.stereo_to_mono("take_left", s))
    assert np.array_equal(ctx.stereo_to_mono_i16(s, True), oracle.stereo_to_mono("avg", s))


@pytest.mark.parametrize("bits,fmt", [(16, 0), (24, 0), (32, 0), (32, 1)])
@pytest.mark.parametrize("channels", [1, 2, 3, 4, 6, 8, 17])  # 3 .. 16 channels: the shared-memory tile kernel
def test_planar_f32_to_pcm(ctx, oracle, bits, fmt, channels):
    for frames in (0, 5, 64, 4100, 33001):
        x = np.frombuffer(_flat_input("f32le", frames * channels).tobytes(), np.float32).reshape(channels, frames).copy()
        if frames > 10:
            x[0, :8] = [0.5 / 32767, 1.5 / 32767, -0.5 / 32768, -1.5 / 32768, 2.5 / 32767, 1e-30, -1e-30, 0.99999994]
        got, want = ctx.planar_f32_to_pcm(x, bits, fmt), oracle.f32_channels_to_bytes(x, bits, fmt)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("bits", [24, 32])
@pytest.mark.parametrize("be", [False, True])
def test_pcm_narrow_i16(ctx, oracle, bits, be):
    data = RNG.integers(0, 256, (bits // 8) * 10007, dtype=np.uint8)
    assert np.array_equal(ctx.pcm_narrow_i16(data, bits, be), oracle.exact_signed_pcm_to_i16(data, bits, be))


@pytest.mark.parametrize("channels", [1, 2, 3, 6])
def test_mixdown_orders(ctx, oracle, channels):
    x = RNG.uniform(-1, 1, (channels, 10007)).astype(np.float32)
    x[0, 0] = -0.0
    got = ctx.mixdown_mono(x, capi.MIX_SUM_THEN_SCALE)
    assert np.array_equal(got.view(np.uint32), oracle.mixdown_to_mono_f32(x).view(np.uint32))
    got = ctx.mixdown_mono(x, capi.MIX_SCALE_THEN_SUM)
    assert np.array_equal(got.view(np.uint32), oracle.downmix_channels(x, 1)[0].view(np.uint32))


@pytest.mark.parametrize("channels", [3, 5, 6, 8])
@pytest.mark.parametrize("amp", [0.3, 1.0])
def test_downmix_stereo_with_peak_normalisation(ctx, oracle, channels, amp):
    x = (amp * RNG.uniform(-1, 1, (channels, 5001))).astype(np.float32)
    got, want = ctx.downmix_stereo(x), oracle.downmix_channels(x, 2)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


# ---------------------------------------------------------------------------------------------- resampler
RATES_INT = [(48000, 16000), (16000, 8000), (32000, 8000), (48000, 8000), (96000, 8000)]
RATES_GEN = [(44100, 48000), (44100, 16000), (8000, 16000), (16000, 44100), (88200, 8000), (48000, 44100)]


def test_sinc_table_identical(ctx, oracle):
    for i, o in ((48000, 16000), (44100, 48000), (8000, 16000)):
        r = ctx.resampler(i, o, 1, 1024)
        t = oracle.Resampler(i, o, 1024, 1).table()
        assert np.array_equal(r.table().view(np.uint32), t.view(np.uint32))
        r.close()


@pytest.mark.parametrize("rates", RATES_INT + RATES_GEN)
@pytest.mark.parametrize("channels", [1, 2, 3])
def test_stateful_resampler_chunks(ctx, oracle, rates, channels):
    """SincFixedIn semantics chunk by chunk: same per-chunk lengths, same last_index, same samples."""
    i, o = rates
    chunk = 4096
    r = ctx.resampler(i, o, channels, chunk)
    ro = oracle.Resampler(i, o, chunk, channels)
    for k in range(4):
        x = RNG.uniform(-0.5, 0.5, (channels, chunk)).astype(np.float32)
        got, want = r.process(x), ro.process(x)
        assert_resample_close(want, got)
        assert r.last_index == ro.last_index
    tail = RNG.uniform(-0.5, 0.5, (channels, 1000)).astype(np.float32)
    assert_resample_close(ro.process_partial(tail), r.process_partial(tail))
    assert_resample_close(ro.process_partial(None), r.process_partial(None))
    r.close()


def test_stateful_resampler_index_walk_known_answers(ctx):
    """outputs per 4096-frame chunk and last_index after six chunks."""
    table = {(48000, 16000): ([1322, 1365, 1365, 1366, 1365, 1365], -260.0),
             (44100, 48000): ([4317, 4458, 4459, 4458, 4458, 4458], -257.9000000017045),
             (8000, 16000): ([7932, 8192, 8192, 8192, 8192, 8192], -258.0)}
    for (i, o), (counts, last) in table.items():
        r = ctx.resampler(i, o, 1, 4096)
        x = np.zeros((1, 4096), np.float32)
        assert [r.process(x).shape[1] for _ in range(6)] == counts
        assert r.last_index == last
        r.close()


@pytest.mark.parametrize("bits,fmt", [(16, 0), (24, 0), (32, 0), (32, 1)])
@pytest.mark.parametrize("rates", [(48000, 16000), (44100, 48000), (48000, 8000)])
def test_downsample_audio_oneshot(ctx, oracle, bits, fmt, rates):
    """soundkit::audio_pipeline::downsample_audio: mode-A convert + one rubato chunk, per channel."""
    i, o = rates
    for channels, frames in ((2, 30011), (1, 5000), (3, 2048)):
        if fmt == 1:
            data = RNG.uniform(-1, 1, frames * channels).astype(np.float32).view(np.uint8)
        else:
            data = RNG.integers(0, 256, frames * channels * bits // 8, dtype=np.uint8)
        got = ctx.normalize_stream(data, bits, fmt, False, channels, capi.MODE_A, i, o, capi.SEM_ONESHOT, capi.MIX_NONE)
        want = oracle.downsample_audio(data, bits, fmt, channels, i, o)
        assert_resample_close(want, got)


def test_oneshot_lengths_known_answers(ctx):
    for (i, o, n), want in {(44100, 16000, 132437): 48002, (48000, 16000, 2880000): 959956,
                            (44100, 48000, 2646000): 2879859, (16000, 8000, 47360): 23615,
                            (48000, 16000, 28800000): 9599956}.items():
        assert ctx.batch_plan(i, o, n, capi.SEM_ONESHOT) == want
    # streaming total == one-shot total (soundkit-decoder/src/lib.rs:5187-5238)
    assert ctx.batch_plan(44100, 16000, 132437, capi.SEM_STREAMING) == 48002


@pytest.mark.parametrize("rates", [(48000, 16000), (44100, 16000), (16000, 8000)])
@pytest.mark.parametrize("bits,be", [(16, False), (16, True), (24, False), (24, True), (32, False)])
def test_streaming_chain_to_mono(ctx, oracle, rates, bits, be):
    """apply_output_options' rate-changing arm (mode-B convert, chunk-4096 streaming resampler with
    flush + trim, downmix_channels to mono) for a whole stream, fused on the device."""
    i, o = rates
    for channels, frames in ((2, 50000), (2, 8192), (1, 9000), (4, 6000)):
        data = RNG.integers(0, 256, frames * channels * bits // 8, dtype=np.uint8)
        got = ctx.normalize_stream(data, bits, 0, be, channels, capi.MODE_B, i, o, capi.SEM_STREAMING, capi.MIX_SCALE_THEN_SUM)
        want = oracle.normalize_chain_mono(data, bits, 0, be, channels, i, o)
        assert got.shape[0] == 1
        assert_resample_close(want, got[0])


def test_float_source_with_nonfinite_and_huge_values(ctx, oracle):
    """mode A passes NaN/Inf/huge floats into the FIR: the integer-step fast path must still equal the
    reference's p0 + 0*(p1 - p0)."""
    frames = 20000
    x = RNG.uniform(-1, 1, frames * 2).astype(np.float32)
    x[5000] = np.inf; x[9001] = np.nan; x[12000] = 3e38; x[12003] = -3e38; x[15000] = 1e35
    got = ctx.normalize_stream(x.view(np.uint8), 32, 1, False, 2, capi.MODE_A, 48000, 16000, capi.SEM_ONESHOT, capi.MIX_NONE)
    want = oracle.downsample_audio(x.view(np.uint8), 32, 1, 2, 48000, 16000)
    assert got.shape == want.shape
    assert np.array_equal(np.isnan(got), np.isnan(want))
    m = ~np.isnan(want)
    assert np.array_equal(got[m].view(np.uint32), want[m].view(np.uint32))


def test_batch_dev_matches_per_stream_and_linearity():
    """device-resident batch entry point (what bench.py times): every stream of a batch equals the
    single-stream result, and the FIR is linear: resample(a) + resample(b) ~= resample(a + b)."""
    import torch
    ctx = capi.Context(0)
    ctx.use_torch_stream()
    frames, S = 48000, 5
    raw = torch.randint(0, 256, (S, frames * 2 * 3), dtype=torch.uint8, device="cuda")
    n = ctx.batch_plan(48000, 16000, frames, capi.SEM_STREAMING)
    out = torch.empty((S, n), dtype=torch.float32, device="cuda")
    d = capi.BatchDesc(raw.data_ptr(), raw.stride(0), S, frames, 24, 0, 0, 2, capi.MODE_B, 48000, 16000, capi.SEM_STREAMING,
                       capi.MIX_SCALE_THEN_SUM, out.data_ptr(), out.stride(0), 0)
    ctx.normalize_batch_dev(d)
    torch.cuda.synchronize()
    host = raw.cpu().numpy()
    for s in range(S):
        one = ctx.normalize_stream(host[s], 24, 0, False, 2, capi.MODE_B, 48000, 16000, capi.SEM_STREAMING, capi.MIX_SCALE_THEN_SUM)[0]
        assert np.array_equal(out[s].cpu().numpy().view(np.uint32), one.view(np.uint32))
    # linearity on planar-equivalent f32 input
    a = torch.rand((1, frames * 2), device="cuda") - 0.5
    b = torch.rand((1, frames * 2), device="cuda") - 0.5
    res = []
    for x in (a, b, a + b):
        o = torch.empty((1, 2, n), dtype=torch.float32, device="cuda")
        d = capi.BatchDesc(x.data_ptr(), x.stride(0) * 4, 1, frames, 32, 1, 0, 2, capi.MODE_A, 48000, 16000, capi.SEM_STREAMING,
                           capi.MIX_NONE, o.data_ptr(), 2 * n, n)
        ctx.normalize_batch_dev(d)
        res.append(o)
    torch.cuda.synchronize()
    assert torch.max(torch.abs(res[0] + res[1] - res[2])).item() < 1e-5
    ctx.close()


@pytest.mark.parametrize("channels,bits,fmt", [(1, 16, 0), (1, 24, 0), (1, 32, 1), (3, 16, 0), (6, 24, 0), (6, 32, 1)])
@pytest.mark.parametrize("rates", [(48000, 16000), (48000, 8000), (32000, 16000)])
@pytest.mark.parametrize("frames", [20000, 20011])
def test_batch_dev_mono_and_multichannel_two_pass_chain(oracle, channels, bits, fmt, rates, frames):
    """mono and > 2-channel sources at integer ratios take two passes (interleaved PCM -> planar f32 rows, then the persistent
    TMA FIR for planar rows; mono streams are convolved in pairs, an odd last stream alone): every stream of a batch must equal
    the oracle bit for bit, unmixed and mixed down, both semantics"""
    import torch
    i, o = rates
    c = capi.Context(0)
    c.use_torch_stream()
    S = 5  # two pairs + a lone stream for mono; 20011 frames: the planes of the row scratch are padded to a 16-byte multiple
    fb = channels * bits // 8
    stride = (frames * fb + 15) // 16 * 16
    rng = np.random.default_rng(channels * 100 + bits + i)
    host = np.zeros((S, stride), np.uint8)
    for st in range(S):
        if fmt == 1:
            host[st, : frames * fb] = rng.uniform(-1, 1, frames * channels).astype(np.float32).view(np.uint8)
        else:
            host[st, : frames * fb] = rng.integers(0, 256, frames * fb, dtype=np.uint8)
    d_in = torch.from_numpy(host).cuda()
    for sem in (capi.SEM_ONESHOT, capi.SEM_STREAMING):
        n = c.batch_plan(i, o, frames, sem)
        plane = (n + 3) // 4 * 4
        # unmixed: one plane per channel (downsample_audio / the streaming resampler per channel)
        out = torch.full((S, channels, plane), float("nan"), dtype=torch.float32, device="cuda")
        c.normalize_batch_dev(capi.BatchDesc(d_in.data_ptr(), stride, S, frames, bits, fmt, 0, channels, capi.MODE_A, i, o, sem, capi.MIX_NONE,
                                             out.data_ptr(), channels * plane, plane))
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        for st in range(S):
            data = host[st, : frames * fb]
            if sem == capi.SEM_ONESHOT:
                want = oracle.downsample_audio(data, bits, fmt, channels, i, o)
            else:
                planar = oracle.pcm_to_planar_f32(data, bits, fmt, False, channels, capi.MODE_A)
                r = oracle.Stream(i, o, channels)
                want = np.concatenate([r.process(planar), r.flush()], axis=1)
            assert want.shape == (channels, n)
            assert np.array_equal(got[st, :, :n].view(np.uint32), want.view(np.uint32)), (sem, st)
        if channels > 1 and fmt == 0:
            # mixed down to mono (apply_output_options' arm): decoder convention, scale-then-sum
            mono = torch.full((S, plane), float("nan"), dtype=torch.float32, device="cuda")
            c.normalize_batch_dev(capi.BatchDesc(d_in.data_ptr(), stride, S, frames, bits, fmt, 0, channels, capi.MODE_B, i, o, capi.SEM_STREAMING,
                                                 capi.MIX_SCALE_THEN_SUM, mono.data_ptr(), plane, 0))
            torch.cuda.synchronize()
            gm = mono.cpu().numpy()
            nm = c.batch_plan(i, o, frames, capi.SEM_STREAMING)
            for st in (0, S - 1):
                want = oracle.normalize_chain_mono(host[st, : frames * fb], bits, 0, False, channels, i, o)
                assert np.array_equal(gm[st, :nm].view(np.uint32), want.view(np.uint32)), st
    c.close()


# ---------------------------------------------------------------------------------------------- 8(f) row 2: 48 kHz stereo normalizer
def _n48_source(bits, fmt, be, channels, frames, rng):
    if fmt == 1:
        x = rng.uniform(-1.3, 1.3, frames * channels).astype(np.float32)
        x[::53] = np.nan; x[1::59] = np.inf; x[2::61] = -np.inf; x[3::67] = -0.0
        b = x.view(np.uint8).reshape(-1, 4)
        return np.ascontiguousarray(b[:, ::-1] if be else b).reshape(-1)
    n = bits // 8
    raw = rng.integers(0, 256, frames * channels * n, dtype=np.uint8)
    raw[: 2 * n] = 0x7f if not be else raw[: 2 * n]  # a couple of near-full-scale samples
    return raw


@pytest.mark.parametrize("bits,fmt", [(16, 0), (24, 0), (32, 0), (32, 1)])
@pytest.mark.parametrize("be", [False, True])
@pytest.mark.parametrize("channels", [1, 2, 6])
@pytest.mark.parametrize("rate", [8000, 44100, 48000, 96000])
def test_normalizer48k_streaming_matches_oracle(ctx, oracle, bits, fmt, be, channels, rate):
    rng = np.random.default_rng(bits * 1000 + channels * 10 + rate % 97 + int(be))
    frames = 1237
    data = _n48_source(bits, fmt, be, channels, frames, rng)
    fb = (bits // 8) * channels
    cuts = [0, 1, 2, 37, 400, 401, 1000, frames]  # includes 1-frame pushes (Ok(None) for most rates)
    gn, on = ctx.normalizer48k(), oracle.Normalizer48k()
    for a, b in zip(cuts[:-1], cuts[1:]):
        gl, gr = gn.push(data[a * fb:b * fb], bits, fmt, be, channels, rate)
        ol, orr = on.push(data[a * fb:b * fb], bits, fmt, be, channels, rate)
        assert np.array_equal(gl.view(np.uint32), ol.view(np.uint32)) and np.array_equal(gr.view(np.uint32), orr.view(np.uint32)), (a, b)
        assert gn.source_frames() == on.source_frames() and gn.output_frames() == on.output_frames()
    gl, gr = gn.finish()
    ol, orr = on.finish()
    assert np.array_equal(gl.view(np.uint32), ol.view(np.uint32)) and np.array_equal(gr.view(np.uint32), orr.view(np.uint32))
    assert gn.output_frames() == on.output_frames() == (frames if rate == 48000 else max(1, int(np.floor(frames * 48000.0 / rate + 0.5))))
    gn.close()


def test_normalizer48k_error_states(ctx):
    n = ctx.normalizer48k()
    with pytest.raises(capi.SoundkitError, match="decoded source contained no PCM"):
        n.finish()
    n = ctx.normalizer48k()
    n.push(np.zeros(8, np.uint8), 16, 0, False, 2, 8000)
    with pytest.raises(capi.SoundkitError, match=r"geometry changed from 8000 Hz/2 ch to 16000 Hz/2 ch"):
        n.push(np.zeros(8, np.uint8), 16, 0, False, 2, 16000)
    with pytest.raises(capi.SoundkitError, match="incomplete PCM frame"):
        n.push(np.zeros(7, np.uint8), 16, 0, False, 2, 8000)
    with pytest.raises(capi.SoundkitError, match="unsupported PCM bit depth 8"):
        n.push(np.zeros(8, np.uint8), 8, 0, False, 2, 8000)
    n.finish()
    with pytest.raises(capi.SoundkitError, match="already finished"):
        n.push(np.zeros(8, np.uint8), 16, 0, False, 2, 8000)


@pytest.mark.parametrize("bits,fmt,channels,rate,be", [
    (16, 0, 2, 44100, False), (24, 0, 2, 96000, False), (32, 1, 1, 16000, False), (16, 0, 2, 48000, False),
    # every sample format through whole ("plain") tiles of the 16-outputs-per-thread form, both byte orders
    (16, 0, 2, 44100, True), (16, 0, 1, 22050, False), (16, 0, 1, 32000, True), (24, 0, 2, 44100, False), (24, 0, 2, 32000, True),
    (24, 0, 1, 8000, False), (32, 0, 2, 44100, False), (32, 0, 2, 24000, True), (32, 1, 2, 44100, True), (32, 1, 2, 48000, False),
    # runs too wide for 16 outputs per thread (4-output form) and for any staged form (direct kernel)
    (32, 1, 2, 96000, False), (32, 0, 2, 88200, True), (32, 1, 6, 96000, False)])
def test_normalizer48k_batch_dev_matches_oracle_per_stream(oracle, bits, fmt, channels, rate, be):
    """device-resident batch of whole streams == push(whole) + finish() per stream; also the chunking-independence
    property of soundkit/src/audio_pipeline.rs:806-823 at a larger size (streams of 30 011 frames)"""
    import torch
    c = capi.Context(0)
    c.use_torch_stream()
    S, frames = 5, 30011
    rng = np.random.default_rng(rate + bits)
    fb = (bits // 8) * channels
    stride = (frames * fb + 15) // 16 * 16
    host = np.zeros((S, stride), np.uint8)
    for s in range(S):
        host[s, : frames * fb] = _n48_source(bits, fmt, be, channels, frames, rng)
    d_in = torch.from_numpy(host).cuda()
    n_out = c.normalize48k_out_frames(frames, rate)
    plane = (n_out + 3) // 4 * 4
    d_out = torch.full((S, 2, plane), float("nan"), dtype=torch.float32, device="cuda")
    got_n = c.normalize48k_batch_dev(d_in, stride, S, frames, bits, fmt, be, channels, rate, d_out, 2 * plane, plane)
    torch.cuda.synchronize()
    assert got_n == n_out
    out = d_out.cpu().numpy()
    for s in range(S):
        wl, wr = oracle.normalize48k([(host[s, : frames * fb], bits, fmt, be, channels, rate)])
        assert len(wl) == n_out
        assert np.array_equal(out[s, 0, :n_out].view(np.uint32), wl.view(np.uint32))
        assert np.array_equal(out[s, 1, :n_out].view(np.uint32), wr.view(np.uint32))
    # chunked == whole through the stateful handle
    nz = c.normalizer48k()
    ls = []
    for a in range(0, frames, 4097):
        ls.append(nz.push(host[0, a * fb:min(frames, a + 4097) * fb], bits, fmt, be, channels, rate)[0].copy())
    ls.append(nz.finish()[0].copy())
    assert np.array_equal(np.concatenate(ls).view(np.uint32), out[0, 0, :n_out].view(np.uint32))
    c.close()


# ---------------------------------------------------------------------------------------------- 8(f) row 3: unpackers
@pytest.mark.parametrize("bits,channels,nbytes", [(16, 2, 0), (16, 2, 4 * 1001), (16, 6, 12 * 333), (24, 1, 12 * 7), (24, 2, 12 * 4099),
                                                   (24, 8, 24 * 1000), (24, 3, 36 * 501), (24, 4, 12 * 65537)])
def test_dvd_lpcm_unpack_bit_exact(ctx, oracle, bits, channels, nbytes):
    packed = RNG.integers(0, 256, nbytes, dtype=np.uint8)
    assert np.array_equal(ctx.dvd_lpcm_unpack(packed, bits, channels), oracle.unpack_dvd_lpcm(packed, bits, channels))


def test_dvd_lpcm_unpack_errors(ctx):
    with pytest.raises(capi.SoundkitError, match="DVD LPCM ends with a partial sample block"):
        ctx.dvd_lpcm_unpack(np.zeros(6, np.uint8), 16, 2)
    with pytest.raises(capi.SoundkitError, match="DVD LPCM ends with a partial 24-bit sample block"):
        ctx.dvd_lpcm_unpack(np.zeros(12, np.uint8), 24, 8)
    with pytest.raises(capi.SoundkitError, match="20-bit DVD LPCM decoding is not implemented"):
        ctx.dvd_lpcm_unpack(np.zeros(12, np.uint8), 20, 2)


def test_dvd_lpcm24_is_an_involution_free_permutation_at_full_size(ctx):
    """size-independent property on 48 MiB: the unpacker is a byte permutation (multiset preserved per group) and its
    output reinterpreted as s24le equals (hi << 16 | mid << 8 | low) of the packed fields"""
    n = 4 << 20  # groups
    packed = RNG.integers(0, 256, n * 12, dtype=np.uint8)
    out = ctx.dvd_lpcm_unpack(packed, 24, 2).reshape(n, 4, 3)
    g = packed.reshape(n, 12)
    assert np.array_equal(out[:, :, 2], g[:, 0:8:2]) and np.array_equal(out[:, :, 1], g[:, 1:8:2]) and np.array_equal(out[:, :, 0], g[:, 8:12])


# ---------------------------------------------------------------------------------------------- 8(f) row 4: G.726 ADPCM
G726_RATES = [("16", 2), ("24", 3), ("32", 4), ("40", 5)]


@pytest.mark.parametrize("suffix,bits", G726_RATES)
def test_g726_golden_fixture_decode(ctx, suffix, bits):
    """the reference's own fixtures: testdata/g726/*.g726 -> golden/g726/*.decoded.wav, pushed in 127-byte chunks"""
    fixture = np.frombuffer(golden(f"g726_{suffix}.g726"), np.uint8)
    want = np.frombuffer(golden(f"g726_{suffix}.decoded.s16le"), "<i2")
    gs, gb = capi.G726_GROUP[bits]
    st, pending, out = ctx.g726_state(), np.zeros(0, np.uint8), []
    for i in range(0, fixture.size, 127):
        buf = np.concatenate([pending, fixture[i:i + 127]])
        whole = buf.size // gb * gb
        out.append(ctx.g726_decode(bits, capi.G726_LEFT, st, buf[:whole]))
        pending = buf[whole:]
    assert pending.size == 0
    assert np.array_equal(np.concatenate(out), want)


@pytest.mark.parametrize("suffix,bits", G726_RATES)
@pytest.mark.parametrize("packing", [0, 1])
def test_g726_encode_decode_match_oracle(ctx, oracle, suffix, bits, packing):
    gs, gb = capi.G726_GROUP[bits]
    speech = np.frombuffer(golden("linear16_8.s16le"), "<i2")
    noise = RNG.integers(-32768, 32768, 8 * 1000).astype(np.int16)  # full-range noise drives the clamps and the transition detector
    for x in (speech[: speech.size // 8 * 8], noise):
        want_codes = oracle.G726(bits, packing).encode(x)
        st = ctx.g726_state()
        got_codes = ctx.g726_encode(bits, packing, st, x)
        assert np.array_equal(got_codes, want_codes)
        og = oracle.G726(bits, packing)
        og.encode(x)
        assert np.array_equal(st, og.state)  # the carried predictor state too
        for kind, name in ((capi.KIND_I16, "i16"), (capi.KIND_I32, "i32"), (capi.KIND_F32, "f32")):
            got = ctx.g726_decode(bits, packing, ctx.g726_state(), want_codes, kind)
            ref = oracle.G726(bits, packing).decode(want_codes)
            if name == "i32":
                ref = ref.astype(np.int32) << 16
            elif name == "f32":
                ref = ref.astype(np.float32) / np.float32(32768.0)
            assert np.array_equal(got.view(np.uint8), ref.view(np.uint8)), name
    # random code streams (not produced by any encoder) exercise states an encoder never reaches
    codes = RNG.integers(0, 256, gb * 4001, dtype=np.uint8)
    assert np.array_equal(ctx.g726_decode(bits, packing, ctx.g726_state(), codes), oracle.G726(bits, packing).decode(codes))
    # i32 input: (s >> 16) as i16
    x32 = RNG.integers(-2 ** 31, 2 ** 31, gs * 500, dtype=np.int64).astype(np.int32)
    assert np.array_equal(ctx.g726_encode(bits, packing, ctx.g726_state(), x32), oracle.G726(bits, packing).encode((x32 >> 16).astype(np.int16)))


@pytest.mark.parametrize("suffix,bits", G726_RATES)
@pytest.mark.parametrize("packing", [0, 1])
def test_g726_batch_dev_every_layout_takes_the_same_values(oracle, suffix, bits, packing):
    """the batch kernels have three code paths per stream — whole 32-sample blocks of code words, the per-sample tail, and the
    all-tail form of a stream that does not start on a word / vector boundary — and three output kinds: every combination of
    stream stride (aligned, +1, +2 bytes), output stride (aligned, odd) and a length that is not a multiple of 32 samples must
    give the oracle's values, in both directions"""
    import torch
    c = capi.Context(0)
    c.use_torch_stream()
    gs, gb = capi.G726_GROUP[bits]
    S = 37
    groups = (32 * 5 + 3 * gs) // gs + 1   # a few whole blocks plus a tail that is not a block
    nbytes, n = groups * gb, groups * gs
    for in_pad, out_pad in ((0, 0), (1, 1), (2, 3), (0, 1), (3, 0)):
        stride = (nbytes + 15) // 16 * 16 + in_pad
        codes = RNG.integers(0, 256, (S, stride), dtype=np.uint8)
        d_codes = torch.from_numpy(codes).cuda()
        ostride = (n + 7) // 8 * 8 + out_pad
        want = [oracle.G726(bits, packing).decode(codes[s, :nbytes]) for s in range(S)]
        for kind, dt in ((capi.KIND_I16, torch.int16), (capi.KIND_I32, torch.int32), (capi.KIND_F32, torch.float32)):
            d_out = torch.zeros((S, ostride), dtype=dt, device="cuda")
            c.g726_decode_batch_dev(bits, packing, d_codes, stride, S, nbytes, kind, d_out, ostride)
            torch.cuda.synchronize()
            out = d_out.cpu().numpy()
            for s in range(S):
                ref = want[s]
                if kind == capi.KIND_I32:
                    ref = ref.astype(np.int32) << 16
                elif kind == capi.KIND_F32:
                    ref = ref.astype(np.float32) / np.float32(32768.0)
                assert np.array_equal(out[s, :n].view(np.uint8), ref.view(np.uint8)), (in_pad, out_pad, kind, s)
                assert not out[s, n:].any()  # nothing written past the stream's samples
        # encode: i16 and i32 samples at the same odd strides
        x = RNG.integers(-32768, 32768, (S, ostride)).astype(np.int16)
        for kind in (capi.KIND_I16, capi.KIND_I32):
            d_x = torch.from_numpy(x if kind == capi.KIND_I16 else (x.astype(np.int32) << 16) | 0x1234).cuda()
            d_enc = torch.zeros((S, stride), dtype=torch.uint8, device="cuda")
            c.g726_encode_batch_dev(bits, packing, kind, d_x, ostride, S, n, d_enc, stride)
            torch.cuda.synchronize()
            enc = d_enc.cpu().numpy()
            for s in range(S):
                assert np.array_equal(enc[s, :nbytes], oracle.G726(bits, packing).encode(x[s, :n])), (in_pad, out_pad, kind, s)
                assert not enc[s, nbytes:].any()
    c.close()


def test_g726_argument_errors(ctx):
    st = ctx.g726_state()
    with pytest.raises(capi.SoundkitError, match="whole number of 3-byte code groups"):
        ctx.g726_decode(3, 0, st, np.zeros(4, np.uint8))
    with pytest.raises(capi.SoundkitError, match="whole number of 8-sample code groups"):
        ctx.g726_encode(5, 0, st, np.zeros(12, np.int16))
    with pytest.raises(capi.SoundkitError, match="code size must be 2, 3, 4 or 5"):
        ctx.g726_decode(6, 0, st, np.zeros(4, np.uint8))


@pytest.mark.parametrize("suffix,bits", G726_RATES)
def test_g726_batch_dev_streams_are_independent_and_chunkable(oracle, suffix, bits):
    """device-resident batch: every stream equals the oracle on its own data; splitting the batch call in two with the
    carried state equals the single call (chunk-boundary invariance, the property soundkit-g726's streaming tests pin)"""
    import torch
    c = capi.Context(0)
    c.use_torch_stream()
    gs, gb = capi.G726_GROUP[bits]
    S, groups = 300, 257
    nbytes, n = groups * gb, groups * gs
    stride = (nbytes + 15) // 16 * 16 + 16
    codes = RNG.integers(0, 256, (S, stride), dtype=np.uint8)
    d_codes = torch.from_numpy(codes).cuda()
    ostride = (n + 7) // 8 * 8
    d_out = torch.zeros((S, ostride), dtype=torch.int16, device="cuda")
    c.g726_decode_batch_dev(bits, 0, d_codes, stride, S, nbytes, capi.KIND_I16, d_out, ostride)
    torch.cuda.synchronize()
    out = d_out.cpu().numpy()
    for s in (0, 1, 31, 32, 150, S - 1):
        assert np.array_equal(out[s, :n], oracle.G726(bits).decode(codes[s, :nbytes])), s
    # two calls with carried state == one call
    d_state = torch.zeros((S, 24), dtype=torch.int32, device="cuda")
    d_state[:] = torch.from_numpy(c.g726_state()).cuda()
    half = (groups // 2) * gb
    d_out2 = torch.zeros_like(d_out)
    c.g726_decode_batch_dev(bits, 0, d_codes, stride, S, half, capi.KIND_I16, d_out2, ostride, d_state)
    c.g726_decode_batch_dev(bits, 0, d_codes[:, half:], stride, S, nbytes - half, capi.KIND_I16, d_out2[:, half // gb * gs:], ostride, d_state)
    torch.cuda.synchronize()
    assert torch.equal(d_out2[:, :n], d_out[:, :n])
    # encode batch: re-encode the decoded streams, compare with the oracle per stream
    d_enc = torch.zeros((S, stride), dtype=torch.uint8, device="cuda")
    c.g726_encode_batch_dev(bits, 0, capi.KIND_I16, d_out, ostride, S, n, d_enc, stride)
    torch.cuda.synchronize()
    enc = d_enc.cpu().numpy()
    for s in (0, 7, 33, S - 1):
        assert np.array_equal(enc[s, :nbytes], oracle.G726(bits).encode(out[s, :n])), s
    c.close()


# ---------------------------------------------------------------------------------------------- general-ratio kernel v2 (2 x 2 register tile)
@pytest.mark.parametrize("rates", [(44100, 48000), (44100, 16000), (48000, 44100), (32000, 48000), (8000, 44100), (22050, 16000)])
@pytest.mark.parametrize("semantics", ["oneshot", "streaming"])
def test_general_ratio_long_streams_bit_exact(ctx, oracle, rates, semantics):
    """streams long enough (>= 64 phase periods of outputs) to run resample_gen2_kernel, including its ragged last tile
    and the warps that fall back to per-output arithmetic; white noise plus a near-Nyquist tone (exposes index drift)"""
    i, o = rates
    frames = int(i * 1.7) + 173
    rng = np.random.default_rng(i + o)
    t = np.arange(frames)
    x = np.stack([rng.uniform(-0.5, 0.5, frames), 0.8 * np.sin(2 * np.pi * 0.41 * min(i, o) * t / i)], axis=1).astype(np.float32)
    data = x.reshape(-1).view(np.uint8)
    if semantics == "oneshot":
        want = oracle.downsample_audio(data, 32, 1, 2, i, o)
        got = ctx.normalize_stream(data, 32, capi.FMT_FLOAT, False, 2, capi.MODE_A, i, o, capi.SEM_ONESHOT, capi.MIX_NONE)
    else:
        st = oracle.Stream(i, o, 2)
        want = np.concatenate([st.process(np.ascontiguousarray(x.T)), st.flush()], axis=1)
        got = ctx.normalize_stream(data, 32, capi.FMT_FLOAT, False, 2, capi.MODE_B, i, o, capi.SEM_STREAMING, capi.MIX_NONE)
    assert got.shape == want.shape and got.shape[1] > 64 * 33
    for c in range(2):
        assert_resample_close(want[c], got[c])


# ---------------------------------------------------------------------------------------------- guard bands
def test_device_entry_points_do_not_write_outside_their_buffers():
    """compute-sanitizer is closed on this pool, so out-of-bounds WRITES are checked the blunt way: every device-pointer
    entry point writes into the middle of a canary-filled arena at ragged sizes; the canaries must survive."""
    import torch
    c = capi.Context(0)
    c.use_torch_stream()
    GUARD = 4096

    def arena(nbytes):
        buf = torch.full((GUARD + nbytes + GUARD + 64,), 0xA5, dtype=torch.uint8, device="cuda")
        return buf, buf[GUARD:GUARD + nbytes]

    def intact(buf, nbytes):
        torch.cuda.synchronize()
        return bool((buf[:GUARD] == 0xA5).all()) and bool((buf[GUARD + nbytes:] == 0xA5).all())

    g = torch.Generator(device="cuda"); g.manual_seed(5)
    rnd = lambda n: torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda", generator=g)
    for n in (1, 511, 512, 513, 70001):
        src = rnd(n)
        for kind, ob in ((capi.KIND_I16, 2), (capi.KIND_F32, 4)):
            buf, out = arena(n * ob)
            c.g711_decode_dev(capi.LAW_ALAW, src, n, kind, out)
            assert intact(buf, n * ob), ("g711 decode", n, kind)
        pcm = rnd(2 * n)
        buf, out = arena(n)
        c.g711_encode_dev(capi.LAW_MULAW, capi.KIND_I16, pcm, n, out)
        assert intact(buf, n), ("g711 encode", n)
    for name in ("s24le_to_i32", "f32le_to_i16", "u8_to_s16", "pkt_s16le_to_f32"):
        iw, ow = c.flat_widths(capi.FLAT_OP[name])
        for n in (1, 31 * 16, 32 * 16 + 5, 50001):
            src = rnd(n * iw)
            buf, out = arena(n * ow)
            c.flat_convert_dev(name, src, n * iw, out)
            assert intact(buf, n * ow), (name, n)
    for bits, fmt in ((16, 0), (24, 0), (32, 1)):
        for frames in (4, 252, 256, 260, 8196):
            S, fb = 3, (bits // 8) * 2
            stride = (frames * fb + 15) // 16 * 16
            src = rnd(S * stride)
            buf, out = arena(S * 2 * frames * 4)
            c.pcm_to_planar_f32_dev(src, stride, S, frames, bits, fmt, False, 2, capi.MODE_B, out, 2 * frames)
            assert intact(buf, S * 2 * frames * 4), ("planar", bits, frames)
            planar = torch.rand((2, frames), device="cuda") * 2 - 1
            buf, out = arena(frames * fb)
            c.planar_f32_to_pcm_dev(planar, frames, 2, bits, fmt, out)
            assert intact(buf, frames * fb), ("interleave", bits, frames)
    for rate in (44100, 48000, 96000):
        S, frames = 3, 5003
        src = rnd(S * frames * 4 + 16)
        n_out = c.normalize48k_out_frames(frames, rate)
        plane = (n_out + 3) // 4 * 4
        buf, out = arena(S * 2 * plane * 4)
        c.normalize48k_batch_dev(src, (frames * 4 + 15) // 16 * 16, S, frames, 16, 0, False, 2, rate, out.view(torch.float32), 2 * plane, plane)
        assert intact(buf, S * 2 * plane * 4), ("norm48k", rate)
    for i_rate, o_rate in ((48000, 16000), (44100, 48000)):
        S, frames = 2, 30011
        stride = (frames * 6 + 15) // 16 * 16
        src = rnd(S * stride)
        n_out = c.batch_plan(i_rate, o_rate, frames, capi.SEM_STREAMING)
        plane = (n_out + 3) // 4 * 4
        buf, out = arena(S * 2 * plane * 4)
        d = capi.BatchDesc(src.data_ptr(), stride, S, frames, 24, capi.FMT_SIGNED, 0, 2, capi.MODE_B, i_rate, o_rate, capi.SEM_STREAMING,
                           capi.MIX_NONE, out.data_ptr(), 2 * plane, plane)
        c.normalize_batch_dev(d)
        assert intact(buf, S * 2 * plane * 4), ("normalize", i_rate, o_rate)
    for bits in (2, 3, 4, 5):
        gs, gb = capi.G726_GROUP[bits]
        S, groups = 70, 333
        src = rnd(S * groups * gb)
        buf, out = arena(S * groups * gs * 2)
        c.g726_decode_batch_dev(bits, 0, src, groups * gb, S, groups * gb, capi.KIND_I16, out, groups * gs)
        assert intact(buf, S * groups * gs * 2), ("g726 decode", bits)
        buf2, enc = arena(S * groups * gb)
        c.g726_encode_batch_dev(bits, 0, capi.KIND_I16, out, groups * gs, S, groups * gs, enc, groups * gb)
        assert intact(buf2, S * groups * gb), ("g726 encode", bits)
    c.close()


# ---------------------------------------------------------------------------------------------- threads
def test_contexts_on_concurrent_threads_do_not_interfere(oracle):
    """the reference runs one worker thread per DecodePipeline and many pipelines at once (soundkit-decoder/src/lib.rs:
    2764): one sk_ctx per thread, calls overlapping in time (ctypes releases the GIL), every result equal to the oracle"""
    import threading
    errors = []

    def worker(seed):
        try:
            rng = np.random.default_rng(seed)
            c = capi.Context(0)
            for it in range(6):
                codes = rng.integers(0, 256, 20000 + seed, dtype=np.uint8)
                law = (seed + it) & 1
                assert np.array_equal(c.g711_decode(law, codes), oracle.g711_decode(law, codes))
                frames = 6000 + 7 * seed
                v = rng.integers(-(1 << 23), 1 << 23, frames * 2, dtype=np.int64)
                raw = np.stack([v & 0xff, (v >> 8) & 0xff, (v >> 16) & 0xff], axis=1).astype(np.uint8).reshape(-1)
                got = c.normalize_stream(raw, 24, capi.FMT_SIGNED, False, 2, capi.MODE_B, 48000, 16000, capi.SEM_STREAMING,
                                         capi.MIX_SCALE_THEN_SUM)[0]
                want = oracle.normalize_chain_mono(raw, 24, oracle.SIGNED, False, 2, 48000, 16000)
                assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
                st = c.g726_state()
                g = c.g726_decode(4, 0, st, codes[:8000])
                assert np.array_equal(g, oracle.G726(4).decode(codes[:8000]))
            c.close()
        except Exception as e:  # noqa: BLE001
            errors.append((seed, repr(e)))

    threads = [threading.Thread(target=worker, args=(s,)) for s in range(1, 7)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


# ---------------------------------------------------------------------------------------------- unaligned device buffers
def test_device_entry_points_with_unaligned_pointers_match_oracle(oracle):
    """callers may hand in device pointers at any offset (a Rust slice into a larger buffer): the vector / TMA kernels
    need 16-byte alignment, everything else must take the scalar kernels and still be exact"""
    import torch
    c = capi.Context(0)
    c.use_torch_stream()
    rng = np.random.default_rng(99)

    def dev(a, off):  # device copy of `a` placed `off` bytes into an allocation
        raw = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        buf = torch.zeros(raw.size + off + 64, dtype=torch.uint8, device="cuda")
        buf[off:off + raw.size] = torch.from_numpy(raw).cuda()
        return buf, buf[off:off + raw.size]

    for off_in, off_out in ((1, 0), (0, 2), (3, 6), (16, 16)):
        n = 5003
        codes = rng.integers(0, 256, n, dtype=np.uint8)
        _, d_in = dev(codes, off_in)
        obuf, d_out = dev(np.zeros(n, np.int16), off_out & ~1)
        c.g711_decode_dev(capi.LAW_MULAW, d_in, n, capi.KIND_I16, d_out)
        torch.cuda.synchronize()
        assert np.array_equal(d_out.cpu().numpy().view(np.int16), oracle.g711_decode(0, codes))
        pcm = rng.integers(-32768, 32768, n).astype(np.int16)
        _, d_pcm = dev(pcm, off_in & ~1)
        _, d_codes = dev(np.zeros(n, np.uint8), off_out + 1)
        c.g711_encode_dev(capi.LAW_ALAW, capi.KIND_I16, d_pcm, n, d_codes)
        torch.cuda.synchronize()
        assert np.array_equal(d_codes.cpu().numpy(), oracle.g711_encode(1, pcm))
        s24 = rng.integers(0, 256, 3 * n, dtype=np.uint8)
        _, d_s24 = dev(s24, off_in)
        _, d_i32 = dev(np.zeros(n, np.int32), off_out & ~3)
        c.flat_convert_dev("s24le_to_i32", d_s24, 3 * n, d_i32)
        torch.cuda.synchronize()
        assert np.array_equal(d_i32.cpu().numpy().view(np.int32), oracle.flat("s24le_to_i32", s24))
    # fused chain and 48 kHz normalizer from a stream that starts 2 bytes into an allocation (no TMA / no staged path)
    frames = 20011
    v = rng.integers(-(1 << 15), 1 << 15, frames * 2).astype(np.int16)
    _, d_v = dev(v, 2)
    n_out = c.batch_plan(48000, 16000, frames, capi.SEM_STREAMING)
    out = torch.zeros(n_out + 8, dtype=torch.float32, device="cuda")
    d = capi.BatchDesc(d_v.data_ptr(), frames * 4, 1, frames, 16, capi.FMT_SIGNED, 0, 2, capi.MODE_B, 48000, 16000, capi.SEM_STREAMING,
                       capi.MIX_SCALE_THEN_SUM, out.data_ptr(), n_out, 0)
    c.normalize_batch_dev(d)
    torch.cuda.synchronize()
    want = oracle.normalize_chain_mono(v.view(np.uint8), 16, oracle.SIGNED, False, 2, 48000, 16000)
    assert np.array_equal(out[:n_out].cpu().numpy().view(np.uint32), want.view(np.uint32))
    # other widths at offsets that break the paired loads (4) and the sample alignment itself (1, 2)
    for bits, fmt, off in ((16, 0, 1), (24, 0, 1), (32, 0, 4), (32, 1, 2), (32, 1, 1)):
        fr2 = 9001
        raw = rng.integers(0, 256, fr2 * 2 * (bits // 8), dtype=np.uint8)
        if fmt == 1:
            raw = rng.uniform(-1, 1, fr2 * 2).astype(np.float32).view(np.uint8)
        _, d_raw = dev(raw, off)
        for (ir, orr) in ((48000, 16000), (44100, 48000)):
            n2 = c.batch_plan(ir, orr, fr2, capi.SEM_ONESHOT)
            o2 = torch.zeros((2, n2), dtype=torch.float32, device="cuda")
            dd = capi.BatchDesc(d_raw.data_ptr(), raw.size, 1, fr2, bits, fmt, 0, 2, capi.MODE_A, ir, orr, capi.SEM_ONESHOT, capi.MIX_NONE,
                                o2.data_ptr(), 2 * n2, n2)
            c.normalize_batch_dev(dd)
            torch.cuda.synchronize()
            w2 = oracle.downsample_audio(raw, bits, fmt, 2, ir, orr)
            assert np.array_equal(o2.cpu().numpy().view(np.uint32), w2.view(np.uint32)), (bits, fmt, off, ir)
    n48 = c.normalize48k_out_frames(frames, 44100)
    o48 = torch.zeros((2, n48 + 3), dtype=torch.float32, device="cuda")
    c.normalize48k_batch_dev(d_v, frames * 4, 1, frames, 16, 0, False, 2, 44100, o48[:, 1:], 2 * (n48 + 3), n48 + 3)  # output 4 bytes off alignment too
    torch.cuda.synchronize()
    wl, wr = oracle.normalize48k([(v.view(np.uint8), 16, 0, False, 2, 44100)])
    got = o48.cpu().numpy()
    assert np.array_equal(got[0, 1:1 + n48].view(np.uint32), wl.view(np.uint32)) and np.array_equal(got[1, 1:1 + n48].view(np.uint32), wr.view(np.uint32))
    c.close()


# ---------------------------------------------------------------------------------------------- full-size properties
def test_f32_deinterleave_interleave_round_trip_at_config_size():
    """BASELINE configs[1] wave (64 streams x 60 s f32le stereo 48 kHz, 1.5 GB): deinterleave_vecs_f32 followed by
    interleave_vecs_f32 (f32_channels_to_bytes, PCMFloat) is the identity on the BYTES, NaN payloads and -0.0 included;
    s16 -> planar f32 -> float_sample_to_i16 is checked the same way through its closed form."""
    import torch
    c = capi.Context(0)
    c.use_torch_stream()
    S, frames = 64, 2880000
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    raw = torch.randint(0, 256, (S, frames * 8), dtype=torch.uint8, device="cuda", generator=g)  # arbitrary bit patterns incl. NaNs
    planar = torch.empty((S, 2, frames), dtype=torch.float32, device="cuda")
    c.pcm_to_planar_f32_dev(raw, raw.stride(0), S, frames, 32, capi.FMT_FLOAT, False, 2, capi.MODE_A, planar, 2 * frames)
    back = torch.empty((frames * 8,), dtype=torch.uint8, device="cuda")
    for s in (0, 17, S - 1):
        c.planar_f32_to_pcm_dev(planar[s], frames, 2, 32, capi.FMT_FLOAT, back)
        torch.cuda.synchronize()
        assert torch.equal(back, raw[s]), s
    del planar, back
    # s16: x / 32768 then float_sample_to_i16: negative values scale back by 32768 (identity), positive by 32767:
    # round(v * 32767 / 32768) = round(v - v / 32768) = v - (1 if v > 16384 else 0) for v >= 0 (v = 16384 is the one tie:
    # 16383.5 rounds half away from zero, back to 16384)
    raw16 = torch.randint(-32768, 32768, (frames * 2,), dtype=torch.int16, device="cuda", generator=g)
    p16 = torch.empty((2, frames), dtype=torch.float32, device="cuda")
    c.pcm_to_planar_f32_dev(raw16, frames * 4, 1, frames, 16, capi.FMT_SIGNED, False, 2, capi.MODE_B, p16, 2 * frames)
    b16 = torch.empty((frames * 2,), dtype=torch.int16, device="cuda")
    c.planar_f32_to_pcm_dev(p16, frames, 2, 16, capi.FMT_SIGNED, b16)
    torch.cuda.synchronize()
    want = raw16.to(torch.int32)
    want = torch.where(want > 16384, want - 1, want).to(torch.int16)
    assert torch.equal(b16, want)
    c.close()


# ---------------------------------------------------------------------------------------------- bytes out (apply_output_options in one call)
@pytest.mark.parametrize("out_bits,out_fmt", [(16, 0), (24, 0), (32, 0), (32, 1)])
@pytest.mark.parametrize("case", [("s24 stereo 48->16 (fused epilogue)", 24, 0, 2, 48000, 16000),
                                  ("s16 stereo 48->8 (fused epilogue)", 16, 0, 2, 48000, 8000),
                                  ("f32 stereo 44.1->16 (scratch + pass)", 32, 1, 2, 44100, 16000),
                                  ("s16 mono 48->16 (scratch + pass)", 16, 0, 1, 48000, 16000),
                                  ("s24 5.1 48->16 (scratch + pass)", 24, 0, 6, 48000, 16000)])
def test_normalize_batch_pcm_dev_matches_output_stage(oracle, case, out_bits, out_fmt):
    """convert -> resample -> downmix_channels(.., 1) -> f32_channels_to_bytes, the rate-changing arm of
    apply_output_options (soundkit-decoder/src/lib.rs:3324-3456), for a device-resident batch with ragged lengths"""
    import torch
    _, bits, fmt, channels, i_rate, o_rate = case
    c = capi.Context(0)
    c.use_torch_stream()
    S, frames = 3, 20011
    rng = np.random.default_rng(out_bits + bits + channels)
    fb = (bits // 8) * channels
    stride = (frames * fb + 15) // 16 * 16
    host = np.zeros((S, stride), np.uint8)
    for s in range(S):
        if fmt == 1:
            host[s, :frames * fb] = rng.uniform(-1.1, 1.1, frames * channels).astype(np.float32).view(np.uint8)
        else:
            host[s, :frames * fb] = rng.integers(0, 256, frames * fb, dtype=np.uint8)
    d_in = torch.from_numpy(host).cuda()
    n_out = c.batch_plan(i_rate, o_rate, frames, capi.SEM_STREAMING)
    ob = out_bits // 8
    ostride = (n_out * ob + 15) // 16 * 16
    d_out = torch.full((S, ostride), 0xEE, dtype=torch.uint8, device="cuda")
    d = capi.BatchDesc(d_in.data_ptr(), stride, S, frames, bits, fmt, 0, channels, capi.MODE_B, i_rate, o_rate, capi.SEM_STREAMING,
                       capi.MIX_SCALE_THEN_SUM, 0, 0, 0)
    c.normalize_batch_pcm_dev(d, out_bits, out_fmt, d_out, ostride)
    torch.cuda.synchronize()
    out = d_out.cpu().numpy()
    for s in range(S):
        planar = oracle.pcm_to_planar_f32(host[s, :frames * fb], bits, fmt, False, channels, capi.MODE_B)
        st = oracle.Stream(i_rate, o_rate, channels)
        res = np.concatenate([st.process(planar), st.flush()], axis=1)
        mono = oracle.downmix_channels(res, 1) if channels > 1 else res
        want = oracle.f32_channels_to_bytes(mono, out_bits, out_fmt)
        assert want.size == n_out * ob
        assert np.array_equal(out[s, :n_out * ob], want), (s, case[0])
        assert np.all(out[s, n_out * ob:] == 0xEE)  # nothing written past the stream's bytes
    c.close()


def test_float_sample_to_i16_without_f64_is_exact_over_whole_binades(ctx, oracle):
    """the device evaluates the reference's f64 rule (soundkit-decoder/src/lib.rs:1815-1827) in f32 with an exact
    error term: every 5th float of every binade that can produce a non-zero sample, both signs, must match"""
    for e in range(100, 128):
        bits = (np.uint32(e) << np.uint32(23)) + np.arange(e % 5, 1 << 23, 5, dtype=np.uint32)
        x = bits.view(np.float32)
        x = np.concatenate([x, -x])[None, :]
        assert np.array_equal(ctx.planar_f32_to_pcm(x, 16, capi.FMT_SIGNED), oracle.f32_channels_to_bytes(x, 16, 0)), e


# ---------------------------------------------------------------------------------------------- poisoned shared memory
@pytest.mark.parametrize("pattern", [0x7FC00000, 0xFF800000])
def test_resamplers_do_not_read_unstaged_shared_memory(oracle, pattern):
    """Shared memory keeps its contents between kernels.  With every SM's shared memory full of NaN (or -Inf) bit patterns, a
    kernel that multiplies a word it never staged — even by a zero coefficient — produces NaN.  (Found the hard way: the
    scalar-staging 48->16 kHz kernel paired tap 255 of a tile's last output with the word one past its window.)"""
    c = capi.Context(0)
    rng = np.random.default_rng(pattern & 0xFFFF)
    cases = [(48000, 16000, 1, 16), (48000, 16000, 6, 24), (48000, 16000, 2, 24), (48000, 24000, 3, 16), (48000, 8000, 2, 16),
             (44100, 16000, 2, 16), (44100, 48000, 1, 24), (96000, 8000, 1, 16), (48000, 12000, 2, 32)]
    for i_rate, o_rate, channels, bits in cases:
        frames = 20011
        data = rng.integers(0, 256, frames * channels * (bits // 8), dtype=np.uint8)
        planar = oracle.pcm_to_planar_f32(data, bits, 0, False, channels, capi.MODE_B)
        st = oracle.Stream(i_rate, o_rate, channels)
        want = np.concatenate([st.process(planar), st.flush()], axis=1)
        c.poison_shared_memory(pattern)
        got = c.normalize_stream(data, bits, capi.FMT_SIGNED, False, channels, capi.MODE_B, i_rate, o_rate, capi.SEM_STREAMING, capi.MIX_NONE)
        assert got.shape == want.shape
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (i_rate, o_rate, channels, bits)
    # the 48 kHz normalizer stages through shared memory as well
    data = rng.integers(0, 256, 30011 * 4, dtype=np.uint8)
    wl, wr = oracle.normalize48k([(data, 16, 0, False, 2, 44100)])
    c.poison_shared_memory(pattern)
    gn = c.normalizer48k()
    gl, gr = gn.push(data, 16, 0, False, 2, 44100)
    fl, fr = gn.finish()
    gn.close()
    assert np.array_equal(np.concatenate([gl, fl]).view(np.uint32), wl.view(np.uint32))
    assert np.array_equal(np.concatenate([gr, fr]).view(np.uint32), wr.view(np.uint32))
    c.close()


@pytest.mark.parametrize("dtype", [np.int16, np.int32, np.float32])
@pytest.mark.parametrize("channels,frames", [(1, 0), (1, 5), (2, 4096), (2, 100003), (3, 1000), (6, 777)])
def test_push_planar_interleave_is_a_bit_copy(ctx, dtype, channels, frames):
    """WavStreamEncoder::push_planar_{i16,i32,f32} (soundkit/src/wav.rs:444-494): out[(f * ch + c)] = planar[c * frames + f], little
    endian; i32 values and float NaN payloads must survive untouched"""
    if dtype == np.float32:
        bits = RNG.integers(0, 2 ** 32, (channels, frames), dtype=np.uint64).astype(np.uint32)  # every bit pattern incl. NaNs / Infs
        planar = bits.view(np.float32)
    else:
        info = np.iinfo(dtype)
        planar = RNG.integers(info.min, info.max + 1, (channels, frames), dtype=np.int64).astype(dtype)
    got = ctx.interleave_planar(planar)
    want = np.ascontiguousarray(planar.T).view(np.uint8).reshape(-1)
    assert np.array_equal(got, want)
