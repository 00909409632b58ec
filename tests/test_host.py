"""Host-layer tests: the C++ mirror of soundkit's Rust interface (soundkit_b200/csrc/host/soundkit_host.hpp,
exposed to pytest through the pybind11 module soundkit_b200._host) driven the way the reference's own unit
tests drive the Rust functions.  Each test cites the reference test it restates.

CPU part (no marker): framing, validation and error strings, which happen before the C ABI is crossed, and
the "no device -> loud failure" contract.  GPU part (@gpu): the arithmetic, through the C ABI."""
import math
import struct

import numpy as np
import pytest

from conftest import golden

H = pytest.importorskip("soundkit_b200._host", reason="run `make -C soundkit_b200/csrc host` (done by __graft_entry__.build())")
AB, AP, RP, G, D = H.audio_bytes, H.audio_pipeline, H.raw_pcm, H.g711, H.decoder
LE, BE = H.Endianness.LittleEndian, H.Endianness.BigEndian
SIGNED, FLOAT = H.EncodingFlag.PCMSigned, H.EncodingFlag.PCMFloat


def audio(bits, ch, rate, data, fmt=SIGNED, endian=LE):
    return H.AudioData(bits, ch, rate, bytes(data), fmt, endian)


# ---------------------------------------------------------------------------------------------- CPU: raw_pcm
def test_linear16_buffers_partial_frames_until_complete():  # soundkit/src/raw_pcm.rs:215-232
    p = RP.RawPcmStreamProcessor(RP.RawPcmFormat.linear16(8000, 1))
    assert p.add(bytes([0x34])) is None
    assert p.buffered_len() == 1
    a = p.add(bytes([0x12, 0x78, 0x56]))
    assert (a.bits_per_sample(), a.channel_count(), a.sampling_rate()) == (16, 1, 8000)
    assert a.audio_format() == SIGNED and a.endianness() == LE
    assert a.data() == bytes([0x34, 0x12, 0x78, 0x56])
    assert p.buffered_len() == 0


def test_stereo_streams_only_emit_complete_interleaved_frames():  # raw_pcm.rs:234-246
    p = RP.RawPcmStreamProcessor(RP.RawPcmFormat.linear16(16000, 2))
    assert p.add(bytes([1, 2, 3])) is None
    assert p.buffered_len() == 3
    a = p.add(bytes([4, 5, 6, 7, 8, 9]))
    assert a.data() == bytes([1, 2, 3, 4, 5, 6, 7, 8])
    assert p.buffered_len() == 1
    with pytest.raises(H.SoundkitError, match="partial-frame"):
        p.flush()


def test_l16_uses_big_endian_metadata_and_linear32_is_float():  # raw_pcm.rs:248-269
    a = RP.RawPcmStreamProcessor(RP.RawPcmFormat.l16(8000, 1)).add(bytes([0x12, 0x34]))
    assert a.endianness() == BE and a.data() == bytes([0x12, 0x34])
    sample = struct.pack("<f", 0.25)
    a = RP.RawPcmStreamProcessor(RP.RawPcmFormat.linear32(48000, 1)).add(sample)
    assert a.bits_per_sample() == 32 and a.audio_format() == FLOAT and a.data() == sample


def test_rejects_invalid_format_metadata_and_oversized_chunks():  # raw_pcm.rs:271-284
    with pytest.raises(H.SoundkitError, match="sample rate must be > 0"):
        RP.RawPcmFormat.linear16(0, 1)
    with pytest.raises(H.SoundkitError, match="channel count must be > 0"):
        RP.RawPcmFormat.linear16(8000, 0)
    p = RP.RawPcmStreamProcessor(RP.RawPcmFormat.linear16(48000, 2))
    with pytest.raises(H.SoundkitError, match="streaming budget"):
        p.add(bytes(RP.MAX_INPUT_CHUNK_BYTES + 1))


def test_decode_linear16_fixture_is_identity_in_333_byte_pushes():  # raw_pcm.rs:286-318
    fixture = golden("linear16.s16le")
    p = RP.RawPcmStreamProcessor(RP.RawPcmFormat.linear16(16000, 1))
    out = b""
    for i in range(0, len(fixture), 333):
        a = p.add(fixture[i:i + 333])
        if a is not None:
            assert (a.bits_per_sample(), a.channel_count(), a.sampling_rate()) == (16, 1, 16000)
            out += a.data()
    assert p.flush() is None
    assert out == fixture == golden("linear16.decoded.s16le")


def test_format_descriptor_arithmetic():  # raw_pcm.rs:15-35, 89-115
    F = RP.RawPcmSampleFormat
    for sf, bits, bps, flag in ((F.I16, 16, 2, SIGNED), (F.I24, 24, 3, SIGNED), (F.I32, 32, 4, SIGNED), (F.F32, 32, 4, FLOAT)):
        f = RP.RawPcmFormat.new(44100, 6, sf, BE)
        assert f.bits_per_sample() == bits and f.bytes_per_frame() == bps * 6 and f.encoding_flag() == flag
        assert f.sample_rate() == 44100 and f.channels() == 6 and f.endianness() == BE and f.sample_format() == sf


# ---------------------------------------------------------------------------------------------- CPU: validation before the ABI
def test_error_strings_produced_before_the_abi():
    # soundkit/src/audio_pipeline.rs:423-436
    with pytest.raises(H.SoundkitError, match=r"invalid f32le byte length 3; expected multiple of 4"):
        AP.f32s_from_le_bytes(bytes([0, 1, 2]))
    assert AP.f32s_from_le_bytes(AP.f32s_to_le_bytes([0.0, 0.25, -0.5, 1.0])).tolist() == [0.0, 0.25, -0.5, 1.0]  # :757-763
    # downsample_audio argument checks, audio_pipeline.rs:440-466 (order matters)
    with pytest.raises(H.SoundkitError, match="Channel count must be > 0"):
        AP.downsample_audio(audio(16, 0, 48000, b""), 16000)
    with pytest.raises(H.SoundkitError, match="Unsupported bits_per_sample: 8"):
        AP.downsample_audio(audio(8, 1, 48000, b"\0" * 8), 16000)
    with pytest.raises(H.SoundkitError, match="sampling_rate must be > 0"):
        AP.downsample_audio(audio(16, 1, 48000, b"\0" * 8), 0)
    with pytest.raises(H.SoundkitError, match="Unsupported input sample_rate: 11025"):
        AP.downsample_audio(audio(16, 1, 11025, b"\0" * 8), 16000)
    with pytest.raises(H.SoundkitError, match="Unsupported output sample_rate: 12345"):
        AP.downsample_audio(audio(16, 1, 48000, b"\0" * 8), 12345)
    with pytest.raises(H.SoundkitError, match="Channel count must be > 0"):
        AP.audio_to_f32_channels(audio(16, 0, 48000, b""))
    with pytest.raises(H.SoundkitError, match="deserialize_audio failed: unsuporrted type"):
        AP.audio_to_f32_channels(audio(8, 1, 48000, b"\0"))
    with pytest.raises(H.SoundkitError, match="unsuporrted type"):
        AP.deserialize_audio(b"\0" * 4, 12, 1)
    with pytest.raises(H.SoundkitError, match="channel length mismatch"):
        AP.mixdown_to_mono_f32([np.zeros(3, np.float32), np.zeros(4, np.float32)])
    # soundkit-g711/src/lib.rs:36-66
    with pytest.raises(H.SoundkitError, match=r"Output buffer too small for G.711 encode: need 4, have 3"):
        G.encode_i16(G.G711Law.MuLaw, [1, 2, 3, 4], 3)
    with pytest.raises(H.SoundkitError, match=r"Output buffer too small for G.711 decode: need 4, have 2"):
        G.decode_i16(G.G711Law.ALaw, bytes(4), 2)
    # soundkit-decoder/src/lib.rs:3563-3617, 3619-3683, 3458-3489
    with pytest.raises(H.SoundkitError, match="Channel count must be > 0"):
        D.audio_data_to_f32_channels(audio(16, 0, 48000, b""))
    with pytest.raises(H.SoundkitError, match="PCM data is unsupported or contains a partial frame"):
        D.audio_data_to_f32_channels(audio(16, 2, 48000, b"\0" * 6))
    with pytest.raises(H.SoundkitError, match="floating-point PCM must contain 32-bit samples"):
        D.audio_data_to_f32_channels(audio(16, 1, 48000, b"\0" * 4, FLOAT))
    with pytest.raises(H.SoundkitError, match="Channel length mismatch"):
        D.f32_channels_to_bytes([np.zeros(3, np.float32), np.zeros(4, np.float32)], 16, SIGNED)
    with pytest.raises(H.SoundkitError, match="PCMFloat output requires 32-bit samples"):
        D.f32_channels_to_bytes([np.zeros(3, np.float32)], 16, FLOAT)
    with pytest.raises(H.SoundkitError, match="Unsupported output bits per sample: 12"):
        D.f32_channels_to_bytes([np.zeros(3, np.float32)], 12, SIGNED)
    with pytest.raises(H.SoundkitError, match="integer PCM contains a partial sample"):
        D.exact_signed_pcm_to_i16(audio(24, 1, 48000, b"\0" * 4))
    # apply_output_options argument checks (lib.rs:3324-3339), wrapped in DecodeError::DecodingFailed
    a = audio(16, 1, 48000, b"\0" * 8)
    for opts, msg in ((D.DecodeOptions(output_sample_rate=0), "Output sample rate must be > 0"),
                      (D.DecodeOptions(output_bits_per_sample=12), "Unsupported output bits per sample: 12"),
                      (D.DecodeOptions(output_channels=0), "Output channels must be > 0")):
        with pytest.raises(H.DecodeError, match=msg):
            D.apply_output_options(a, opts, D.ResamplerSlot())
    # the fast path (nothing to do, lib.rs:3340-3345) hands the frame back untouched and needs no device
    out = D.apply_output_options(a, D.DecodeOptions(), D.ResamplerSlot())
    assert len(out) == 1 and out[0].data() == a.data()


@pytest.mark.gpu
def test_float_sample_to_i16_scalar_rule():  # soundkit-decoder/src/lib.rs:1815-1827 (device arithmetic: the host layer has none)
    f = D.float_sample_to_i16
    assert [f(x) for x in (0.0, 1.0, -1.0, 2.0, -2.0, float("nan"), float("inf"), 0.5, -0.5)] == \
           [0, 32767, -32768, 32767, -32768, 0, 0, 16384, -16384]
    # round() is half away from zero; x * 32768 is exact for these inputs, so the tie is real
    assert [f(-0.5 / 32768.0), f(-1.5 / 32768.0), f(-2.5 / 32768.0)] == [-1, -2, -3]


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_compute_calls_fail_loudly_without_a_device():
    """There is no CPU fallback: every arithmetic entry point raises when sk_ctx_create finds no device."""
    with pytest.raises(H.SoundkitError, match="no CUDA device|no CPU fallback"):
        AB.i16le_to_f32(bytes(4))
    with pytest.raises(H.SoundkitError, match="no CUDA device|no CPU fallback"):
        G.decode_i16(G.G711Law.MuLaw, bytes(4), 4)
    with pytest.raises(H.SoundkitError, match="no CUDA device|no CPU fallback"):
        AP.downsample_audio(audio(16, 1, 48000, bytes(9600)), 16000)


# ---------------------------------------------------------------------------------------------- GPU: audio_bytes
@pytest.mark.gpu
def test_deinterleave_and_interleave_unit_vectors():  # soundkit/src/audio_bytes.rs:380-415
    r = AB.deinterleave_vecs_i16(bytes([1, 0, 2, 0, 3, 0, 4, 0, 5, 0, 6, 0]), 2)
    assert [c.tolist() for c in r] == [[1, 3, 5], [2, 4, 6]]
    assert AB.interleave_vecs_i16([[1, 3, 5], [2, 4, 6]]) == bytes([1, 0, 2, 0, 3, 0, 4, 0, 5, 0, 6, 0])
    r = AB.deinterleave_vecs_s24(bytes([1, 0, 0, 2, 0, 0, 3, 0, 0, 4, 0, 0, 5, 0, 0, 6, 0, 0]), 2)
    assert [c.tolist() for c in r] == [[1, 3, 5], [2, 4, 6]]
    r = AB.deinterleave_vecs_f32(bytes([0, 0, 128, 63, 0, 0, 0, 64, 0, 0, 64, 64, 0, 0, 128, 64, 0, 0, 160, 64, 0, 0, 192, 64]), 2)
    assert [c.tolist() for c in r] == [[1.0, 3.0, 5.0], [2.0, 4.0, 6.0]]


@pytest.mark.gpu
def test_i16le_to_f32_unit_vector():  # audio_bytes.rs:417-440
    got = AB.i16le_to_f32(bytes([0, 0, 0, 64, 255, 127, 0, 192, 0, 128]))
    assert np.allclose(got, [0.0, 0.5, 0.9999694, -0.5, -1.0], atol=1e-4)
    assert got.tolist() == [0.0, 0.5, 32767 / 32768, -0.5, -1.0]  # and in fact exact
    with pytest.raises(H.SoundkitPanic):  # the reference asserts an even length (:4)
        AB.i16le_to_f32(bytes(3))


@pytest.mark.gpu
def test_stereo_to_mono_unit_vectors():  # audio_bytes.rs:442-468
    assert AB.stereo_to_mono_take_left([10, 20, -30, -40, 50, 60]).tolist() == [10, -30, 50]
    assert AB.stereo_to_mono_inplace_take_left([10, 20, -30, -40, 50, 60]).tolist() == [10, -30, 50]
    assert AB.stereo_to_mono_avg([100, -100, 50, 150, -200, 200]).tolist() == [0, 100, 0]
    assert AB.stereo_to_mono_inplace_avg([100, -100, 50, 150, -200, 200]).tolist() == [0, 100, 0]
    assert AB.stereo_to_mono_avg([-3, -4, 3, 4]).tolist() == [-3, 3]  # Rust `/` truncates toward zero
    with pytest.raises(H.SoundkitPanic):
        AB.stereo_to_mono_avg([1, 2, 3])


@pytest.mark.gpu
def test_flat_functions_against_oracle(oracle):
    rng = np.random.default_rng(7)
    raw = rng.integers(0, 256, 12 * 1001, dtype=np.uint8)
    for name in ("i16le_to_i16", "s16be_to_i16", "s16le_to_i32", "s24le_to_i32", "s24le_to_i16", "s24be_to_i16", "s32le_to_i32",
                 "s32be_to_i32", "s32le_to_s24", "s32be_to_s24", "s32le_to_f32", "s32be_to_f32", "s32le_to_i16", "s32be_to_i16"):
        got = getattr(AB, name)(raw.tobytes())
        want = oracle.flat({"i16le_to_i16": "s16le_to_i16"}.get(name, name), raw)  # :25 and :231 are the same function
        assert np.array_equal(np.asarray(got).view(np.uint8), want.view(np.uint8)), name
    x = rng.uniform(-1.25, 1.25, 3001).astype(np.float32)
    x[::97] = np.nan
    for name in ("f32le_to_i16", "f32le_to_i32", "f32le_to_s24"):
        assert np.array_equal(getattr(AB, name)(x.tobytes()), oracle.flat(name, x.view(np.uint8))), name


# ---------------------------------------------------------------------------------------------- GPU: audio_pipeline
@pytest.mark.gpu
def test_audio_to_mono_f32_averages_channels():  # soundkit/src/audio_pipeline.rs:698-713
    a = audio(16, 2, 48000, AB.interleave_vecs_i16([[32767, -32768], [-32768, 32767]]))
    mono = AP.audio_to_mono_f32(a)
    assert len(mono) == 2 and abs(mono[0]) < 0.01 and abs(mono[1]) < 0.01


@pytest.mark.gpu
def test_audio_to_f32_channels_keeps_the_24_bit_quirk():  # audio_pipeline.rs:74-98 (24-bit / 2^31, always LE)
    a = audio(24, 1, 48000, bytes([0xff, 0xff, 0x7f, 0x00, 0x00, 0x80]), SIGNED, BE)
    got = AP.audio_to_f32_channels(a)[0]
    assert got.tolist() == [8388607 / 2 ** 31, -8388608 / 2 ** 31]


@pytest.mark.gpu
def test_downsample_audio_wav32f_fixture_to_8k(oracle):  # audio_pipeline.rs:645-696 (fixture -> 8 kHz, non-empty)
    raw = golden("wav_32f.f32le")
    a = audio(32, 1, 16000, raw, FLOAT)
    got = AP.downsample_audio(a, 8000)
    assert len(got) == 1 and len(got[0]) > 0
    want = oracle.downsample_audio(np.frombuffer(raw, np.uint8), 32, oracle.FLOAT, 1, 16000, 8000)
    assert np.array_equal(got[0].view(np.uint32), want[0].view(np.uint32))


# ---------------------------------------------------------------------------------------------- GPU: G.711
SAMPLES = [-32768, -30000, -20000, -12000, -4096, -1024, -32, -1, 0, 1, 32, 1024, 4096, 12000, 20000, 30000, 32767]
MU, A = G.G711Law.MuLaw, G.G711Law.ALaw


@pytest.mark.gpu
def test_g711_known_zero_codes():  # soundkit-g711/src/lib.rs:402-414
    assert G.decode_sample(MU, 0xff) == 0 and G.decode_sample(MU, 0x7f) == 0 and G.encode_sample(MU, 0) == 0xff
    assert G.decode_sample(A, 0xd5) == 8 and G.decode_sample(A, 0x55) == -8 and G.encode_sample(A, 0) == 0xd5


@pytest.mark.gpu
@pytest.mark.parametrize("law", [MU, A])
def test_g711_roundtrip_is_stable_after_quantization(law):  # lib.rs:417-434
    for s in SAMPLES:
        d = G.decode_sample(law, G.encode_sample(law, s))
        assert G.decode_sample(law, G.encode_sample(law, d)) == d


@pytest.mark.gpu
def test_g711_streaming_chunk_boundaries_do_not_change_encoded_output():  # lib.rs:437-451
    enc = G.G711Encoder.new_mulaw(8000, 1)
    n, whole = enc.encode_i16(SAMPLES, len(SAMPLES))
    chunked = []
    for i in range(0, len(SAMPLES), 3):
        w, out = enc.encode_i16(SAMPLES[i:i + 3], 4)
        chunked += out[:w].tolist()
    assert chunked == whole[:n].tolist()
    assert bytes(chunked).hex() == "00020c182f4d7b7ffffffbcdaf988c8280"  # SURVEY.md 8a known answer


@pytest.mark.gpu
def test_g711_decoder_trait_supports_i16_i32_and_f32_output():  # lib.rs:454-479
    encoded = bytes([0xff, 0xd7, 0xb7, 0x37, 0x17, 0x7f])
    dec = G.G711Decoder.new_mulaw(8000, 1)
    n16, i16 = dec.decode_i16(encoded, 6, False)
    n32, i32 = dec.decode_i32(encoded, 6, False)
    nf, f32 = dec.decode_f32(encoded, 6, False)
    assert (n16, n32, nf) == (6, 6, 6)
    for k in range(6):
        assert int(i32[k]) == int(i16[k]) << 16
        assert abs(float(f32[k]) - int(i16[k]) / 32768.0) < np.finfo(np.float32).eps


@pytest.mark.gpu
def test_g711_fixture_roundtrips_and_golden_decode():  # lib.rs:482-530 + golden/g711_*/*.decoded.wav
    samples = np.frombuffer(golden("linear16.s16le"), "<i2")
    for law in (MU, A):
        enc, dec = G.G711Encoder.new_with_law(law, 16000, 1), G.G711Decoder.new_with_law(law, 16000, 1)
        encoded = b""
        for i in range(0, len(samples), 257):
            w, out = enc.encode_i16(samples[i:i + 257].tolist(), 257)
            encoded += out[:w].tobytes()
        assert len(encoded) == len(samples)
        decoded = []
        for i in range(0, len(encoded), 113):
            w, out = dec.decode_i16(encoded[i:i + 113], 113, False)
            decoded.append(out[:w])
        decoded = np.concatenate(decoded)
        assert len(decoded) == len(samples) and np.any(decoded != 0)
    for law, stem in ((MU, "g711_ulaw"), (A, "g711_alaw")):
        fixture = golden(stem + "." + stem.split("_")[1])
        dec = G.G711Decoder.new_with_law(law, 8000, 1)
        decoded = []
        for i in range(0, len(fixture), 137):
            w, out = dec.decode_i16(fixture[i:i + 137], 137, False)
            decoded.append(out[:w])
        assert np.concatenate(decoded).astype("<i2").tobytes() == golden(stem + ".decoded.s16le")


# ---------------------------------------------------------------------------------------------- GPU: decoder output stage
def f32_audio(rate, channels):
    """create_f32_audio of the reference tests: planar f32 -> interleaved PCMFloat AudioData"""
    inter = np.stack(channels, axis=1).astype("<f4").reshape(-1)
    return audio(32, len(channels), rate, inter.tobytes(), FLOAT)


@pytest.mark.gpu
def test_streaming_resampler_matches_single_pass_length():  # soundkit-decoder/src/lib.rs:5187-5238
    in_rate, out_rate = 44100, 16000
    n = in_rate * 3 + 137
    idx = np.arange(n, dtype=np.float32)
    samples = (0.5 * np.sin(np.float32(2.0 * math.pi * 440.0) * idx / np.float32(in_rate))).astype(np.float32)
    reference = AP.downsample_audio(f32_audio(in_rate, [samples]), out_rate)
    reference_len = len(reference[0])
    opts = D.DecodeOptions(output_bits_per_sample=16, output_sample_rate=out_rate, output_channels=1)
    slot = D.ResamplerSlot()
    streaming_len = 0
    for i in range(0, n, 997):
        for frame in D.apply_output_options(f32_audio(in_rate, [samples[i:i + 997]]), opts, slot):
            assert (frame.bits_per_sample(), frame.channel_count(), frame.sampling_rate()) == (16, 1, out_rate)
            streaming_len += len(D.audio_data_to_f32_channels(frame)[0])
    assert slot.is_some()
    for frame in D.flush_resampler_frames(slot):
        streaming_len += len(D.audio_data_to_f32_channels(frame)[0])
    assert streaming_len == reference_len == 48002


@pytest.mark.gpu
def test_decoder_scale_conventions():  # audio_pipeline.rs:715-755 vectors, decoder convention (lib.rs:3563-3617)
    be = audio(16, 2, 48000, struct.pack(">hh", 32767, -32768), SIGNED, BE)
    l, r = D.audio_data_to_f32_channels(be)
    assert l[0] == np.float32(32767 / 32768) and r[0] == -1.0
    s24 = audio(24, 2, 48000, bytes([0xff, 0xff, 0x7f, 0x00, 0x00, 0x80]))
    l, r = D.audio_data_to_f32_channels(s24)
    assert l[0] == np.float32(8388607 / 8388608) and r[0] == -1.0
    nan = audio(32, 1, 48000, struct.pack("<f", float("nan")), FLOAT)
    assert D.audio_data_to_f32_channels(nan)[0].tolist() == [0.0]


@pytest.mark.gpu
def test_apply_output_options_width_and_channel_arms(oracle):
    rng = np.random.default_rng(11)
    # exact integer narrowing: 24-bit -> 16-bit without touching float (lib.rs:3349-3356)
    v = rng.integers(-(1 << 23), 1 << 23, 2 * 480, dtype=np.int64)
    raw = np.stack([v & 0xff, (v >> 8) & 0xff, (v >> 16) & 0xff], axis=1).astype(np.uint8).reshape(-1)
    out = D.apply_output_options(audio(24, 2, 48000, raw.tobytes()), D.DecodeOptions(output_bits_per_sample=16), D.ResamplerSlot())
    assert len(out) == 1 and out[0].bits_per_sample() == 16 and out[0].channel_count() == 2
    assert np.array_equal(np.frombuffer(out[0].data(), "<i2"), (v >> 8).astype(np.int16))
    # stereo -> mono at the same rate: scale-then-sum downmix, then float_sample_to_i16 (lib.rs:3433-3455)
    x = rng.uniform(-1.0, 1.0, (2, 333)).astype(np.float32)
    out = D.apply_output_options(f32_audio(48000, [x[0], x[1]]), D.DecodeOptions(output_bits_per_sample=16, output_channels=1), D.ResamplerSlot())
    mono = (np.float32(0.0) + x[0] * np.float32(0.5)) + x[1] * np.float32(0.5)
    want = np.array([D.float_sample_to_i16(float(s)) for s in mono], np.int16)
    assert np.array_equal(np.frombuffer(out[0].data(), "<i2"), want)


# ---------------------------------------------------------------------------------------------- GPU: 48 kHz stereo normalizer
def pcm16_audio(rate, channels):  # audio_pipeline.rs:770-779
    return audio(16, len(channels), rate, AB.interleave_vecs_i16([list(map(int, c)) for c in channels]))


def normalize_blocks(blocks):  # audio_pipeline.rs:781-795
    nz = AP.StreamingStereo48kNormalizer()
    left, right = [], []
    for b in blocks:
        out = nz.push(b)
        if out is not None:
            left += out[0].tolist(); right += out[1].tolist()
    out = nz.finish()
    if out is not None:
        left += out[0].tolist(); right += out[1].tolist()
    return left, right


@pytest.mark.gpu
def test_streaming_normalizer_converts_pcm_without_intermediate_channel_planes():  # audio_pipeline.rs:715-755
    eps = float(np.finfo(np.float32).eps)
    l, r = AP.StreamingStereo48kNormalizer().push(audio(16, 2, 48000, struct.pack(">hh", 32767, -32768), SIGNED, BE))
    assert abs(l[0] - 32767.0 / 32768.0) < eps and r[0] == -1.0
    l, r = AP.StreamingStereo48kNormalizer().push(audio(24, 2, 48000, bytes([0xff, 0xff, 0x7f, 0x00, 0x00, 0x80])))
    assert abs(l[0] - 8388607.0 / 8388608.0) < eps and r[0] == -1.0
    l, r = AP.StreamingStereo48kNormalizer().push(audio(32, 1, 48000, struct.pack("<f", float("nan")), FLOAT))
    assert l.tolist() == [0.0] and r.tolist() == l.tolist()


@pytest.mark.gpu
def test_streaming_stereo_normalizer_duplicates_mono_without_resampling():  # audio_pipeline.rs:797-804
    left, right = normalize_blocks([pcm16_audio(48000, [[-32768, 0, 32767]])])
    assert left == right and len(left) == 3 and left[0] == -1.0 and left[1] == 0.0


@pytest.mark.gpu
def test_streaming_stereo_normalizer_is_independent_of_decoder_chunking():  # audio_pipeline.rs:806-823
    left = [s * 211 - 10000 for s in range(100)]
    right = [10000 - s * 173 for s in range(100)]
    whole = pcm16_audio(44100, [left, right])
    chunks = [pcm16_audio(44100, [left[:37], right[:37]]), pcm16_audio(44100, [left[37:73], right[37:73]]),
              pcm16_audio(44100, [left[73:], right[73:]])]
    expected, actual = normalize_blocks([whole]), normalize_blocks(chunks)
    assert actual == expected and len(actual[0]) == 109


@pytest.mark.gpu
def test_streaming_normalizer_error_strings():  # audio_pipeline.rs:288-309, 366-373, 133-148
    nz = AP.StreamingStereo48kNormalizer()
    with pytest.raises(H.SoundkitError, match="decoded source contained no PCM"):
        nz.finish()
    with pytest.raises(H.SoundkitError, match="48 kHz stereo normalizer is already finished"):
        nz.push(pcm16_audio(8000, [[1, 2]]))
    nz = AP.StreamingStereo48kNormalizer()
    with pytest.raises(H.SoundkitError, match="decoded audio has invalid PCM geometry"):
        nz.push(audio(16, 0, 8000, b"\0\0"))
    with pytest.raises(H.SoundkitError, match="unsupported PCM bit depth 8"):
        nz.push(audio(8, 1, 8000, b"\0\0"))
    with pytest.raises(H.SoundkitError, match="decoded audio contained an incomplete PCM frame"):
        nz.push(audio(16, 1, 8000, b"\0\0\0"))
    assert nz.push(pcm16_audio(8000, [[5]])) is None  # one frame: nothing to interpolate yet
    with pytest.raises(H.SoundkitError, match=r"decoded PCM geometry changed from 8000 Hz/1 ch to 8000 Hz/2 ch"):
        nz.push(pcm16_audio(8000, [[1], [2]]))
    assert (nz.source_sample_rate(), nz.source_channels(), nz.source_frames()) == (8000, 1, 1)
    tail = nz.finish()
    assert len(tail[0]) == 6 and nz.output_frames() == 6  # round(1 * 48000 / 8000)


# ---------------------------------------------------------------------------------------------- GPU: soundkit_g726 (8f row 4)
G6 = H.g726
G726_RATES = [(G6.G726Rate.Rate16000, "16"), (G6.G726Rate.Rate24000, "24"), (G6.G726Rate.Rate32000, "32"), (G6.G726Rate.Rate40000, "40")]


def g726_samples():  # soundkit-g726/src/lib.rs:1031-1038
    idx = np.arange(397, dtype=np.float32)
    return (np.sin(idx / np.float32(80.0) * np.float32(2.0 * math.pi)).astype(np.float32) * np.float32(10000.0)).astype(np.int16).tolist()


def g726_encode_samples(x, rate, packing):  # encode_samples, lib.rs:1069-1077
    enc = G6.G726Encoder.try_new(rate, packing)
    out = b""
    for i in range(0, len(x), 37):
        out += enc.encode_to_vec(x[i:i + 37])
    return out + enc.flush_to_vec()


@pytest.mark.gpu
def test_g726_streaming_encoder_matches_padded_whole_encode_for_all_rates():  # lib.rs:1079-1108
    x = g726_samples()
    for rate, _ in G726_RATES:
        gs, gb = G6.samples_per_byte_group(rate), G6.bytes_per_group(rate)
        padded = x + [0] * ((-len(x)) % gs)
        whole = G6.G726Encoder.try_new(rate, G6.G726Packing.Left).encode_i16(padded, len(padded) // gs * gb)
        enc = G6.G726Encoder.try_new(rate, G6.G726Packing.Left)
        chunked = b""
        for i in range(0, len(x), 37):
            chunked += enc.encode_i16(x[i:i + 37], 64)
        chunked += enc.flush_to_vec()
        assert chunked == whole, rate


@pytest.mark.gpu
def test_g726_streaming_decoder_matches_whole_decode_for_all_rates():  # lib.rs:1110-1140
    for rate, _ in G726_RATES:
        gs, gb = G6.samples_per_byte_group(rate), G6.bytes_per_group(rate)
        encoded = g726_encode_samples(g726_samples(), rate, G6.G726Packing.Left)
        whole = G6.G726Decoder.try_new(rate, G6.G726Packing.Left).decode_i16(encoded, len(encoded) // gb * gs, False)
        dec = G6.G726Decoder.try_new(rate, G6.G726Packing.Left)
        chunked = []
        for i in range(len(encoded)):
            chunked += dec.decode_i16(encoded[i:i + 1], 64, False).tolist()
        dec.flush()
        assert chunked == whole.tolist(), rate


@pytest.mark.gpu
def test_g726_right_packing_round_trips_and_trait_outputs():  # lib.rs:1142-1190
    encoded = g726_encode_samples(g726_samples(), G6.G726Rate.Rate32000, G6.G726Packing.Right)
    decoded = G6.G726Decoder.new_32k_right().decode_i16(encoded, len(encoded) * 2, False)
    assert len(decoded) == len(encoded) * 2 and np.any(decoded != 0)
    encoded = g726_encode_samples(g726_samples(), G6.G726Rate.Rate32000, G6.G726Packing.Left)
    n = len(encoded) * 2
    i16 = G6.G726Decoder.new_32k_left().decode_i16(encoded, n, False)
    i32 = G6.G726Decoder.new_32k_left().decode_i32(encoded, n, False)
    f32 = G6.G726Decoder.new_32k_left().decode_f32(encoded, n, False)
    assert len(i16) == len(i32) == len(f32) == n
    assert np.array_equal(i32, i16.astype(np.int32) << 16)
    assert np.all(np.abs(f32 - i16.astype(np.float32) / np.float32(32768.0)) < np.finfo(np.float32).eps)


@pytest.mark.gpu
def test_g726_decode_fixture_golden_and_error_strings():  # lib.rs:1113-1145 of the reference tests; error texts :644-668, :700-709, :802-833
    for rate, suffix in G726_RATES:
        gs, gb = G6.samples_per_byte_group(rate), G6.bytes_per_group(rate)
        fixture = golden(f"g726_{suffix}.g726")
        dec = G6.G726Decoder.try_new(rate, G6.G726Packing.Left)
        decoded = []
        for i in range(0, len(fixture), 127):
            decoded.append(dec.decode_i16(fixture[i:i + 127], 512, False))
        dec.flush()
        decoded = np.concatenate(decoded)
        assert len(decoded) == len(fixture) // gb * gs
        assert decoded.astype("<i2").tobytes() == golden(f"g726_{suffix}.decoded.s16le")
    with pytest.raises(H.SoundkitError, match=r"Output buffer too small for G.726 decode: need 8, have 7"):
        G6.G726Decoder.try_new(G6.G726Rate.Rate24000, G6.G726Packing.Left).decode_i16(bytes(3), 7, False)
    with pytest.raises(H.SoundkitError, match=r"Output buffer too small for G.726 encode: need 5, have 4"):
        G6.G726Encoder.try_new(G6.G726Rate.Rate40000, G6.G726Packing.Left).encode_i16([0] * 8, 4)
    d = G6.G726Decoder.try_new(G6.G726Rate.Rate40000, G6.G726Packing.Left)
    assert len(d.decode_i16(bytes(7), 64, False)) == 8
    with pytest.raises(H.SoundkitError, match=r"G.726 stream ended with 2 trailing partial-packet byte\(s\)"):
        d.flush()
    e = G6.G726Encoder.try_new(G6.G726Rate.Rate16000, G6.G726Packing.Left)
    assert e.encode_i16([1, 2, 3], 8) == b""
    with pytest.raises(H.SoundkitError, match=r"Output buffer too small for G.726 flush: need 1, have 0"):
        e.flush_into(0)
    assert G6.from_bitrate(24000) == G6.G726Rate.Rate24000 and G6.from_bitrate(12345) == G6.G726Rate.Rate32000


@pytest.mark.gpu
def test_apply_output_options_mid_stream_errors_and_surround(oracle):
    """lib.rs:3395-3403 (the resampler is bound to the first frame's geometry) and the > 2 channel arm: resample per channel,
    then downmix_channels(.., 2) with its block-wide peak normalisation, then bytes"""
    rng = np.random.default_rng(3)
    opts = D.DecodeOptions(output_bits_per_sample=16, output_sample_rate=16000, output_channels=1)
    slot = D.ResamplerSlot()
    x = rng.uniform(-0.5, 0.5, (2, 5000)).astype(np.float32)
    D.apply_output_options(f32_audio(48000, [x[0], x[1]]), opts, slot)
    with pytest.raises(H.DecodeError, match="Resampler configuration changed mid-stream"):
        D.apply_output_options(f32_audio(44100, [x[0], x[1]]), opts, slot)          # input rate changed
    with pytest.raises(H.DecodeError, match="Resampler configuration changed mid-stream"):
        D.apply_output_options(f32_audio(48000, [x[0]]), opts, slot)                # channel count changed
    with pytest.raises(H.DecodeError, match="Resampler configuration changed mid-stream"):
        D.apply_output_options(f32_audio(48000, [x[0], x[1]]), D.DecodeOptions(output_bits_per_sample=16, output_sample_rate=8000, output_channels=1), slot)
    # 5.1 float -> stereo s24 at 24 kHz, whole stream in one push + flush
    ch = rng.uniform(-0.9, 0.9, (6, 9000)).astype(np.float32)
    opts = D.DecodeOptions(output_bits_per_sample=24, output_sample_rate=24000, output_channels=2)
    slot = D.ResamplerSlot()
    frames = D.apply_output_options(f32_audio(48000, list(ch)), opts, slot) + D.flush_resampler_frames(slot)
    got = b"".join(f.data() for f in frames)
    st = oracle.Stream(48000, 24000, 6)
    want = b""
    for part in (st.process(ch), st.flush()):   # the reference emits (and peak-normalises) one block per resampler chunk batch
        if part.shape[1]:
            want += oracle.f32_channels_to_bytes(oracle.downmix_channels(part, 2), 24, 0).tobytes()
    assert all(f.bits_per_sample() == 24 and f.channel_count() == 2 and f.sampling_rate() == 24000 for f in frames)
    assert len(got) == len(want)


# ---------------------------------------------------------------------------------------------- soundkit_g722 (8f row 4)
G2 = H.g722


def g722_samples():  # soundkit-g722/src/lib.rs:238-245
    idx = np.arange(161, dtype=np.float32)
    phase = idx / np.float32(160.0) * np.float32(2.0 * math.pi) * np.float32(3.0)
    return (np.sin(phase).astype(np.float32) * np.float32(12000.0)).astype(np.int16).tolist()


def test_g722_encoder_buffer_checks_and_pending_sample_need_no_device():  # :83-100, 33-46: checked before the ABI is crossed
    e = G2.G722Encoder.new_64k()
    with pytest.raises(H.SoundkitError, match=r"Output buffer too small for G\.722 encode: need 2, have 1"):
        e.encode_i16([1, 2, 3, 4], 1)
    assert e.encode_i16([7], 0) == b""           # a single sample is kept pending, nothing is encoded
    assert e.encode_i16([], 0) == b""
    d = G2.G722Decoder.new_64k()
    assert (d.sample_rate(), d.channels()) == (16000, 1)
    with pytest.raises(H.SoundkitError, match=r"Output buffer too small for G\.722 decode: need 6, have 5"):
        d.decode_i16(bytes(3), 5, False)
    e2 = G2.G722Encoder(16000, 16, 1, 320, 64000)  # Encoder::new ignores every argument (:68-76)
    assert e2.flush_to_vec() == b""


@pytest.mark.gpu
def test_g722_streaming_encoder_matches_padded_whole_encode():  # :246-268
    x = g722_samples()
    padded = x + [0]
    whole = G2.G722Encoder.new_64k().encode_i16(padded, len(padded) // 2)
    enc = G2.G722Encoder.new_64k()
    chunked = b""
    for i in range(0, len(x), 5):
        chunked += enc.encode_i16(x[i:i + 5], 8)
    chunked += enc.flush_to_vec()
    assert chunked == whole and len(whole) == 81
    with pytest.raises(H.SoundkitError, match=r"Output buffer too small for G\.722 flush: need 1, have 0"):
        e = G2.G722Encoder.new_64k()
        e.encode_i16([5], 0)
        e.flush_into(0)


@pytest.mark.gpu
def test_g722_streaming_decoder_matches_whole_decode_and_trait_outputs():  # :270-328
    x = g722_samples()
    enc = G2.G722Encoder.new_64k()
    encoded = enc.encode_to_vec(x) + enc.flush_to_vec()
    whole = G2.G722Decoder.new_64k().decode_i16(encoded, len(encoded) * 2, False)
    dec = G2.G722Decoder.new_64k()
    chunked = np.concatenate([dec.decode_i16(encoded[i:i + 3], 6, False) for i in range(0, len(encoded), 3)])
    assert np.array_equal(chunked, whole)
    i32 = G2.G722Decoder.new_64k().decode_i32(encoded, len(encoded) * 2, False)
    f32 = G2.G722Decoder.new_64k().decode_f32(encoded, len(encoded) * 2, False)
    assert np.array_equal(i32, whole.astype(np.int32) << 16)
    assert np.max(np.abs(f32 - whole.astype(np.float32) / np.float32(32768.0))) < np.finfo(np.float32).eps


@pytest.mark.gpu
def test_g722_decode_fixture_golden():  # :330-360: 127-byte chunks into a 254-sample scratch
    raw = golden("g722.g722")
    want = np.frombuffer(golden("g722.decoded.s16le"), "<i2")
    dec = G2.G722Decoder.new_64k()
    got = np.concatenate([dec.decode_i16(raw[i:i + 127], 254, False) for i in range(0, len(raw), 127)])
    assert got.size == len(raw) * 2 and np.any(got != 0)
    assert np.array_equal(got, want)
