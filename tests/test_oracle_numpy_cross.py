"""A second, independent restatement (vectorised numpy, written from the reference source, not from oracle/*.c) of the
integer / exact-float conversions and G.711, checked against the C oracle on random data.  Two restatements that agree
bit for bit make a transcription slip in either unlikely; the reference's own unit vectors pin both (test_oracle_golden)."""
import numpy as np
import pytest

RNG = np.random.default_rng(20261018)


def rust_as_i32(x):  # f32 -> i32 `as`: truncate toward zero, saturate, NaN -> 0
    x = np.asarray(x, np.float32)
    y = np.where(np.isnan(x), np.float32(0), x)
    y = np.clip(np.trunc(y.astype(np.float64)), -2147483648.0, 2147483647.0)
    return y.astype(np.int64).astype(np.int32)


def s24(raw, be=False):
    b = raw[: raw.size // 3 * 3].reshape(-1, 3).astype(np.int64)
    u = (b[:, 0] << 16 | b[:, 1] << 8 | b[:, 2]) if be else (b[:, 2] << 16 | b[:, 1] << 8 | b[:, 0])
    return np.where(u & 0x800000, u - (1 << 24), u)


def f32s(raw, be=False):
    return raw[: raw.size // 4 * 4].view(">f4" if be else "<f4").astype(np.float32)


def clamp1(x):  # Rust f32::clamp keeps NaN
    return np.where(np.isnan(x), x, np.minimum(np.maximum(x, np.float32(-1)), np.float32(1))).astype(np.float32)


NUMPY_FLAT = {
    # soundkit/src/audio_bytes.rs
    "i16le_to_f32": lambda r: (r[: r.size // 2 * 2].view("<i2").astype(np.float32) / np.float32(32768.0)),
    "s16le_to_i16": lambda r: r[: r.size // 2 * 2].view("<i2").astype(np.int16),
    "s16be_to_i16": lambda r: r[: r.size // 2 * 2].view(">i2").astype(np.int16),
    "s16le_to_i32": lambda r: r[: r.size // 2 * 2].view("<i2").astype(np.int32),
    "s24le_to_i32": lambda r: s24(r).astype(np.int32),
    "s24le_to_i16": lambda r: (s24(r) >> 8).astype(np.int16),
    "s24be_to_i16": lambda r: (s24(r, True) >> 8).astype(np.int16),
    "s32le_to_i32": lambda r: r[: r.size // 4 * 4].view("<i4").astype(np.int32),
    "s32be_to_i32": lambda r: r[: r.size // 4 * 4].view(">i4").astype(np.int32),
    "s32le_to_s24": lambda r: (r[: r.size // 4 * 4].view("<i4").astype(np.int64) & 0xFFFFFF).astype(np.int32),
    "s32be_to_s24": lambda r: (r[: r.size // 4 * 4].view(">i4").astype(np.int64) & 0xFFFFFF).astype(np.int32),
    "s32le_to_f32": lambda r: r[: r.size // 4 * 4].view("<i4").astype(np.float32) / (np.float32(2.0) ** 31 - np.float32(1.0)),
    "s32be_to_f32": lambda r: r[: r.size // 4 * 4].view(">i4").astype(np.float32) / (np.float32(2.0) ** 31 - np.float32(1.0)),
    "s32le_to_i16": lambda r: (r[: r.size // 4 * 4].view("<i4").astype(np.int64) >> 16).astype(np.int16),
    "s32be_to_i16": lambda r: (r[: r.size // 4 * 4].view(">i4").astype(np.int64) >> 16).astype(np.int16),
    "f32le_to_i16": lambda r: rust_as_i32(clamp1(f32s(r)) * np.float32(32767.0)).astype(np.int16),
    "f32be_to_i16": lambda r: rust_as_i32(clamp1(f32s(r, True)) * np.float32(32767.0)).astype(np.int16),
    "f32le_to_i32": lambda r: rust_as_i32(clamp1(f32s(r)) * np.float32(2147483648.0)),
    "f32le_to_s24": lambda r: rust_as_i32(np.where(clamp1(f32s(r)) >= 0, clamp1(f32s(r)) * np.float32(8388607.0), clamp1(f32s(r)) * np.float32(8388608.0))),
    # soundkit/src/audio_packet.rs, soundkit-decoder/src/lib.rs:362-367
    "pkt_s16le_to_f32": lambda r: r[: r.size // 2 * 2].view("<i2").astype(np.float32) / np.float32(32767.0),
    "pkt_s24le_to_f32": lambda r: s24(r).astype(np.float32) / np.float32(8388608.0),
    "pkt_s32le_to_i16": lambda r: np.clip(np.trunc(r[: r.size // 4 * 4].view("<i4").astype(np.int64) * 32767 / 2147483647.0), -32768, 32767).astype(np.int16),
    "pkt_f32le_to_i16": lambda r: np.clip(rust_as_i32(f32s(r) * np.float32(32767.0)), -32768, 32767).astype(np.int16),
    "u8_to_s16": lambda r: ((r.astype(np.int32) - 128) << 8).astype(np.int16),
}


def float_source(n):
    x = RNG.uniform(-1.3, 1.3, n).astype(np.float32)
    x[::31] = np.nan; x[1::37] = np.inf; x[2::41] = -np.inf; x[3::43] = -0.0; x[4::47] = 1.0; x[5::53] = -1.0
    return x.view(np.uint8)


@pytest.mark.parametrize("name", sorted(NUMPY_FLAT))
def test_flat_ops_two_restatements_agree(oracle, name):
    assert set(NUMPY_FLAT) == set(oracle.FLAT_NAMES)
    n = 20011
    src = float_source(n) if "f32" in name.split("_to_")[0] else RNG.integers(0, 256, n * 4 + 3, dtype=np.uint8)
    if name == "pkt_s32le_to_i16":  # exact integer division, not float: redo in Python ints on a subset
        v = src[: src.size // 4 * 4].view("<i4").astype(np.int64)
        want = np.array([max(-32768, min(32767, int(abs(int(x)) * 32767 // 2147483647) * (1 if x >= 0 else -1))) for x in v[:5000]], np.int16)
        assert np.array_equal(oracle.flat(name, src)[:5000], want)
        return
    got, want = oracle.flat(name, src), NUMPY_FLAT[name](src)
    assert got.dtype == want.dtype and np.array_equal(got.view(np.uint8), want.view(np.uint8))


def g711_numpy_decode(law, codes):  # soundkit-g711/src/lib.rs:256-315, vectorised
    b = codes.astype(np.int32)
    if law == 0:
        u = (~b) & 0xFF
        m = (((u & 0x0F) << 3) + 0x84) << ((u & 0x70) >> 4)
        return np.where(u & 0x80, 0x84 - m, m - 0x84).astype(np.int16)
    a = b ^ 0x55
    seg, m = (a & 0x70) >> 4, (a & 0x0F) << 4
    m = np.where(seg == 0, m + 8, np.where(seg == 1, m + 0x108, (m + 0x108) << np.maximum(seg - 1, 0)))
    return np.where(a & 0x80, m, -m).astype(np.int16)


def g711_numpy_encode(law, s):  # :235-294 with segment_for (:317-322) as a searchsorted over SEGMENT_ENDS
    ends = np.array([0xFF, 0x1FF, 0x3FF, 0x7FF, 0xFFF, 0x1FFF, 0x3FFF, 0x7FFF])
    p = s.astype(np.int32)
    if law == 0:
        mask = np.where(p < 0, 0x7F, 0xFF)
        p = np.minimum(np.where(p < 0, 0x84 - p, p + 0x84), 32635)
        seg = np.searchsorted(ends, p, side="left")
        enc = np.where(seg >= 8, 0x7F, (seg << 4) | ((p >> (seg + 3)) & 0x0F))
        return ((enc ^ mask) & 0xFF).astype(np.uint8)
    mask = np.where(p >= 0, 0xD5, 0x55)
    p = np.where(p >= 0, p, -p - 1)
    seg = np.searchsorted(ends, p, side="left")
    enc = np.where(seg >= 8, 0x7F, (seg << 4) | ((p >> np.where(seg < 2, 4, seg + 3)) & 0x0F))
    return ((enc ^ mask) & 0xFF).astype(np.uint8)


@pytest.mark.parametrize("law", [0, 1])
def test_g711_two_restatements_agree_exhaustively(oracle, law):
    codes = np.arange(256, dtype=np.uint8)
    assert np.array_equal(oracle.g711_decode(law, codes), g711_numpy_decode(law, codes))
    samples = np.arange(-32768, 32768).astype(np.int16)
    assert np.array_equal(oracle.g711_encode(law, samples), g711_numpy_encode(law, samples))


def test_float_sample_to_i16_two_restatements_agree(oracle):  # soundkit-decoder/src/lib.rs:1815-1827
    x = np.concatenate([f32s(float_source(20000)), np.array([0.5 / 32767, 1.5 / 32767, -0.5 / 32768, -1.5 / 32768, 0.99999994, -0.99999994], np.float32)])
    f = np.where(np.isfinite(x), np.clip(x, -1, 1), 0).astype(np.float64)
    scaled = np.where(f < 0, f * 32768.0, f * 32767.0)
    want = np.clip(np.sign(scaled) * np.floor(np.abs(scaled) + 0.5), -32768, 32767).astype(np.int16)  # f64::round: half away from zero
    got = oracle.f32_channels_to_bytes(x.reshape(1, -1), 16, 0).view("<i2")
    assert np.array_equal(got, want)
