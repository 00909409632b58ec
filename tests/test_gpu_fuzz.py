"""Randomised differential tests (hypothesis): the CUDA path through the C ABI against the CPU oracle on random
formats, sizes, channel counts and chunkings.  Every comparison is bit-exact."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from soundkit_b200 import capi

pytestmark = pytest.mark.gpu
COMMON = dict(deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow], derandomize=True)
RATES = [8000, 16000, 22050, 24000, 32000, 44100, 48000, 88200, 96000]


def pcm_bytes(rng, bits, fmt, n):
    if fmt == 1:
        x = rng.uniform(-1.5, 1.5, n).astype(np.float32)
        if n:
            k = max(1, n // 17)
            x[rng.integers(0, n, k)] = rng.choice(np.array([np.nan, np.inf, -np.inf, -0.0, 1.0, -1.0, 1e-40], np.float32), k)
        return x.view(np.uint8)
    return rng.integers(0, 256, n * (bits // 8), dtype=np.uint8)


@settings(max_examples=60, **COMMON)
@given(seed=st.integers(0, 2 ** 31), bits_fmt=st.sampled_from([(16, 0), (24, 0), (32, 0), (32, 1)]), channels=st.integers(1, 8),
       frames=st.integers(0, 3000), be=st.booleans(), mode=st.sampled_from([capi.MODE_A, capi.MODE_B]))
def test_pcm_to_planar_and_back(ctx, oracle, seed, bits_fmt, channels, frames, be, mode):
    bits, fmt = bits_fmt
    rng = np.random.default_rng(seed)
    data = pcm_bytes(rng, bits, fmt, frames * channels)
    got, want = ctx.pcm_to_planar_f32(data, bits, fmt, be, channels, mode), oracle.pcm_to_planar_f32(data, bits, fmt, be, channels, mode)
    assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    if frames:
        x = rng.uniform(-1.2, 1.2, (channels, frames)).astype(np.float32)
        for ob, of in ((16, 0), (24, 0), (32, 0), (32, 1)):
            assert np.array_equal(ctx.planar_f32_to_pcm(x, ob, of), oracle.f32_channels_to_bytes(x, ob, of))
        if channels > 1:
            for mixmode, name in ((capi.MIX_SUM_THEN_SCALE, "mixdown_to_mono_f32"),):
                assert np.array_equal(ctx.mixdown_mono(x, mixmode).view(np.uint32), getattr(oracle, name)(x).view(np.uint32))


@settings(max_examples=40, **COMMON)
@given(seed=st.integers(0, 2 ** 31), n=st.integers(0, 70000), law=st.sampled_from([0, 1]), name=st.sampled_from(list(capi.FLAT_OPS)))
def test_g711_and_flat_ops(ctx, oracle, seed, n, law, name):
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, 256, n, dtype=np.uint8)
    pcm = rng.integers(-32768, 32768, n).astype(np.int16)
    assert np.array_equal(ctx.g711_decode(law, codes), oracle.g711_decode(law, codes))
    assert np.array_equal(ctx.g711_encode(law, pcm), oracle.g711_encode(law, pcm))
    iw, _ = ctx.flat_widths(capi.FLAT_OP[name])
    src = pcm_bytes(rng, 32, 1, n) if name.startswith(("f32", "pkt_f32")) else rng.integers(0, 256, n * iw + (n % 3), dtype=np.uint8)
    got, want = ctx.flat_convert(name, src), oracle.flat(name, src)
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8))


@settings(max_examples=25, **COMMON)
@given(seed=st.integers(0, 2 ** 31), bits=st.sampled_from([2, 3, 4, 5]), packing=st.sampled_from([0, 1]), groups=st.integers(1, 400),
       cuts=st.lists(st.integers(1, 399), max_size=4))
def test_g726_chunked_state_carry(ctx, oracle, seed, bits, packing, groups, cuts):
    rng = np.random.default_rng(seed)
    gs, gb = capi.G726_GROUP[bits]
    codes = rng.integers(0, 256, groups * gb, dtype=np.uint8)
    want = oracle.G726(bits, packing).decode(codes)
    st_, out = ctx.g726_state(), []
    edges = sorted({0, groups, *[c for c in cuts if c < groups]})
    for a, b in zip(edges[:-1], edges[1:]):
        out.append(ctx.g726_decode(bits, packing, st_, codes[a * gb:b * gb]))
    assert np.array_equal(np.concatenate(out), want)
    x = rng.integers(-32768, 32768, groups * gs).astype(np.int16)
    assert np.array_equal(ctx.g726_encode(bits, packing, ctx.g726_state(), x), oracle.G726(bits, packing).encode(x))


@settings(max_examples=30, **COMMON)
@given(seed=st.integers(0, 2 ** 31), bits_fmt=st.sampled_from([(16, 0), (24, 0), (32, 0), (32, 1)]), channels=st.integers(1, 4),
       rate=st.sampled_from(RATES + [11025, 12000, 192000]), be=st.booleans(), cuts=st.lists(st.integers(1, 2999), max_size=5),
       frames=st.integers(1, 3000))
def test_normalizer48k_random_chunking(ctx, oracle, seed, bits_fmt, channels, rate, be, cuts, frames):
    bits, fmt = bits_fmt
    rng = np.random.default_rng(seed)
    data = pcm_bytes(rng, bits, fmt, frames * channels)
    fb = (bits // 8) * channels
    gn, on = ctx.normalizer48k(), oracle.Normalizer48k()
    edges = sorted({0, frames, *[c for c in cuts if c < frames]})
    for a, b in zip(edges[:-1], edges[1:]):
        gl, gr = gn.push(data[a * fb:b * fb], bits, fmt, be, channels, rate)
        ol, orr = on.push(data[a * fb:b * fb], bits, fmt, be, channels, rate)
        assert np.array_equal(gl.view(np.uint32), ol.view(np.uint32)) and np.array_equal(gr.view(np.uint32), orr.view(np.uint32))
    gl, gr = gn.finish()
    ol, orr = on.finish()
    assert np.array_equal(gl.view(np.uint32), ol.view(np.uint32)) and np.array_equal(gr.view(np.uint32), orr.view(np.uint32))
    assert gn.output_frames() == on.output_frames()
    gn.close()


@settings(max_examples=25, **COMMON)
@given(seed=st.integers(0, 2 ** 31), bits_fmt=st.sampled_from([(16, 0), (24, 0), (32, 0), (32, 1)]), channels=st.integers(1, 3),
       rates=st.tuples(st.sampled_from(RATES), st.sampled_from(RATES)).filter(lambda r: r[0] != r[1]), frames=st.integers(300, 9000),
       streaming=st.booleans())
def test_resample_chain_random(ctx, oracle, seed, bits_fmt, channels, rates, frames, streaming):
    bits, fmt = bits_fmt
    i, o = rates
    rng = np.random.default_rng(seed)
    data = pcm_bytes(rng, bits, fmt, frames * channels) if fmt == 0 else rng.uniform(-1, 1, frames * channels).astype(np.float32).view(np.uint8)
    if streaming:
        planar = oracle.pcm_to_planar_f32(data, bits, fmt, False, channels, capi.MODE_B)
        s = oracle.Stream(i, o, channels)
        want = np.concatenate([s.process(planar), s.flush()], axis=1)
        got = ctx.normalize_stream(data, bits, fmt, False, channels, capi.MODE_B, i, o, capi.SEM_STREAMING, capi.MIX_NONE)
    else:
        want = oracle.downsample_audio(data, bits, fmt, channels, i, o)
        got = ctx.normalize_stream(data, bits, fmt, False, channels, capi.MODE_A, i, o, capi.SEM_ONESHOT, capi.MIX_NONE)
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
