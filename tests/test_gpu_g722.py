"""G.722 64 kbit/s on the device (sk_g722.cu) against the CPU restatement and the reference's golden fixture: bit-exact."""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu

from soundkit_b200 import capi  # noqa: E402

RNG = np.random.default_rng(722)


def test_g722_golden_fixture_decode(ctx):
    """testdata/g722/*.g722 -> golden/g722/*.decoded.wav (soundkit-g722/src/lib.rs:330-360): whole and in the reference test's 127-byte chunks"""
    raw = np.frombuffer(golden("g722.g722"), np.uint8)
    want = np.frombuffer(golden("g722.decoded.s16le"), "<i2")
    assert np.array_equal(ctx.g722_decode(ctx.g722_state(), raw), want)
    st = ctx.g722_state()
    parts = [ctx.g722_decode(st, raw[i:i + 127]) for i in range(0, raw.size, 127)]
    assert np.array_equal(np.concatenate(parts), want)
    # i32 / f32 forms of the Decoder trait (:183-232)
    i32 = ctx.g722_decode(ctx.g722_state(), raw[:999], capi.KIND_I32)
    f32 = ctx.g722_decode(ctx.g722_state(), raw[:999], capi.KIND_F32)
    assert np.array_equal(i32, want[:1998].astype(np.int32) << 16)
    assert np.array_equal(f32.view(np.uint32), (want[:1998].astype(np.float32) / np.float32(32768.0)).view(np.uint32))


@pytest.mark.parametrize("n", [0, 2, 6, 8, 322, 4098, 100000])
def test_g722_encode_decode_match_oracle(ctx, oracle, n):
    x = RNG.integers(-32768, 32768, n).astype(np.int16)
    x[: n // 3] = (x[: n // 3] // 64)  # a quiet stretch, a loud one (saturating predictors), full-scale noise
    codes = ctx.g722_encode(ctx.g722_state(), x)
    assert np.array_equal(codes, oracle.G722().encode(x))
    rnd = RNG.integers(0, 256, n // 2, dtype=np.uint8)  # arbitrary code bytes drive the decoder through every table entry
    assert np.array_equal(ctx.g722_decode(ctx.g722_state(), rnd), oracle.G722().decode(rnd))
    x32 = x.astype(np.int32) << 16 | RNG.integers(0, 65536, n).astype(np.int32)
    assert np.array_equal(ctx.g722_encode(ctx.g722_state(), x32), codes)  # encode_i32 = (s >> 16) as i16 (:119-122)


def test_g722_argument_errors(ctx):
    with pytest.raises(capi.SoundkitError) as e:
        ctx.g722_encode(ctx.g722_state(), np.zeros(3, np.int16))
    assert "whole number of sample pairs" in str(e.value)


def test_g722_batch_dev_streams_are_independent_and_chunkable(oracle):
    """65 streams (more than one warp, not a multiple of the block size) in one launch, state carried on the device across two
    calls with unaligned split points == each stream through the oracle alone"""
    import torch
    S, nbytes = 65, 3001
    codes = RNG.integers(0, 256, (S, nbytes), dtype=np.uint8)
    c = capi.Context(0)
    c.use_torch_stream()
    try:
        d_codes = torch.from_numpy(codes).cuda()
        d_state = torch.zeros((S, capi.G722_STATE_INTS), dtype=torch.int32, device="cuda")
        out = torch.zeros((S, 2 * nbytes), dtype=torch.int16, device="cuda")
        split = 1003
        c.g722_decode_batch_dev(d_codes, nbytes, S, split, capi.KIND_I16, out, 2 * nbytes, d_state)
        c.g722_decode_batch_dev(d_codes[:, split:], nbytes, S, nbytes - split, capi.KIND_I16, out[:, 2 * split:], 2 * nbytes, d_state)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        for s in (0, 1, 31, 32, 64):
            assert np.array_equal(got[s], oracle.G722().decode(codes[s])), s
        # encode the decoded audio again, stateless form (state = NULL: every stream from reset)
        pcm = out
        enc = torch.zeros((S, nbytes), dtype=torch.uint8, device="cuda")
        c.g722_encode_batch_dev(capi.KIND_I16, pcm, 2 * nbytes, S, 2 * nbytes, enc, nbytes, None)
        torch.cuda.synchronize()
        e = enc.cpu().numpy()
        for s in (0, 17, 64):
            assert np.array_equal(e[s], oracle.G722().encode(got[s])), s
    finally:
        c.close()
