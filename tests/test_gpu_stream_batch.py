"""sk_resampler_batch_*: N StreamingResamplers in lock step with device-resident state (soundkit-decoder/src/lib.rs:1917-2060).
Every push of every stream must equal the CPU restatement of StreamingResampler::process for that stream (same frames per push,
same samples bit for bit), flush included; the concatenation equals the whole-stream result of the fused batch chain."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from soundkit_b200 import capi  # noqa: E402


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


PUSHES = [1000, 4096, 5000, 300, 8192, 1, 4095, 2, 6000]


@pytest.mark.parametrize("rates", [(48000, 16000), (44100, 48000), (44100, 16000), (8000, 16000), (48000, 8000)])
@pytest.mark.parametrize("n_streams,channels", [(3, 2), (70, 1), (33, 2)])
def test_batch_streaming_matches_oracle_push_by_push(oracle, rates, n_streams, channels):
    import torch
    i, o = rates
    rng = np.random.default_rng(i + 3 * o + n_streams)
    total = sum(PUSHES)
    x = rng.uniform(-0.5, 0.5, (n_streams, channels, total)).astype(np.float32)
    c = capi.Context(0)
    c.use_torch_stream()
    try:
        rb = c.resampler_batch(i, o, channels, n_streams, max_push=8192)
        xd = torch.from_numpy(x).cuda()
        out = torch.zeros((n_streams, channels, rb.out_max), dtype=torch.float32, device="cuda")
        check = sorted({0, n_streams // 2, n_streams - 1})
        refs = {s: oracle.Stream(i, o, channels) for s in check}
        pos, pieces = 0, {s: [] for s in check}
        for k, n in enumerate(PUSHES):
            got_n = rb.process_dev(xd[:, :, pos:pos + n], out)
            torch.cuda.synchronize()
            got = out[:, :, :got_n].cpu().numpy()
            for s in check:
                want = refs[s].process(x[s, :, pos:pos + n])
                assert want.shape == (channels, got_n), (k, want.shape, got_n)
                assert np.array_equal(_bits(got[s]), _bits(want)), f"push {k} stream {s}"
                pieces[s].append(got[s])
            pos += n
        assert rb.pending == total % 4096
        got_n = rb.flush_dev(out)
        torch.cuda.synchronize()
        got = out[:, :, :got_n].cpu().numpy()
        for s in check:
            want = refs[s].flush()
            assert want.shape == (channels, got_n)
            assert np.array_equal(_bits(got[s]), _bits(want)), f"flush stream {s}"
            pieces[s].append(got[s])
        # the concatenation is the whole-stream result (streaming semantics of the fused batch chain)
        n_whole = c.batch_plan(i, o, total, capi.SEM_STREAMING)
        for s in check:
            assert np.concatenate(pieces[s], axis=1).shape == (channels, n_whole)
        inter = np.ascontiguousarray(np.transpose(x[check[0]], (1, 0))).reshape(-1)
        whole = c.normalize_stream(inter.view(np.uint8), 32, capi.FMT_FLOAT, False, channels, capi.MODE_A, i, o, capi.SEM_STREAMING, capi.MIX_NONE)
        assert np.array_equal(_bits(whole), _bits(np.concatenate(pieces[check[0]], axis=1)))
        # a second flush with nothing pending: one chunk of zeros through the carried history (process_partial(None))
        got_n2 = rb.flush_dev(out)
        torch.cuda.synchronize()
        want2 = refs[check[0]].flush()
        assert want2.shape[1] == got_n2
        assert np.array_equal(_bits(out[check[0], :, :got_n2].cpu().numpy()), _bits(want2))
        rb.close()
    finally:
        c.close()


class _DevRows:
    """lanes x frames f32 view of device memory at a raw address (torch reads __cuda_array_interface__)"""

    def __init__(self, ptr, lanes, frames, pitch_floats):
        self.__cuda_array_interface__ = {"shape": (lanes, frames), "typestr": "<f4", "data": (ptr, False), "version": 2, "strides": (pitch_floats * 4, 4)}


@pytest.mark.parametrize("rates", [(48000, 16000), (44100, 48000), (48000, 24000)])
def test_batch_streaming_in_place_append_equals_the_copy_form(rates):
    """sk_resampler_batch_input_dev: a device-side producer writes each push straight into the handle's rows and calls process with
    in == NULL; every push must produce exactly what the copying form produces (which the test above pins to the oracle)"""
    import torch
    i, o = rates
    n_streams, channels = 9, 2
    rng = np.random.default_rng(i + o)
    total = sum(PUSHES)
    x = torch.from_numpy(rng.uniform(-0.5, 0.5, (n_streams, channels, total)).astype(np.float32)).cuda()
    c = capi.Context(0)
    c.use_torch_stream()
    try:
        ra, rb = c.resampler_batch(i, o, channels, n_streams, max_push=8192), c.resampler_batch(i, o, channels, n_streams, max_push=8192)
        oa = torch.zeros((n_streams, channels, ra.out_max), dtype=torch.float32, device="cuda")
        ob = torch.zeros_like(oa)
        pos = 0
        for k, n in enumerate(PUSHES):
            na = ra.process_dev(x[:, :, pos:pos + n], oa)
            ptr, pitch = rb.input_dev()
            rows = torch.as_tensor(_DevRows(ptr, n_streams * channels, n, pitch), device="cuda")
            rows.copy_(x[:, :, pos:pos + n].reshape(n_streams * channels, n))
            nb = rb.process_in_place_dev(n, ob)
            torch.cuda.synchronize()
            assert na == nb, k
            assert torch.equal(oa[:, :, :na].view(torch.int32), ob[:, :, :nb].view(torch.int32)), k
            pos += n
        assert ra.pending == rb.pending
        na, nb = ra.flush_dev(oa), rb.flush_dev(ob)
        torch.cuda.synchronize()
        assert na == nb and torch.equal(oa[:, :, :na].view(torch.int32), ob[:, :, :nb].view(torch.int32))
        ra.close(); rb.close()
    finally:
        c.close()


def test_batch_streaming_argument_errors():
    c = capi.Context(0)
    try:
        with pytest.raises(capi.SoundkitError) as e:
            c.resampler_batch(0, 16000, 2, 4)
        assert "Input sample rate must be > 0" in str(e.value)
        with pytest.raises(capi.SoundkitError) as e:
            c.resampler_batch(48000, 0, 2, 4)
        assert "Output sample rate must be > 0" in str(e.value)
        with pytest.raises(capi.SoundkitError) as e:
            c.resampler_batch(48000, 16000, 0, 4)
        assert "Channel count must be > 0" in str(e.value)
        import torch
        rb = c.resampler_batch(48000, 16000, 1, 2, max_push=100)
        x = torch.zeros((2, 1, 101), device="cuda")
        o = torch.zeros((2, 1, rb.out_max), device="cuda")
        with pytest.raises(capi.SoundkitError):
            rb.process_dev(x, o)
        rb.close()
    finally:
        c.close()
