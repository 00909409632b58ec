"""Float64 evaluation of the algorithm rubato 0.14.1 `SincFixedIn` publishes (sinc.rs make_sincs + windows.rs blackman_harris,
asynchro_sinc.rs process_into_buffer / get_nearest_times_2, interpolation.rs interp_lin) — TEST INFRASTRUCTURE.

It is an independent check of oracle/rubato_sinc.c (the f32 restatement the GPU kernels are compared with bit for bit): the
table is built in f64 with f64 sin / cos, every dot product is summed in f64, so the only thing it shares with the f32 oracle is
the published formula.  A wrong tap alignment, cutoff scaling, window denominator or interpolation weight in the f32 oracle
shows up as an error orders of magnitude above f32 rounding noise (tests/test_oracle_golden.py bounds it at 2e-6 / 125 dB)."""
import math

import numpy as np

SINC_LEN, FACTOR = 256, 256


def make_sincs_f64(ratio):
    """[FACTOR][SINC_LEN] table in f64.  f_cutoff follows make_interpolator: 0.95 (f32 literal) scaled by the ratio when
    downsampling (the f32 product, as rubato computes it, then widened)."""
    fc = float(np.float32(0.95)) if ratio >= 1.0 else float(np.float32(0.95) * np.float32(ratio))
    tot = SINC_LEN * FACTOR
    x = np.arange(tot, dtype=np.float64)
    w = 0.35875 - 0.48829 * np.cos(2 * np.pi * x / tot) + 0.14128 * np.cos(4 * np.pi * x / tot) - 0.01168 * np.cos(6 * np.pi * x / tot)
    w = w * w                                     # BlackmanHarris2
    v = (x - tot // 2) * fc / FACTOR
    y = w * np.sinc(v)                            # numpy sinc(v) = sin(pi v) / (pi v), 1 at 0
    y /= y.sum() / FACTOR
    sincs = np.empty((FACTOR, SINC_LEN), np.float64)
    for n in range(FACTOR):
        sincs[FACTOR - n - 1, :] = y[n::FACTOR][:SINC_LEN]  # sincs[factor - 1 - n][p] = y[factor * p + n]
    return sincs


def walk(last_index, ratio, chunk):
    """the sequential f64 index walk of one chunk -> (idx values, new last_index)"""
    t_ratio = 1.0 / ratio
    end_idx = chunk - (SINC_LEN + 1) - math.ceil(t_ratio)
    out, idx = [], last_index
    while idx < end_idx:
        idx += t_ratio
        out.append(idx)
    return np.array(out, np.float64), idx - chunk


def resample_chunk_f64(sincs, buf, idxs):
    """buf: f64 [2 * SINC_LEN + chunk] (history then chunk); idxs relative to the chunk start"""
    fl = np.floor(idxs)
    i0 = fl.astype(np.int64)
    s0 = np.floor((idxs - fl) * FACTOR).astype(np.int64)
    f = idxs * FACTOR
    frac = f - np.floor(f)
    s1 = s0 + 1
    i1 = np.where(s1 < FACTOR, i0, i0 + 1)
    s1 = np.where(s1 < FACTOR, s1, 0)
    j = np.arange(SINC_LEN)
    p0 = np.einsum("nj,nj->n", buf[(i0 + 2 * SINC_LEN)[:, None] + j], sincs[s0])
    p1 = np.einsum("nj,nj->n", buf[(i1 + 2 * SINC_LEN)[:, None] + j], sincs[s1])
    return p0 + frac * (p1 - p0)


def downsample_oneshot_f64(x, in_rate, out_rate):
    """one channel, downsample_audio semantics (one chunk = the whole buffer, no flush); x: 1-D array"""
    ratio = out_rate / in_rate
    sincs = make_sincs_f64(ratio)
    n = len(x)
    buf = np.zeros(2 * SINC_LEN + n + 8, np.float64)
    buf[2 * SINC_LEN:2 * SINC_LEN + n] = x
    idxs, _ = walk(-(SINC_LEN / 2), ratio, n)
    return resample_chunk_f64(sincs, buf, idxs)


def stream_f64(x, in_rate, out_rate, chunk=4096):
    """one channel, StreamingResampler semantics (chunks of 4096 + flush with trim)"""
    ratio = out_rate / in_rate
    sincs = make_sincs_f64(ratio)
    hist = np.zeros(2 * SINC_LEN, np.float64)
    last = -(SINC_LEN / 2)
    outs, pos = [], 0
    x = np.asarray(x, np.float64)

    def run(block):
        nonlocal hist, last
        buf = np.concatenate([hist, block, np.zeros(8)])
        idxs, last = walk(last, ratio, chunk)
        y = resample_chunk_f64(sincs, buf, idxs)
        hist = buf[chunk:chunk + 2 * SINC_LEN].copy()
        return y

    while len(x) - pos >= chunk:
        outs.append(run(x[pos:pos + chunk]))
        pos += chunk
    rem = len(x) - pos
    block = np.zeros(chunk)
    block[:rem] = x[pos:]
    y = run(block)
    if rem > 0:
        trim = int(math.floor((chunk - rem) * out_rate / in_rate + 0.5))
        y = y[:max(0, len(y) - trim)]
    outs.append(y)
    return np.concatenate(outs)
