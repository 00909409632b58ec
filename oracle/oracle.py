"""ctypes view of oracle/libsk_oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module (see oracle/sk_oracle.h).  The product package soundkit_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsk_oracle.so")

MULAW, ALAW = 0, 1
SIGNED, FLOAT = 0, 1
MODE_A, MODE_B = 0, 1
ERR = C.c_size_t(-1).value


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("g711.c", "pcm.c", "rubato_sinc.c", "pipeline.c", "sk_oracle.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libsk_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _proto(_lib)
    return _lib


_u8p, _i16p, _i32p, _f32p = (C.POINTER(t) for t in (C.c_uint8, C.c_int16, C.c_int32, C.c_float))


def _proto(L):
    L.sko_g711_encode_sample.restype = C.c_uint8
    L.sko_g711_encode_sample.argtypes = [C.c_int, C.c_int16]
    L.sko_g711_decode_sample.restype = C.c_int16
    L.sko_g711_decode_sample.argtypes = [C.c_int, C.c_uint8]
    for name in ("sko_g711_encode_i16", "sko_g711_encode_i32", "sko_g711_decode_i16", "sko_g711_decode_i32",
                 "sko_g711_decode_f32"):
        getattr(L, name).restype = None
        getattr(L, name).argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
    for name in _FLAT:
        getattr(L, "sko_" + name).restype = C.c_size_t
        getattr(L, "sko_" + name).argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    L.sko_interleave_vecs_i16.restype = None
    L.sko_interleave_vecs_i16.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
    for name in ("sko_deinterleave_vecs_i16", "sko_deinterleave_vecs_s24", "sko_deinterleave_vecs_f32"):
        getattr(L, name).restype = C.c_size_t
        getattr(L, name).argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
    for name in ("sko_stereo_to_mono_take_left", "sko_stereo_to_mono_avg", "sko_vec_f32_to_i16",
                 "sko_vec_i16_to_f32", "sko_vec_i32_to_f32", "sko_i16_to_i16le"):
        getattr(L, name).restype = None
        getattr(L, name).argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    L.sko_pcm_to_planar_f32.restype = C.c_size_t
    L.sko_pcm_to_planar_f32.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int,
                                        C.c_void_p]
    L.sko_mixdown_to_mono_f32.restype = None
    L.sko_mixdown_to_mono_f32.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
    L.sko_downmix_channels.restype = C.c_size_t
    L.sko_downmix_channels.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
    L.sko_float_sample_to_i16.restype = C.c_int16
    L.sko_float_sample_to_i16.argtypes = [C.c_float]
    L.sko_f32_channels_to_bytes.restype = C.c_size_t
    L.sko_f32_channels_to_bytes.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
    L.sko_exact_signed_pcm_to_i16.restype = C.c_size_t
    L.sko_exact_signed_pcm_to_i16.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
    L.sko_resampler_new.restype = C.c_void_p
    L.sko_resampler_new.argtypes = [C.c_uint32, C.c_uint32, C.c_size_t, C.c_size_t]
    L.sko_resampler_free.restype = None
    L.sko_resampler_free.argtypes = [C.c_void_p]
    L.sko_resampler_out_max.restype = C.c_size_t
    L.sko_resampler_out_max.argtypes = [C.c_void_p]
    L.sko_resampler_table.restype = C.c_void_p
    L.sko_resampler_table.argtypes = [C.c_void_p]
    L.sko_resampler_last_index.restype = C.c_double
    L.sko_resampler_last_index.argtypes = [C.c_void_p]
    L.sko_resampler_process.restype = C.c_size_t
    L.sko_resampler_process.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.sko_resampler_process_partial.restype = C.c_size_t
    L.sko_resampler_process_partial.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
    L.sko_resampler_walk.restype = C.c_size_t
    L.sko_resampler_walk.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    L.sko_resampler_uses_avx.restype = C.c_int
    L.sko_downsample_audio.restype = C.c_size_t
    L.sko_downsample_audio.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_size_t, C.c_uint32, C.c_uint32,
                                       C.c_void_p, C.c_size_t, C.c_int]
    L.sko_stream_new.restype = C.c_void_p
    L.sko_stream_new.argtypes = [C.c_uint32, C.c_uint32, C.c_size_t]
    L.sko_stream_free.restype = None
    L.sko_stream_free.argtypes = [C.c_void_p]
    L.sko_stream_process.restype = C.c_int
    L.sko_stream_process.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                     C.POINTER(C.c_size_t)]
    L.sko_stream_flush.restype = C.c_int
    L.sko_stream_flush.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.sko_norm48k_new.restype = C.c_void_p
    L.sko_norm48k_free.restype = None
    L.sko_norm48k_free.argtypes = [C.c_void_p]
    L.sko_norm48k_source_frames.restype = C.c_uint64
    L.sko_norm48k_source_frames.argtypes = [C.c_void_p]
    L.sko_norm48k_output_frames.restype = C.c_uint64
    L.sko_norm48k_output_frames.argtypes = [C.c_void_p]
    L.sko_norm48k_push.restype = C.c_int
    L.sko_norm48k_push.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p,
                                   C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.sko_norm48k_finish.restype = C.c_int
    L.sko_norm48k_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.sko_normalize_chain_mono.restype = C.c_size_t
    L.sko_normalize_chain_mono.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_uint32,
                                           C.c_uint32, C.c_void_p, C.c_size_t]


# name -> (bytes per input sample, output dtype)
_FLAT = {
    "i16le_to_f32": (2, np.float32), "s16le_to_i16": (2, np.int16), "s16be_to_i16": (2, np.int16),
    "s16le_to_i32": (2, np.int32), "s24le_to_i32": (3, np.int32), "s24le_to_i16": (3, np.int16),
    "s24be_to_i16": (3, np.int16), "s32le_to_i32": (4, np.int32), "s32be_to_i32": (4, np.int32),
    "s32le_to_s24": (4, np.int32), "s32be_to_s24": (4, np.int32), "s32le_to_f32": (4, np.float32),
    "s32be_to_f32": (4, np.float32), "s32le_to_i16": (4, np.int16), "s32be_to_i16": (4, np.int16),
    "f32le_to_i16": (4, np.int16), "f32be_to_i16": (4, np.int16), "f32le_to_i32": (4, np.int32),
    "f32le_to_s24": (4, np.int32),
    "pkt_s16le_to_f32": (2, np.float32), "pkt_s24le_to_f32": (3, np.float32), "pkt_s32le_to_i16": (4, np.int16),
    "pkt_f32le_to_i16": (4, np.int16), "u8_to_s16": (1, np.int16),
}
FLAT_NAMES = tuple(_FLAT)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _bytes(b):
    return np.ascontiguousarray(np.frombuffer(bytes(b), dtype=np.uint8)) if not isinstance(b, np.ndarray) \
        else np.ascontiguousarray(b.view(np.uint8).reshape(-1))


# ---- G.711 ------------------------------------------------------------
def g711_decode(law, data, kind="i16"):
    d = _bytes(data)
    out = np.empty(d.size, {"i16": np.int16, "i32": np.int32, "f32": np.float32}[kind])
    getattr(lib(), "sko_g711_decode_" + kind)(law, _ptr(d), d.size, _ptr(out))
    return out


def g711_encode(law, samples):
    s = np.ascontiguousarray(samples)
    out = np.empty(s.size, np.uint8)
    fn = lib().sko_g711_encode_i16 if s.dtype == np.int16 else lib().sko_g711_encode_i32
    assert s.dtype in (np.int16, np.int32)
    fn(law, _ptr(s), s.size, _ptr(out))
    return out


# ---- flat conversions --------------------------------------------------
def flat(name, data):
    bps, dt = _FLAT[name]
    d = _bytes(data)
    out = np.empty(d.size // bps, dt)
    n = getattr(lib(), "sko_" + name)(_ptr(d), d.size, _ptr(out))
    return out[:n]


def interleave_vecs_i16(planar):
    p = np.ascontiguousarray(planar, np.int16)
    ch, fr = p.shape
    out = np.empty(ch * fr * 2, np.uint8)
    lib().sko_interleave_vecs_i16(_ptr(p), ch, fr, _ptr(out))
    return out


def deinterleave(kind, data, channels):
    bps, dt = {"i16": (2, np.int16), "s24": (3, np.int32), "f32": (4, np.float32)}[kind]
    d = _bytes(data)
    frames = d.size // (bps * channels)
    out = np.empty((channels, frames), dt)
    getattr(lib(), "sko_deinterleave_vecs_" + kind)(_ptr(d), d.size, channels, _ptr(out))
    return out


def stereo_to_mono(kind, samples):
    s = np.ascontiguousarray(samples, np.int16)
    out = np.empty(s.size // 2, np.int16)
    getattr(lib(), "sko_stereo_to_mono_" + kind)(_ptr(s), s.size, _ptr(out))
    return out


def vec(name, x):
    dt_in, dt_out = {"f32_to_i16": (np.float32, np.int16), "i16_to_f32": (np.int16, np.float32),
                     "i32_to_f32": (np.int32, np.float32)}[name]
    a = np.ascontiguousarray(x, dt_in)
    out = np.empty(a.size, dt_out)
    getattr(lib(), "sko_vec_" + name)(_ptr(a), a.size, _ptr(out))
    return out


def pcm_to_planar_f32(data, bits, fmt, big_endian, channels, mode):
    d = _bytes(data)
    bps = (bits + 7) // 8
    frames = d.size // (bps * max(channels, 1)) if bps else 0
    out = np.empty((max(channels, 1), frames), np.float32)
    n = lib().sko_pcm_to_planar_f32(_ptr(d), d.size, bits, fmt, int(big_endian), channels, mode, _ptr(out))
    if n == ERR:
        return None
    assert n == frames
    return out


def mixdown_to_mono_f32(planar):
    p = np.ascontiguousarray(planar, np.float32)
    out = np.empty(p.shape[1], np.float32)
    lib().sko_mixdown_to_mono_f32(_ptr(p), p.shape[0], p.shape[1], _ptr(out))
    return out


def downmix_channels(planar, target):
    p = np.ascontiguousarray(planar, np.float32)
    out = np.empty((max(target, 1), p.shape[1]), np.float32)
    n = lib().sko_downmix_channels(_ptr(p), p.shape[0], p.shape[1], target, _ptr(out))
    return out[:n]


def float_sample_to_i16(x):
    return int(lib().sko_float_sample_to_i16(C.c_float(x)))


def f32_channels_to_bytes(planar, bits, fmt):
    p = np.ascontiguousarray(planar, np.float32)
    out = np.empty(p.size * 4, np.uint8)
    n = lib().sko_f32_channels_to_bytes(_ptr(p), p.shape[0], p.shape[1], bits, fmt, _ptr(out))
    return None if n == ERR else out[:n]


def exact_signed_pcm_to_i16(data, bits, big_endian):
    d = _bytes(data)
    out = np.empty(d.size, np.uint8)
    n = lib().sko_exact_signed_pcm_to_i16(_ptr(d), d.size, bits, int(big_endian), _ptr(out))
    return None if n == ERR else out[:n]


# ---- resampler ----------------------------------------------------------
class Resampler:
    """rubato SincFixedIn<f32> restatement (fixed chunk)."""

    def __init__(self, in_rate, out_rate, chunk, channels):
        self.h = lib().sko_resampler_new(in_rate, out_rate, chunk, channels)
        self.chunk, self.channels = chunk, channels
        self.out_max = lib().sko_resampler_out_max(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().sko_resampler_free(self.h)
            self.h = None

    def table(self):
        p = lib().sko_resampler_table(self.h)
        return np.ctypeslib.as_array(C.cast(p, _f32p), shape=(256, 256)).copy()

    @property
    def last_index(self):
        return lib().sko_resampler_last_index(self.h)

    def walk(self):
        cap = self.out_max
        i0 = np.empty(cap, np.int32); s0 = np.empty(cap, np.uint8); fr = np.empty(cap, np.float32)
        n = lib().sko_resampler_walk(self.h, _ptr(i0), _ptr(s0), _ptr(fr), cap)
        return i0[:n], s0[:n], fr[:n]

    def _run(self, planar, frames, partial, threads):
        out = np.zeros((self.channels, max(self.out_max, 1)), np.float32)
        op = (C.c_void_p * self.channels)(*[out[c].ctypes.data for c in range(self.channels)])
        if planar is None:
            n = lib().sko_resampler_process_partial(self.h, None, 0, op, threads)
        else:
            p = np.ascontiguousarray(planar, np.float32)
            ip = (C.c_void_p * self.channels)(*[p[c].ctypes.data for c in range(self.channels)])
            if partial:
                n = lib().sko_resampler_process_partial(self.h, ip, frames, op, threads)
            else:
                assert p.shape[1] >= self.chunk
                n = lib().sko_resampler_process(self.h, ip, op, threads)
        return out[:, :n].copy()

    def process(self, planar, threads=1):
        return self._run(planar, self.chunk, False, threads)

    def process_partial(self, planar_or_none, threads=1):
        frames = 0 if planar_or_none is None else np.asarray(planar_or_none).shape[1]
        return self._run(planar_or_none, frames, True, threads)


def downsample_audio(data, bits, fmt, channels, in_rate, out_rate, threads=1):
    d = _bytes(data)
    bps = bits // 8 if bits in (16, 24, 32) else 1
    frames = d.size // (bps * max(channels, 1))
    cap = int(frames * out_rate / max(in_rate, 1) * 2.0 + 16)
    out = np.empty((max(channels, 1), cap), np.float32)
    n = lib().sko_downsample_audio(_ptr(d), d.size, bits, fmt, channels, in_rate, out_rate, _ptr(out), cap, threads)
    return None if n == ERR else out[:, :n].copy()


class Stream:
    """StreamingResampler restatement (chunk 4096 + flush/trim)."""

    def __init__(self, in_rate, out_rate, channels):
        self.h = lib().sko_stream_new(in_rate, out_rate, channels)
        self.channels, self.in_rate, self.out_rate = channels, in_rate, out_rate

    def __del__(self):
        if getattr(self, "h", None):
            lib().sko_stream_free(self.h)
            self.h = None

    def process(self, planar):
        p = np.ascontiguousarray(planar, np.float32)
        cap = int((p.shape[1] + 8192) * self.out_rate / self.in_rate * 2 + 64)
        out = np.empty((self.channels, cap), np.float32)
        n = C.c_size_t(0)
        rc = lib().sko_stream_process(self.h, _ptr(p), p.shape[1], _ptr(out), cap, C.byref(n))
        assert rc == 0
        return out[:, :n.value].copy()

    def flush(self):
        cap = int(8192 * self.out_rate / self.in_rate * 2 + 64)
        out = np.empty((self.channels, cap), np.float32)
        n = C.c_size_t(0)
        rc = lib().sko_stream_flush(self.h, _ptr(out), cap, C.byref(n))
        assert rc == 0
        return out[:, :n.value].copy()


def normalize_chain_mono(data, bits, fmt, big_endian, channels, in_rate, out_rate):
    d = _bytes(data)
    bps = (bits + 7) // 8
    frames = d.size // (bps * max(channels, 1))
    cap = int((frames + 8192) * out_rate / in_rate + 64)
    out = np.empty(cap, np.float32)
    n = lib().sko_normalize_chain_mono(_ptr(d), d.size, bits, fmt, int(big_endian), channels, in_rate, out_rate,
                                       _ptr(out), cap)
    return None if n == ERR else out[:n].copy()


def unpack_dvd_lpcm(packed, bits, channels):
    """unpack_dvd_lpcm (soundkit-decoder/src/lib.rs:750-804); None for the Err arms"""
    d = _bytes(packed)
    out = np.empty(d.size, np.uint8)
    lib().sko_unpack_dvd_lpcm.restype = C.c_size_t
    lib().sko_unpack_dvd_lpcm.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
    n = lib().sko_unpack_dvd_lpcm(_ptr(d), d.size, bits, channels, _ptr(out))
    return None if n == ERR else out[:n]


G726_BITS = {16000: 2, 24000: 3, 32000: 4, 40000: 5}
G726_GROUP = {2: (4, 1), 3: (8, 3), 4: (2, 1), 5: (8, 5)}  # bits -> (samples, bytes) per group


class G726:
    """G726Core + whole-group packing (soundkit-g726/src/lib.rs:215-596); state persists across calls"""

    def __init__(self, bits, packing=0):
        self.bits, self.packing = bits, packing
        self.state = np.zeros(24, np.int32)
        L = lib()
        L.sko_g726_state_init.restype = None
        L.sko_g726_state_init.argtypes = [C.c_void_p]
        for f in (L.sko_g726_decode, L.sko_g726_encode):
            f.restype = C.c_size_t
            f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.sko_g726_state_init(_ptr(self.state))

    def decode(self, data):
        d = _bytes(data)
        gs, gb = G726_GROUP[self.bits]
        out = np.empty(d.size // gb * gs, np.int16)
        n = lib().sko_g726_decode(self.bits, self.packing, _ptr(self.state), _ptr(d), d.size, _ptr(out))
        return None if n == ERR else out[:n]

    def encode(self, samples):
        s = np.ascontiguousarray(samples, np.int16)
        gs, gb = G726_GROUP[self.bits]
        out = np.empty(s.size // gs * gb, np.uint8)
        n = lib().sko_g726_encode(self.bits, self.packing, _ptr(self.state), _ptr(s), s.size, _ptr(out))
        return None if n == ERR else out[:n]


class G722:
    """G.722 64 kbit/s restatement (g722.c): one object = one stream's codec state"""

    def __init__(self):
        L = lib()
        L.sko_g722_decode.restype = C.c_size_t
        L.sko_g722_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.sko_g722_encode.restype = C.c_size_t
        L.sko_g722_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        self.state = np.zeros(114, np.int32)
        L.sko_g722_state_init(_ptr(self.state))

    def decode(self, data):
        d = _bytes(data)
        out = np.empty(d.size * 2, np.int16)
        n = lib().sko_g722_decode(_ptr(self.state), _ptr(d), d.size, _ptr(out))
        return out[:n]

    def encode(self, samples):
        s = np.ascontiguousarray(samples, np.int16)
        out = np.empty(s.size // 2 + 1, np.uint8)
        n = lib().sko_g722_encode(_ptr(self.state), _ptr(s), s.size, _ptr(out))
        return out[:n]


class Normalizer48k:
    """StreamingStereo48kNormalizer restatement (soundkit/src/audio_pipeline.rs:256-394).  push/finish return
    (left, right) float32 arrays (empty = Ok(None)) or raise ValueError(code) for the reference's Err arms."""

    def __init__(self):
        self.h = lib().sko_norm48k_new()

    def __del__(self):
        if getattr(self, "h", None):
            lib().sko_norm48k_free(self.h)
            self.h = None

    def source_frames(self):
        return lib().sko_norm48k_source_frames(self.h)

    def output_frames(self):
        return lib().sko_norm48k_output_frames(self.h)

    def push(self, data, bits, fmt, big_endian, channels, rate):
        d = _bytes(data)
        frames = d.size // max(1, (bits // 8) * max(channels, 1))
        cap = frames * 48000 // max(rate, 1) + 8 if rate != 48000 else frames + 1
        l, r = np.empty(cap, np.float32), np.empty(cap, np.float32)
        n = C.c_size_t(0)
        rc = lib().sko_norm48k_push(self.h, _ptr(d), d.size, bits, fmt, int(big_endian), channels, rate, _ptr(l), _ptr(r), cap, C.byref(n))
        if rc:
            raise ValueError(rc)
        return l[:n.value].copy(), r[:n.value].copy()

    def finish(self, cap=1 << 16):
        l, r = np.empty(cap, np.float32), np.empty(cap, np.float32)
        n = C.c_size_t(0)
        rc = lib().sko_norm48k_finish(self.h, _ptr(l), _ptr(r), cap, C.byref(n))
        if rc:
            raise ValueError(rc)
        return l[:n.value].copy(), r[:n.value].copy()


def normalize48k(blocks):
    """normalize_blocks of the reference tests (audio_pipeline.rs:781-795): blocks = [(data, bits, fmt, be, channels, rate)]"""
    nz = Normalizer48k()
    ls, rs = [], []
    for b in blocks:
        l, r = nz.push(*b)
        ls.append(l); rs.append(r)
    l, r = nz.finish()
    ls.append(l); rs.append(r)
    return np.concatenate(ls), np.concatenate(rs)
