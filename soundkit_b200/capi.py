"""ctypes binding of libsoundkit_cuda.so (include/soundkit_cuda.h) — the C ABI, nothing else.

This is the same surface the Rust `soundkit-cuda-sys` crate binds (rust/soundkit-cuda-sys/src/lib.rs);
tests and bench.py call the kernels through it.  There is no CPU path: if the library is missing the
import fails loudly, and without a CUDA device `Context()` raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SOUNDKIT_CUDA_LIB") or os.path.join(_HERE, "libsoundkit_cuda.so")  # env override: A/B builds

# enums of include/soundkit_cuda.h
OK, E_ARG, E_ALLOC, E_CUDA, E_STATE, E_NODEVICE = 0, -1, -2, -3, -4, -5
LAW_MULAW, LAW_ALAW = 0, 1
KIND_I16, KIND_I32, KIND_F32 = 0, 1, 2
FMT_SIGNED, FMT_FLOAT = 0, 1
MODE_A, MODE_B = 0, 1
MIX_NONE, MIX_SUM_THEN_SCALE, MIX_SCALE_THEN_SUM = 0, 1, 2
SEM_ONESHOT, SEM_STREAMING = 0, 1
RESAMPLE_BIT_EXACT, RESAMPLE_FAST = 0, 1
FLAT_OPS = ("i16le_to_f32", "s16le_to_i16", "s16be_to_i16", "s16le_to_i32", "s24le_to_i32", "s24le_to_i16",
            "s24be_to_i16", "s32le_to_i32", "s32be_to_i32", "s32le_to_s24", "s32be_to_s24", "s32le_to_f32",
            "s32be_to_f32", "s32le_to_i16", "s32be_to_i16", "f32le_to_i16", "f32be_to_i16", "f32le_to_i32",
            "f32le_to_s24", "pkt_s16le_to_f32", "pkt_s24le_to_f32", "pkt_s32le_to_i16", "pkt_f32le_to_i16", "u8_to_s16")
FLAT_OP = {name: i for i, name in enumerate(FLAT_OPS)}
G726_GROUP = {2: (4, 1), 3: (8, 3), 4: (2, 1), 5: (8, 5)}  # code bits -> (samples, bytes) per group
G726_LEFT, G726_RIGHT = 0, 1
G722_STATE_INTS = 68

EXPORTS = (
    "sk_abi_version", "sk_ctx_create", "sk_ctx_destroy", "sk_last_error", "sk_ctx_set_stream", "sk_ctx_reset_stream",
    "sk_ctx_synchronize",
    "sk_ctx_launch_count", "sk_ctx_set_batch_slab_bytes", "sk_ctx_set_resample_mode", "sk_ctx_resample_fallback_outputs", "sk_ctx_measure_ffma_peak", "sk_ctx_measure_ffma_peaks", "sk_ctx_poison_shared_memory",
    "sk_g711_decode", "sk_g711_decode_dev", "sk_g711_encode", "sk_g711_encode_dev",
    "sk_pcm_flat_convert", "sk_pcm_flat_convert_dev", "sk_flat_op_widths", "sk_dvd_lpcm_unpack", "sk_dvd_lpcm_unpack_dev",
    "sk_deinterleave", "sk_interleave_i16", "sk_interleave_planar", "sk_interleave_planar_dev",
    "sk_stereo_to_mono_i16", "sk_pcm_to_planar_f32", "sk_pcm_to_planar_f32_dev", "sk_planar_f32_to_pcm",
    "sk_planar_f32_to_pcm_dev", "sk_pcm_narrow_i16", "sk_mixdown_mono", "sk_mixdown_mono_dev", "sk_downmix_stereo",
    "sk_resampler_create", "sk_resampler_destroy", "sk_resampler_out_max", "sk_resampler_last_index",
    "sk_resampler_process", "sk_resampler_process_partial", "sk_resampler_copy_table",
    "sk_resampler_batch_create", "sk_resampler_batch_destroy", "sk_resampler_batch_out_cap", "sk_resampler_batch_pending",
    "sk_resampler_batch_last_index", "sk_resampler_batch_process_dev", "sk_resampler_batch_input_dev", "sk_resampler_batch_flush_dev",
    "sk_normalize_batch_plan", "sk_normalize_batch_dev", "sk_normalize_batch_pcm_dev", "sk_normalize_batch", "sk_normalize_stream",
    "sk_norm48k_create", "sk_norm48k_destroy", "sk_norm48k_push", "sk_norm48k_finish", "sk_norm48k_source_frames",
    "sk_norm48k_output_frames", "sk_norm48k_source_sample_rate", "sk_norm48k_source_channels", "sk_norm48k_max_output",
    "sk_normalize48k_out_frames", "sk_normalize48k_batch_dev",
    "sk_g726_state_init", "sk_g726_decode", "sk_g726_encode", "sk_g726_decode_batch_dev", "sk_g726_encode_batch_dev",
    "sk_g722_state_init", "sk_g722_decode", "sk_g722_encode", "sk_g722_decode_batch_dev", "sk_g722_encode_batch_dev",
)


class BatchDesc(C.Structure):
    _fields_ = [("in_", C.c_void_p), ("in_stride", C.c_size_t), ("n_streams", C.c_size_t), ("frames", C.c_size_t),
                ("bits", C.c_int), ("fmt", C.c_int), ("big_endian", C.c_int), ("channels", C.c_int),
                ("scale_mode", C.c_int), ("in_rate", C.c_uint32), ("out_rate", C.c_uint32), ("semantics", C.c_int),
                ("mix_mode", C.c_int), ("out", C.c_void_p), ("out_stride", C.c_size_t), ("out_stride_ch", C.c_size_t)]


class ResampleParams(C.Structure):
    _fields_ = [("sinc_len", C.c_uint32), ("oversampling", C.c_uint32), ("f_cutoff", C.c_float), ("interp", C.c_int),
                ("window", C.c_int), ("max_ratio_rel", C.c_double)]


class SoundkitError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsoundkit_cuda error {code}: {msg}")
        self.code, self.msg = code, msg


_lib = None


def load():
    """Load the shared library (no CUDA call is made here)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, sz, i, u32 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint32
    psz = C.POINTER(C.c_size_t)
    proto = {
        "sk_abi_version": (i, []),
        "sk_ctx_create": (i, [i, C.POINTER(vp)]),
        "sk_ctx_destroy": (None, [vp]),
        "sk_last_error": (C.c_char_p, [vp]),
        "sk_ctx_set_stream": (i, [vp, vp]),
        "sk_ctx_reset_stream": (i, [vp]),
        "sk_ctx_synchronize": (i, [vp]),
        "sk_ctx_launch_count": (C.c_uint64, [vp]),
        "sk_ctx_set_batch_slab_bytes": (i, [vp, sz]),
        "sk_ctx_set_resample_mode": (i, [vp, i]),
        "sk_ctx_resample_fallback_outputs": (i, [vp, C.POINTER(C.c_uint64)]),
        "sk_ctx_measure_ffma_peak": (i, [vp, C.POINTER(C.c_double)]),
        "sk_ctx_measure_ffma_peaks": (i, [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "sk_ctx_poison_shared_memory": (i, [vp, C.c_uint32]),
        "sk_g711_decode": (i, [vp, i, vp, sz, i, vp]),
        "sk_g711_decode_dev": (i, [vp, i, vp, sz, i, vp]),
        "sk_g711_encode": (i, [vp, i, i, vp, sz, vp]),
        "sk_g711_encode_dev": (i, [vp, i, i, vp, sz, vp]),
        "sk_pcm_flat_convert": (i, [vp, i, vp, sz, vp, psz]),
        "sk_pcm_flat_convert_dev": (i, [vp, i, vp, sz, vp]),
        "sk_flat_op_widths": (i, [i, C.POINTER(i), C.POINTER(i)]),
        "sk_dvd_lpcm_unpack": (i, [vp, vp, sz, i, i, vp]),
        "sk_dvd_lpcm_unpack_dev": (i, [vp, vp, sz, i, i, vp]),
        "sk_deinterleave": (i, [vp, i, vp, sz, i, vp, psz]),
        "sk_interleave_i16": (i, [vp, vp, i, sz, vp]),
        "sk_interleave_planar": (i, [vp, i, vp, i, sz, vp]),
        "sk_interleave_planar_dev": (i, [vp, i, vp, i, sz, vp]),
        "sk_stereo_to_mono_i16": (i, [vp, i, vp, sz, vp]),
        "sk_pcm_to_planar_f32": (i, [vp, vp, sz, i, i, i, i, i, vp, sz, psz]),
        "sk_pcm_to_planar_f32_dev": (i, [vp, vp, sz, sz, sz, i, i, i, i, i, vp, sz]),
        "sk_planar_f32_to_pcm": (i, [vp, vp, sz, i, i, i, vp]),
        "sk_planar_f32_to_pcm_dev": (i, [vp, vp, sz, i, i, i, vp]),
        "sk_pcm_narrow_i16": (i, [vp, vp, sz, i, i, vp]),
        "sk_mixdown_mono": (i, [vp, vp, sz, i, i, vp]),
        "sk_mixdown_mono_dev": (i, [vp, vp, sz, i, i, vp]),
        "sk_downmix_stereo": (i, [vp, vp, sz, i, vp]),
        "sk_resampler_create": (i, [vp, u32, u32, i, sz, C.POINTER(ResampleParams), C.POINTER(vp)]),
        "sk_resampler_destroy": (None, [vp]),
        "sk_resampler_out_max": (sz, [vp]),
        "sk_resampler_last_index": (C.c_double, [vp]),
        "sk_resampler_process": (i, [vp, vp, vp, psz]),
        "sk_resampler_process_partial": (i, [vp, vp, sz, vp, psz]),
        "sk_resampler_copy_table": (i, [vp, vp]),
        "sk_resampler_batch_create": (i, [vp, u32, u32, i, sz, sz, C.POINTER(vp)]),
        "sk_resampler_batch_destroy": (None, [vp]),
        "sk_resampler_batch_out_cap": (sz, [vp]),
        "sk_resampler_batch_pending": (sz, [vp]),
        "sk_resampler_batch_last_index": (C.c_double, [vp]),
        "sk_resampler_batch_process_dev": (i, [vp, vp, sz, sz, sz, vp, sz, sz, psz]),
        "sk_resampler_batch_flush_dev": (i, [vp, vp, sz, sz, psz]),
        "sk_resampler_batch_input_dev": (i, [vp, C.POINTER(vp), psz]),
        "sk_normalize_batch_plan": (i, [vp, u32, u32, sz, i, psz]),
        "sk_normalize_batch_dev": (i, [vp, C.POINTER(BatchDesc)]),
        "sk_normalize_batch": (i, [vp, C.POINTER(BatchDesc)]),
        "sk_normalize_batch_pcm_dev": (i, [vp, C.POINTER(BatchDesc), i, i, vp, sz]),
        "sk_normalize_stream": (i, [vp, vp, sz, i, i, i, i, i, u32, u32, i, i, vp, sz, psz]),
        "sk_norm48k_create": (i, [vp, C.POINTER(vp)]),
        "sk_norm48k_destroy": (None, [vp]),
        "sk_norm48k_push": (i, [vp, vp, sz, i, i, i, i, u32, vp, vp, sz, psz]),
        "sk_norm48k_finish": (i, [vp, vp, vp, sz, psz]),
        "sk_norm48k_source_frames": (C.c_uint64, [vp]),
        "sk_norm48k_output_frames": (C.c_uint64, [vp]),
        "sk_norm48k_source_sample_rate": (u32, [vp]),
        "sk_norm48k_source_channels": (i, [vp]),
        "sk_norm48k_max_output": (sz, [sz, u32]),
        "sk_normalize48k_out_frames": (sz, [sz, u32]),
        "sk_normalize48k_batch_dev": (i, [vp, vp, sz, sz, sz, i, i, i, i, u32, vp, sz, sz, psz]),
        "sk_g726_state_init": (None, [vp]),
        "sk_g726_decode": (i, [vp, i, i, vp, vp, sz, i, vp]),
        "sk_g726_encode": (i, [vp, i, i, vp, i, vp, sz, vp]),
        "sk_g726_decode_batch_dev": (i, [vp, i, i, vp, sz, sz, sz, i, vp, sz, vp]),
        "sk_g726_encode_batch_dev": (i, [vp, i, i, i, vp, sz, sz, sz, vp, sz, vp]),
        "sk_g722_state_init": (None, [vp]),
        "sk_g722_decode": (i, [vp, vp, vp, sz, i, vp]),
        "sk_g722_encode": (i, [vp, vp, i, vp, sz, vp]),
        "sk_g722_decode_batch_dev": (i, [vp, vp, sz, sz, sz, i, vp, sz, vp]),
        "sk_g722_encode_batch_dev": (i, [vp, i, vp, sz, sz, sz, vp, sz, vp]),
    }
    for name in EXPORTS:
        fn = getattr(L, name)  # AttributeError if the header and the library disagree
        fn.restype, fn.argtypes = proto[name]
    _lib = L
    return L


def _p(a):
    """void* of a numpy array / torch tensor / int address / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if isinstance(a, int):
        return C.c_void_p(a)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


def _u8(b):
    if isinstance(b, np.ndarray):
        return np.ascontiguousarray(b).view(np.uint8).reshape(-1)
    return np.frombuffer(bytes(b), dtype=np.uint8)


class Context:
    """One sk_ctx: one per host thread and device."""

    def __init__(self, device=0):
        self.L = load()
        h = C.c_void_p()
        rc = self.L.sk_ctx_create(device, C.byref(h))
        if rc != OK:
            raise SoundkitError(rc, "sk_ctx_create failed (no sm_100 CUDA device? the library has no CPU path)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.sk_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != OK:
            raise SoundkitError(rc, self.L.sk_last_error(self.h).decode())

    # -- plumbing
    def set_stream(self, stream_ptr):
        """launch on this cudaStream_t; 0 / None is CUDA's legacy default stream"""
        self._ck(self.L.sk_ctx_set_stream(self.h, C.c_void_p(stream_ptr) if stream_ptr else None))

    def reset_stream(self):
        self._ck(self.L.sk_ctx_reset_stream(self.h))

    def use_torch_stream(self):
        import torch
        self.set_stream(torch.cuda.current_stream().cuda_stream)

    def synchronize(self):
        self._ck(self.L.sk_ctx_synchronize(self.h))

    @property
    def launches(self):
        return int(self.L.sk_ctx_launch_count(self.h))

    def set_batch_slab_bytes(self, nbytes):
        """input bytes per slab of the host-pointer batch pipeline (normalize_batch_host)"""
        self._ck(self.L.sk_ctx_set_batch_slab_bytes(self.h, nbytes))

    def set_resample_mode(self, fast):
        """general-ratio resampler arithmetic: False = bit-exact (default), True = the opt-in tolerance mode (1e-5 / 120 dB)"""
        self._ck(self.L.sk_ctx_set_resample_mode(self.h, RESAMPLE_FAST if fast else RESAMPLE_BIT_EXACT))

    def resample_fallback_outputs(self):
        """outputs of the general-ratio resampler that took the per-output exact path so far (synchronises)"""
        v = C.c_uint64()
        self._ck(self.L.sk_ctx_resample_fallback_outputs(self.h, C.byref(v)))
        return v.value

    def ffma_peak_tflops(self):
        v = C.c_double()
        self._ck(self.L.sk_ctx_measure_ffma_peak(self.h, C.byref(v)))
        return v.value

    def poison_shared_memory(self, pattern=0x7FC00000):
        """diagnostic: every SM's shared memory is left full of `pattern` (default: quiet NaN) for the next kernels to find"""
        self._ck(self.L.sk_ctx_poison_shared_memory(self.h, pattern))

    def ffma_peaks_tflops(self):
        a, b = C.c_double(), C.c_double()
        self._ck(self.L.sk_ctx_measure_ffma_peaks(self.h, C.byref(a), C.byref(b)))
        return {"ffma": a.value, "ffma2": b.value}

    # -- G.711 (host buffers)
    def g711_decode(self, law, data, kind=KIND_I16):
        d = _u8(data)
        out = np.empty(d.size, {KIND_I16: np.int16, KIND_I32: np.int32, KIND_F32: np.float32}[kind])
        self._ck(self.L.sk_g711_decode(self.h, law, _p(d), d.size, kind, _p(out)))
        return out

    def g711_encode(self, law, samples):
        s = np.ascontiguousarray(samples)
        kind = {np.dtype(np.int16): KIND_I16, np.dtype(np.int32): KIND_I32}[s.dtype]
        out = np.empty(s.size, np.uint8)
        self._ck(self.L.sk_g711_encode(self.h, law, kind, _p(s), s.size, _p(out)))
        return out

    def g711_decode_dev(self, law, d_in, n, kind, d_out):
        self._ck(self.L.sk_g711_decode_dev(self.h, law, _p(d_in), n, kind, _p(d_out)))

    def g711_encode_dev(self, law, kind, d_in, n, d_out):
        self._ck(self.L.sk_g711_encode_dev(self.h, law, kind, _p(d_in), n, _p(d_out)))

    # -- flat conversions
    def flat_widths(self, op):
        a, b = C.c_int(), C.c_int()
        self._ck(self.L.sk_flat_op_widths(op, C.byref(a), C.byref(b)))
        return a.value, b.value

    def flat_convert(self, name, data):
        op = FLAT_OP[name]
        iw, ow = self.flat_widths(op)
        d = _u8(data)
        dt = np.float32 if name.endswith("f32") else (np.int16 if ow == 2 else np.int32)
        out = np.empty(d.size // iw, dt)
        n = C.c_size_t()
        self._ck(self.L.sk_pcm_flat_convert(self.h, op, _p(d), d.size, _p(out), C.byref(n)))
        return out[:n.value]

    def dvd_lpcm_unpack(self, packed, bits, channels):
        d = _u8(packed)
        out = np.empty(d.size, np.uint8)
        self._ck(self.L.sk_dvd_lpcm_unpack(self.h, _p(d), d.size, bits, channels, _p(out)))
        return out

    def flat_convert_dev(self, name, d_in, nbytes, d_out):
        self._ck(self.L.sk_pcm_flat_convert_dev(self.h, FLAT_OP[name], _p(d_in), nbytes, _p(d_out)))

    def deinterleave(self, kind, data, channels):
        d = _u8(data)
        bps, dt = {"i16": (2, np.int16), "s24": (3, np.int32), "f32": (4, np.float32)}[kind]
        frames = d.size // (bps * channels)
        out = np.empty((channels, frames), dt)
        n = C.c_size_t()
        self._ck(self.L.sk_deinterleave(self.h, ("i16", "s24", "f32").index(kind), _p(d), d.size, channels, _p(out),
                                        C.byref(n)))
        assert n.value == frames
        return out

    def interleave_i16(self, planar):
        p = np.ascontiguousarray(planar, np.int16)
        out = np.empty(p.size * 2, np.uint8)
        self._ck(self.L.sk_interleave_i16(self.h, _p(p), p.shape[0], p.shape[1], _p(out)))
        return out

    def interleave_planar(self, planar):
        """WavStreamEncoder::push_planar_{i16,i32,f32}: [channels][frames] of i16 / i32 / f32 -> interleaved LE bytes (bit copy)"""
        p = np.ascontiguousarray(planar)
        assert p.ndim == 2 and p.dtype.itemsize in (2, 4)
        out = np.empty(p.size * p.dtype.itemsize, np.uint8)
        self._ck(self.L.sk_interleave_planar(self.h, p.dtype.itemsize, _p(p), p.shape[0], p.shape[1], _p(out)))
        return out

    def stereo_to_mono_i16(self, samples, avg):
        s = np.ascontiguousarray(samples, np.int16)
        out = np.empty(s.size // 2, np.int16)
        self._ck(self.L.sk_stereo_to_mono_i16(self.h, int(avg), _p(s), s.size, _p(out)))
        return out

    def pcm_to_planar_f32(self, data, bits, fmt, big_endian, channels, mode):
        d = _u8(data)
        frames = d.size // ((bits // 8) * channels) if bits in (16, 24, 32) and channels > 0 else 0
        out = np.empty((max(channels, 1), frames), np.float32)
        n = C.c_size_t()
        self._ck(self.L.sk_pcm_to_planar_f32(self.h, _p(d), d.size, bits, fmt, int(big_endian), channels, mode,
                                             _p(out), frames, C.byref(n)))
        return out[:, :n.value]

    def pcm_to_planar_f32_dev(self, d_in, in_stride, n_streams, frames, bits, fmt, big_endian, channels, mode, d_out,
                              out_stride):
        self._ck(self.L.sk_pcm_to_planar_f32_dev(self.h, _p(d_in), in_stride, n_streams, frames, bits, fmt,
                                                 int(big_endian), channels, mode, _p(d_out), out_stride))

    def planar_f32_to_pcm(self, planar, bits, fmt):
        p = np.ascontiguousarray(planar, np.float32)
        out = np.empty(p.size * (bits // 8), np.uint8)
        self._ck(self.L.sk_planar_f32_to_pcm(self.h, _p(p), p.shape[1], p.shape[0], bits, fmt, _p(out)))
        return out

    def planar_f32_to_pcm_dev(self, d_planar, frames, channels, bits, fmt, d_out):
        self._ck(self.L.sk_planar_f32_to_pcm_dev(self.h, _p(d_planar), frames, channels, bits, fmt, _p(d_out)))

    def pcm_narrow_i16(self, data, bits, big_endian):
        d = _u8(data)
        out = np.empty(d.size // (bits // 8) * 2, np.uint8)
        self._ck(self.L.sk_pcm_narrow_i16(self.h, _p(d), d.size, bits, int(big_endian), _p(out)))
        return out

    def mixdown_mono(self, planar, mode):
        p = np.ascontiguousarray(planar, np.float32)
        out = np.empty(p.shape[1], np.float32)
        self._ck(self.L.sk_mixdown_mono(self.h, _p(p), p.shape[1], p.shape[0], mode, _p(out)))
        return out

    def downmix_stereo(self, planar):
        p = np.ascontiguousarray(planar, np.float32)
        out = np.empty((2, p.shape[1]), np.float32)
        self._ck(self.L.sk_downmix_stereo(self.h, _p(p), p.shape[1], p.shape[0], _p(out)))
        return out

    # -- batch chain
    def batch_plan(self, in_rate, out_rate, frames, semantics):
        n = C.c_size_t()
        self._ck(self.L.sk_normalize_batch_plan(self.h, in_rate, out_rate, frames, semantics, C.byref(n)))
        return n.value

    def normalize_batch_dev(self, desc):
        self._ck(self.L.sk_normalize_batch_dev(self.h, C.byref(desc)))

    def normalize_batch_pcm_dev(self, desc, out_bits, out_fmt, d_out, out_stride_bytes):
        self._ck(self.L.sk_normalize_batch_pcm_dev(self.h, C.byref(desc), out_bits, out_fmt, _p(d_out), out_stride_bytes))

    def normalize_batch_host(self, desc):
        self._ck(self.L.sk_normalize_batch(self.h, C.byref(desc)))

    def normalize_stream(self, data, bits, fmt, big_endian, channels, mode, in_rate, out_rate, semantics, mix):
        d = _u8(data)
        frames = d.size // ((bits // 8) * channels)
        n = self.batch_plan(in_rate, out_rate, frames, semantics)
        planes = channels if (mix == MIX_NONE or channels == 1) else 1
        out = np.empty((planes, max(n, 1)), np.float32)
        got = C.c_size_t()
        self._ck(self.L.sk_normalize_stream(self.h, _p(d), d.size, bits, fmt, int(big_endian), channels, mode, in_rate,
                                            out_rate, semantics, mix, _p(out), max(n, 1), C.byref(got)))
        return out[:, :got.value]

    def resampler(self, in_rate, out_rate, channels, chunk):
        return Resampler(self, in_rate, out_rate, channels, chunk)

    # -- G.726 ADPCM
    def resampler_batch(self, in_rate, out_rate, channels, n_streams, max_push=4096):
        """N lock-step StreamingResamplers with device-resident state (sk_resampler_batch_*)"""
        return ResamplerBatch(self, in_rate, out_rate, channels, n_streams, max_push)

    def g726_state(self):
        st = np.zeros(24, np.int32)
        self.L.sk_g726_state_init(_p(st))
        return st

    def g726_decode(self, bits, packing, state, data, kind=KIND_I16):
        d = _u8(data)
        gs, gb = G726_GROUP.get(bits, (1, 1))  # an unknown code size is the library's error to report
        out = np.empty(d.size // gb * gs, (np.int16, np.int32, np.float32)[kind])
        self._ck(self.L.sk_g726_decode(self.h, bits, packing, _p(state), _p(d), d.size, kind, _p(out)))
        return out

    def g726_encode(self, bits, packing, state, samples):
        s = np.ascontiguousarray(samples)
        kind = KIND_I16 if s.dtype == np.int16 else KIND_I32
        gs, gb = G726_GROUP.get(bits, (1, 1))
        out = np.empty(s.size // gs * gb, np.uint8)
        self._ck(self.L.sk_g726_encode(self.h, bits, packing, _p(state), kind, _p(s), s.size, _p(out)))
        return out

    def g726_decode_batch_dev(self, bits, packing, d_in, in_stride, n_streams, nbytes, kind, d_out, out_stride, d_state=None):
        self._ck(self.L.sk_g726_decode_batch_dev(self.h, bits, packing, _p(d_in), in_stride, n_streams, nbytes, kind, _p(d_out), out_stride,
                                                 _p(d_state)))

    def g726_encode_batch_dev(self, bits, packing, kind, d_in, in_stride, n_streams, n, d_out, out_stride, d_state=None):
        self._ck(self.L.sk_g726_encode_batch_dev(self.h, bits, packing, kind, _p(d_in), in_stride, n_streams, n, _p(d_out), out_stride,
                                                 _p(d_state)))

    # -- G.722 (codec state: numpy int32[G722_STATE_INTS], all zero at reset)
    def g722_state(self):
        st = np.zeros(G722_STATE_INTS, np.int32)
        self.L.sk_g722_state_init(_p(st))
        return st

    def g722_decode(self, state, data, kind=KIND_I16):
        d = _u8(data)
        out = np.empty(d.size * 2, {KIND_I16: np.int16, KIND_I32: np.int32, KIND_F32: np.float32}[kind])
        self._ck(self.L.sk_g722_decode(self.h, _p(state), _p(d), d.size, kind, _p(out)))
        return out

    def g722_encode(self, state, samples):
        s = np.ascontiguousarray(samples)
        kind = KIND_I16 if s.dtype == np.int16 else KIND_I32
        out = np.empty(s.size // 2, np.uint8)
        self._ck(self.L.sk_g722_encode(self.h, _p(state), kind, _p(s), s.size, _p(out)))
        return out

    def g722_decode_batch_dev(self, d_in, in_stride, n_streams, nbytes, kind, d_out, out_stride, d_state=None):
        self._ck(self.L.sk_g722_decode_batch_dev(self.h, C.c_void_p(d_in.data_ptr()), in_stride, n_streams, nbytes, kind, C.c_void_p(d_out.data_ptr()),
                                                 out_stride, C.c_void_p(d_state.data_ptr()) if d_state is not None else None))

    def g722_encode_batch_dev(self, kind, d_in, in_stride, n_streams, n, d_out, out_stride, d_state=None):
        self._ck(self.L.sk_g722_encode_batch_dev(self.h, kind, C.c_void_p(d_in.data_ptr()), in_stride, n_streams, n, C.c_void_p(d_out.data_ptr()),
                                                 out_stride, C.c_void_p(d_state.data_ptr()) if d_state is not None else None))

    # -- 48 kHz stereo normalizer (StreamingStereo48kNormalizer)
    def normalizer48k(self):
        return Normalizer48k(self)

    def normalize48k_out_frames(self, frames, rate):
        return self.L.sk_normalize48k_out_frames(frames, rate)

    def normalize48k_batch_dev(self, d_in, in_stride, n_streams, frames, bits, fmt, big_endian, channels, rate, d_out, out_stride,
                               out_stride_ch):
        n = C.c_size_t()
        self._ck(self.L.sk_normalize48k_batch_dev(self.h, _p(d_in), in_stride, n_streams, frames, bits, fmt, int(big_endian), channels,
                                                  rate, _p(d_out), out_stride, out_stride_ch, C.byref(n)))
        return n.value


class Normalizer48k:
    """sk_norm48k: soundkit's StreamingStereo48kNormalizer, arithmetic on the device."""

    def __init__(self, ctx):
        self.ctx = ctx
        h = C.c_void_p()
        ctx._ck(ctx.L.sk_norm48k_create(ctx.h, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.ctx.L.sk_norm48k_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def source_frames(self):
        return self.ctx.L.sk_norm48k_source_frames(self.h)

    def output_frames(self):
        return self.ctx.L.sk_norm48k_output_frames(self.h)

    def push(self, data, bits, fmt, big_endian, channels, rate):
        d = _u8(data)
        frames = d.size // max(1, (bits // 8) * max(channels, 1))
        cap = max(1, self.ctx.L.sk_norm48k_max_output(frames, rate))
        l, r = np.empty(cap, np.float32), np.empty(cap, np.float32)
        n = C.c_size_t()
        self.ctx._ck(self.ctx.L.sk_norm48k_push(self.h, _p(d), d.size, bits, fmt, int(big_endian), channels, rate, _p(l), _p(r), cap,
                                                C.byref(n)))
        return l[:n.value], r[:n.value]

    def finish(self, cap=1 << 16):
        l, r = np.empty(cap, np.float32), np.empty(cap, np.float32)
        n = C.c_size_t()
        self.ctx._ck(self.ctx.L.sk_norm48k_finish(self.h, _p(l), _p(r), cap, C.byref(n)))
        return l[:n.value], r[:n.value]


class ResamplerBatch:
    """sk_resampler_batch: StreamingResampler::process / flush for n_streams streams at once, torch CUDA tensors in and out
    (planar f32 [n_streams][channels][frames])."""

    def __init__(self, ctx, in_rate, out_rate, channels, n_streams, max_push):
        self.ctx, self.channels, self.n_streams = ctx, channels, n_streams
        h = C.c_void_p()
        ctx._ck(ctx.L.sk_resampler_batch_create(ctx.h, in_rate, out_rate, channels, n_streams, max_push, C.byref(h)))
        self.h = h
        self.out_max = int(ctx.L.sk_resampler_batch_out_cap(h))

    def close(self):
        if getattr(self, "h", None):
            self.ctx.L.sk_resampler_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def pending(self):
        return int(self.ctx.L.sk_resampler_batch_pending(self.h))

    @property
    def last_index(self):
        return float(self.ctx.L.sk_resampler_batch_last_index(self.h))

    def process_dev(self, x, out):
        """x: [n_streams, channels, frames] f32 CUDA tensor; out: [n_streams, channels, >= out_max]; returns frames produced"""
        assert x.dim() == 3 and x.stride(2) == 1 and out.stride(2) == 1 and out.shape[2] >= self.out_max
        n = C.c_size_t(0)
        self.ctx._ck(self.ctx.L.sk_resampler_batch_process_dev(self.h, C.c_void_p(x.data_ptr()), x.stride(0), x.stride(1), x.shape[2],
                                                              C.c_void_p(out.data_ptr()), out.stride(0), out.stride(1), C.byref(n)))
        return n.value

    def input_dev(self):
        """(device address, lane pitch in floats) of where the next push's frames go inside the handle's rows (zero-copy append)"""
        ptr, pitch = C.c_void_p(), C.c_size_t(0)
        self.ctx._ck(self.ctx.L.sk_resampler_batch_input_dev(self.h, C.byref(ptr), C.byref(pitch)))
        return ptr.value, pitch.value

    def process_in_place_dev(self, frames, out):
        """the frames were written at input_dev(): StreamingResampler::process without the append copy"""
        assert out.stride(2) == 1 and out.shape[2] >= self.out_max
        n = C.c_size_t(0)
        self.ctx._ck(self.ctx.L.sk_resampler_batch_process_dev(self.h, None, 0, 0, frames, C.c_void_p(out.data_ptr()), out.stride(0), out.stride(1),
                                                              C.byref(n)))
        return n.value

    def flush_dev(self, out):
        assert out.stride(2) == 1 and out.shape[2] >= self.out_max
        n = C.c_size_t(0)
        self.ctx._ck(self.ctx.L.sk_resampler_batch_flush_dev(self.h, C.c_void_p(out.data_ptr()), out.stride(0), out.stride(1), C.byref(n)))
        return n.value


class Resampler:
    """sk_resampler: rubato SincFixedIn<f32> with soundkit's parameters, state on the device."""

    def __init__(self, ctx, in_rate, out_rate, channels, chunk, params=None):
        self.ctx, self.channels, self.chunk = ctx, channels, chunk
        h = C.c_void_p()
        ctx._ck(ctx.L.sk_resampler_create(ctx.h, in_rate, out_rate, channels, chunk,
                                          C.byref(params) if params is not None else None, C.byref(h)))
        self.h = h
        self.out_max = ctx.L.sk_resampler_out_max(h)

    def close(self):
        if getattr(self, "h", None):
            self.ctx.L.sk_resampler_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def last_index(self):
        return self.ctx.L.sk_resampler_last_index(self.h)

    def table(self):
        t = np.empty((256, 256), np.float32)
        self.ctx._ck(self.ctx.L.sk_resampler_copy_table(self.h, _p(t)))
        return t

    def _run(self, planar, partial):
        out = np.zeros((self.channels, max(self.out_max, 1)), np.float32)
        op = (C.c_void_p * self.channels)(*[out[c].ctypes.data for c in range(self.channels)])
        n = C.c_size_t()
        if planar is None:
            self.ctx._ck(self.ctx.L.sk_resampler_process_partial(self.h, None, 0, op, C.byref(n)))
        else:
            p = np.ascontiguousarray(planar, np.float32)
            ip = (C.c_void_p * self.channels)(*[p[c].ctypes.data for c in range(self.channels)])
            if partial:
                self.ctx._ck(self.ctx.L.sk_resampler_process_partial(self.h, ip, p.shape[1], op, C.byref(n)))
            else:
                assert p.shape[1] >= self.chunk
                self.ctx._ck(self.ctx.L.sk_resampler_process(self.h, ip, op, C.byref(n)))
        return out[:, :n.value].copy()

    def process(self, planar):
        return self._run(planar, False)

    def process_partial(self, planar_or_none):
        return self._run(planar_or_none, True)
