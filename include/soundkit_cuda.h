/*
 * soundkit_cuda.h — C ABI of libsoundkit_cuda.so (hand-written sm_100a kernels).
 *
 * The reference (wavey-ai/soundkit) has NO FFI layer for this path: its callers use plain Rust
 * functions and two traits.  This header is the `extern "C"` surface a thin `soundkit-cuda-sys`
 * crate binds; each entry point cites the reference interface it replaces
 * (paths relative to the reference repo root).  Argument VALIDATION that produces the reference's
 * exact error strings stays in the host layer above this ABI (rust/soundkit-cuda and soundkit_b200/csrc/host/soundkit_host.hpp);
 * this layer returns negative status codes + sk_last_error().
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types.
 *   - entry points without a suffix take HOST pointers: they stage through pinned memory, run the
 *     kernel on the context's stream and return after the result is back in the caller's buffer.
 *   - `_dev` entry points take DEVICE pointers, enqueue on the context's stream and return
 *     immediately (call sk_ctx_synchronize or synchronize the stream yourself).
 *   - one sk_ctx per host thread; handles are not shared across threads (the reference runs one
 *     worker thread per DecodePipeline, soundkit-decoder/src/lib.rs:2764).
 *   - there is NO CPU fallback: without a CUDA device sk_ctx_create fails with SK_E_NODEVICE.
 */
#ifndef SOUNDKIT_CUDA_H
#define SOUNDKIT_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SK_ABI_VERSION 1

typedef struct sk_ctx sk_ctx;
typedef struct sk_resampler sk_resampler;

enum sk_status {
    SK_OK = 0,
    SK_E_ARG = -1,      /* invalid argument (message says which) */
    SK_E_ALLOC = -2,    /* host or device allocation failed */
    SK_E_CUDA = -3,     /* CUDA runtime error (message carries cudaGetErrorString) */
    SK_E_STATE = -4,    /* handle used in the wrong state */
    SK_E_NODEVICE = -5  /* no usable CUDA device: the library has no CPU path */
};

/* soundkit-g711/src/lib.rs:17 `enum G711Law { MuLaw, ALaw }` */
enum sk_law { SK_LAW_MULAW = 0, SK_LAW_ALAW = 1 };
/* sample container of a typed slice (`&[i16]`, `&[i32]`, `&[f32]`) */
enum sk_kind { SK_KIND_I16 = 0, SK_KIND_I32 = 1, SK_KIND_F32 = 2 };
/* frame-header EncodingFlag::{PCMSigned, PCMFloat} (soundkit/src/audio_packet.rs:43-50) */
enum sk_format { SK_FMT_SIGNED = 0, SK_FMT_FLOAT = 1 };
/* scale convention: A = soundkit::audio_pipeline::audio_to_f32_channels (audio_pipeline.rs:74-98;
 * always LE, 24-bit / 2^31, no NaN scrub), B = soundkit-decoder audio_data_to_f32_channels
 * (soundkit-decoder/src/lib.rs:3563-3617; endian-aware, /2^15 /2^23 /2^31, non-finite -> 0) */
enum sk_scale_mode { SK_MODE_A = 0, SK_MODE_B = 1 };
/* mono mixdown order: SUM_THEN_SCALE = mixdown_to_mono_f32 (audio_pipeline.rs:105-131),
 * SCALE_THEN_SUM = downmix_channels(.., 1) (soundkit-decoder/src/lib.rs:3500-3510) */
enum sk_mix_mode { SK_MIX_NONE = 0, SK_MIX_SUM_THEN_SCALE = 1, SK_MIX_SCALE_THEN_SUM = 2 };
/* resampler call semantics: ONESHOT = downsample_audio (one rubato chunk = whole buffer, no tail
 * flush; audio_pipeline.rs:482-491), STREAMING = StreamingResampler (chunk 4096 + flush + trim;
 * soundkit-decoder/src/lib.rs:1970-2059) */
enum sk_resample_semantics { SK_SEM_ONESHOT = 0, SK_SEM_STREAMING = 1 };

/* flat element-wise conversions of soundkit/src/audio_bytes.rs (line = `pub fn` in the reference) */
enum sk_flat_op {
    SK_OP_I16LE_TO_F32 = 0,  /* :3   i16le_to_f32  (also vec_i16_to_f32, audio_pipeline.rs:29) */
    SK_OP_S16LE_TO_I16 = 1,  /* :25 i16le_to_i16, :231 s16le_to_i16 */
    SK_OP_S16BE_TO_I16 = 2,  /* :222 */
    SK_OP_S16LE_TO_I32 = 3,  /* :240 */
    SK_OP_S24LE_TO_I32 = 4,  /* :36 */
    SK_OP_S24LE_TO_I16 = 5,  /* :51 */
    SK_OP_S24BE_TO_I16 = 6,  /* :66 */
    SK_OP_S32LE_TO_I32 = 7,  /* :81 */
    SK_OP_S32BE_TO_I32 = 8,  /* :91 */
    SK_OP_S32LE_TO_S24 = 9,  /* :101 keeps the LOW 24 bits */
    SK_OP_S32BE_TO_S24 = 10, /* :112 */
    SK_OP_S32LE_TO_F32 = 11, /* :123 (also vec_i32_to_f32, audio_pipeline.rs:40) */
    SK_OP_S32BE_TO_F32 = 12, /* :134 */
    SK_OP_S32LE_TO_I16 = 13, /* :145 */
    SK_OP_S32BE_TO_I16 = 14, /* :156 */
    SK_OP_F32LE_TO_I16 = 15, /* :167 (also vec_f32_to_i16, audio_pipeline.rs:17) */
    SK_OP_F32BE_TO_I16 = 16, /* :178 */
    SK_OP_F32LE_TO_I32 = 17, /* :189 */
    SK_OP_F32LE_TO_S24 = 18, /* :205 */
    /* SURVEY.md 8(f) row 3 — the packet scalers of soundkit/src/audio_packet.rs (a third scale convention) and the
     * container-PCM unpackers that feed the output stage */
    SK_OP_PKT_S16LE_TO_F32 = 19, /* audio_packet.rs:176-180, 243-247: f32::from(i16) / f32::from(i16::MAX) (true division by 32767) */
    SK_OP_PKT_S24LE_TO_F32 = 20, /* audio_packet.rs:182-186, 250-254: i24 as f32 / (1 << 23) */
    SK_OP_PKT_S32LE_TO_I16 = 21, /* audio_packet.rs:108-112: (s as i64 * 32767 / 2147483647) clamped to i16 */
    SK_OP_PKT_F32LE_TO_I16 = 22, /* audio_packet.rs:113-121: ((x * 32767.0) as i32) clamped to i16 (no pre-clamp of x) */
    SK_OP_U8_TO_S16 = 23,        /* soundkit-decoder/src/lib.rs:362-367 (AVI 8-bit PCM): (i16::from(u8) - 128) << 8 */
    SK_OP_COUNT = 24
};

/* ------------------------------------------------------------------ context */
int  sk_abi_version(void);
int  sk_ctx_create(int device, sk_ctx** out);
void sk_ctx_destroy(sk_ctx* ctx);
/* last error message of this context (thread-confined); "" when none */
const char* sk_last_error(const sk_ctx* ctx);
/* run on a caller-owned cudaStream_t (e.g. torch's current stream); NULL = CUDA's legacy default stream */
int  sk_ctx_set_stream(sk_ctx* ctx, void* cuda_stream);
/* back to the context's own non-blocking stream (the default after sk_ctx_create) */
int  sk_ctx_reset_stream(sk_ctx* ctx);
int  sk_ctx_synchronize(sk_ctx* ctx);
/* number of kernels this context has launched so far (bench.py reports it as gpu_launches) */
uint64_t sk_ctx_launch_count(const sk_ctx* ctx);
/* input bytes per slab of the host-pointer batch pipeline (sk_normalize_batch keeps three slabs of this size plus their
 * outputs on the device; default 256 MiB).  Smaller slabs bound the device footprint, larger ones amortise launches. */
int  sk_ctx_set_batch_slab_bytes(sk_ctx* ctx, size_t bytes);
/* Arithmetic of the general-ratio (non-integer) resampler.  BIT_EXACT (default): every output is rubato's p0 + frac (p1 - p0) with
 * its 8-lane accumulation order, bit-identical to the CPU restatement.  FAST (opt-in): batches of >= 64 (stream, channel) lanes
 * are convolved against the interpolated row h_s + frac (h_{s+1} - h_s) — one product per output, half the work; results stay
 * within the north-star contract (max-abs 1e-5 of full scale, SNR >= 120 dB) but are not bit-identical.  Integer ratios (48 -> 16
 * kHz ...) and everything else are unaffected. */
enum sk_resample_mode { SK_RESAMPLE_BIT_EXACT = 0, SK_RESAMPLE_FAST = 1 };
int  sk_ctx_set_resample_mode(sk_ctx* ctx, int mode);
/* Diagnostic: how many output samples of the general-ratio resampler so far left the tiled fast path and were computed by the
 * per-output exact path (non-uniform table rows inside a warp tile, ragged tile ends).  Results are identical either way;
 * the parity tests report and bound the share.  Synchronises the context's stream. */
int  sk_ctx_resample_fallback_outputs(sk_ctx* ctx, uint64_t* count);
/* measured FP32 FMA peak of this device in TFLOP/s: register-resident microbenchmarks of the scalar
 * FFMA and the packed sm_100 FFMA2 instruction; _peak returns the larger of the two */
int  sk_ctx_measure_ffma_peak(sk_ctx* ctx, double* tflops);
int  sk_ctx_measure_ffma_peaks(sk_ctx* ctx, double* scalar_tflops, double* packed_tflops);
/* Diagnostic: fill the shared memory of every SM with `pattern` (e.g. 0x7fc00000, a quiet NaN). Shared memory keeps its
 * contents between kernels, so a parity run after this call exposes any kernel that reads a word it did not stage. */
int  sk_ctx_poison_shared_memory(sk_ctx* ctx, uint32_t pattern);

/* ------------------------------------------------------------------ kernel 2: G.711 */
/* replaces soundkit_g711::decode_i16 (soundkit-g711/src/lib.rs:52) and
 * G711Decoder::decode_{i16,i32,f32} (:183-232); out_kind selects the output slice type. */
int sk_g711_decode(sk_ctx*, int law, const uint8_t* in, size_t n, int out_kind, void* out);
int sk_g711_decode_dev(sk_ctx*, int law, const uint8_t* in, size_t n, int out_kind, void* out);
/* replaces soundkit_g711::encode_i16 (:36) and G711Encoder::encode_{i16,i32} (:120-138) */
int sk_g711_encode(sk_ctx*, int law, int in_kind, const void* in, size_t n, uint8_t* out);
int sk_g711_encode_dev(sk_ctx*, int law, int in_kind, const void* in, size_t n, uint8_t* out);

/* ------------------------------------------------------------------ kernel 1: conversion */
/* the 19 flat functions of soundkit/src/audio_bytes.rs; writes floor(bytes / in_width) elements */
int sk_pcm_flat_convert(sk_ctx*, int op, const uint8_t* in, size_t bytes, void* out, size_t* n_out);
int sk_pcm_flat_convert_dev(sk_ctx*, int op, const uint8_t* in, size_t bytes, void* out);
/* bytes per input / output element of a flat op */
int sk_flat_op_widths(int op, int* in_width, int* out_width);

/* unpack_dvd_lpcm (soundkit-decoder/src/lib.rs:750-804): DVD / MPEG-PS LPCM payload -> little-endian interleaved PCM of
 * the same byte length.  bits 16: byte swap per sample; bits 24: every 12-byte group (4 x 16 high bits big-endian, then
 * 4 low bytes) -> 4 s24le samples.  A partial block is SK_E_ARG with the reference's message in sk_last_error()
 * ("DVD LPCM ends with a partial sample block" / "... partial 24-bit sample block"; bits 20: "20-bit DVD LPCM decoding is
 * not implemented"). */
int sk_dvd_lpcm_unpack(sk_ctx*, const uint8_t* packed, size_t bytes, int bits, int channels, uint8_t* out);
int sk_dvd_lpcm_unpack_dev(sk_ctx*, const uint8_t* packed, size_t bytes, int bits, int channels, uint8_t* out);

/* deinterleave_vecs_{i16,s24,f32} (audio_bytes.rs:264-310): interleaved LE bytes -> planar
 * [channels][frames] of i16 / sign-extended i32 / f32 bits.  kind: 0 = i16, 1 = s24, 2 = f32. */
int sk_deinterleave(sk_ctx*, int kind, const uint8_t* in, size_t bytes, int channels, void* out,
                    size_t* frames);
/* interleave_vecs_i16 (audio_bytes.rs:250-262): planar [channels][frames] i16 -> LE bytes */
int sk_interleave_i16(sk_ctx*, const int16_t* planar, int channels, size_t frames, uint8_t* out);
/* WavStreamEncoder::push_planar_{i16,i32,f32} (soundkit/src/wav.rs:444-494; also PcmData -> bytes in push, :496-540): planar
 * [channels][frames] elements of elem_bytes = 2 (i16) or 4 (i32 / f32) -> interleaved little-endian frames.  A bit copy (i32
 * samples and NaN payloads pass unchanged); validation (format, channel count, reserve_chunk) stays in the host layer. */
int sk_interleave_planar(sk_ctx*, int elem_bytes, const void* planar, int channels, size_t frames, uint8_t* out);
int sk_interleave_planar_dev(sk_ctx*, int elem_bytes, const void* planar, int channels, size_t frames, uint8_t* out);
/* stereo_to_mono_{take_left,avg} (audio_bytes.rs:317-373): avg != 0 -> ((l + r) / 2) truncating */
int sk_stereo_to_mono_i16(sk_ctx*, int avg, const int16_t* in, size_t n, int16_t* out);

/* audio_to_f32_channels (mode A) / audio_data_to_f32_channels (mode B): interleaved PCM bytes ->
 * planar f32 [channels][frames].  `frames` = whole frames in `bytes` (mode A silently drops a
 * trailing partial frame like the reference; mode B rejects it: SK_E_ARG). */
int sk_pcm_to_planar_f32(sk_ctx*, const uint8_t* in, size_t bytes, int bits, int fmt, int big_endian,
                         int channels, int scale_mode, float* out, size_t frames_cap, size_t* frames);
/* batched, device resident: n_streams buffers of `frames` frames each, in_stride bytes apart;
 * out[s][c][f] with out_stride floats between streams (channel planes are `frames` apart). */
int sk_pcm_to_planar_f32_dev(sk_ctx*, const uint8_t* in, size_t in_stride, size_t n_streams,
                             size_t frames, int bits, int fmt, int big_endian, int channels,
                             int scale_mode, float* out, size_t out_stride);
/* f32_channels_to_bytes (soundkit-decoder/src/lib.rs:3619-3701): planar f32 -> interleaved LE
 * bytes; bits 16 uses float_sample_to_i16 (:1815), 24/32 the clamp-and-truncate scalers,
 * fmt FLOAT (bits 32) is a bit copy (interleave_vecs_f32, :3685). */
int sk_planar_f32_to_pcm(sk_ctx*, const float* planar, size_t frames, int channels, int bits, int fmt,
                         uint8_t* out);
int sk_planar_f32_to_pcm_dev(sk_ctx*, const float* planar, size_t frames, int channels, int bits,
                             int fmt, uint8_t* out);
/* exact_signed_pcm_to_i16 (soundkit-decoder/src/lib.rs:3458-3489): 24/32-bit LE/BE -> s16 LE */
int sk_pcm_narrow_i16(sk_ctx*, const uint8_t* in, size_t bytes, int bits, int big_endian, uint8_t* out);
/* mixdown_to_mono_f32 / downmix_channels(.., 1): planar [channels][frames] -> mono */
int sk_mixdown_mono(sk_ctx*, const float* planar, size_t frames, int channels, int mix_mode, float* out);
int sk_mixdown_mono_dev(sk_ctx*, const float* planar, size_t frames, int channels, int mix_mode,
                        float* out);
/* downmix_channels(.., 2) for > 2 channels (soundkit-decoder/src/lib.rs:3512-3557), including the
 * block-wide peak normalisation; out is planar [2][frames]. */
int sk_downmix_stereo(sk_ctx*, const float* planar, size_t frames, int channels, float* out);

/* ------------------------------------------------------------------ kernel 3: sinc resampler */
/* rubato SincInterpolationParameters as soundkit hard-codes them (audio_pipeline.rs:474-480,
 * soundkit-decoder/src/lib.rs:1939-1945).  Only these values are implemented; anything else is
 * SK_E_ARG.  NULL = defaults. */
typedef struct {
    uint32_t sinc_len;      /* 256 */
    uint32_t oversampling;  /* 256 */
    float    f_cutoff;      /* 0.95 */
    int      interp;        /* 0 = Linear */
    int      window;        /* 0 = BlackmanHarris2 */
    double   max_ratio_rel; /* 2.0 */
} sk_resample_params;

/* SincFixedIn::<f32>::new(out/in, 2.0, params, chunk, channels) — stateful, one per stream */
int    sk_resampler_create(sk_ctx*, uint32_t in_rate, uint32_t out_rate, int channels, size_t chunk,
                           const sk_resample_params* params, sk_resampler** out);
void   sk_resampler_destroy(sk_resampler*);
size_t sk_resampler_out_max(const sk_resampler*);        /* (chunk * ratio * 2.0 + 10) as usize */
double sk_resampler_last_index(const sk_resampler*);
/* Resampler::process(&wave_in, None): in[c] -> chunk frames, out[c] has room for out_max frames */
int    sk_resampler_process(sk_resampler*, const float* const* in, float* const* out, size_t* n_out);
/* Resampler::process_partial(Some(x) | None, None): zero-pads `frames` to one whole chunk */
int    sk_resampler_process_partial(sk_resampler*, const float* const* in_or_null, size_t frames,
                                    float* const* out, size_t* n_out);
/* the [256][256] windowed-sinc table this resampler uses (for parity tests) */
int    sk_resampler_copy_table(const sk_resampler*, float* out65536);

/* N StreamingResamplers in lock step (soundkit-decoder/src/lib.rs:1917-2060: `StreamingResampler`, rubato chunk 4096): the state
 * of n_streams live streams — per (stream, channel) the 512-sample rubato history and the not yet consumed input; the f64 index
 * of the walk is shared, it is data independent — stays on the device between calls.  Every call serves all streams with one
 * kernel launch.  `in` / `out` are DEVICE pointers to planar f32: in[s][c][0 .. frames), out[s][c][0 .. *n_out), strides in floats. */
typedef struct sk_resampler_batch sk_resampler_batch;
int    sk_resampler_batch_create(sk_ctx*, uint32_t in_rate, uint32_t out_rate, int channels, size_t n_streams,
                                 size_t max_push_frames, sk_resampler_batch** out);
void   sk_resampler_batch_destroy(sk_resampler_batch*);
/* frames per channel one process / flush call can produce at most: every `out` row must hold this many */
size_t sk_resampler_batch_out_cap(const sk_resampler_batch*);
size_t sk_resampler_batch_pending(const sk_resampler_batch*);      /* frames pushed but not yet consumed (< 4096) */
double sk_resampler_batch_last_index(const sk_resampler_batch*);
/* StreamingResampler::process (:1970-2015) for every stream: append `frames` (<= max_push_frames) frames per channel, run every
 * whole 4096-frame chunk that is pending; *n_out = frames produced per channel (0 while less than a chunk is pending) */
int    sk_resampler_batch_process_dev(sk_resampler_batch*, const float* in, size_t in_stream_stride, size_t in_ch_stride,
                                      size_t frames, float* out, size_t out_stream_stride, size_t out_ch_stride, size_t* n_out);
/* Zero-copy form of the append inside process (`accum.extend_from_slice`, :1985-1992): *in = where the next push's frames of lane
 * (stream 0, channel 0) go inside the handle's device rows, *lane_pitch = floats between consecutive (stream, channel) lanes.  A
 * device-side producer writes up to max_push_frames frames per lane there and then calls sk_resampler_batch_process_dev with
 * in == NULL (strides ignored).  Valid until the next process / flush call on this handle. */
int    sk_resampler_batch_input_dev(sk_resampler_batch*, float** in, size_t* lane_pitch);
/* StreamingResampler::flush (:2017-2059): zero-pad the remainder to one chunk, drop the padding's share of the output
 * (round(padded * out / in) frames); with nothing pending: one chunk of zeros, untrimmed */
int    sk_resampler_batch_flush_dev(sk_resampler_batch*, float* out, size_t out_stream_stride, size_t out_ch_stride, size_t* n_out);

/* ------------------------------------------------------------------ kernel 4: fused batch chain */
/* Batched, device-resident normalisation of n_streams equal-length streams:
 *   interleaved PCM (bits/fmt/endianness/channels) --convert(scale_mode)--> f32
 *   --sinc resample in_rate -> out_rate per channel--> [--mono mixdown(mix_mode)-->] f32 planar
 * in one kernel (one HBM read of the input, one HBM write of the output).
 * semantics ONESHOT reproduces soundkit::audio_pipeline::downsample_audio per stream;
 * STREAMING reproduces apply_output_options' rate-changing arm (chunk-4096 StreamingResampler,
 * flush and trim included) for a whole stream.
 * Output: mix_mode NONE -> out[s][c][n_out] (channel planes out_stride_ch floats apart),
 *         otherwise     -> out[s][n_out] mono.  out_stride = floats between streams. */
typedef struct {
    const void* in;        /* device pointer */
    size_t   in_stride;    /* bytes between consecutive streams */
    size_t   n_streams;
    size_t   frames;       /* frames per stream */
    int      bits;         /* 16 | 24 | 32 */
    int      fmt;          /* sk_format */
    int      big_endian;
    int      channels;
    int      scale_mode;   /* sk_scale_mode */
    uint32_t in_rate, out_rate;
    int      semantics;    /* sk_resample_semantics */
    int      mix_mode;     /* sk_mix_mode */
    float*   out;          /* device pointer */
    size_t   out_stride;   /* floats between consecutive streams */
    size_t   out_stride_ch;/* floats between channel planes (mix_mode NONE) */
} sk_batch_desc;

/* frames each stream produces for (in_rate, out_rate, frames, semantics); builds and caches the
 * index plan so the first sk_normalize_batch_dev call does not pay for it */
int sk_normalize_batch_plan(sk_ctx*, uint32_t in_rate, uint32_t out_rate, size_t frames, int semantics,
                            size_t* n_out);
int sk_normalize_batch_dev(sk_ctx*, const sk_batch_desc* desc);
/* The same chain with BYTES out: apply_output_options' rate-changing arm down to one channel (soundkit-decoder/src/lib.rs:
 * 3324-3456), i.e. convert -> resample -> downmix -> f32_channels_to_bytes (:3619-3683) for a device-resident batch.
 * desc->out / out_stride / out_stride_ch are ignored; the result must be mono (channels == 1 or mix_mode != NONE).
 * out[s] = n_out samples of out_bits / 8 little-endian bytes, out_stride_bytes between streams; out_fmt FLOAT needs
 * out_bits 32 (the reference only emits PCMFloat when the INPUT was float — that choice stays with the caller, :3377-3382).
 * Two-channel integer-ratio sources (48 -> 16 kHz ...) are quantised inside the fused kernel; other combinations take an
 * f32 scratch and one more pass. */
int sk_normalize_batch_pcm_dev(sk_ctx*, const sk_batch_desc* desc, int out_bits, int out_fmt, void* out, size_t out_stride_bytes);
/* host-pointer form of the batch call: desc->in / desc->out are HOST pointers (pinned memory gives
 * full PCIe rate).  Streams are processed in slabs; the H2D copy of the next slab and the D2H copy of
 * the previous one overlap the kernel of the current one.  Returns when every result is in `out`. */
int sk_normalize_batch(sk_ctx*, const sk_batch_desc* desc);
/* host-pointer form of the same call for ONE stream (what downsample_audio / the decoder's output
 * stage would call per AudioData): copies in, runs the fused kernel, copies out. */
int sk_normalize_stream(sk_ctx*, const uint8_t* in, size_t bytes, int bits, int fmt, int big_endian,
                        int channels, int scale_mode, uint32_t in_rate, uint32_t out_rate, int semantics,
                        int mix_mode, float* out, size_t out_cap_per_plane, size_t* n_out);

/* ------------------------------------------------------------------ G.726 ADPCM */
/* soundkit-g726/src/lib.rs: G726Core (:198-462) as SK_G726_STATE_INTS int32 words in the order
 * yl yu dms dml ap a[2] b[6] pk[2] dq[6] sr[2] td.  bits = code size 2 / 3 / 4 / 5 (G726Rate::bits_per_sample, :84-91:
 * 16 / 24 / 32 / 40 kbit/s); packing 0 = G726Packing::Left (`ffmpeg -f g726`), 1 = Right (`g726le`) (:191-197).
 * Inputs must be whole code groups (1 / 3 / 1 / 5 bytes = 4 / 8 / 2 / 8 samples, :114-130): the Rust G726Decoder /
 * G726Encoder keep the partial group in pending_bytes / pending_samples (:631-742, :821-860) and stay in the host layer. */
#define SK_G726_STATE_INTS 24
void sk_g726_state_init(int32_t* state);                 /* Default for G726Core, :214-230 */
/* G726Decoder::decode_{i16,i32,f32} on whole groups (:821-951); out gets bytes / group_bytes * group_samples elements */
int  sk_g726_decode(sk_ctx*, int bits, int packing, int32_t* state, const uint8_t* in, size_t bytes, int out_kind, void* out);
/* G726Encoder::encode_{i16,i32} on whole groups (:700-738); out gets n / group_samples * group_bytes bytes */
int  sk_g726_encode(sk_ctx*, int bits, int packing, int32_t* state, int in_kind, const void* in, size_t n, uint8_t* out);
/* batched, device resident: n_streams independent streams (one thread per stream); in_stride / out_stride in bytes for
 * code buffers and in ELEMENTS for sample buffers; state = [n_streams][SK_G726_STATE_INTS] device words carried across
 * calls, or NULL to start every stream from the default state and discard it */
int  sk_g726_decode_batch_dev(sk_ctx*, int bits, int packing, const uint8_t* in, size_t in_stride, size_t n_streams, size_t bytes,
                              int out_kind, void* out, size_t out_stride, int32_t* state);
int  sk_g726_encode_batch_dev(sk_ctx*, int bits, int packing, int in_kind, const void* in, size_t in_stride, size_t n_streams, size_t n,
                              uint8_t* out, size_t out_stride, int32_t* state);

/* ------------------------------------------------------------------ G.722 sub-band ADPCM */
/* soundkit-g722/src/lib.rs:1-230 over `ezk-g722 0.1.1` libg722::{Decoder, Encoder}::new(Bitrate::Mode1_64000, false, false):
 * ITU-T G.722 mode 1 (64 kbit/s), 16 kHz mono: one code byte <-> two samples.  The codec state (both band predictors + the
 * 24-tap QMF delay line) is SK_G722_STATE_INTS int32 words, all zero at reset — the crate starts both bands from a zero scale
 * factor, which is what the reference's golden WAV pins (see sk_g722.cu).  The Rust G722Encoder keeps an odd pending sample
 * (:10-13, 83-117) and G722Decoder has no pending state; both stay in the host layer and hand whole pairs / bytes to these calls. */
#define SK_G722_STATE_INTS 68
void sk_g722_state_init(int32_t* state);
/* G722Decoder::decode_{i16,i32,f32} (:161-232): out gets bytes * 2 elements */
int  sk_g722_decode(sk_ctx*, int32_t* state, const uint8_t* in, size_t bytes, int out_kind, void* out);
/* G722Encoder::encode_{i16,i32} on whole sample pairs (:83-122): n must be even, out gets n / 2 bytes */
int  sk_g722_encode(sk_ctx*, int32_t* state, int in_kind, const void* in, size_t n, uint8_t* out);
/* batched, device resident: n_streams independent streams (one thread per stream); in_stride / out_stride in bytes for code
 * buffers and in ELEMENTS for sample buffers; state = [n_streams][SK_G722_STATE_INTS] device words carried across calls, or NULL
 * to start every stream from the reset state and discard it */
int  sk_g722_decode_batch_dev(sk_ctx*, const uint8_t* in, size_t in_stride, size_t n_streams, size_t bytes, int out_kind, void* out,
                              size_t out_stride, int32_t* state);
int  sk_g722_encode_batch_dev(sk_ctx*, int in_kind, const void* in, size_t in_stride, size_t n_streams, size_t n, uint8_t* out,
                              size_t out_stride, int32_t* state);

/* ------------------------------------------------------------------ 48 kHz stereo normalizer */
/* StreamingStereo48kNormalizer (soundkit/src/audio_pipeline.rs:256-394): decoder-convention conversion, mono
 * duplicated / first two channels, finite_pcm clamp, 2-tap linear interpolation to 48 kHz over the whole-stream
 * timeline, exact rounded-duration tail.  Stateful, one per stream; the reference's error strings are produced by the
 * host layer, this layer reports SK_E_ARG / SK_E_STATE with the same text in sk_last_error(). */
typedef struct sk_norm48k sk_norm48k;
int      sk_norm48k_create(sk_ctx*, sk_norm48k** out);
void     sk_norm48k_destroy(sk_norm48k*);
/* push(&AudioData) (:288-363): writes *n_out frames to left/right (0 = Ok(None)); cap >= sk_norm48k_max_output() */
int      sk_norm48k_push(sk_norm48k*, const uint8_t* data, size_t bytes, int bits, int fmt, int big_endian, int channels,
                         uint32_t rate, float* left, float* right, size_t cap, size_t* n_out);
/* finish() (:366-394): the repeated-final-sample tail */
int      sk_norm48k_finish(sk_norm48k*, float* left, float* right, size_t cap, size_t* n_out);
uint64_t sk_norm48k_source_frames(const sk_norm48k*);      /* :280 */
uint64_t sk_norm48k_output_frames(const sk_norm48k*);      /* :284 */
uint32_t sk_norm48k_source_sample_rate(const sk_norm48k*); /* :272 */
int      sk_norm48k_source_channels(const sk_norm48k*);    /* :276 */
size_t   sk_norm48k_max_output(size_t frames, uint32_t rate); /* (frames * 48000).div_ceil(rate) + 4, :318-320 */
/* whole streams, batched and device resident: push(whole) + finish() for n_streams equal-length streams in one
 * launch.  out[s][0][n] = left, out[s][1][n] = right (out_stride floats between streams, out_stride_ch between the
 * two planes); *n_out = sk_normalize48k_out_frames(frames, rate). */
size_t   sk_normalize48k_out_frames(size_t frames, uint32_t rate);
int      sk_normalize48k_batch_dev(sk_ctx*, const void* in, size_t in_stride, size_t n_streams, size_t frames, int bits,
                                   int fmt, int big_endian, int channels, uint32_t rate, float* out, size_t out_stride,
                                   size_t out_stride_ch, size_t* n_out);

#ifdef __cplusplus
}
#endif
#endif /* SOUNDKIT_CUDA_H */
