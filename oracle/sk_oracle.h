/*
 * sk_oracle.h — CPU restatement of soundkit's canonical audio normalization path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the timed CPU baseline.
 * The product (soundkit_b200/) never links, imports or calls it.
 *
 * Every function cites the reference file:line (relative to /root/reference) it
 * restates.  The Rust reference cannot be built in this image (no cargo/rustc),
 * so there is no oracle/_ref; the restatement is pinned against the reference's
 * own fixtures and unit vectors (tests/test_oracle_golden.py):
 *   - G.711 decode: golden/g711_{ulaw,alaw}/<name>.decoded.wav  (bit-exact)
 *   - G.711 unit asserts soundkit-g711/src/lib.rs:402-479
 *   - conversion unit vectors soundkit/src/audio_bytes.rs:380-468,
 *     soundkit/src/audio_pipeline.rs:698-769
 *   - CAF/MOV endian-swap SHA-256 (soundkit-audio-demux tests)
 *   - resampler LENGTH relation soundkit-decoder/src/lib.rs:5187-5238
 * RESAMPLER SAMPLE VALUES: PARITY UNPINNED.  rubato 0.14.1 is a crates.io
 * dependency (soundkit/Cargo.toml:15) that is not vendored under /root/reference
 * and no reference test or golden file fixes a single resampled value; the
 * restatement in rubato_sinc.c follows rubato 0.14.1's published algorithm
 * (SincFixedIn<f32>, AVX+FMA interpolator order).
 */
#ifndef SK_ORACLE_H
#define SK_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- formats --------------------------------------------------------- */
enum { SKO_LAW_MULAW = 0, SKO_LAW_ALAW = 1 };
enum { SKO_FMT_SIGNED = 0, SKO_FMT_FLOAT = 1 };      /* EncodingFlag PCMSigned / PCMFloat */
enum { SKO_MODE_A = 0, SKO_MODE_B = 1 };             /* soundkit / decoder scale convention */

/* ---- R2: G.711 (soundkit-g711/src/lib.rs) ---------------------------- */
uint8_t sko_g711_encode_sample(int law, int16_t s);
int16_t sko_g711_decode_sample(int law, uint8_t b);
void sko_g711_encode_i16(int law, const int16_t* in, size_t n, uint8_t* out);
void sko_g711_encode_i32(int law, const int32_t* in, size_t n, uint8_t* out);
void sko_g711_decode_i16(int law, const uint8_t* in, size_t n, int16_t* out);
void sko_g711_decode_i32(int law, const uint8_t* in, size_t n, int32_t* out);
void sko_g711_decode_f32(int law, const uint8_t* in, size_t n, float* out);

/* ---- R1: flat conversions (soundkit/src/audio_bytes.rs) -------------- */
/* All take `bytes` input bytes and return the number of samples written.  */
size_t sko_i16le_to_f32(const uint8_t* in, size_t bytes, float* out);
size_t sko_s16le_to_i16(const uint8_t* in, size_t bytes, int16_t* out);
size_t sko_s16be_to_i16(const uint8_t* in, size_t bytes, int16_t* out);
size_t sko_s16le_to_i32(const uint8_t* in, size_t bytes, int32_t* out);
size_t sko_s24le_to_i32(const uint8_t* in, size_t bytes, int32_t* out);
size_t sko_s24le_to_i16(const uint8_t* in, size_t bytes, int16_t* out);
size_t sko_s24be_to_i16(const uint8_t* in, size_t bytes, int16_t* out);
size_t sko_s32le_to_i32(const uint8_t* in, size_t bytes, int32_t* out);
size_t sko_s32be_to_i32(const uint8_t* in, size_t bytes, int32_t* out);
size_t sko_s32le_to_s24(const uint8_t* in, size_t bytes, int32_t* out);
size_t sko_s32be_to_s24(const uint8_t* in, size_t bytes, int32_t* out);
size_t sko_s32le_to_f32(const uint8_t* in, size_t bytes, float* out);
size_t sko_s32be_to_f32(const uint8_t* in, size_t bytes, float* out);
size_t sko_s32le_to_i16(const uint8_t* in, size_t bytes, int16_t* out);
size_t sko_s32be_to_i16(const uint8_t* in, size_t bytes, int16_t* out);
size_t sko_f32le_to_i16(const uint8_t* in, size_t bytes, int16_t* out);
size_t sko_f32be_to_i16(const uint8_t* in, size_t bytes, int16_t* out);
size_t sko_f32le_to_i32(const uint8_t* in, size_t bytes, int32_t* out);
size_t sko_f32le_to_s24(const uint8_t* in, size_t bytes, int32_t* out);
void   sko_i16_to_i16le(const int16_t* in, size_t n, uint8_t* out);

/* planar <-> interleaved; planar buffers are [channels][frames] contiguous. */
void   sko_interleave_vecs_i16(const int16_t* planar, size_t channels, size_t frames, uint8_t* out);
size_t sko_deinterleave_vecs_i16(const uint8_t* in, size_t bytes, size_t channels, int16_t* planar);
size_t sko_deinterleave_vecs_s24(const uint8_t* in, size_t bytes, size_t channels, int32_t* planar);
size_t sko_deinterleave_vecs_f32(const uint8_t* in, size_t bytes, size_t channels, float* planar);
void   sko_stereo_to_mono_take_left(const int16_t* in, size_t n, int16_t* out);
void   sko_stereo_to_mono_avg(const int16_t* in, size_t n, int16_t* out);

/* soundkit/src/audio_pipeline.rs:17-50 */
void sko_vec_f32_to_i16(const float* in, size_t n, int16_t* out);
void sko_vec_i16_to_f32(const int16_t* in, size_t n, float* out);
void sko_vec_i32_to_f32(const int32_t* in, size_t n, float* out);

/* audio_to_f32_channels (mode A, audio_pipeline.rs:74-98) and
 * audio_data_to_f32_channels (mode B, soundkit-decoder/src/lib.rs:3563-3617).
 * out is planar [channels][frames]; returns frames, or (size_t)-1 on the
 * reference's error conditions. */
size_t sko_pcm_to_planar_f32(const uint8_t* in, size_t bytes, int bits, int fmt, int big_endian,
                             size_t channels, int mode, float* out);
/* mixdown_to_mono_f32 (audio_pipeline.rs:105-131): sum then scale */
void sko_mixdown_to_mono_f32(const float* planar, size_t channels, size_t frames, float* out);
/* downmix_channels (soundkit-decoder/src/lib.rs:3492-3561); returns output channels */
size_t sko_downmix_channels(const float* planar, size_t channels, size_t frames, int target_channels,
                            float* out);
/* float_sample_to_i16 (soundkit-decoder/src/lib.rs:1815-1827) */
int16_t sko_float_sample_to_i16(float s);
/* f32_channels_to_bytes (soundkit-decoder/src/lib.rs:3619-3701); returns bytes written or (size_t)-1 */
size_t sko_f32_channels_to_bytes(const float* planar, size_t channels, size_t frames, int bits,
                                 int fmt, uint8_t* out);
/* exact_signed_pcm_to_i16 (soundkit-decoder/src/lib.rs:3458-3489); returns bytes written or -1 */
size_t sko_exact_signed_pcm_to_i16(const uint8_t* in, size_t bytes, int bits, int big_endian,
                                   uint8_t* out);

/* ---- R3/R4: rubato 0.14.1 SincFixedIn<f32> restatement ----------------- */
typedef struct sko_resampler sko_resampler;
/* SincFixedIn::new(out/in, 2.0, {256, 0.95, Linear, 256, BlackmanHarris2}, chunk, channels) */
sko_resampler* sko_resampler_new(uint32_t in_rate, uint32_t out_rate, size_t chunk, size_t channels);
void   sko_resampler_free(sko_resampler*);
size_t sko_resampler_out_max(const sko_resampler*);
const float* sko_resampler_table(const sko_resampler*);   /* [256][256] sincs[subindex][tap] */
double sko_resampler_last_index(const sko_resampler*);
/* process(): in = channels pointers to >= chunk frames, out = channels pointers with
 * room for out_max frames.  threads > 1 splits the OUTPUT index range over pthreads
 * after a serial index walk (same arithmetic, only for the timed CPU baseline).
 * executes both interpolation products like the reference (P = 2). Returns n_out. */
size_t sko_resampler_process(sko_resampler*, const float* const* in, float* const* out, int threads);
/* process_partial(Some(x)|None): zero-pads `frames` (< = chunk) to chunk. */
size_t sko_resampler_process_partial(sko_resampler*, const float* const* in_or_null, size_t frames,
                                     float* const* out, int threads);
/* index walk only: fills i0[n], s0[n], frac[n] for one chunk starting from last_index
 * without touching state; any output pointer may be NULL. Returns n_out. */
size_t sko_resampler_walk(const sko_resampler*, int32_t* i0, uint8_t* s0, float* frac, size_t cap);
/* which dot-product body runs: 1 = AVX2+FMA intrinsics, 0 = scalar fmaf (same results) */
int    sko_resampler_uses_avx(void);

/* downsample_audio (audio_pipeline.rs:438-493): mode-A convert + one-shot resample.
 * out planar [channels][out_cap]; returns n_out per channel or (size_t)-1 on validation error. */
size_t sko_downsample_audio(const uint8_t* in, size_t bytes, int bits, int fmt, size_t channels,
                            uint32_t in_rate, uint32_t out_rate, float* out, size_t out_cap,
                            int threads);

/* StreamingResampler (soundkit-decoder/src/lib.rs:1917-2060): accum + chunk-4096 processing. */
typedef struct sko_stream sko_stream;
sko_stream* sko_stream_new(uint32_t in_rate, uint32_t out_rate, size_t channels);
void   sko_stream_free(sko_stream*);
/* push planar input [channels][frames]; appends produced frames to out planar
 * [channels][out_cap] at offset *out_len; returns 0 or -1 */
int sko_stream_process(sko_stream*, const float* planar, size_t frames, float* out, size_t out_cap,
                       size_t* out_len);
int sko_stream_flush(sko_stream*, float* out, size_t out_cap, size_t* out_len);

/* full chain used by bench config C5 / apply_output_options with rate change:
 * mode-B convert -> streaming resample per channel -> downmix_channels(target 1) ->
 * f32 mono. out_cap frames; returns mono frames written. */
size_t sko_normalize_chain_mono(const uint8_t* in, size_t bytes, int bits, int fmt, int big_endian,
                                size_t channels, uint32_t in_rate, uint32_t out_rate, float* out,
                                size_t out_cap);

/* ---- 8(f) next row 4: G.726 ADPCM 16/24/32/40 kbit/s (soundkit-g726/src/lib.rs:215-860), g726.c ---- */
#define SKO_G726_STATE_INTS 24
void   sko_g726_state_init(int32_t* state);
size_t sko_g726_decode(int bits, int packing, int32_t* state, const uint8_t* in, size_t bytes, int16_t* out);
size_t sko_g726_encode(int bits, int packing, int32_t* state, const int16_t* in, size_t n, uint8_t* out);

/* ---- 8(f) next row 3: packet scalers (soundkit/src/audio_packet.rs:105-124, 176-186, 243-254), AVI u8 (soundkit-decoder/
 * src/lib.rs:362-367) and unpack_dvd_lpcm (soundkit-decoder/src/lib.rs:750-804); pcm.c ---- */
size_t sko_pkt_s16le_to_f32(const uint8_t* in, size_t bytes, float* out);
size_t sko_pkt_s24le_to_f32(const uint8_t* in, size_t bytes, float* out);
size_t sko_pkt_s32le_to_i16(const uint8_t* in, size_t bytes, int16_t* out);
size_t sko_pkt_f32le_to_i16(const uint8_t* in, size_t bytes, int16_t* out);
size_t sko_u8_to_s16(const uint8_t* in, size_t bytes, int16_t* out);
size_t sko_unpack_dvd_lpcm(const uint8_t* packed, size_t bytes, int bits, int channels, uint8_t* out);

/* ---- 8(f) next row 2: StreamingStereo48kNormalizer (soundkit/src/audio_pipeline.rs:133-413), norm48k.c ---- */
typedef struct sko_norm48k sko_norm48k;
sko_norm48k* sko_norm48k_new(void);
void sko_norm48k_free(sko_norm48k*);
uint64_t sko_norm48k_source_frames(const sko_norm48k*);
uint64_t sko_norm48k_output_frames(const sko_norm48k*);
int sko_norm48k_push(sko_norm48k*, const uint8_t* data, size_t bytes, int bits, int fmt, int be, int channels, uint32_t rate,
                     float* left, float* right, size_t cap, size_t* n_out);
int sko_norm48k_finish(sko_norm48k*, float* left, float* right, size_t cap, size_t* n_out);

/* ---- 8(f) next row 4': G.722 64 kbit/s sub-band ADPCM (soundkit-g722/src/lib.rs:1-230 over ezk-g722 0.1.1), g722.c ---- */
#define SKO_G722_STATE_INTS 114
void   sko_g722_state_init(int32_t* state);
size_t sko_g722_decode(int32_t* state, const uint8_t* in, size_t bytes, int16_t* out);
size_t sko_g722_encode(int32_t* state, const int16_t* in, size_t n, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif
