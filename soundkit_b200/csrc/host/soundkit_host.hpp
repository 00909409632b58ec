// soundkit_host.hpp — host-side mirror of the reference's Rust interface for the normalization path,
// on top of the C ABI (include/soundkit_cuda.h).  The reference is Rust and there is no Rust toolchain
// in this image, so this layer is C++: same module / function / type names, same argument meaning,
// same error strings (Rust `Err(String)` -> soundkit::Error, Rust panic -> soundkit::Panic), so the
// parity tests read like the reference's own tests.  rust/soundkit-cuda/ holds the equivalent Rust
// sources (cannot be compiled here).
//
// All arithmetic happens in libsoundkit_cuda.so on the GPU; this file only validates arguments,
// frames bytes and moves buffers.  There is no CPU fallback: without a device every call throws.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>

#include "../../../include/soundkit_cuda.h"

namespace soundkit {

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };   // Rust Err(String)
struct Panic : std::logic_error { using std::logic_error::logic_error; };         // Rust panic!/assert!/unwrap

// ---- one sk_ctx per host thread (soundkit-decoder runs one worker thread per pipeline) -------------
class Gpu {
public:
    static sk_ctx* ctx() {
        thread_local Gpu g;
        return g.ctx_;
    }
    static void check(int rc) {
        if (rc != SK_OK) throw Error(std::string("soundkit-cuda: ") + sk_last_error(ctx()));
    }
private:
    Gpu() {
        int rc = sk_ctx_create(0, &ctx_);
        if (rc != SK_OK) throw Error("soundkit-cuda: no CUDA device (sk_ctx_create failed); there is no CPU fallback");
    }
    ~Gpu() { sk_ctx_destroy(ctx_); }
    sk_ctx* ctx_ = nullptr;
};

// ---- frame-header wire enums (soundkit/src/audio_packet.rs:43-50) -----------------------------------
enum class EncodingFlag : int { PCMSigned = 0, PCMFloat = 1, Opus = 2, FLAC = 3, AAC = 4 };
enum class Endianness : int { LittleEndian = 0, BigEndian = 1 };

namespace audio_types {
// soundkit/src/audio_types.rs:3-7
using PcmData = std::variant<std::vector<std::vector<int16_t>>, std::vector<std::vector<int32_t>>, std::vector<std::vector<float>>>;
// soundkit/src/audio_types.rs:9-61
class AudioData {
public:
    AudioData(uint8_t bits_per_sample, uint8_t channel_count, uint32_t sampling_rate, std::vector<uint8_t> data, EncodingFlag audio_format,
              Endianness endianness)
        : bits_(bits_per_sample), ch_(channel_count), rate_(sampling_rate), data_(std::move(data)), fmt_(audio_format), end_(endianness) {}
    uint8_t bits_per_sample() const { return bits_; }
    uint8_t channel_count() const { return ch_; }
    uint32_t sampling_rate() const { return rate_; }
    const std::vector<uint8_t>& data() const { return data_; }
    EncodingFlag audio_format() const { return fmt_; }
    Endianness endianness() const { return end_; }
private:
    uint8_t bits_, ch_; uint32_t rate_; std::vector<uint8_t> data_; EncodingFlag fmt_; Endianness end_;
};
}  // namespace audio_types
using audio_types::AudioData;
using audio_types::PcmData;

// ---- soundkit::audio_bytes (soundkit/src/audio_bytes.rs) ----------------------------------------------
namespace audio_bytes {
namespace detail {
template <typename T>
inline std::vector<T> flat(int op, const std::vector<uint8_t>& in) {
    int iw = 0, ow = 0;
    sk_flat_op_widths(op, &iw, &ow);
    std::vector<T> out(in.size() / (size_t)iw);
    size_t n = 0;
    Gpu::check(sk_pcm_flat_convert(Gpu::ctx(), op, in.data(), in.size(), out.data(), &n));
    out.resize(n);
    return out;
}
template <typename T>
inline std::vector<std::vector<T>> split_planes(const std::vector<T>& flat, size_t channels, size_t frames) {
    std::vector<std::vector<T>> out(channels);
    for (size_t c = 0; c < channels; c++) out[c].assign(flat.begin() + c * frames, flat.begin() + (c + 1) * frames);
    return out;
}
template <typename T>
inline std::vector<std::vector<T>> deinterleave(int kind, const std::vector<uint8_t>& in, size_t channel_count, size_t bytes_per_sample) {
    if (channel_count == 0) throw Panic("attempt to divide by zero");  // input.len() / (channel_count * N)
    size_t frames = in.size() / (channel_count * bytes_per_sample);
    std::vector<T> flat(channel_count * frames);
    size_t got = 0;
    Gpu::check(sk_deinterleave(Gpu::ctx(), kind, in.data(), in.size(), (int)channel_count, flat.data(), &got));
    return split_planes(flat, channel_count, frames);
}
}  // namespace detail

inline std::vector<float> i16le_to_f32(const std::vector<uint8_t>& bytes) {                        // :3
    if (bytes.size() % 2) throw Panic("Bytes length must be a multiple of 2");
    return detail::flat<float>(SK_OP_I16LE_TO_F32, bytes);
}
inline std::vector<uint8_t> i16_to_i16le(const std::vector<int16_t>& data) {                       // :17 (byte view on a LE host)
    std::vector<uint8_t> out(data.size() * 2);
    std::memcpy(out.data(), data.data(), out.size());
    return out;
}
inline std::vector<int16_t> i16le_to_i16(const std::vector<uint8_t>& bytes) {                      // :25
    if (bytes.size() % 2) throw Panic("Bytes length must be a multiple of 2");
    return detail::flat<int16_t>(SK_OP_S16LE_TO_I16, bytes);
}
inline std::vector<int32_t> s24le_to_i32(const std::vector<uint8_t>& d) { return detail::flat<int32_t>(SK_OP_S24LE_TO_I32, d); }   // :36
inline std::vector<int16_t> s24le_to_i16(const std::vector<uint8_t>& d) { return detail::flat<int16_t>(SK_OP_S24LE_TO_I16, d); }   // :51
inline std::vector<int16_t> s24be_to_i16(const std::vector<uint8_t>& d) { return detail::flat<int16_t>(SK_OP_S24BE_TO_I16, d); }   // :66
inline std::vector<int32_t> s32le_to_i32(const std::vector<uint8_t>& d) { return detail::flat<int32_t>(SK_OP_S32LE_TO_I32, d); }   // :81
inline std::vector<int32_t> s32be_to_i32(const std::vector<uint8_t>& d) { return detail::flat<int32_t>(SK_OP_S32BE_TO_I32, d); }   // :91
inline std::vector<int32_t> s32le_to_s24(const std::vector<uint8_t>& d) { return detail::flat<int32_t>(SK_OP_S32LE_TO_S24, d); }   // :101
inline std::vector<int32_t> s32be_to_s24(const std::vector<uint8_t>& d) { return detail::flat<int32_t>(SK_OP_S32BE_TO_S24, d); }   // :112
inline std::vector<float> s32le_to_f32(const std::vector<uint8_t>& d) { return detail::flat<float>(SK_OP_S32LE_TO_F32, d); }       // :123
inline std::vector<float> s32be_to_f32(const std::vector<uint8_t>& d) { return detail::flat<float>(SK_OP_S32BE_TO_F32, d); }       // :134
inline std::vector<int16_t> s32le_to_i16(const std::vector<uint8_t>& d) { return detail::flat<int16_t>(SK_OP_S32LE_TO_I16, d); }   // :145
inline std::vector<int16_t> s32be_to_i16(const std::vector<uint8_t>& d) { return detail::flat<int16_t>(SK_OP_S32BE_TO_I16, d); }   // :156
inline std::vector<int16_t> f32le_to_i16(const std::vector<uint8_t>& d) { return detail::flat<int16_t>(SK_OP_F32LE_TO_I16, d); }   // :167
inline std::vector<int16_t> f32be_to_i16(const std::vector<uint8_t>& d) { return detail::flat<int16_t>(SK_OP_F32BE_TO_I16, d); }   // :178
inline std::vector<int32_t> f32le_to_i32(const std::vector<uint8_t>& d) { return detail::flat<int32_t>(SK_OP_F32LE_TO_I32, d); }   // :189
inline std::vector<int32_t> f32le_to_s24(const std::vector<uint8_t>& d) { return detail::flat<int32_t>(SK_OP_F32LE_TO_S24, d); }   // :205
inline std::vector<int16_t> s16be_to_i16(const std::vector<uint8_t>& d) { return detail::flat<int16_t>(SK_OP_S16BE_TO_I16, d); }   // :222
inline std::vector<int16_t> s16le_to_i16(const std::vector<uint8_t>& d) { return detail::flat<int16_t>(SK_OP_S16LE_TO_I16, d); }   // :231
inline std::vector<int32_t> s16le_to_i32(const std::vector<uint8_t>& d) { return detail::flat<int32_t>(SK_OP_S16LE_TO_I32, d); }   // :240

inline std::vector<uint8_t> interleave_vecs_i16(const std::vector<std::vector<int16_t>>& channels) {   // :250
    if (channels.empty()) throw Panic("index out of bounds: the len is 0 but the index is 0");  // channels[0]
    const size_t frames = channels[0].size();
    std::vector<int16_t> planar(channels.size() * frames);
    for (size_t c = 0; c < channels.size(); c++) {
        if (channels[c].size() < frames) throw Panic("index out of bounds");  // channel[i]
        std::memcpy(planar.data() + c * frames, channels[c].data(), frames * 2);
    }
    std::vector<uint8_t> out(planar.size() * 2);
    Gpu::check(sk_interleave_i16(Gpu::ctx(), planar.data(), (int)channels.size(), frames, out.data()));
    return out;
}
inline std::vector<std::vector<int16_t>> deinterleave_vecs_i16(const std::vector<uint8_t>& in, size_t ch) { return detail::deinterleave<int16_t>(0, in, ch, 2); }  // :264
inline std::vector<std::vector<int32_t>> deinterleave_vecs_s24(const std::vector<uint8_t>& in, size_t ch) { return detail::deinterleave<int32_t>(1, in, ch, 3); }  // :280
inline std::vector<std::vector<float>> deinterleave_vecs_f32(const std::vector<uint8_t>& in, size_t ch) { return detail::deinterleave<float>(2, in, ch, 4); }      // :296
inline int32_t s24le_to_i32_sample(const uint8_t (&b)[3]) {                                           // :312
    return s24le_to_i32(std::vector<uint8_t>{b[0], b[1], b[2]}).at(0);
}
inline std::vector<int16_t> stereo_to_mono_take_left(const std::vector<int16_t>& input) {            // :317
    if (input.size() % 2) throw Panic("Stereo buffer must contain an even number of samples");
    std::vector<int16_t> out(input.size() / 2);
    Gpu::check(sk_stereo_to_mono_i16(Gpu::ctx(), 0, input.data(), input.size(), out.data()));
    return out;
}
inline std::vector<int16_t> stereo_to_mono_avg(const std::vector<int16_t>& input) {                  // :344
    if (input.size() % 2) throw Panic("Stereo buffer must contain an even number of samples");
    std::vector<int16_t> out(input.size() / 2);
    Gpu::check(sk_stereo_to_mono_i16(Gpu::ctx(), 1, input.data(), input.size(), out.data()));
    return out;
}
// the in-place variants (:331, :360) return the leading `frames` samples of the same buffer
inline size_t stereo_to_mono_inplace_take_left(std::vector<int16_t>& samples) {
    auto m = stereo_to_mono_take_left(samples);
    std::memcpy(samples.data(), m.data(), m.size() * 2);
    return m.size();
}
inline size_t stereo_to_mono_inplace_avg(std::vector<int16_t>& samples) {
    auto m = stereo_to_mono_avg(samples);
    std::memcpy(samples.data(), m.data(), m.size() * 2);
    return m.size();
}
}  // namespace audio_bytes

// ---- soundkit::raw_pcm (soundkit/src/raw_pcm.rs) — host-side framing, no arithmetic ---------------------
namespace raw_pcm {
constexpr size_t MAX_INPUT_CHUNK_BYTES = 4 * 1024 * 1024;  // :3
enum class RawPcmSampleFormat { I16, I24, I32, F32 };      // :6-11
inline uint8_t bits_per_sample(RawPcmSampleFormat f) { return f == RawPcmSampleFormat::I16 ? 16 : (f == RawPcmSampleFormat::I24 ? 24 : 32); }  // :15
inline size_t bytes_per_sample(RawPcmSampleFormat f) { return bits_per_sample(f) / 8; }                                                     // :23
inline EncodingFlag encoding_flag(RawPcmSampleFormat f) { return f == RawPcmSampleFormat::F32 ? EncodingFlag::PCMFloat : EncodingFlag::PCMSigned; }  // :27

class RawPcmFormat {  // :38-126
public:
    static RawPcmFormat make(uint32_t sample_rate, uint8_t channels, RawPcmSampleFormat sample_format, Endianness endianness) {  // new, :46
        RawPcmFormat f(sample_rate, channels, sample_format, endianness);
        f.validate();
        return f;
    }
    static RawPcmFormat linear16(uint32_t r, uint8_t c) { return make(r, c, RawPcmSampleFormat::I16, Endianness::LittleEndian); }  // :62
    static RawPcmFormat l16(uint32_t r, uint8_t c) { return make(r, c, RawPcmSampleFormat::I16, Endianness::BigEndian); }          // :71
    static RawPcmFormat linear32(uint32_t r, uint8_t c) { return make(r, c, RawPcmSampleFormat::F32, Endianness::LittleEndian); }  // :80
    size_t bytes_per_frame() const { return raw_pcm::bytes_per_sample(fmt_) * channels_; }  // :89
    uint32_t sample_rate() const { return rate_; }
    uint8_t channels() const { return channels_; }
    RawPcmSampleFormat sample_format() const { return fmt_; }
    Endianness endianness() const { return end_; }
    uint8_t bits_per_sample() const { return raw_pcm::bits_per_sample(fmt_); }
    EncodingFlag encoding_flag() const { return raw_pcm::encoding_flag(fmt_); }
    void validate() const {  // :117-125
        if (rate_ == 0) throw Error("Raw PCM sample rate must be > 0");
        if (channels_ == 0) throw Error("Raw PCM channel count must be > 0");
    }
private:
    RawPcmFormat(uint32_t r, uint8_t c, RawPcmSampleFormat f, Endianness e) : rate_(r), channels_(c), fmt_(f), end_(e) {}
    uint32_t rate_; uint8_t channels_; RawPcmSampleFormat fmt_; Endianness end_;
};

class RawPcmStreamProcessor {  // :128-191
public:
    explicit RawPcmStreamProcessor(RawPcmFormat format) : format_(format) {}
    RawPcmFormat format() const { return format_; }
    size_t buffered_len() const { return buffer_.size(); }
    std::optional<AudioData> add(const std::vector<uint8_t>& chunk) {  // :150
        if (chunk.size() > MAX_INPUT_CHUNK_BYTES)
            throw Error("raw PCM input chunk exceeds the " + std::to_string(MAX_INPUT_CHUNK_BYTES) + " byte streaming budget");
        buffer_.insert(buffer_.end(), chunk.begin(), chunk.end());
        const size_t bpf = format_.bytes_per_frame();
        if (bpf == 0) throw Error("Raw PCM bytes per frame must be > 0");
        const size_t complete = buffer_.size() / bpf * bpf;
        if (complete == 0) return std::nullopt;
        std::vector<uint8_t> data(buffer_.begin(), buffer_.begin() + complete);
        buffer_.erase(buffer_.begin(), buffer_.begin() + complete);
        return AudioData(format_.bits_per_sample(), format_.channels(), format_.sample_rate(), std::move(data), format_.encoding_flag(),
                         format_.endianness());
    }
    std::optional<AudioData> flush() {  // :181
        if (buffer_.empty()) return std::nullopt;
        throw Error("Raw PCM stream ended with " + std::to_string(buffer_.size()) + " trailing partial-frame byte(s)");
    }
private:
    RawPcmFormat format_;
    std::vector<uint8_t> buffer_;
};
}  // namespace raw_pcm

// ---- soundkit::audio_pipeline hot subset (soundkit/src/audio_pipeline.rs) ---------------------------------
namespace audio_pipeline {
inline std::vector<uint8_t> as_bytes(const void* p, size_t n) { return std::vector<uint8_t>((const uint8_t*)p, (const uint8_t*)p + n); }

inline std::vector<int16_t> vec_f32_to_i16(const std::vector<float>& input) {   // :17
    return audio_bytes::detail::flat<int16_t>(SK_OP_F32LE_TO_I16, as_bytes(input.data(), input.size() * 4));
}
inline std::vector<float> vec_i16_to_f32(const std::vector<int16_t>& input) {   // :29
    return audio_bytes::detail::flat<float>(SK_OP_I16LE_TO_F32, as_bytes(input.data(), input.size() * 2));
}
inline std::vector<float> vec_i32_to_f32(const std::vector<int32_t>& input) {   // :40 (2147483647.0f32 == 2^31)
    return audio_bytes::detail::flat<float>(SK_OP_S32LE_TO_F32, as_bytes(input.data(), input.size() * 4));
}
inline PcmData deserialize_audio(const std::vector<uint8_t>& data, uint8_t bits_per_sample, uint8_t channel_count) {  // :52
    switch (bits_per_sample) {
        case 16: return audio_bytes::deinterleave_vecs_i16(data, channel_count);
        case 24: return audio_bytes::deinterleave_vecs_s24(data, channel_count);
        case 32: return audio_bytes::deinterleave_vecs_f32(data, channel_count);
        default: throw Error("unsuporrted type");
    }
}
inline std::vector<std::vector<float>> audio_to_f32_channels(const AudioData& audio) {  // :74 (scale convention A)
    const size_t ch = audio.channel_count();
    if (ch == 0) throw Error("Channel count must be > 0");
    const int bits = audio.bits_per_sample();
    if (bits != 16 && bits != 24 && bits != 32) throw Error("deserialize_audio failed: unsuporrted type");
    const int fmt = (bits == 32 && audio.audio_format() == EncodingFlag::PCMFloat) ? SK_FMT_FLOAT : SK_FMT_SIGNED;
    const size_t frames = audio.data().size() / ((size_t)(bits / 8) * ch);
    std::vector<float> flat(ch * frames);
    size_t got = 0;
    Gpu::check(sk_pcm_to_planar_f32(Gpu::ctx(), audio.data().data(), audio.data().size(), bits, fmt, 0, (int)ch, SK_MODE_A, flat.data(), frames, &got));
    return audio_bytes::detail::split_planes(flat, ch, frames);
}
inline std::vector<float> mixdown_to_mono_f32(const std::vector<std::vector<float>>& channels) {  // :105
    if (channels.empty()) return {};
    if (channels.size() == 1) return channels[0];
    const size_t n = channels[0].size();
    for (auto& c : channels)
        if (c.size() != n) throw Error("channel length mismatch");
    std::vector<float> planar(channels.size() * n), out(n);
    for (size_t c = 0; c < channels.size(); c++) std::memcpy(planar.data() + c * n, channels[c].data(), n * 4);
    Gpu::check(sk_mixdown_mono(Gpu::ctx(), planar.data(), n, (int)channels.size(), SK_MIX_SUM_THEN_SCALE, out.data()));
    return out;
}
inline std::vector<float> audio_to_mono_f32(const AudioData& audio) { return mixdown_to_mono_f32(audio_to_f32_channels(audio)); }  // :100
inline std::vector<uint8_t> f32s_to_le_bytes(const std::vector<float>& samples) { return as_bytes(samples.data(), samples.size() * 4); }  // :415 bit copy
inline std::vector<float> f32s_from_le_bytes(const std::vector<uint8_t>& bytes) {  // :423
    if (bytes.size() % 4) throw Error("invalid f32le byte length " + std::to_string(bytes.size()) + "; expected multiple of 4");
    std::vector<float> out(bytes.size() / 4);
    std::memcpy(out.data(), bytes.data(), bytes.size());
    return out;
}
// downsample_audio, :438-493 (feature sinc-resampler): mode-A convert + one rubato chunk, fused on the device
inline std::vector<std::vector<float>> downsample_audio(const AudioData& audio, size_t sampling_rate) {
    static const uint32_t kRates[9] = {8000, 16000, 22050, 24000, 32000, 44100, 48000, 88200, 96000};  // :12-13
    auto common = [&](uint32_t r) { for (uint32_t k : kRates) if (k == r) return true; return false; };
    const size_t ch = audio.channel_count();
    if (ch == 0) throw Error("Channel count must be > 0");
    const int bits = audio.bits_per_sample();
    if (bits != 16 && bits != 24 && bits != 32) throw Error("Unsupported bits_per_sample: " + std::to_string(bits));
    if (sampling_rate > 0xffffffffull) throw Error("sampling_rate out of range");
    const uint32_t out_rate = (uint32_t)sampling_rate, in_rate = audio.sampling_rate();
    if (in_rate == 0 || out_rate == 0) throw Error("sampling_rate must be > 0");
    if (!common(in_rate)) throw Error("Unsupported input sample_rate: " + std::to_string(in_rate));
    if (!common(out_rate)) throw Error("Unsupported output sample_rate: " + std::to_string(out_rate));
    const int fmt = (bits == 32 && audio.audio_format() == EncodingFlag::PCMFloat) ? SK_FMT_FLOAT : SK_FMT_SIGNED;
    const size_t frames = audio.data().size() / ((size_t)(bits / 8) * ch);
    size_t n = 0;
    Gpu::check(sk_normalize_batch_plan(Gpu::ctx(), in_rate, out_rate, frames, SK_SEM_ONESHOT, &n));
    std::vector<float> flat(ch * (n ? n : 1));
    size_t got = 0;
    Gpu::check(sk_normalize_stream(Gpu::ctx(), audio.data().data(), audio.data().size(), bits, fmt, 0, (int)ch, SK_MODE_A, in_rate, out_rate,
                                   SK_SEM_ONESHOT, SK_MIX_NONE, flat.data(), n ? n : 1, &got));
    std::vector<std::vector<float>> out(ch);
    for (size_t c = 0; c < ch; c++) out[c].assign(flat.begin() + c * (n ? n : 1), flat.begin() + c * (n ? n : 1) + got);
    return out;
}
// ---- StreamingStereo48kNormalizer (soundkit/src/audio_pipeline.rs:229-394) ---------------------------------
struct Stereo48kBlock {  // :234-248
    std::vector<float> left, right;
    size_t frame_count() const { return left.size(); }
    bool is_empty() const { return left.empty(); }
};
class StreamingStereo48kNormalizer {  // :256-394; conversion + interpolation run on the device (sk_norm48k)
public:
    StreamingStereo48kNormalizer() {
        sk_norm48k* h = nullptr;
        Gpu::check(sk_norm48k_create(Gpu::ctx(), &h));
        h_.reset(h, sk_norm48k_destroy);
    }
    uint32_t source_sample_rate() const { return source_sample_rate_; }
    uint8_t source_channels() const { return source_channels_; }
    uint64_t source_frames() const { return sk_norm48k_source_frames(h_.get()); }
    uint64_t output_frames() const { return sk_norm48k_output_frames(h_.get()); }

    std::optional<Stereo48kBlock> push(const AudioData& audio) {  // :288
        if (finished_) throw Error("48 kHz stereo normalizer is already finished");
        const uint32_t sample_rate = audio.sampling_rate();
        const uint8_t channel_count = audio.channel_count();
        if (sample_rate == 0 || channel_count == 0) throw Error("decoded audio has invalid PCM geometry");
        if (source_sample_rate_ == 0) { source_sample_rate_ = sample_rate; source_channels_ = channel_count; }
        else if (source_sample_rate_ != sample_rate || source_channels_ != channel_count)
            throw Error("decoded PCM geometry changed from " + std::to_string(source_sample_rate_) + " Hz/" + std::to_string(source_channels_) +
                        " ch to " + std::to_string(sample_rate) + " Hz/" + std::to_string(channel_count) + " ch");
        // audio_to_stereo_f32, :133-174
        const int bits = audio.bits_per_sample();
        if (bits != 16 && bits != 24 && bits != 32) throw Error("unsupported PCM bit depth " + std::to_string(bits));
        const size_t frame_bytes = (size_t)(bits / 8) * channel_count;
        if (audio.data().empty() || audio.data().size() % frame_bytes) throw Error("decoded audio contained an incomplete PCM frame");
        const size_t frames = audio.data().size() / frame_bytes;
        const size_t cap = sk_norm48k_max_output(frames, sample_rate);
        Stereo48kBlock block;
        block.left.resize(cap); block.right.resize(cap);
        size_t n = 0;
        Gpu::check(sk_norm48k_push(h_.get(), audio.data().data(), audio.data().size(), bits,
                                   audio.audio_format() == EncodingFlag::PCMFloat ? SK_FMT_FLOAT : SK_FMT_SIGNED,
                                   audio.endianness() == Endianness::BigEndian, channel_count, sample_rate, block.left.data(), block.right.data(), cap, &n));
        block.left.resize(n); block.right.resize(n);
        if (block.is_empty()) return std::nullopt;
        return block;
    }
    std::optional<Stereo48kBlock> finish() {  // :366
        if (finished_) throw Error("48 kHz stereo normalizer is already finished");
        finished_ = true;
        if (source_frames() == 0 || source_sample_rate_ == 0) throw Error("decoded source contained no PCM");
        if (source_sample_rate_ == 48000) { size_t n = 0; Gpu::check(sk_norm48k_finish(h_.get(), nullptr, nullptr, 0, &n)); return std::nullopt; }
        // remaining <= round(frames * 48000 / rate) - emitted <= 48000 / rate + 5
        const size_t cap = (size_t)(48000 / source_sample_rate_) + 8;
        Stereo48kBlock block;
        block.left.resize(cap); block.right.resize(cap);
        size_t n = 0;
        Gpu::check(sk_norm48k_finish(h_.get(), block.left.data(), block.right.data(), cap, &n));
        block.left.resize(n); block.right.resize(n);
        if (block.is_empty()) return std::nullopt;
        return block;
    }
private:
    std::shared_ptr<sk_norm48k> h_;
    uint32_t source_sample_rate_ = 0;
    uint8_t source_channels_ = 0;
    bool finished_ = false;
};
}  // namespace audio_pipeline
}  // namespace soundkit

// ---- soundkit_g711 (soundkit-g711/src/lib.rs) ----------------------------------------------------------------
namespace soundkit_g711 {
using soundkit::Error;
using soundkit::Gpu;
enum class G711Law { MuLaw = 0, ALaw = 1 };  // :17

inline size_t encode_i16(G711Law law, const int16_t* input, size_t n_in, uint8_t* output, size_t n_out) {  // :36
    if (n_out < n_in) throw Error("Output buffer too small for G.711 encode: need " + std::to_string(n_in) + ", have " + std::to_string(n_out));
    Gpu::check(sk_g711_encode(Gpu::ctx(), (int)law, SK_KIND_I16, input, n_in, output));
    return n_in;
}
inline size_t decode_i16(G711Law law, const uint8_t* input, size_t n_in, int16_t* output, size_t n_out) {  // :52
    if (n_out < n_in) throw Error("Output buffer too small for G.711 decode: need " + std::to_string(n_in) + ", have " + std::to_string(n_out));
    Gpu::check(sk_g711_decode(Gpu::ctx(), (int)law, input, n_in, SK_KIND_I16, output));
    return n_in;
}
inline uint8_t encode_sample(G711Law law, int16_t sample) { uint8_t o = 0; encode_i16(law, &sample, 1, &o, 1); return o; }  // :22
inline int16_t decode_sample(G711Law law, uint8_t sample) { int16_t o = 0; decode_i16(law, &sample, 1, &o, 1); return o; }  // :29

class G711Encoder {  // :69-143, impl soundkit::audio_packet::Encoder
public:
    G711Encoder(uint32_t sample_rate, uint32_t /*bits_per_sample*/, uint32_t channels, uint32_t /*frame_size*/, uint32_t /*bitrate*/)
        : law_(G711Law::MuLaw), rate_(sample_rate), ch_((uint8_t)channels) {}  // Encoder::new defaults to mu-law (:106-114)
    static G711Encoder new_with_law(G711Law law, uint32_t r, uint8_t c) { G711Encoder e(r, 16, c, 0, 0); e.law_ = law; return e; }
    static G711Encoder new_mulaw(uint32_t r, uint8_t c) { return new_with_law(G711Law::MuLaw, r, c); }
    static G711Encoder new_alaw(uint32_t r, uint8_t c) { return new_with_law(G711Law::ALaw, r, c); }
    G711Law law() const { return law_; }
    uint32_t sample_rate() const { return rate_; }
    uint8_t channels() const { return ch_; }
    void init() {}
    void reset() {}
    size_t encode_i16(const int16_t* in, size_t n, uint8_t* out, size_t cap) { return soundkit_g711::encode_i16(law_, in, n, out, cap); }
    size_t encode_i32(const int32_t* in, size_t n, uint8_t* out, size_t cap) {  // (sample >> 16) as i16, :124-138
        if (cap < n) throw Error("Output buffer too small for G.711 encode: need " + std::to_string(n) + ", have " + std::to_string(cap));
        Gpu::check(sk_g711_encode(Gpu::ctx(), (int)law_, SK_KIND_I32, in, n, out));
        return n;
    }
private:
    G711Law law_; uint32_t rate_; uint8_t ch_;
};

class G711Decoder {  // :146-233, impl soundkit::audio_packet::Decoder (fec is ignored)
public:
    static G711Decoder new_with_law(G711Law law, uint32_t r, uint8_t c) { return G711Decoder(law, r, c); }
    static G711Decoder new_mulaw(uint32_t r, uint8_t c) { return G711Decoder(G711Law::MuLaw, r, c); }
    static G711Decoder new_alaw(uint32_t r, uint8_t c) { return G711Decoder(G711Law::ALaw, r, c); }
    G711Law law() const { return law_; }
    uint32_t sample_rate() const { return rate_; }
    uint8_t channels() const { return ch_; }
    size_t decode_i16(const uint8_t* in, size_t n, int16_t* out, size_t cap, bool /*fec*/) { return soundkit_g711::decode_i16(law_, in, n, out, cap); }
    size_t decode_i32(const uint8_t* in, size_t n, int32_t* out, size_t cap, bool /*fec*/) { return run(in, n, out, cap, SK_KIND_I32); }
    size_t decode_f32(const uint8_t* in, size_t n, float* out, size_t cap, bool /*fec*/) { return run(in, n, out, cap, SK_KIND_F32); }
private:
    G711Decoder(G711Law l, uint32_t r, uint8_t c) : law_(l), rate_(r), ch_(c) {}
    size_t run(const uint8_t* in, size_t n, void* out, size_t cap, int kind) {
        if (cap < n) throw Error("Output buffer too small for G.711 decode: need " + std::to_string(n) + ", have " + std::to_string(cap));
        Gpu::check(sk_g711_decode(Gpu::ctx(), (int)law_, in, n, kind, out));
        return n;
    }
    G711Law law_; uint32_t rate_; uint8_t ch_;
};
}  // namespace soundkit_g711

// ---- soundkit_g726 (soundkit-g726/src/lib.rs) — SURVEY.md 8(f) row 4 -----------------------------------------------
namespace soundkit_g726 {
using soundkit::Error;
using soundkit::Gpu;
constexpr uint32_t G726_SAMPLE_RATE = 8000;  // :7
constexpr uint8_t G726_CHANNELS = 1;         // :8
enum class G726Rate { Rate16000, Rate24000, Rate32000, Rate40000 };  // :62-72
enum class G726Packing { Left = 0, Right = 1 };                      // :191-197
inline G726Rate rate_from_bitrate(uint32_t b) {                      // :75-82
    return b == 16000 ? G726Rate::Rate16000 : (b == 24000 ? G726Rate::Rate24000 : (b == 40000 ? G726Rate::Rate40000 : G726Rate::Rate32000));
}
inline size_t bits_per_sample(G726Rate r) { return 2 + (size_t)r; }   // :84-91
inline uint32_t bit_rate(G726Rate r) { return 16000u + 8000u * (uint32_t)r; }
inline size_t samples_per_byte_group(G726Rate r) { static const size_t t[4] = {4, 8, 2, 8}; return t[(int)r]; }  // :104-111
inline size_t bytes_per_group(G726Rate r) { static const size_t t[4] = {1, 3, 1, 5}; return t[(int)r]; }         // :113-120

class G726Encoder {  // :598-742, impl soundkit::audio_packet::Encoder; the partial group stays here (pending_samples)
public:
    static G726Encoder try_new(G726Rate rate, G726Packing packing) { return G726Encoder(rate, packing); }
    static G726Encoder new_32k_left() { return try_new(G726Rate::Rate32000, G726Packing::Left); }
    static G726Encoder new_32k_right() { return try_new(G726Rate::Rate32000, G726Packing::Right); }
    static G726Encoder new_16k_left() { return try_new(G726Rate::Rate16000, G726Packing::Left); }
    static G726Encoder new_24k_left() { return try_new(G726Rate::Rate24000, G726Packing::Left); }
    static G726Encoder new_40k_left() { return try_new(G726Rate::Rate40000, G726Packing::Left); }
    // Encoder::new(sample_rate, bits_per_sample, channels, frame_size, bitrate), :681-692
    G726Encoder(uint32_t, uint32_t, uint32_t, uint32_t, uint32_t bitrate) : G726Encoder(rate_from_bitrate(bitrate), G726Packing::Left) {}
    G726Rate rate() const { return rate_; }
    G726Packing packing() const { return packing_; }
    void init() {}
    void reset() { sk_g726_state_init(core_); pending_.clear(); }  // :740-744
    size_t encode_i16(const int16_t* input, size_t n, uint8_t* output, size_t cap) {  // :700-733
        const size_t required = required_encode_bytes(n, false);
        if (cap < required) throw Error("Output buffer too small for G.726 encode: need " + std::to_string(required) + ", have " + std::to_string(cap));
        const size_t gs = samples_per_byte_group(rate_);
        std::vector<int16_t> samples(pending_);
        samples.insert(samples.end(), input, input + n);
        const size_t complete = samples.size() / gs * gs;
        size_t written = 0;
        if (complete) written = encode_complete(samples.data(), complete, output);
        pending_.assign(samples.begin() + complete, samples.end());
        return written;
    }
    size_t encode_i32(const int32_t* input, size_t n, uint8_t* output, size_t cap) {  // :735-738
        std::vector<int16_t> s(n);
        for (size_t i = 0; i < n; i++) s[i] = (int16_t)(input[i] >> 16);
        return encode_i16(s.data(), n, output, cap);
    }
    size_t encode_to_vec(const std::vector<int16_t>& input, std::vector<uint8_t>& output) {  // :635-642
        const size_t start = output.size();
        output.resize(start + required_encode_bytes(input.size(), false));
        const size_t w = encode_i16(input.data(), input.size(), output.data() + start, output.size() - start);
        output.resize(start + w);
        return w;
    }
    size_t flush_into(uint8_t* output, size_t cap) {  // :644-662: zero-pads the pending samples to one group
        const size_t required = required_encode_bytes(0, true);
        if (cap < required) throw Error("Output buffer too small for G.726 flush: need " + std::to_string(required) + ", have " + std::to_string(cap));
        if (pending_.empty()) return 0;
        std::vector<int16_t> samples(pending_);
        samples.resize(samples_per_byte_group(rate_), 0);
        pending_.clear();
        return encode_complete(samples.data(), samples.size(), output);
    }
    size_t flush_to_vec(std::vector<uint8_t>& output) {  // :664-671
        const size_t start = output.size();
        output.resize(start + required_encode_bytes(0, true));
        const size_t w = flush_into(output.data() + start, output.size() - start);
        output.resize(start + w);
        return w;
    }
private:
    G726Encoder(G726Rate r, G726Packing p) : rate_(r), packing_(p) { sk_g726_state_init(core_); }
    size_t required_encode_bytes(size_t input_samples, bool flush) const {  // :681-690
        const size_t gs = samples_per_byte_group(rate_), total = pending_.size() + input_samples;
        return (flush ? (total + gs - 1) / gs : total / gs) * bytes_per_group(rate_);
    }
    size_t encode_complete(const int16_t* s, size_t n, uint8_t* out) {
        Gpu::check(sk_g726_encode(Gpu::ctx(), (int)bits_per_sample(rate_), (int)packing_, core_, SK_KIND_I16, s, n, out));
        return n / samples_per_byte_group(rate_) * bytes_per_group(rate_);
    }
    G726Rate rate_; G726Packing packing_;
    int32_t core_[SK_G726_STATE_INTS];
    std::vector<int16_t> pending_;
};

class G726Decoder {  // :746-951, impl soundkit::audio_packet::Decoder (fec ignored); pending_bytes stays here
public:
    static G726Decoder try_new(G726Rate rate, G726Packing packing) { return G726Decoder(rate, packing); }
    static G726Decoder new_32k_left() { return try_new(G726Rate::Rate32000, G726Packing::Left); }
    static G726Decoder new_32k_right() { return try_new(G726Rate::Rate32000, G726Packing::Right); }
    static G726Decoder new_16k_left() { return try_new(G726Rate::Rate16000, G726Packing::Left); }
    static G726Decoder new_24k_left() { return try_new(G726Rate::Rate24000, G726Packing::Left); }
    static G726Decoder new_40k_left() { return try_new(G726Rate::Rate40000, G726Packing::Left); }
    uint32_t sample_rate() const { return G726_SAMPLE_RATE; }
    uint8_t channels() const { return G726_CHANNELS; }
    G726Rate rate() const { return rate_; }
    G726Packing packing() const { return packing_; }
    void flush() const {  // :802-811
        if (!pending_.empty()) throw Error("G.726 stream ended with " + std::to_string(pending_.size()) + " trailing partial-packet byte(s)");
    }
    size_t decode_i16(const uint8_t* in, size_t n, int16_t* out, size_t cap, bool) { return run(in, n, out, cap, SK_KIND_I16); }
    size_t decode_i32(const uint8_t* in, size_t n, int32_t* out, size_t cap, bool) { return run(in, n, out, cap, SK_KIND_I32); }
    size_t decode_f32(const uint8_t* in, size_t n, float* out, size_t cap, bool) { return run(in, n, out, cap, SK_KIND_F32); }
private:
    G726Decoder(G726Rate r, G726Packing p) : rate_(r), packing_(p) { sk_g726_state_init(core_); }
    size_t run(const uint8_t* in, size_t n, void* out, size_t cap, int kind) {  // :821-860
        const size_t gb = bytes_per_group(rate_), total = pending_.size() + n;
        const size_t complete = total / gb * gb, required = complete / gb * samples_per_byte_group(rate_);
        if (cap < required) throw Error("Output buffer too small for G.726 decode: need " + std::to_string(required) + ", have " + std::to_string(cap));
        std::vector<uint8_t> bytes(pending_);
        bytes.insert(bytes.end(), in, in + n);
        if (complete) Gpu::check(sk_g726_decode(Gpu::ctx(), (int)bits_per_sample(rate_), (int)packing_, core_, bytes.data(), complete, kind, out));
        pending_.assign(bytes.begin() + complete, bytes.end());
        return required;
    }
    G726Rate rate_; G726Packing packing_;
    int32_t core_[SK_G726_STATE_INTS];
    std::vector<uint8_t> pending_;
};
}  // namespace soundkit_g726

// ---- soundkit-decoder output-conversion stage (soundkit-decoder/src/lib.rs) ---------------------------------------
namespace soundkit_g722 {  // soundkit-g722/src/lib.rs
using soundkit::Error;
using soundkit::Gpu;
constexpr uint32_t G722_SAMPLE_RATE = 16000;  // :6
constexpr uint8_t G722_CHANNELS = 1;          // :7
inline size_t max_encoded_len(size_t input_samples) { return (input_samples + 1) / 2; }  // :234-236

class G722Encoder {  // :9-129, impl soundkit::audio_packet::Encoder; the odd pending sample stays here
public:
    static G722Encoder new_64k() { return G722Encoder(); }
    G722Encoder() { sk_g722_state_init(state_); }
    // Encoder::new(sample_rate, bits_per_sample, channels, frame_size, bitrate), :68-76: every argument ignored
    G722Encoder(uint32_t, uint32_t, uint32_t, uint32_t, uint32_t) : G722Encoder() {}
    void init() {}
    void reset() { sk_g722_state_init(state_); has_pending_ = false; }  // :124-128
    size_t encode_i16(const int16_t* input, size_t n, uint8_t* output, size_t cap) {  // :83-117
        const size_t total = n + (has_pending_ ? 1 : 0);
        const size_t complete = total - (total % 2), required = complete / 2;
        if (cap < required) throw Error("Output buffer too small for G.722 encode: need " + std::to_string(required) + ", have " + std::to_string(cap));
        if (total < 2) {
            if (!has_pending_ && n) { pending_ = input[0]; has_pending_ = true; }
            return 0;
        }
        std::vector<int16_t> samples;
        samples.reserve(complete + 1);
        if (has_pending_) { samples.push_back(pending_); has_pending_ = false; }
        samples.insert(samples.end(), input, input + n);
        if (samples.size() % 2) { pending_ = samples.back(); has_pending_ = true; samples.pop_back(); }
        Gpu::check(sk_g722_encode(Gpu::ctx(), state_, SK_KIND_I16, samples.data(), samples.size(), output));
        return samples.size() / 2;
    }
    size_t encode_i32(const int32_t* input, size_t n, uint8_t* output, size_t cap) {  // :119-122
        std::vector<int16_t> s(n);
        for (size_t i = 0; i < n; i++) s[i] = (int16_t)(input[i] >> 16);
        return encode_i16(s.data(), n, output, cap);
    }
    size_t encode_to_vec(const std::vector<int16_t>& input, std::vector<uint8_t>& output) {  // :23-31
        const size_t start = output.size();
        std::vector<uint8_t> scratch(max_encoded_len(input.size() + (has_pending_ ? 1 : 0)));
        const size_t w = encode_i16(input.data(), input.size(), scratch.data(), scratch.size());
        output.insert(output.end(), scratch.begin(), scratch.begin() + w);
        return output.size() - start;
    }
    size_t flush_into(uint8_t* output, size_t cap) {  // :33-46: the pending sample is paired with a zero
        if (!has_pending_) return 0;
        if (cap == 0) throw Error("Output buffer too small for G.722 flush: need 1, have 0");
        const int16_t pair[2] = {pending_, 0};
        has_pending_ = false;
        Gpu::check(sk_g722_encode(Gpu::ctx(), state_, SK_KIND_I16, pair, 2, output));
        return 1;
    }
    size_t flush_to_vec(std::vector<uint8_t>& output) {  // :48-54
        const size_t start = output.size();
        output.resize(start + 1, 0);
        const size_t w = flush_into(output.data() + start, 1);
        output.resize(start + w);
        return w;
    }
private:
    int32_t state_[SK_G722_STATE_INTS];
    int16_t pending_ = 0;
    bool has_pending_ = false;
};

class G722Decoder {  // :131-232, impl soundkit::audio_packet::Decoder (fec ignored)
public:
    static G722Decoder new_64k() { return G722Decoder(); }
    G722Decoder() { sk_g722_state_init(state_); }
    uint32_t sample_rate() const { return G722_SAMPLE_RATE; }
    uint8_t channels() const { return G722_CHANNELS; }
    size_t decode_i16(const uint8_t* in, size_t n, int16_t* out, size_t cap, bool) { return run(in, n, out, cap, SK_KIND_I16); }
    size_t decode_i32(const uint8_t* in, size_t n, int32_t* out, size_t cap, bool) { return run(in, n, out, cap, SK_KIND_I32); }
    size_t decode_f32(const uint8_t* in, size_t n, float* out, size_t cap, bool) { return run(in, n, out, cap, SK_KIND_F32); }
private:
    size_t run(const uint8_t* in, size_t n, void* out, size_t cap, int kind) {
        const size_t required = n * 2;
        if (cap < required) throw Error("Output buffer too small for G.722 decode: need " + std::to_string(required) + ", have " + std::to_string(cap));
        if (n) Gpu::check(sk_g722_decode(Gpu::ctx(), state_, in, n, kind, out));
        return required;
    }
    int32_t state_[SK_G722_STATE_INTS];
};
}  // namespace soundkit_g722

namespace soundkit_decoder {
using soundkit::AudioData;
using soundkit::EncodingFlag;
using soundkit::Endianness;
using soundkit::Gpu;
struct DecodeError : soundkit::Error { using soundkit::Error::Error; };  // DecodeError::DecodingFailed(String)

constexpr size_t RESAMPLE_CHUNK_SIZE = 4096;  // :79

struct DecodeOptions {  // :146-151
    std::optional<uint8_t> output_bits_per_sample;
    std::optional<uint32_t> output_sample_rate;
    std::optional<uint8_t> output_channels;
};

using Channels = std::vector<std::vector<float>>;

inline std::vector<float> flatten(const Channels& ch, size_t frames) {
    std::vector<float> p(ch.size() * frames);
    for (size_t c = 0; c < ch.size(); c++) std::memcpy(p.data() + c * frames, ch[c].data(), frames * 4);
    return p;
}

// audio_data_to_f32_channels, :3563-3617 (scale convention B)
inline Channels audio_data_to_f32_channels(const AudioData& a) {
    const size_t ch = a.channel_count();
    if (ch == 0) throw soundkit::Error("Channel count must be > 0");
    const int bits = a.bits_per_sample();
    const size_t bps = (size_t)(bits + 7) / 8;
    if (!(bits == 16 || bits == 24 || bits == 32) || a.data().size() % bps != 0 || (a.data().size() / bps) % ch != 0)
        throw soundkit::Error("PCM data is unsupported or contains a partial frame");
    if (a.audio_format() == EncodingFlag::PCMFloat && bits != 32) throw soundkit::Error("floating-point PCM must contain 32-bit samples");
    const size_t frames = a.data().size() / bps / ch;
    std::vector<float> flat(ch * frames);
    size_t got = 0;
    Gpu::check(sk_pcm_to_planar_f32(Gpu::ctx(), a.data().data(), a.data().size(), bits, a.audio_format() == EncodingFlag::PCMFloat ? SK_FMT_FLOAT : SK_FMT_SIGNED,
                                    a.endianness() == Endianness::BigEndian, (int)ch, SK_MODE_B, flat.data(), frames, &got));
    return soundkit::audio_bytes::detail::split_planes(flat, ch, frames);
}

// downmix_channels, :3492-3561
inline Channels downmix_channels(const Channels& channels, uint8_t target_channels) {
    if (channels.empty() || target_channels == 0) return {};
    const size_t n = channels[0].size();
    if (target_channels == 1) {
        std::vector<float> out(n), planar = flatten(channels, n);
        Gpu::check(sk_mixdown_mono(Gpu::ctx(), planar.data(), n, (int)channels.size(), SK_MIX_SCALE_THEN_SUM, out.data()));
        return {out};
    }
    if (target_channels == 2 && channels.size() > 2) {
        std::vector<float> out(2 * n), planar = flatten(channels, n);
        Gpu::check(sk_downmix_stereo(Gpu::ctx(), planar.data(), n, (int)channels.size(), out.data()));
        return soundkit::audio_bytes::detail::split_planes(out, 2, n);
    }
    return Channels(channels.begin(), channels.begin() + target_channels);  // first N channels
}

// f32_channels_to_bytes, :3619-3683
inline std::vector<uint8_t> f32_channels_to_bytes(const Channels& channels, uint8_t bits_per_sample, EncodingFlag output_format) {
    if (channels.empty()) return {};
    const size_t n = channels[0].size();
    for (auto& c : channels)
        if (c.size() != n) throw soundkit::Error("Channel length mismatch");
    if (output_format == EncodingFlag::PCMFloat && bits_per_sample != 32) throw soundkit::Error("PCMFloat output requires 32-bit samples");
    if (output_format != EncodingFlag::PCMFloat && bits_per_sample != 16 && bits_per_sample != 24 && bits_per_sample != 32)
        throw soundkit::Error("Unsupported output bits per sample: " + std::to_string(bits_per_sample));
    std::vector<uint8_t> out(n * channels.size() * (bits_per_sample / 8));
    std::vector<float> planar = flatten(channels, n);
    Gpu::check(sk_planar_f32_to_pcm(Gpu::ctx(), planar.data(), n, (int)channels.size(), bits_per_sample,
                                    output_format == EncodingFlag::PCMFloat ? SK_FMT_FLOAT : SK_FMT_SIGNED, out.data()));
    return out;
}

// exact_signed_pcm_to_i16, :3458-3489
inline AudioData exact_signed_pcm_to_i16(const AudioData& audio) {
    const size_t bps = audio.bits_per_sample() / 8;
    if (audio.data().size() % bps != 0) throw DecodeError("integer PCM contains a partial sample");
    std::vector<uint8_t> out(audio.data().size() / bps * 2);
    Gpu::check(sk_pcm_narrow_i16(Gpu::ctx(), audio.data().data(), audio.data().size(), audio.bits_per_sample(),
                                 audio.endianness() == Endianness::BigEndian, out.data()));
    return AudioData(16, audio.channel_count(), audio.sampling_rate(), std::move(out), EncodingFlag::PCMSigned, Endianness::LittleEndian);
}

// float_sample_to_i16, :1815-1827 — one sample through the 16-bit output kernel
inline int16_t float_sample_to_i16(float sample) {
    auto b = f32_channels_to_bytes({{sample}}, 16, EncodingFlag::PCMSigned);
    return (int16_t)(b[0] | (b[1] << 8));
}

// StreamingResampler, :1917-2060: rubato SincFixedIn(chunk 4096) state lives on the device (sk_resampler)
class StreamingResampler {
public:
    StreamingResampler(uint32_t input_sample_rate, uint32_t output_sample_rate, size_t channels, uint8_t target_bits_per_sample, uint8_t target_channels,
                       EncodingFlag output_format)
        : chunk_size(RESAMPLE_CHUNK_SIZE), channels(channels), input_sample_rate(input_sample_rate), output_sample_rate(output_sample_rate),
          target_bits_per_sample(target_bits_per_sample), target_channels(target_channels), output_format(output_format), accum_(channels) {
        sk_resampler* r = nullptr;
        if (sk_resampler_create(Gpu::ctx(), input_sample_rate, output_sample_rate, (int)channels, RESAMPLE_CHUNK_SIZE, nullptr, &r) != SK_OK)
            throw soundkit::Error(std::string("Failed to create resampler: ") + sk_last_error(Gpu::ctx()));
        r_.reset(r, sk_resampler_destroy);
        out_max_ = sk_resampler_out_max(r);
    }
    std::vector<Channels> process(const Channels& input) {  // :1970
        if (input.size() != channels)
            throw soundkit::Error("Channel count changed mid-stream: expected " + std::to_string(channels) + ", got " + std::to_string(input.size()));
        for (size_t c = 0; c < input.size(); c++) {
            if (input[c].size() != input[0].size()) throw soundkit::Error("Channel sample counts changed mid-stream");
            accum_[c].insert(accum_[c].end(), input[c].begin(), input[c].end());
        }
        std::vector<Channels> outputs;
        while (accum_[0].size() - accum_start_ >= chunk_size) {
            Channels res = run(accum_start_, chunk_size, false);
            accum_start_ += chunk_size;
            if (!res.empty() && !res[0].empty()) outputs.push_back(std::move(res));
        }
        if (accum_start_ >= chunk_size * 8 && accum_start_ * 2 >= accum_[0].size()) {
            for (auto& c : accum_) c.erase(c.begin(), c.begin() + accum_start_);
            accum_start_ = 0;
        }
        return outputs;
    }
    std::vector<Channels> flush() {  // :2017
        std::vector<Channels> outputs;
        const size_t remaining = accum_[0].size() - accum_start_;
        if (remaining > 0) {
            const size_t padded = chunk_size - remaining;
            Channels res = run(accum_start_, remaining, true);
            if (padded > 0) {
                const size_t trim = (size_t)std::llround(((double)padded * (double)output_sample_rate) / (double)input_sample_rate);
                for (auto& c : res) c.resize(c.size() > trim ? c.size() - trim : 0);
            }
            if (!res.empty() && !res[0].empty()) outputs.push_back(std::move(res));
            for (auto& c : accum_) c.clear();
            accum_start_ = 0;
        } else {
            Channels res = run(0, 0, true);
            if (!res.empty() && !res[0].empty()) outputs.push_back(std::move(res));
        }
        return outputs;
    }
    size_t chunk_size, channels;
    uint32_t input_sample_rate, output_sample_rate;
    uint8_t target_bits_per_sample, target_channels;
    EncodingFlag output_format;
private:
    Channels run(size_t start, size_t frames, bool partial) {
        std::vector<const float*> ip(channels);
        Channels out(channels, std::vector<float>(out_max_));
        std::vector<float*> op(channels);
        for (size_t c = 0; c < channels; c++) { ip[c] = accum_[c].data() + start; op[c] = out[c].data(); }
        size_t n = 0;
        int rc = partial ? sk_resampler_process_partial(r_.get(), frames ? ip.data() : nullptr, frames, op.data(), &n)
                         : sk_resampler_process(r_.get(), ip.data(), op.data(), &n);
        if (rc != SK_OK) throw soundkit::Error(std::string(partial ? "Resample partial failed: " : "Resample failed: ") + sk_last_error(Gpu::ctx()));
        for (auto& c : out) c.resize(n);
        return out;
    }
    std::shared_ptr<sk_resampler> r_;
    size_t out_max_ = 0;
    Channels accum_;
    size_t accum_start_ = 0;
};

// emit_resampled_chunks, :3261-3290
inline std::vector<AudioData> emit_resampled_chunks(std::vector<Channels> chunks, uint8_t target_bits, uint8_t target_channels, uint32_t target_rate,
                                                    EncodingFlag output_format) {
    std::vector<AudioData> out;
    for (auto& channels : chunks) {
        uint8_t count = (uint8_t)channels.size();
        if (target_channels < (uint8_t)channels.size()) { channels = downmix_channels(channels, target_channels); count = target_channels; }
        std::vector<uint8_t> bytes;
        try { bytes = f32_channels_to_bytes(channels, target_bits, output_format); }
        catch (const soundkit::Error& e) { throw DecodeError(std::string("Output conversion failed: ") + e.what()); }
        out.emplace_back(target_bits, count, target_rate, std::move(bytes), output_format, Endianness::LittleEndian);
    }
    return out;
}

// apply_output_options, :3324-3456
inline std::vector<AudioData> apply_output_options(AudioData audio_data, const DecodeOptions& options, std::optional<StreamingResampler>& resampler) {
    const uint32_t target_rate = options.output_sample_rate.value_or(audio_data.sampling_rate());
    const uint8_t target_bits = options.output_bits_per_sample.value_or(audio_data.bits_per_sample());
    const uint8_t target_channels = options.output_channels.value_or(audio_data.channel_count());
    if (target_rate == audio_data.sampling_rate() && target_bits == audio_data.bits_per_sample() && target_channels == audio_data.channel_count())
        return {std::move(audio_data)};  // fast path
    if (target_rate == audio_data.sampling_rate() && target_channels == audio_data.channel_count() && target_bits == 16 &&
        audio_data.audio_format() == EncodingFlag::PCMSigned && (audio_data.bits_per_sample() == 24 || audio_data.bits_per_sample() == 32))
        return {exact_signed_pcm_to_i16(audio_data)};
    if (target_rate == 0) throw DecodeError("Output sample rate must be > 0");
    if (target_bits != 16 && target_bits != 24 && target_bits != 32) throw DecodeError("Unsupported output bits per sample: " + std::to_string(target_bits));
    if (target_channels == 0) throw DecodeError("Output channels must be > 0");
    const EncodingFlag output_format =
        (target_bits == 32 && audio_data.audio_format() == EncodingFlag::PCMFloat) ? EncodingFlag::PCMFloat : EncodingFlag::PCMSigned;
    auto to_f32 = [&]() {
        try { return audio_data_to_f32_channels(audio_data); }
        catch (const soundkit::Error& e) { throw DecodeError(std::string("Output conversion failed: ") + e.what()); }
    };
    if (target_rate != audio_data.sampling_rate()) {
        if (audio_data.sampling_rate() == 0) throw DecodeError("Input sample rate must be > 0");
        Channels input = to_f32();
        if (resampler) {
            if (resampler->input_sample_rate != audio_data.sampling_rate() || resampler->channels != input.size() ||
                resampler->output_sample_rate != target_rate)
                throw DecodeError("Resampler configuration changed mid-stream");
        } else {
            try { resampler.emplace(audio_data.sampling_rate(), target_rate, input.size(), target_bits, target_channels, output_format); }
            catch (const soundkit::Error& e) { throw DecodeError(e.what()); }
        }
        std::vector<Channels> pending;
        try { pending = resampler->process(input); }
        catch (const soundkit::Error& e) { throw DecodeError(e.what()); }
        return emit_resampled_chunks(std::move(pending), target_bits, target_channels, target_rate, resampler->output_format);
    }
    Channels channels = to_f32();
    uint8_t count = (uint8_t)channels.size();
    if (target_channels < (uint8_t)channels.size()) { channels = downmix_channels(channels, target_channels); count = target_channels; }
    std::vector<uint8_t> bytes;
    try { bytes = f32_channels_to_bytes(channels, target_bits, output_format); }
    catch (const soundkit::Error& e) { throw DecodeError(std::string("Output conversion failed: ") + e.what()); }
    return {AudioData(target_bits, count, target_rate, std::move(bytes), output_format, Endianness::LittleEndian)};
}

// flush_resampler_frames, :3292-3305
inline std::vector<AudioData> flush_resampler_frames(StreamingResampler& resampler) {
    std::vector<Channels> chunks;
    try { chunks = resampler.flush(); }
    catch (const soundkit::Error& e) { throw DecodeError(std::string("Resampler flush failed: ") + e.what()); }
    return emit_resampled_chunks(std::move(chunks), resampler.target_bits_per_sample, resampler.target_channels, resampler.output_sample_rate,
                                 resampler.output_format);
}
}  // namespace soundkit_decoder
