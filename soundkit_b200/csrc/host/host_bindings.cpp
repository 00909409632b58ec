// host_bindings.cpp — pybind11 view of soundkit_host.hpp so pytest can drive the reference-shaped
// C++ host layer (same names / errors as the Rust API).  Test plumbing only: no arithmetic here.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "soundkit_host.hpp"

namespace py = pybind11;
using namespace soundkit;

static std::vector<uint8_t> to_bytes(const py::buffer& b) {
    py::buffer_info info = b.request();
    const uint8_t* p = static_cast<const uint8_t*>(info.ptr);
    return std::vector<uint8_t>(p, p + info.size * info.itemsize);
}
template <typename T>
static py::array_t<T> arr(const std::vector<T>& v) { return py::array_t<T>(v.size(), v.data()); }
template <typename T>
static py::list arrs(const std::vector<std::vector<T>>& v) {
    py::list out;
    for (auto& c : v) out.append(arr(c));
    return out;
}
template <typename T>
static std::vector<T> from_arr(const py::array_t<T, py::array::c_style | py::array::forcecast>& a) {
    return std::vector<T>(a.data(), a.data() + a.size());
}
static soundkit_decoder::Channels channels_from(const py::list& l) {
    soundkit_decoder::Channels out;
    for (auto item : l) out.push_back(from_arr<float>(py::cast<py::array_t<float, py::array::c_style | py::array::forcecast>>(item)));
    return out;
}

PYBIND11_MODULE(_host, m) {
    m.doc() = "C++ host mirror of soundkit's Rust interface for the normalization path (GPU arithmetic via libsoundkit_cuda.so)";
    // translators run in reverse registration order: the more specific DecodeError is registered last
    auto& ex_error = py::register_exception<Error>(m, "SoundkitError", PyExc_RuntimeError);
    py::register_exception<Panic>(m, "SoundkitPanic", PyExc_AssertionError);
    py::register_exception<soundkit_decoder::DecodeError>(m, "DecodeError", ex_error.ptr());

    py::enum_<EncodingFlag>(m, "EncodingFlag").value("PCMSigned", EncodingFlag::PCMSigned).value("PCMFloat", EncodingFlag::PCMFloat)
        .value("Opus", EncodingFlag::Opus).value("FLAC", EncodingFlag::FLAC).value("AAC", EncodingFlag::AAC);
    py::enum_<Endianness>(m, "Endianness").value("LittleEndian", Endianness::LittleEndian).value("BigEndian", Endianness::BigEndian);

    py::class_<AudioData>(m, "AudioData")
        .def(py::init([](uint8_t bits, uint8_t ch, uint32_t rate, const py::buffer& data, EncodingFlag f, Endianness e) {
            return AudioData(bits, ch, rate, to_bytes(data), f, e);
        }))
        .def("bits_per_sample", &AudioData::bits_per_sample)
        .def("channel_count", &AudioData::channel_count)
        .def("sampling_rate", &AudioData::sampling_rate)
        .def("data", [](const AudioData& a) { return py::bytes((const char*)a.data().data(), a.data().size()); })
        .def("audio_format", &AudioData::audio_format)
        .def("endianness", &AudioData::endianness);

    // ---- audio_bytes
    auto ab = m.def_submodule("audio_bytes");
#define FLAT(name, T) ab.def(#name, [](const py::buffer& b) { return arr<T>(audio_bytes::name(to_bytes(b))); })
    FLAT(i16le_to_f32, float); FLAT(i16le_to_i16, int16_t); FLAT(s24le_to_i32, int32_t); FLAT(s24le_to_i16, int16_t);
    FLAT(s24be_to_i16, int16_t); FLAT(s32le_to_i32, int32_t); FLAT(s32be_to_i32, int32_t); FLAT(s32le_to_s24, int32_t);
    FLAT(s32be_to_s24, int32_t); FLAT(s32le_to_f32, float); FLAT(s32be_to_f32, float); FLAT(s32le_to_i16, int16_t);
    FLAT(s32be_to_i16, int16_t); FLAT(f32le_to_i16, int16_t); FLAT(f32be_to_i16, int16_t); FLAT(f32le_to_i32, int32_t);
    FLAT(f32le_to_s24, int32_t); FLAT(s16be_to_i16, int16_t); FLAT(s16le_to_i16, int16_t); FLAT(s16le_to_i32, int32_t);
#undef FLAT
    ab.def("i16_to_i16le", [](const std::vector<int16_t>& d) { auto v = audio_bytes::i16_to_i16le(d); return py::bytes((const char*)v.data(), v.size()); });
    ab.def("interleave_vecs_i16", [](const std::vector<std::vector<int16_t>>& c) { auto v = audio_bytes::interleave_vecs_i16(c); return py::bytes((const char*)v.data(), v.size()); });
    ab.def("deinterleave_vecs_i16", [](const py::buffer& b, size_t ch) { return arrs(audio_bytes::deinterleave_vecs_i16(to_bytes(b), ch)); });
    ab.def("deinterleave_vecs_s24", [](const py::buffer& b, size_t ch) { return arrs(audio_bytes::deinterleave_vecs_s24(to_bytes(b), ch)); });
    ab.def("deinterleave_vecs_f32", [](const py::buffer& b, size_t ch) { return arrs(audio_bytes::deinterleave_vecs_f32(to_bytes(b), ch)); });
    ab.def("s24le_to_i32_sample", [](const std::vector<uint8_t>& b) { uint8_t a[3] = {b.at(0), b.at(1), b.at(2)}; return audio_bytes::s24le_to_i32_sample(a); });
    ab.def("stereo_to_mono_take_left", [](const std::vector<int16_t>& s) { return arr(audio_bytes::stereo_to_mono_take_left(s)); });
    ab.def("stereo_to_mono_avg", [](const std::vector<int16_t>& s) { return arr(audio_bytes::stereo_to_mono_avg(s)); });
    ab.def("stereo_to_mono_inplace_take_left", [](std::vector<int16_t> s) { size_t n = audio_bytes::stereo_to_mono_inplace_take_left(s); s.resize(n); return arr(s); });
    ab.def("stereo_to_mono_inplace_avg", [](std::vector<int16_t> s) { size_t n = audio_bytes::stereo_to_mono_inplace_avg(s); s.resize(n); return arr(s); });

    // ---- audio_pipeline
    auto ap = m.def_submodule("audio_pipeline");
    ap.def("vec_f32_to_i16", [](const std::vector<float>& v) { return arr(audio_pipeline::vec_f32_to_i16(v)); });
    ap.def("vec_i16_to_f32", [](const std::vector<int16_t>& v) { return arr(audio_pipeline::vec_i16_to_f32(v)); });
    ap.def("vec_i32_to_f32", [](const std::vector<int32_t>& v) { return arr(audio_pipeline::vec_i32_to_f32(v)); });
    ap.def("deserialize_audio", [](const py::buffer& b, uint8_t bits, uint8_t ch) -> py::object {
        PcmData d = audio_pipeline::deserialize_audio(to_bytes(b), bits, ch);
        if (auto* p = std::get_if<0>(&d)) return py::make_tuple("I16", arrs(*p));
        if (auto* p = std::get_if<1>(&d)) return py::make_tuple("I32", arrs(*p));
        return py::make_tuple("F32", arrs(std::get<2>(d)));
    });
    ap.def("audio_to_f32_channels", [](const AudioData& a) { return arrs(audio_pipeline::audio_to_f32_channels(a)); });
    ap.def("audio_to_mono_f32", [](const AudioData& a) { return arr(audio_pipeline::audio_to_mono_f32(a)); });
    ap.def("mixdown_to_mono_f32", [](const py::list& c) { return arr(audio_pipeline::mixdown_to_mono_f32(channels_from(c))); });
    ap.def("f32s_to_le_bytes", [](const std::vector<float>& v) { auto b = audio_pipeline::f32s_to_le_bytes(v); return py::bytes((const char*)b.data(), b.size()); });
    ap.def("f32s_from_le_bytes", [](const py::buffer& b) { return arr(audio_pipeline::f32s_from_le_bytes(to_bytes(b))); });
    ap.def("downsample_audio", [](const AudioData& a, size_t rate) { return arrs(audio_pipeline::downsample_audio(a, rate)); });

    py::class_<audio_pipeline::StreamingStereo48kNormalizer>(ap, "StreamingStereo48kNormalizer")
        .def(py::init<>())
        .def("source_sample_rate", &audio_pipeline::StreamingStereo48kNormalizer::source_sample_rate)
        .def("source_channels", &audio_pipeline::StreamingStereo48kNormalizer::source_channels)
        .def("source_frames", &audio_pipeline::StreamingStereo48kNormalizer::source_frames)
        .def("output_frames", &audio_pipeline::StreamingStereo48kNormalizer::output_frames)
        // Option<Stereo48kBlock> -> None | (left, right)
        .def("push", [](audio_pipeline::StreamingStereo48kNormalizer& n, const AudioData& a) -> py::object {
            auto b = n.push(a);
            if (!b) return py::none();
            return py::make_tuple(arr(b->left), arr(b->right));
        })
        .def("finish", [](audio_pipeline::StreamingStereo48kNormalizer& n) -> py::object {
            auto b = n.finish();
            if (!b) return py::none();
            return py::make_tuple(arr(b->left), arr(b->right));
        });

    // ---- raw_pcm
    auto rp = m.def_submodule("raw_pcm");
    rp.attr("MAX_INPUT_CHUNK_BYTES") = raw_pcm::MAX_INPUT_CHUNK_BYTES;
    py::enum_<raw_pcm::RawPcmSampleFormat>(rp, "RawPcmSampleFormat").value("I16", raw_pcm::RawPcmSampleFormat::I16)
        .value("I24", raw_pcm::RawPcmSampleFormat::I24).value("I32", raw_pcm::RawPcmSampleFormat::I32).value("F32", raw_pcm::RawPcmSampleFormat::F32);
    py::class_<raw_pcm::RawPcmFormat>(rp, "RawPcmFormat")
        .def_static("new", &raw_pcm::RawPcmFormat::make)
        .def_static("linear16", &raw_pcm::RawPcmFormat::linear16)
        .def_static("l16", &raw_pcm::RawPcmFormat::l16)
        .def_static("linear32", &raw_pcm::RawPcmFormat::linear32)
        .def("bytes_per_frame", &raw_pcm::RawPcmFormat::bytes_per_frame)
        .def("sample_rate", &raw_pcm::RawPcmFormat::sample_rate)
        .def("channels", &raw_pcm::RawPcmFormat::channels)
        .def("sample_format", &raw_pcm::RawPcmFormat::sample_format)
        .def("endianness", &raw_pcm::RawPcmFormat::endianness)
        .def("bits_per_sample", &raw_pcm::RawPcmFormat::bits_per_sample)
        .def("encoding_flag", &raw_pcm::RawPcmFormat::encoding_flag);
    py::class_<raw_pcm::RawPcmStreamProcessor>(rp, "RawPcmStreamProcessor")
        .def(py::init<raw_pcm::RawPcmFormat>())
        .def("format", &raw_pcm::RawPcmStreamProcessor::format)
        .def("buffered_len", &raw_pcm::RawPcmStreamProcessor::buffered_len)
        .def("add", [](raw_pcm::RawPcmStreamProcessor& p, const py::buffer& b) { return p.add(to_bytes(b)); })
        .def("flush", &raw_pcm::RawPcmStreamProcessor::flush);

    // ---- soundkit_g711
    auto g = m.def_submodule("g711");
    using namespace soundkit_g711;
    py::enum_<G711Law>(g, "G711Law").value("MuLaw", G711Law::MuLaw).value("ALaw", G711Law::ALaw);
    g.def("encode_sample", &encode_sample);
    g.def("decode_sample", &decode_sample);
    // slice functions: `out_len` plays the role of output.len(); returns (written, output array)
    g.def("encode_i16", [](G711Law law, const std::vector<int16_t>& in, size_t out_len) {
        std::vector<uint8_t> out(out_len); size_t n = encode_i16(law, in.data(), in.size(), out.data(), out.size()); return py::make_tuple(n, arr(out)); });
    g.def("decode_i16", [](G711Law law, const py::buffer& in, size_t out_len) {
        auto b = to_bytes(in); std::vector<int16_t> out(out_len); size_t n = decode_i16(law, b.data(), b.size(), out.data(), out.size()); return py::make_tuple(n, arr(out)); });
    py::class_<G711Encoder>(g, "G711Encoder")
        .def(py::init<uint32_t, uint32_t, uint32_t, uint32_t, uint32_t>())
        .def_static("new_with_law", &G711Encoder::new_with_law).def_static("new_mulaw", &G711Encoder::new_mulaw).def_static("new_alaw", &G711Encoder::new_alaw)
        .def("law", &G711Encoder::law).def("sample_rate", &G711Encoder::sample_rate).def("channels", &G711Encoder::channels)
        .def("init", &G711Encoder::init).def("reset", &G711Encoder::reset)
        .def("encode_i16", [](G711Encoder& e, const std::vector<int16_t>& in, size_t out_len) {
            std::vector<uint8_t> out(out_len); size_t n = e.encode_i16(in.data(), in.size(), out.data(), out.size()); return py::make_tuple(n, arr(out)); })
        .def("encode_i32", [](G711Encoder& e, const std::vector<int32_t>& in, size_t out_len) {
            std::vector<uint8_t> out(out_len); size_t n = e.encode_i32(in.data(), in.size(), out.data(), out.size()); return py::make_tuple(n, arr(out)); });
    py::class_<G711Decoder>(g, "G711Decoder")
        .def_static("new_with_law", &G711Decoder::new_with_law).def_static("new_mulaw", &G711Decoder::new_mulaw).def_static("new_alaw", &G711Decoder::new_alaw)
        .def("law", &G711Decoder::law).def("sample_rate", &G711Decoder::sample_rate).def("channels", &G711Decoder::channels)
        .def("decode_i16", [](G711Decoder& d, const py::buffer& in, size_t out_len, bool fec) {
            auto b = to_bytes(in); std::vector<int16_t> out(out_len); size_t n = d.decode_i16(b.data(), b.size(), out.data(), out.size(), fec); return py::make_tuple(n, arr(out)); })
        .def("decode_i32", [](G711Decoder& d, const py::buffer& in, size_t out_len, bool fec) {
            auto b = to_bytes(in); std::vector<int32_t> out(out_len); size_t n = d.decode_i32(b.data(), b.size(), out.data(), out.size(), fec); return py::make_tuple(n, arr(out)); })
        .def("decode_f32", [](G711Decoder& d, const py::buffer& in, size_t out_len, bool fec) {
            auto b = to_bytes(in); std::vector<float> out(out_len); size_t n = d.decode_f32(b.data(), b.size(), out.data(), out.size(), fec); return py::make_tuple(n, arr(out)); });

    // ---- soundkit_g726
    {
        auto g6 = m.def_submodule("g726");
        using namespace soundkit_g726;
        py::enum_<G726Rate>(g6, "G726Rate").value("Rate16000", G726Rate::Rate16000).value("Rate24000", G726Rate::Rate24000)
            .value("Rate32000", G726Rate::Rate32000).value("Rate40000", G726Rate::Rate40000);
        py::enum_<G726Packing>(g6, "G726Packing").value("Left", G726Packing::Left).value("Right", G726Packing::Right);
        g6.def("samples_per_byte_group", &samples_per_byte_group);
        g6.def("bytes_per_group", &bytes_per_group);
        g6.def("from_bitrate", &rate_from_bitrate);
        py::class_<G726Encoder>(g6, "G726Encoder")
            .def(py::init<uint32_t, uint32_t, uint32_t, uint32_t, uint32_t>())
            .def_static("try_new", &G726Encoder::try_new)
            .def("rate", &G726Encoder::rate).def("packing", &G726Encoder::packing).def("reset", &G726Encoder::reset)
            .def("encode_i16", [](G726Encoder& e, const std::vector<int16_t>& in, size_t cap) {
                std::vector<uint8_t> out(cap); size_t n = e.encode_i16(in.data(), in.size(), out.data(), cap); out.resize(n);
                return py::bytes((const char*)out.data(), out.size()); })
            .def("encode_i32", [](G726Encoder& e, const std::vector<int32_t>& in, size_t cap) {
                std::vector<uint8_t> out(cap); size_t n = e.encode_i32(in.data(), in.size(), out.data(), cap); out.resize(n);
                return py::bytes((const char*)out.data(), out.size()); })
            .def("encode_to_vec", [](G726Encoder& e, const std::vector<int16_t>& in) {
                std::vector<uint8_t> out; e.encode_to_vec(in, out); return py::bytes((const char*)out.data(), out.size()); })
            .def("flush_to_vec", [](G726Encoder& e) { std::vector<uint8_t> out; e.flush_to_vec(out); return py::bytes((const char*)out.data(), out.size()); })
            .def("flush_into", [](G726Encoder& e, size_t cap) { std::vector<uint8_t> out(cap); size_t n = e.flush_into(out.data(), cap); out.resize(n);
                return py::bytes((const char*)out.data(), out.size()); });
        py::class_<G726Decoder>(g6, "G726Decoder")
            .def_static("try_new", &G726Decoder::try_new)
            .def_static("new_32k_left", &G726Decoder::new_32k_left).def_static("new_32k_right", &G726Decoder::new_32k_right)
            .def("sample_rate", &G726Decoder::sample_rate).def("channels", &G726Decoder::channels).def("flush", &G726Decoder::flush)
            .def("decode_i16", [](G726Decoder& d, const py::buffer& in, size_t cap, bool fec) {
                auto b = to_bytes(in); std::vector<int16_t> out(cap); size_t n = d.decode_i16(b.data(), b.size(), out.data(), cap, fec); out.resize(n); return arr(out); })
            .def("decode_i32", [](G726Decoder& d, const py::buffer& in, size_t cap, bool fec) {
                auto b = to_bytes(in); std::vector<int32_t> out(cap); size_t n = d.decode_i32(b.data(), b.size(), out.data(), cap, fec); out.resize(n); return arr(out); })
            .def("decode_f32", [](G726Decoder& d, const py::buffer& in, size_t cap, bool fec) {
                auto b = to_bytes(in); std::vector<float> out(cap); size_t n = d.decode_f32(b.data(), b.size(), out.data(), cap, fec); out.resize(n); return arr(out); });
    }

    // ---- soundkit_g722
    {
        auto g2 = m.def_submodule("g722");
        using namespace soundkit_g722;
        g2.attr("G722_SAMPLE_RATE") = G722_SAMPLE_RATE;
        g2.attr("G722_CHANNELS") = G722_CHANNELS;
        py::class_<G722Encoder>(g2, "G722Encoder")
            .def(py::init<>())
            .def(py::init<uint32_t, uint32_t, uint32_t, uint32_t, uint32_t>())
            .def_static("new_64k", &G722Encoder::new_64k)
            .def("reset", &G722Encoder::reset)
            .def("encode_i16", [](G722Encoder& e, const std::vector<int16_t>& in, size_t cap) {
                std::vector<uint8_t> out(cap); size_t n = e.encode_i16(in.data(), in.size(), out.data(), cap); out.resize(n);
                return py::bytes((const char*)out.data(), out.size()); })
            .def("encode_i32", [](G722Encoder& e, const std::vector<int32_t>& in, size_t cap) {
                std::vector<uint8_t> out(cap); size_t n = e.encode_i32(in.data(), in.size(), out.data(), cap); out.resize(n);
                return py::bytes((const char*)out.data(), out.size()); })
            .def("encode_to_vec", [](G722Encoder& e, const std::vector<int16_t>& in) {
                std::vector<uint8_t> out; e.encode_to_vec(in, out); return py::bytes((const char*)out.data(), out.size()); })
            .def("flush_to_vec", [](G722Encoder& e) { std::vector<uint8_t> out; e.flush_to_vec(out); return py::bytes((const char*)out.data(), out.size()); })
            .def("flush_into", [](G722Encoder& e, size_t cap) { std::vector<uint8_t> out(cap); size_t n = e.flush_into(out.data(), cap); out.resize(n);
                return py::bytes((const char*)out.data(), out.size()); });
        py::class_<G722Decoder>(g2, "G722Decoder")
            .def(py::init<>())
            .def_static("new_64k", &G722Decoder::new_64k)
            .def("sample_rate", &G722Decoder::sample_rate).def("channels", &G722Decoder::channels)
            .def("decode_i16", [](G722Decoder& d, const py::buffer& in, size_t cap, bool fec) {
                auto b = to_bytes(in); std::vector<int16_t> out(cap); size_t n = d.decode_i16(b.data(), b.size(), out.data(), cap, fec); out.resize(n); return arr(out); })
            .def("decode_i32", [](G722Decoder& d, const py::buffer& in, size_t cap, bool fec) {
                auto b = to_bytes(in); std::vector<int32_t> out(cap); size_t n = d.decode_i32(b.data(), b.size(), out.data(), cap, fec); out.resize(n); return arr(out); })
            .def("decode_f32", [](G722Decoder& d, const py::buffer& in, size_t cap, bool fec) {
                auto b = to_bytes(in); std::vector<float> out(cap); size_t n = d.decode_f32(b.data(), b.size(), out.data(), cap, fec); out.resize(n); return arr(out); });
    }

    // ---- soundkit_decoder output stage
    auto d = m.def_submodule("decoder");
    using namespace soundkit_decoder;
    d.attr("RESAMPLE_CHUNK_SIZE") = RESAMPLE_CHUNK_SIZE;
    py::class_<DecodeOptions>(d, "DecodeOptions")
        .def(py::init([](std::optional<uint8_t> bits, std::optional<uint32_t> rate, std::optional<uint8_t> ch) { return DecodeOptions{bits, rate, ch}; }),
             py::arg("output_bits_per_sample") = py::none(), py::arg("output_sample_rate") = py::none(), py::arg("output_channels") = py::none());
    d.def("audio_data_to_f32_channels", [](const AudioData& a) { return arrs(audio_data_to_f32_channels(a)); });
    d.def("downmix_channels", [](const py::list& c, uint8_t t) { return arrs(downmix_channels(channels_from(c), t)); });
    d.def("f32_channels_to_bytes", [](const py::list& c, uint8_t bits, EncodingFlag f) { auto b = f32_channels_to_bytes(channels_from(c), bits, f); return py::bytes((const char*)b.data(), b.size()); });
    d.def("exact_signed_pcm_to_i16", &exact_signed_pcm_to_i16);
    d.def("float_sample_to_i16", &float_sample_to_i16);
    py::class_<StreamingResampler>(d, "StreamingResampler")
        .def(py::init<uint32_t, uint32_t, size_t, uint8_t, uint8_t, EncodingFlag>())
        .def("process", [](StreamingResampler& r, const py::list& c) { py::list out; for (auto& ch : r.process(channels_from(c))) out.append(arrs(ch)); return out; })
        .def("flush", [](StreamingResampler& r) { py::list out; for (auto& ch : r.flush()) out.append(arrs(ch)); return out; });
    // the pipeline's `Option<StreamingResampler>` slot
    struct Slot { std::optional<StreamingResampler> r; };
    py::class_<Slot>(d, "ResamplerSlot").def(py::init<>()).def("is_some", [](const Slot& s) { return s.r.has_value(); });
    d.def("apply_output_options", [](const AudioData& a, const DecodeOptions& o, Slot& s) { return apply_output_options(a, o, s.r); });
    d.def("flush_resampler_frames", [](Slot& s) {
        if (!s.r) return std::vector<AudioData>();
        auto out = flush_resampler_frames(*s.r);
        s.r.reset();
        return out;
    });
}
