#!/usr/bin/env python3
"""Generate tests/golden/ from the reference's committed fixtures (run in the build container only).

/root/reference does not exist on the GPU box, so everything the tests need is copied / extracted
here and committed together with this script:

  * G.711 fixtures + the reference's golden decoded WAV payloads
    (testdata/g711_{ulaw,alaw}/*, golden/g711_{ulaw,alaw}/*.decoded.wav; soundkit-g711/src/lib.rs:376-400)
  * linear16 raw PCM + golden WAV payload (soundkit/src/raw_pcm.rs:286-318), linear16_8 (8 kHz source)
  * CAF / MOV PCM payloads (4 800 stereo frames each) with the LE-canonical SHA-256 the reference asserts
    (soundkit-audio-demux/src/caf.rs:777-842, soundkit-audio-demux/src/lib.rs:5568-5690)
  * short slices of the s32le / 24-bit WAV / f32 WAV speech fixtures as realistic conversion inputs
  * G.726 fixtures for 16/24/32/40 kbit/s + the reference's golden decoded WAV payloads
    (testdata/g726/*, golden/g726/*.decoded.wav; soundkit-g726/src/lib.rs:1113-1145)

Only DATA is copied, no reference source.  MANIFEST.json records the SHA-256 of every file.
"""
import hashlib
import json
import os
import struct
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
NAME = "A_Tusk_is_used_to_make_costly_gifts"

ENDIAN_HASHES = {  # LE-canonical payload hashes asserted by the reference tests
    "s16": "7126a0d8077e0433c1e78df6c0f1074f78daf3800240ef48c936114ef7cbb563",
    "s24": "c4e54b143defc7ce84e84c3b6fe96b1f87423477a8ca0f6676085b35eba7d087",
    "s32": "1568245a441f20f5c94b0810fa460c820f0e7d65ab2c6091100bf8c23ad7913c",
    "f32": "31522f2b9f09614ab868408fd95f5d826028b61407827b7b112d0ab9981ad65e",
}


def aifc_ssnd(path):
    """sound data of an AIFF-C file: FORM/AIFC chunk walk, SSND body after its offset / blockSize words"""
    b = open(path, "rb").read()
    assert b[:4] == b"FORM" and b[8:12] == b"AIFC"
    pos = 12
    while pos + 8 <= len(b):
        cid, size = b[pos:pos + 4], struct.unpack(">I", b[pos + 4:pos + 8])[0]
        body = b[pos + 8:pos + 8 + size]
        if cid == b"SSND":
            off = struct.unpack(">I", body[:4])[0]
            return body[8 + off:]
        pos += 8 + size + (size & 1)
    raise ValueError("no SSND chunk in " + path)


def caf_payload(path):
    b = open(path, "rb").read()
    assert b[:4] == b"caff"
    off = 8
    while off < len(b):
        kind, size = b[off:off + 4], struct.unpack(">q", b[off + 4:off + 12])[0]
        off += 12
        if kind == b"data":
            end = len(b) if size < 0 else off + size
            return b[off + 4:end]  # skip mEditCount
        off += size
    raise ValueError("no data chunk")


def mov_payload(path):
    b = open(path, "rb").read()
    off = 0
    while off < len(b):
        size, kind = struct.unpack(">I4s", b[off:off + 8])
        hdr = 8
        if size == 1:
            size = struct.unpack(">Q", b[off + 8:off + 16])[0]
            hdr = 16
        if kind == b"mdat":
            return b[off + hdr:off + size]
        off += size
    raise ValueError("no mdat")


def wav_payload(path):
    b = open(path, "rb").read()
    assert b[:4] == b"RIFF" and b[8:12] == b"WAVE"
    off = 12
    while off < len(b):
        kind, size = b[off:off + 4], struct.unpack("<I", b[off + 4:off + 8])[0]
        if kind == b"data":
            return b[off + 8:off + 8 + size]
        off += 8 + size + (size & 1)
    raise ValueError("no data chunk")


def to_le(payload, bps):
    out = bytearray(payload)
    for i in range(0, len(out) - bps + 1, bps):
        out[i:i + bps] = out[i:i + bps][::-1]
    return bytes(out)


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (build container only)")
    files = {}

    def put(name, data):
        with open(os.path.join(HERE, name), "wb") as f:
            f.write(data)
        files[name] = hashlib.sha256(data).hexdigest()

    rd = lambda p: open(os.path.join(REF, p), "rb").read()
    put("g711_ulaw.ulaw", rd(f"testdata/g711_ulaw/{NAME}.ulaw"))
    put("g711_alaw.alaw", rd(f"testdata/g711_alaw/{NAME}.alaw"))
    put("g711_ulaw.decoded.s16le", wav_payload(os.path.join(REF, f"golden/g711_ulaw/{NAME}.decoded.wav")))
    put("g711_alaw.decoded.s16le", wav_payload(os.path.join(REF, f"golden/g711_alaw/{NAME}.decoded.wav")))
    put("linear16.s16le", rd(f"testdata/linear16/{NAME}.s16le"))
    put("linear16.decoded.s16le", wav_payload(os.path.join(REF, f"golden/linear16/{NAME}.decoded.wav")))
    put("linear16_8.s16le", rd(f"testdata/linear16_8/{NAME}.s16le"))
    # realistic conversion inputs, first 16 000 samples each
    put("linear32.s32le", rd(f"testdata/linear32/{NAME}.s32le")[:64000])
    put("wav_24.s24le", wav_payload(os.path.join(REF, f"testdata/wav_24/{NAME}.wav"))[:48000])
    put("wav_32f.f32le", wav_payload(os.path.join(REF, f"testdata/wav_32f/{NAME}.wav"))[:64000])

    # G.726 fixtures (ffmpeg-encoded, left-justified packing) + the reference's golden decodes for all four rates
    # (soundkit-g726/src/lib.rs:1113-1145): pins the ADPCM decoder restatement bit for bit
    for r in ("16", "24", "32", "40"):
        put(f"g726_{r}.g726", rd(f"testdata/g726/{NAME}_{r}.g726"))
        put(f"g726_{r}.decoded.s16le", wav_payload(os.path.join(REF, f"golden/g726/{NAME}_{r}.decoded.wav")))
    # G.722 fixture and the WAV the reference's own decoder wrote (soundkit-g722/src/lib.rs:330-360): pins the sub-band ADPCM
    # decoder restatement bit for bit, including the crate's zero reset state
    put("g722.g722", rd(f"testdata/g722/{NAME}.g722"))
    put("g722.decoded.s16le", wav_payload(os.path.join(REF, f"golden/g722/{NAME}.decoded.wav")))

    # AIFF-C mu-law / A-law (soundkit-aiff/src/lib.rs:529-540 decodes the SSND payload with audio-codec-algorithms' decode_ulaw /
    # decode_alaw): the payloads of testdata/aifc/stream-{ulaw,alaw}.aifc are byte-identical to the raw G.711 fixtures above, whose
    # decodes the reference holds as golden WAVs -> no second copy, the identity is recorded
    aifc = {}
    for law in ("ulaw", "alaw"):
        payload = aifc_ssnd(os.path.join(REF, f"testdata/aifc/stream-{law}.aifc"))
        raw = rd(f"testdata/g711_{law}/{NAME}.{law}")
        assert payload == raw, law
        aifc[f"stream-{law}.aifc"] = {"ssnd_sha256": hashlib.sha256(payload).hexdigest(), "same_bytes_as": f"g711_{law}.{law}",
                                      "decoded_golden": f"g711_{law}.decoded.s16le"}

    endian = {}
    for name, kind, bps, be in (("pcm-s16be.caf", "s16", 2, True), ("pcm-s24le.caf", "s24", 3, False),
                                ("pcm-s32be.caf", "s32", 4, True), ("pcm-f32le.caf", "f32", 4, False)):
        p = caf_payload(os.path.join(REF, "testdata/caf", name))[:4800 * 2 * bps]
        le = to_le(p, bps) if be else p
        assert hashlib.sha256(le).hexdigest() == ENDIAN_HASHES[kind], name
        out = name.replace(".caf", ".caf.bin")
        put(out, p)
        endian[out] = {"kind": kind, "bytes_per_sample": bps, "big_endian": be, "le_sha256": ENDIAN_HASHES[kind]}
    for name, kind, bps, be in (("pcm-s24be.mov", "s24", 3, True), ("pcm-s24le.mov", "s24", 3, False),
                                ("pcm-s32be.mov", "s32", 4, True), ("pcm-s32le.mov", "s32", 4, False),
                                ("pcm-f32be.mov", "f32", 4, True), ("pcm-f32le.mov", "f32", 4, False)):
        p = mov_payload(os.path.join(REF, "testdata/mov-pcm", name))[:4800 * 2 * bps]
        le = to_le(p, bps) if be else p
        assert hashlib.sha256(le).hexdigest() == ENDIAN_HASHES[kind], name
        out = name.replace(".mov", ".mov.bin")
        put(out, p)
        endian[out] = {"kind": kind, "bytes_per_sample": bps, "big_endian": be, "le_sha256": ENDIAN_HASHES[kind]}

    with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
        json.dump({"files": files, "endian_payloads": endian, "aifc_g711": aifc}, f, indent=1, sort_keys=True)
    print(f"wrote {len(files)} fixtures")


if __name__ == "__main__":
    main()
