#!/usr/bin/env python3
"""tools/dump_rubato_inputs.py — writes the seeded inputs of the rubato pin under tests/golden/rubato/ (<case>.json +
<case>.in.f32, planar f32 LE).  rust/rubato-dump turns them into <case>.out.f32 with the real crate; the inputs are small and
deterministic, so only the outputs need to be committed by whoever has cargo (see tests/golden/rubato/README.md)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "rubato")
CASES = [  # (name, in_rate, out_rate, channels, frames, semantics, signal)
    ("c1_48k_16k_oneshot", 48000, 16000, 2, 30011, "oneshot", "noise"),
    ("c5_48k_16k_streaming", 48000, 16000, 2, 50000, "streaming", "noise"),
    ("c4_44k1_48k_oneshot", 44100, 48000, 2, 30011, "oneshot", "noise"),
    ("c4_44k1_48k_streaming", 44100, 48000, 2, 50000, "streaming", "tones"),
    ("asr_44k1_16k_streaming", 44100, 16000, 1, 44100, "streaming", "tones"),
    ("tel_8k_16k_oneshot", 8000, 16000, 1, 12000, "oneshot", "noise"),
]


def signal(kind, rate, channels, frames, seed):
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.uniform(-0.5, 0.5, (channels, frames)).astype(np.float32)
    t = np.arange(frames) / rate
    return np.stack([0.45 * np.sin(2 * np.pi * (997.0 + 211.0 * c) * t) + 0.45 * np.sin(2 * np.pi * 0.43 * rate * t + c)
                     for c in range(channels)]).astype(np.float32)


def main():
    os.makedirs(OUT, exist_ok=True)
    for k, (name, i, o, ch, frames, sem, kind) in enumerate(CASES):
        x = signal(kind, i, ch, frames, 0x5EED0000 + k)
        x.astype("<f4").tofile(os.path.join(OUT, name + ".in.f32"))
        json.dump({"in_rate": i, "out_rate": o, "channels": ch, "frames": frames, "semantics": sem, "signal": kind},
                  open(os.path.join(OUT, name + ".json"), "w"))
        print(name, x.shape)


if __name__ == "__main__":
    sys.exit(main())
