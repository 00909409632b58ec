"""CPU tests: the C-ABI library loads and exports every symbol include/soundkit_cuda.h declares
(no compute calls without a GPU), and the product refuses to run without its CUDA path."""
import os
import re
import subprocess

import pytest

from conftest import ROOT
from soundkit_b200 import capi


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "soundkit_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sk_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_list_the_same_symbols():
    assert declared_symbols() == sorted(capi.EXPORTS)


def test_library_exports_every_declared_symbol():
    L = capi.load()
    for name in declared_symbols():
        assert hasattr(L, name), name
    out = subprocess.check_output(["nm", "-D", "--defined-only", capi.LIB_PATH], text=True)
    exported = set(re.findall(r" T (sk_[a-z0-9_]+)", out))
    assert set(declared_symbols()) <= exported
    assert L.sk_abi_version() == 1


def test_library_is_sm100a_only():
    out = subprocess.check_output(["cuobjdump", "-lelf", capi.LIB_PATH], text=True)
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.SoundkitError) as e:
        capi.Context(0)
    assert e.value.code == capi.E_NODEVICE


def test_product_never_touches_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use oracle/."""
    pkg = os.path.join(ROOT, "soundkit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower().replace("no cpu", ""), os.path.join(dirpath, f)


def test_flat_op_widths_table():
    L = capi.load()
    import ctypes as C
    for name, op in capi.FLAT_OP.items():
        a, b = C.c_int(), C.c_int()
        assert L.sk_flat_op_widths(op, C.byref(a), C.byref(b)) == 0
        src = name.split("_to_")[0].replace("pkt_", "")
        assert a.value == {"i16le": 2, "s16le": 2, "s16be": 2, "s24le": 3, "s24be": 3, "s32le": 4, "s32be": 4,
                           "f32le": 4, "f32be": 4, "u8": 1}[src]
        assert b.value == (2 if name.endswith(("i16", "s16")) else 4)
    assert L.sk_flat_op_widths(99, None, None) == capi.E_ARG


def test_rust_sys_crate_binds_every_declared_symbol():
    """the Rust `-sys` crate cannot be compiled here (no cargo): at least its extern block must name exactly the header's symbols,
    and the state-size constants must agree with the header"""
    src = open(os.path.join(ROOT, "rust", "soundkit-cuda-sys", "src", "lib.rs")).read()
    bound = sorted(set(re.findall(r"pub fn (sk_[a-z0-9_]+)\s*\(", src)))
    assert bound == declared_symbols()
    hdr = open(os.path.join(ROOT, "include", "soundkit_cuda.h")).read()
    for name in ("SK_G726_STATE_INTS", "SK_G722_STATE_INTS"):
        c = int(re.search(rf"#define {name} (\d+)", hdr).group(1))
        r = int(re.search(rf"pub const {name}: usize = (\d+);", src).group(1))
        assert c == r, name
    assert capi.G722_STATE_INTS == int(re.search(r"#define SK_G722_STATE_INTS (\d+)", hdr).group(1))
