"""The C-ABI library loads without a GPU and exports every symbol include/mpshuffle.h declares
(no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header="mpshuffle.h", prefix="mp_"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(pkg):
    lib = ctypes.CDLL(pkg.lib_path)
    names = declared_symbols()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_header(pkg):
    from mental_poker_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_bls12_377_header_symbols_exported_and_bound(pkg):
    # second curve (include/mpshuffle_bls12_377.h): same library, mp377_ prefix
    assert sorted(os.listdir(os.path.join(ROOT, "include"))) == ["mpshuffle.h", "mpshuffle_bls12_377.h"]
    lib = ctypes.CDLL(pkg.lib_path)
    names = declared_symbols("mpshuffle_bls12_377.h", "mp377_")
    assert len(names) >= 15
    assert not [n for n in names if not hasattr(lib, n)]
    assert sorted(pkg.bls12_377.SIGNATURES) == names


def test_status_strings(pkg):
    # the reference's pinned message (tests.rs:223-225)
    assert pkg.lib.mp_verify_status_string(1) == b"Hadamard Product (5.1)"
    assert pkg.lib.mp_verify_status_string(0) == b"ok"
    # masking.rs:103-105 / tests.rs:72-77
    assert pkg.lib.mp_verify_status_string(5) == b"Chaum-Pedersen"
    assert pkg.lib.mp_verify_status_string(6) == b"Schnorr Identification"


def test_no_cpu_fallback_without_device(pkg):
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    with pytest.raises(pkg.MpError):
        pkg.Context(0)
    with pytest.raises(pkg.MpError):
        pkg.bls12_377.Context(0)
