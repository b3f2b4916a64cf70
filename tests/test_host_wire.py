"""CPU test of the host-compiled halves of the wire format: the word-level Tonelli-Shanks of
csrc/fq_sqrt.cuh (the device algorithm, compiled with g++) against the oracle's big-int square root,
and the host-side compression / proof serialisation of csrc/wire_host.hpp against oracle/py/wire.py."""
import ctypes
import json
import os
import random
import subprocess

import pytest

from oracle.py import stark, wire

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "wire_vectors.json")))
SHUF = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_vectors.json")))["shuffle"]
h = bytes.fromhex


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim") / "host_shim.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "host", "host_shim.cpp")])
    return ctypes.CDLL(out)


def test_word_level_sqrt_matches_the_oracle(shim):
    rnd = random.Random(4)
    out = (ctypes.c_uint32 * 8)()
    w = lambda x: (ctypes.c_uint32 * 8)(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)])
    squares = nonsquares = 0
    cases = [0, 1, 4, stark.P - 1, 2, 3] + [rnd.randrange(stark.P) for _ in range(60)] + [rnd.randrange(stark.P) ** 2 % stark.P for _ in range(20)]
    for a in cases:
        ok = shim.h_fq_sqrt(w(a), out)
        r = sum(int(out[i]) << (32 * i) for i in range(8))
        want = stark.fq_sqrt(a)
        assert bool(ok) == (want is not None), a
        if ok:
            assert r * r % stark.P == a and r < stark.P
            squares += 1
        else:
            nonsquares += 1
    assert squares >= 30 and nonsquares >= 15


def test_windowed_sqrt_matches_the_oracle(shim):
    """The uniform-control-flow form the GPU runs (8-bit windows over the 192-bit discrete log)."""
    rnd = random.Random(5)
    out = (ctypes.c_uint32 * 8)()
    keys = ctypes.c_int(0)
    w = lambda x: (ctypes.c_uint32 * 8)(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)])
    squares = nonsquares = 0
    cases = [0, 1, 4, stark.P - 1, 2, 3] + [rnd.randrange(stark.P) for _ in range(80)] + [rnd.randrange(stark.P) ** 2 % stark.P for _ in range(40)]
    # elements of small 2-power order: every window of the discrete log but a few is zero
    t = (stark.P - 1) >> 192
    zeta = pow(3, t, stark.P)
    cases += [pow(zeta, 1 << s, stark.P) for s in (1, 7, 8, 9, 100, 183, 184, 190, 191)]
    for a in cases:
        ok = shim.h_fq_sqrt_win(w(a), out, ctypes.byref(keys))
        assert keys.value == 256
        r = sum(int(out[i]) << (32 * i) for i in range(8))
        assert bool(ok) == (stark.fq_sqrt(a) is not None), a
        if ok:
            assert r * r % stark.P == a and r < stark.P
            squares += 1
        else:
            nonsquares += 1
    assert squares >= 60 and nonsquares >= 20


def test_compression_and_proof_serialisation(shim):
    pts = b"".join(h(fx["point"]) for fx in GOLD["points"])
    out = ctypes.create_string_buffer(32 * len(GOLD["points"]))
    shim.h_wire_compress(pts, ctypes.c_uint64(len(GOLD["points"])), out)
    assert out.raw == b"".join(h(fx["compressed"]) for fx in GOLD["points"])
    for fx in SHUF:
        m, n, proof = fx["m"], fx["n"], h(fx["proof"])
        buf = ctypes.create_string_buffer((11 * m + 8) * 32 + (5 * n + 9) * 32)
        shim.h_wire_proof_serialize.restype = ctypes.c_uint64
        assert shim.h_wire_proof_serialize(m, n, proof, buf) == len(buf.raw)
        # walk the flat layout (include/mpshuffle.h) with the oracle's compressor
        want, pos = b"", 0
        for is_pt, cnt in [(1, 5 * m + 4), (0, 2 * n + 3), (1, 3), (0, 2 * n + 2), (1, 6 * m + 1), (0, n + 4)]:
            for _ in range(cnt):
                if is_pt:
                    want += wire.compress(stark.point_from_bytes64(proof[pos:pos + 64])); pos += 64
                else:
                    want += proof[pos:pos + 32]; pos += 32
        assert pos == len(proof) and buf.raw == want


def test_sqrt_over_bls12_377_matches_big_ints(tmp_path_factory):
    """csrc/fq_sqrt.cuh compiled for the second curve (q - 1 = 2^46 * t: 23 windows of 2 bits, exponent (t-1)/2 by
    square-and-multiply): textbook and windowed forms against big-int arithmetic -- r^2 == a for squares, refusal for
    non-residues, elements of small 2-power order, and the (q-1)/2 constant behind the "larger y" flag."""
    from oracle.py import bls12_377 as bls
    out_so = str(tmp_path_factory.mktemp("shim377w") / "host_shim_bls12_377.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-DMP_CURVE_BLS12_377", "-o", out_so,
                           os.path.join(ROOT, "tests", "host", "host_shim_bls12_377.cpp")])
    lib = ctypes.CDLL(out_so)
    Q = bls.Q
    w = lambda x: (ctypes.c_uint32 * 12)(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(12)])
    assert lib.h_fq_half_is_half(w((Q - 1) // 2)) == 1
    rnd = random.Random(6)
    out = (ctypes.c_uint32 * 12)()
    keys = ctypes.c_int(0)
    t = (Q - 1) >> 46
    zeta = pow(5, t, Q)
    assert pow(zeta, 1 << 45, Q) == Q - 1
    cases = [0, 1, 4, Q - 1, 2, 3, 5] + [rnd.randrange(Q) for _ in range(40)] + [rnd.randrange(Q) ** 2 % Q for _ in range(25)]
    cases += [pow(zeta, 1 << s, Q) for s in (0, 1, 2, 3, 22, 23, 43, 44, 45)] + [pow(zeta, 3 << s, Q) for s in (0, 1, 7, 40)]
    squares = nonsquares = 0
    for a in cases:
        is_sq = a == 0 or pow(a, (Q - 1) // 2, Q) == 1
        for fn in ("plain", "win"):
            ok = lib.h_fq_sqrt(w(a), out) if fn == "plain" else lib.h_fq_sqrt_win(w(a), out, ctypes.byref(keys))
            r = sum(int(out[i]) << (32 * i) for i in range(12))
            assert bool(ok) == is_sq, (fn, a)
            if ok:
                assert r * r % Q == a and r < Q, (fn, a)
        squares += is_sq
        nonsquares += not is_sq
    assert keys.value == 4 and squares >= 30 and nonsquares >= 15
