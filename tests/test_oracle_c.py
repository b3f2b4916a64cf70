"""Pins the C restatement (oracle/c) against the Python big-int oracle, RFC vectors and the
committed golden fixtures -- byte for byte."""
import ctypes
import hashlib
import json
import os
import random

import pytest

from oracle import c_oracle
from oracle.py import stark, bayer_groth as bg
from oracle.py.transcript import FiatShamirRng
from _util import chain_points, scalars, b32, pb

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.json")))


@pytest.fixture(scope="module")
def co():
    c_oracle.build()
    return c_oracle.COracle()


def test_field_mul(co):
    rnd = random.Random(1)
    out = ctypes.create_string_buffer(32)
    for mod, fn in ((stark.P, co.lib.oc_fq_mul), (stark.N, co.lib.oc_fr_mul)):
        for a, b in [(0, 5), (1, 1), (mod - 1, mod - 1)] + [(rnd.randrange(mod), rnd.randrange(mod)) for _ in range(500)]:
            fn(b32(a), b32(b), out)
            assert int.from_bytes(out.raw, "little") == a * b % mod


def test_blake2s_and_transcript(co):
    out = ctypes.create_string_buffer(32)
    for msg in [b"", b"abc", b"x" * 64, b"y" * 65, bytes(range(256)) * 5]:
        co.lib.oc_blake2s(msg, len(msg), out)
        assert out.raw == hashlib.blake2s(msg).digest()
    for data in [b"", b"hello", bytes(1000)]:
        buf = ctypes.create_string_buffer(32 * 5)
        co.lib.oc_fs_challenges(data, len(data), 5, buf)
        fs = FiatShamirRng()
        if data:
            fs.absorb(data)
        want = b"".join(b32(fs.challenge()) for _ in range(5))
        assert buf.raw == want


@pytest.mark.parametrize("mode", [0, 1, 5, 9])
def test_msm_matches_python(co, mode):
    s0, s1, pts, st = chain_points(70, 4)
    for kind in ["uniform", "zero", "max", "small", "same"]:
        ks = scalars(st, 70, kind)
        if kind == "small":
            ks[3] = 1  # arkworks adds scalars equal to one directly
        want = pb(stark.msm(pts, ks))
        assert co.msm(b"".join(map(pb, pts)), b"".join(map(b32, ks)), 1, mode) == want, kind


def test_msm_point_edges(co):
    rnd = random.Random(7)
    Pt = stark.mul(stark.G, 77)
    ks = [rnd.randrange(stark.N) for _ in range(40)]
    for pts in ([Pt] * 40, [Pt if i % 2 else stark.neg(Pt) for i in range(40)], [None if i % 3 == 0 else Pt for i in range(40)]):
        for mode in (0, 1, 4):
            assert co.msm(b"".join(map(pb, pts)), b"".join(map(b32, ks)), 1, mode) == pb(stark.msm(pts, ks))
    assert co.lib.oc_on_curve(pb(Pt)) == 1 and co.lib.oc_on_curve(b32(5) + b32(7)) == 0


def test_golden_msm(co):
    for fx in GOLD["msm"]:
        assert co.msm(bytes.fromhex(fx["points"]), bytes.fromhex(fx["scalars"])).hex() == fx["result"]


def _args(fx):
    h = bytes.fromhex
    return (fx["m"], fx["n"], h(fx["enc_g"]), h(fx["ck_g"]), h(fx["ck_h"]), h(fx["ghat"]), h(fx["pk"]))


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
@pytest.mark.parametrize("msm_mode", [0, 1])
def test_golden_shuffle_bytes(co, idx, msm_mode):
    fx = GOLD["shuffle"][idx]
    h = bytes.fromhex
    co.set(msm_mode=msm_mode)
    a = _args(fx)
    deck2 = co.remask(a[2], a[6], h(fx["deck"]), fx["perm"], h(fx["rho"]))
    assert deck2.hex() == fx["deck2"]
    proof = co.prove(*a, h(fx["deck"]), deck2, fx["perm"], h(fx["rho"]), h(fx["rand"]))
    assert proof.hex() == fx["proof"]
    assert co.verify(*a, h(fx["deck"]), deck2, proof) == 0
    co.set(msm_mode=0)


def test_negative_cases_match_reference_strings(co):
    # tests.rs:213-226 -> "Hadamard Product (5.1)"
    fx = GOLD["shuffle"][2]
    h = bytes.fromhex
    a = _args(fx)
    _, _, pts, _ = chain_points(104, 99)
    wrong = b"".join(pb(p) for p in pts)
    st = co.verify(*a, h(fx["deck"]), wrong, h(fx["proof"]))
    assert st == bg.ERR_HADAMARD and bg.ERR_STRINGS[st] == "Hadamard Product (5.1)"
    # flipping one scalar of each sub-proof trips that sub-argument
    m, n = fx["m"], fx["n"]
    proof = bytearray(h(fx["proof"]))
    off_zero_t = (2 * m + 1 + m + 2 * m + 3) * 64 + (2 * n + 2) * 32
    for off, code in [(off_zero_t, bg.ERR_ZERO), (off_zero_t + 32 + 3 * 64 + (2 * n) * 32, bg.ERR_SVP),
                      (len(proof) - 32, bg.ERR_MULTIEXP)]:
        p2 = bytearray(proof)
        p2[off] ^= 1
        assert co.verify(*a, h(fx["deck"]), h(fx["deck2"]), bytes(p2)) == code
        pf = bg.proof_from_bytes(bytes(p2), m, n)  # same verdict from the Python oracle


def test_threads_do_not_change_results(co):
    fx = GOLD["shuffle"][1]
    h = bytes.fromhex
    co.set(threads=4)
    a = _args(fx)
    assert co.prove(*a, h(fx["deck"]), h(fx["deck2"]), fx["perm"], h(fx["rho"]), h(fx["rand"])).hex() == fx["proof"]
    assert co.verify(*a, h(fx["deck"]), h(fx["deck2"]), h(fx["proof"])) == 0
    co.set(threads=1)


def test_non_canonical_proof_scalars_are_rejected_by_both_oracles(co):
    """`Proof: CanonicalDeserialize` (reference src/lib.rs:45-71): a scalar >= the group order never reaches the
    reference's verifier.  Both restatements reject it (-5 = MP_ERR_NOT_CANONICAL / NonCanonicalScalar)."""
    import pytest
    from oracle.py import bayer_groth as bg, stark
    fx = GOLD["shuffle"][0]
    m, n = fx["m"], fx["n"]
    h = bytes.fromhex
    args = (m, n, h(fx["enc_g"]), h(fx["ck_g"]), h(fx["ck_h"]), h(fx["ghat"]), h(fx["pk"]), h(fx["deck"]), h(fx["deck2"]))
    proof = h(fx["proof"])
    assert co.verify(*args, proof) == 0
    f1 = 64 * (5 * m + 4)
    f2 = f1 + 32 * (2 * n + 3) + 64 * 3
    f3 = f2 + 32 * (2 * n + 2) + 64 * (6 * m + 1)
    for off in (f1, f1 + 32 * (2 * n + 2), f2, f2 + 32 * (2 * n + 1), f3, len(proof) - 32):
        s = int.from_bytes(proof[off:off + 32], "little")
        p2 = proof[:off] + (s + stark.N).to_bytes(32, "little") + proof[off + 32:]
        assert co.verify(*args, p2) == -5
        with pytest.raises(bg.NonCanonicalScalar):
            bg.proof_from_bytes(p2, m, n)
    # points are untouched by the check: the first point run ends where the first scalar run starts
    assert co.verify(*args, proof[:f1 - 1] + bytes([proof[f1 - 1] ^ 0x80]) + proof[f1:]) != -5
