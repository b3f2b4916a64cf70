"""CPU test of the host half of the batched sigma protocols (csrc/sigma_host.hpp, compiled with g++):
the Fiat-Shamir challenges must equal the oracle's for the same statement bytes."""
import ctypes
import os
import subprocess

import pytest

from oracle.py import stark, sigma
from oracle.py.transcript import FiatShamirRng
from _util import chain_points

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pb = stark.point_to_bytes64


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim") / "host_shim.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "host", "host_shim.cpp")])
    return ctypes.CDLL(out)


def test_challenges_match_the_oracle(shim):
    _, _, pts, st = chain_points(6, 31)
    pts[3] = stark.INF  # the identity is encoded as (0, 1, infinity) inside the transcript
    out = ctypes.create_string_buffer(32)
    for which, seed in enumerate([sigma.MASKING_RNG_SEED, sigma.REMASKING_RNG_SEED, sigma.REVEAL_RNG_SEED]):
        shim.h_cp_challenge(which, *[pb(p) for p in pts], out)
        assert int.from_bytes(out.raw, "little") == sigma.cp_challenge(*pts, seed)
    for info in [b"", b"alice", bytes(range(200))]:
        shim.h_schnorr_challenge(info, ctypes.c_uint64(len(info)), pb(pts[0]), pb(pts[1]), pb(pts[2]), out)
        fs = FiatShamirRng(sigma.KEY_OWN_RNG_SEED + info)
        fs.absorb(b"schnorr_identity" + b"".join(stark.point_to_bytes65(p) for p in pts[:3]))
        assert int.from_bytes(out.raw, "little") == fs.challenge()
    for v, want in [(0, 1), (stark.N - 1, 1), (stark.N, 0), (2 ** 256 - 1, 0)]:
        assert shim.h_fr_bytes_canonical(v.to_bytes(32, "little")) == want


def test_challenges_match_the_oracle_over_bls12_377(tmp_path_factory, monkeypatch):
    """The same header compiled for the second curve (-DMP_CURVE_BLS12_377: 97-byte points in the transcript,
    challenges in ark_bls12_377::Fr) against oracle/py/sigma.py run over that curve."""
    from oracle.py import bayer_groth as bg, bls12_377 as bls
    from _util_bls12_377 import chain_points as chain377, pb as pb377
    out_so = str(tmp_path_factory.mktemp("shim377s") / "host_shim_bls12_377.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-DMP_CURVE_BLS12_377", "-o", out_so,
                           os.path.join(ROOT, "tests", "host", "host_shim_bls12_377.cpp")])
    lib = ctypes.CDLL(out_so)
    assert lib.h_sigma_proof_lens(0) == 2 * 96 + 32 and lib.h_sigma_proof_lens(1) == 96 + 32
    _, _, pts, st = chain377(6, 31)
    pts[3] = None
    out = ctypes.create_string_buffer(32)
    with bg.curve("bls12_377") as grp:      # group + challenge field of the oracle
        monkeypatch.setattr(sigma, "stark", grp)
        monkeypatch.setattr(sigma, "P65", grp.point_to_bytes65)
        monkeypatch.setattr(sigma, "Q", bls.N)
        for which, seed in enumerate([sigma.MASKING_RNG_SEED, sigma.REMASKING_RNG_SEED, sigma.REVEAL_RNG_SEED]):
            lib.h_cp_challenge(which, *[pb377(p) for p in pts], out)
            assert int.from_bytes(out.raw, "little") == sigma.cp_challenge(*pts, seed)
        for info in [b"", b"alice", bytes(range(200))]:
            lib.h_schnorr_challenge(info, ctypes.c_uint64(len(info)), pb377(pts[0]), pb377(pts[1]), pb377(pts[2]), out)
            fs = FiatShamirRng(sigma.KEY_OWN_RNG_SEED + info)
            fs.absorb(b"schnorr_identity" + b"".join(grp.point_to_bytes65(p) for p in pts[:3]))
            assert int.from_bytes(out.raw, "little") == fs.challenge()
        # a Chaum-Pedersen proof made and checked by the oracle over this curve (mask: s0 = c1, s1 = c2 - card)
        card, r, om = pts[0], st.scalar(), st.scalar()
        masked, proof = sigma.mask(bls.G, pts[1], card, r, om)
        assert sigma.verify_mask(bls.G, pts[1], card, masked, proof) == sigma.OK
        bad = (proof[0], proof[1], (proof[2] + 1) % bls.N)
        assert sigma.verify_mask(bls.G, pts[1], card, masked, bad) == sigma.ERR_CHAUM_PEDERSEN
        monkeypatch.undo()
    for v, want in [(0, 1), (bls.N - 1, 1), (bls.N, 0), (2 ** 256 - 1, 0)]:
        assert lib.h_fr_bytes_canonical(v.to_bytes(32, "little")) == want
