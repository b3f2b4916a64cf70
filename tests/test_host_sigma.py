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
