"""Generates the committed golden fixtures from the Python oracle.

The reference ships no golden vectors and cannot be built here, so
these fixtures pin the *oracle's* outputs: they guard the oracle itself against regressions,
give the C restatement (oracle/c) and the CUDA path byte-level targets, and travel to the
GPU box where /root/reference does not exist.  Re-run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.py import stark, bayer_groth as bg  # noqa: E402
from _util import instance, chain_points, pb, b32  # noqa: E402


def shuffle_fixture(m, n, seed):
    pp, pk, deck, perm, rho, rnd = instance(m, n, seed)
    deck2, proof = bg.shuffle_and_remask(pp, pk, deck, rho, perm, rnd)
    assert bg.shuffle_verify(pp, pk, deck, deck2, proof) == bg.OK
    return dict(
        m=m, n=n, seed=seed,
        enc_g=pb(pp.enc_g).hex(), ck_g=b"".join(map(pb, pp.ck_g)).hex(), ck_h=pb(pp.ck_h).hex(),
        ghat=pb(pp.ghat).hex(), pk=pb(pk).hex(),
        deck=b"".join(pb(c[0]) + pb(c[1]) for c in deck).hex(),
        perm=perm, rho=b"".join(map(b32, rho)).hex(), rand=b"".join(map(b32, rnd)).hex(),
        deck2=b"".join(pb(c[0]) + pb(c[1]) for c in deck2).hex(),
        proof=bg.proof_to_bytes(proof).hex())


def msm_fixture(n, seed):
    s0, s1, pts, st = chain_points(n, seed)
    ks = [st.scalar() for _ in range(n)]
    return dict(n=n, seed=seed, points=b"".join(map(pb, pts)).hex(), scalars=b"".join(map(b32, ks)).hex(),
                result=pb(stark.msm(pts, ks)).hex())


if __name__ == "__main__":
    out = dict(
        shuffle=[shuffle_fixture(2, 3, 1), shuffle_fixture(3, 4, 2), shuffle_fixture(4, 13, 1), shuffle_fixture(2, 26, 3)],
        msm=[msm_fixture(1, 1), msm_fixture(37, 2), msm_fixture(200, 3)])
    with open(os.path.join(HERE, "oracle_vectors.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", os.path.join(HERE, "oracle_vectors.json"))
