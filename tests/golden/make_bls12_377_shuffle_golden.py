"""Golden fixture of one Bayer-Groth shuffle proof over BLS12-377 G1, generated from the Python oracle
(`with bayer_groth.curve("bls12_377")`).  It pins the oracle for the NEXT row of the second curve -- the protocol
driver over BLS12-377 (upper half; the reference instantiates exactly this in
examples/parameter_selection.rs:25-29) -- the way oracle_vectors.json pins the Stark instantiation.  Layout: the flat
C-ABI proof of include/mpshuffle.h with 96-byte points and 32-byte scalars.  Re-run:
    python tests/golden/make_bls12_377_shuffle_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.py import bayer_groth as bg  # noqa: E402
from _util_bls12_377 import instance, pb, b32  # noqa: E402


def fixture(m, n, seed):
    with bg.curve("bls12_377"):
        pp, pk, deck, perm, rho, rnd = instance(m, n, seed)
        deck2, proof = bg.shuffle_and_remask(pp, pk, deck, rho, perm, rnd)
        assert bg.shuffle_verify(pp, pk, deck, deck2, proof) == bg.OK
        return dict(m=m, n=n, seed=seed, enc_g=pb(pp.enc_g).hex(), ck_g=b"".join(map(pb, pp.ck_g)).hex(),
                    ck_h=pb(pp.ck_h).hex(), ghat=pb(pp.ghat).hex(), pk=pb(pk).hex(),
                    deck=b"".join(pb(c[0]) + pb(c[1]) for c in deck).hex(), perm=perm,
                    rho=b"".join(map(b32, rho)).hex(), rand=b"".join(map(b32, rnd)).hex(),
                    deck2=b"".join(pb(c[0]) + pb(c[1]) for c in deck2).hex(), proof=bg.proof_to_bytes(proof).hex())


if __name__ == "__main__":
    out = dict(about="Bayer-Groth shuffle over BLS12-377 G1, oracle fixture: 96-byte points, 32-byte scalars",
               shuffle=[fixture(2, 3, 1), fixture(3, 4, 2)])
    path = os.path.join(HERE, "bls12_377_shuffle_vectors.json")
    json.dump(out, open(path, "w"), indent=0)
    print("wrote", os.path.getsize(path), "bytes")
