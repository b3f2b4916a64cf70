"""Golden proofs for the reference's own benchmark shape (examples/parameter_selection.rs:31-43): ONE 300-card deck over
BLS12-377 G1 at (m, n) = (10, 30) -- the split the reference names as the proof-size optimum -- and (30, 10), from the
Python oracle (`with bayer_groth.curve("bls12_377")`).  Only the proof bytes and a digest of the shuffled deck are
stored: parameters, deck, permutation and randomness are regenerated from the seed by tests/_util_bls12_377.instance.
Re-run (about two minutes of big-int arithmetic per shape):
    python tests/golden/make_bls12_377_shuffle_300_golden.py
"""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.py import bayer_groth as bg  # noqa: E402
from _util_bls12_377 import instance, pb  # noqa: E402


def fixture(m, n, seed):
    with bg.curve("bls12_377"):
        t0 = time.time()
        pp, pk, deck, perm, rho, rnd = instance(m, n, seed)
        deck2, proof = bg.shuffle_and_remask(pp, pk, deck, rho, perm, rnd)
        assert bg.shuffle_verify(pp, pk, deck, deck2, proof) == bg.OK
        d2 = b"".join(pb(c[0]) + pb(c[1]) for c in deck2)
        print((m, n), "oracle prove + verify: %.1f s" % (time.time() - t0), flush=True)
        return dict(m=m, n=n, seed=seed, deck2_sha256=hashlib.sha256(d2).hexdigest(), proof=bg.proof_to_bytes(proof).hex())


if __name__ == "__main__":
    out = dict(about="Bayer-Groth shuffle of a 300-card deck over BLS12-377 G1 (the reference benchmark's shape), oracle proofs; "
                     "inputs regenerated from the seed by tests/_util_bls12_377.instance",
               shuffle=[fixture(10, 30, 7), fixture(30, 10, 8)])
    path = os.path.join(HERE, "bls12_377_shuffle_300_vectors.json")
    json.dump(out, open(path, "w"), indent=0)
    print("wrote", os.path.getsize(path), "bytes")
