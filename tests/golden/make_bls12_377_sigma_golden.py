"""Generates tests/golden/bls12_377_sigma_vectors.json from the Python oracle (oracle/py/sigma.py run over BLS12-377 G1):
seeded mask / remask / reveal / key-ownership instances with their proofs, in the C-ABI byte formats of
include/mpshuffle_bls12_377.h (96-byte points, 224-byte Chaum-Pedersen proofs, 128-byte Schnorr proofs).
Run from the repository root:  python tests/golden/make_bls12_377_sigma_golden.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.py import bls12_377 as bls, sigma  # noqa: E402
from _util_bls12_377 import chain_points, pb  # noqa: E402

hx = lambda v: "%064x" % v


def main():
    with sigma.curve("bls12_377"):
        s0, s1, pts, st = chain_points(12, 177)
        g = bls.G
        sks = [st.scalar() for _ in range(3)]
        pks = [bls.mul(g, sk) for sk in sks]
        shared = None
        for pk in pks:
            shared = bls.add(shared, pk)
        out = {"g": pb(g).hex(), "shared_key": pb(shared).hex(), "mask": [], "remask": [], "reveal": [], "key_ownership": []}
        cards = pts[:5] + [None]                      # the last card is the identity
        maskeds = []
        for i, card in enumerate(cards):
            r = [st.scalar(), 0, bls.N - 1][i % 3] if i >= 3 else st.scalar()
            omega = st.scalar()
            masked, proof = sigma.mask(g, shared, card, r, omega)
            assert sigma.verify_mask(g, shared, card, masked, proof) == sigma.OK
            maskeds.append(masked)
            out["mask"].append({"card": pb(card).hex(), "r": hx(r), "omega": hx(omega), "masked": (pb(masked[0]) + pb(masked[1])).hex(),
                                "proof": sigma.cp_proof_bytes(proof).hex()})
        remaskeds = []
        for i, masked in enumerate(maskeds):
            alpha, omega = (st.scalar() if i != 2 else 0), st.scalar()
            remasked, proof = sigma.remask(g, shared, masked, alpha, omega)
            assert sigma.verify_remask(g, shared, masked, remasked, proof) == sigma.OK
            remaskeds.append(remasked)
            out["remask"].append({"original": (pb(masked[0]) + pb(masked[1])).hex(), "alpha": hx(alpha), "omega": hx(omega),
                                  "remasked": (pb(remasked[0]) + pb(remasked[1])).hex(), "proof": sigma.cp_proof_bytes(proof).hex()})
        for i, masked in enumerate(remaskeds):
            sk, pk, omega = sks[i % 3], pks[i % 3], st.scalar()
            token, proof = sigma.compute_reveal_token(g, sk, pk, masked, omega)
            assert sigma.verify_reveal(g, pk, token, masked, proof) == sigma.OK
            out["reveal"].append({"masked": (pb(masked[0]) + pb(masked[1])).hex(), "sk": hx(sk), "pk": pb(pk).hex(), "omega": hx(omega),
                                  "token": pb(token).hex(), "proof": sigma.cp_proof_bytes(proof).hex()})
        for i, (sk, pk) in enumerate(zip(sks, pks)):
            info, omega = (b"player-%d" % i) * (i + 1), st.scalar()
            proof = sigma.prove_key_ownership(g, pk, sk, info, omega)
            assert sigma.verify_key_ownership(g, pk, info, proof) == sigma.OK
            out["key_ownership"].append({"sk": hx(sk), "pk": pb(pk).hex(), "info": info.hex(), "omega": hx(omega),
                                         "proof": sigma.schnorr_proof_bytes(proof).hex()})
    with open(os.path.join(ROOT, "tests", "golden", "bls12_377_sigma_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
