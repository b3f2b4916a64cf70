"""Generates tests/golden/sigma_vectors.json from the Python oracle (oracle/py/sigma.py): seeded
mask / remask / reveal / key-ownership instances with their proofs, in the C-ABI byte formats.
Run from the repository root:  python tests/golden/make_sigma_golden.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.py import stark, sigma  # noqa: E402
from _util import chain_points  # noqa: E402

pb = stark.point_to_bytes64
hx = lambda v: "%064x" % v


def main():
    s0, s1, pts, st = chain_points(24, 77)
    g = stark.G
    sks = [st.scalar() for _ in range(3)]
    pks = [stark.mul(g, sk) for sk in sks]
    shared = stark.INF
    for pk in pks:
        shared = stark.add(shared, pk)
    out = {"g": pb(g).hex(), "shared_key": pb(shared).hex(), "mask": [], "remask": [], "reveal": [], "key_ownership": []}
    cards = pts[:6] + [stark.INF]
    maskeds = []
    for i, card in enumerate(cards):
        r = [st.scalar(), 0, stark.N - 1][i % 3] if i >= 4 else st.scalar()
        omega = st.scalar()
        masked, proof = sigma.mask(g, shared, card, r, omega)
        maskeds.append(masked)
        out["mask"].append({"card": pb(card).hex(), "r": hx(r), "omega": hx(omega), "masked": (pb(masked[0]) + pb(masked[1])).hex(),
                            "proof": sigma.cp_proof_bytes(proof).hex()})
    remaskeds = []
    for i, masked in enumerate(maskeds):
        alpha, omega = (st.scalar() if i != 2 else 0), st.scalar()
        remasked, proof = sigma.remask(g, shared, masked, alpha, omega)
        remaskeds.append(remasked)
        out["remask"].append({"original": (pb(masked[0]) + pb(masked[1])).hex(), "alpha": hx(alpha), "omega": hx(omega),
                              "remasked": (pb(remasked[0]) + pb(remasked[1])).hex(), "proof": sigma.cp_proof_bytes(proof).hex()})
    for i, masked in enumerate(remaskeds):
        sk, pk, omega = sks[i % 3], pks[i % 3], st.scalar()
        token, proof = sigma.compute_reveal_token(g, sk, pk, masked, omega)
        out["reveal"].append({"masked": (pb(masked[0]) + pb(masked[1])).hex(), "sk": hx(sk), "pk": pb(pk).hex(), "omega": hx(omega),
                              "token": pb(token).hex(), "proof": sigma.cp_proof_bytes(proof).hex()})
    for i, (sk, pk) in enumerate(zip(sks, pks)):
        info, omega = (b"player-%d" % i) * (i + 1), st.scalar()
        proof = sigma.prove_key_ownership(g, pk, sk, info, omega)
        out["key_ownership"].append({"sk": hx(sk), "pk": pb(pk).hex(), "info": info.hex(), "omega": hx(omega),
                                     "proof": sigma.schnorr_proof_bytes(proof).hex()})
    with open(os.path.join(ROOT, "tests", "golden", "sigma_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
