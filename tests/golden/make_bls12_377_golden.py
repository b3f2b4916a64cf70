"""Golden fixtures of the BLS12-377 G1 group layer, generated from the Python oracle
(oracle/py/bls12_377.py).  Like the Stark fixtures they pin the ORACLE's outputs (the reference ships no
vectors and cannot be built here) and travel to the GPU box.  Re-run:
    python tests/golden/make_bls12_377_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.py import bls12_377 as bls  # noqa: E402
from _util_bls12_377 import chain_points, pb, b32  # noqa: E402


def msm_fixture(n, seed, ncomp=1):
    s0, s1, pts, st = chain_points(n * ncomp, seed)
    ks = [st.scalar() for _ in range(n)]
    res = [bls.msm(pts[c::ncomp], ks) for c in range(ncomp)]
    return dict(n=n, seed=seed, ncomp=ncomp, points=b"".join(map(pb, pts)).hex(), scalars=b"".join(map(b32, ks)).hex(),
                result=b"".join(map(pb, res)).hex())


def pedersen_fixture(length, k, seed):
    s0, s1, pts, st = chain_points(length + 1, seed)
    vals = [[st.scalar() for _ in range(length)] for _ in range(k)]
    blinds = [st.scalar() for _ in range(k)]
    out = [bls.add(bls.mul(pts[0], r), bls.msm(pts[1:], v)) for v, r in zip(vals, blinds)]
    return dict(len=length, k=k, seed=seed, ck=b"".join(map(pb, pts)).hex(),
                values=b"".join(b32(x) for v in vals for x in v).hex(), blinds=b"".join(map(b32, blinds)).hex(),
                result=b"".join(map(pb, out)).hex())


def main():
    g = bls.G
    fix = dict(
        about="BLS12-377 G1 oracle fixtures; field elements 48-byte LE canonical, points x||y (96 B), scalars 32 B",
        generator=pb(g).hex(),
        multiples={str(k): pb(bls.mul(g, k)).hex() for k in (2, 3, 5, bls.N - 1, (1 << 252) + 12345)},
        msm=[msm_fixture(1, 1), msm_fixture(5, 2), msm_fixture(33, 3), msm_fixture(12, 4, ncomp=2)],
        pedersen=[pedersen_fixture(6, 3, 5)],
    )
    with open(os.path.join(HERE, "bls12_377_vectors.json"), "w") as f:
        json.dump(fix, f, indent=0)
    print("wrote", os.path.getsize(os.path.join(HERE, "bls12_377_vectors.json")), "bytes")


if __name__ == "__main__":
    main()
