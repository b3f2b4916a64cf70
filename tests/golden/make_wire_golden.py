"""Generates tests/golden/wire_vectors.json from the Python oracle (oracle/py/wire.py).
Run from the repository root:  python tests/golden/make_wire_golden.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.py import stark, wire  # noqa: E402
from _util import chain_points  # noqa: E402

pb = stark.point_to_bytes64


def main():
    _, _, pts, st = chain_points(24, 404)
    pts = pts[:12] + [stark.neg(p) for p in pts[:4]] + [stark.INF, stark.G, stark.neg(stark.G)]
    out = {"points": [{"point": pb(p).hex(), "compressed": wire.compress(p).hex()} for p in pts]}
    deck = [(pts[2 * i], pts[2 * i + 1]) for i in range(8)]
    out["deck"] = b"".join(pb(a) + pb(b) for a, b in deck).hex()
    out["deck_serialized"] = wire.deck_serialize(deck).hex()
    rejected = []
    x = pts[0][0]
    for _ in range(3):  # abscissas off the curve
        x += 1
        while stark.fq_sqrt((x ** 3 + x + stark.B) % stark.P) is not None:
            x += 1
        rejected.append(stark.fe_to_bytes(x).hex())
    rejected.append(stark.fe_to_bytes(stark.P).hex())            # x = p
    rejected.append((bytes(31) + b"\xc0").hex())                  # infinity + sign flag
    rejected.append((b"\x01" + bytes(30) + b"\x40").hex())        # infinity with x != 0
    out["rejected"] = rejected
    with open(os.path.join(ROOT, "tests", "golden", "wire_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
