"""Generates tests/golden/bls12_377_wire_vectors.json from the Python oracle (oracle/py/wire.py, generic forms over
BLS12-377 G1): compressed points, a serialised deck, and encodings a deserialiser must reject (abscissas off the curve,
x >= q, stray bits, bad infinity encodings, a curve point outside G1).
Run from the repository root:  python tests/golden/make_bls12_377_wire_golden.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.py import bls12_377 as bls, wire  # noqa: E402
from _util_bls12_377 import chain_points, pb  # noqa: E402

C = bls.CURVE


def main():
    _, _, pts, st = chain_points(24, 4377)
    pts = pts[:12] + [bls.neg(p) for p in pts[:4]] + [None, bls.G, bls.neg(bls.G)]
    out = {"points": [{"point": pb(p).hex(), "compressed": wire.compress_generic(p, C).hex()} for p in pts]}
    for p in pts:
        assert wire.decompress_generic(wire.compress_generic(p, C), C) == p
    deck = [(pts[2 * i], pts[2 * i + 1]) for i in range(8)]
    out["deck"] = b"".join(pb(a) + pb(b) for a, b in deck).hex()
    out["deck_serialized"] = wire.deck_serialize_generic(deck, C).hex()
    assert wire.deck_deserialize_generic(bytes.fromhex(out["deck_serialized"]), C) == deck
    rejected, statuses = [], []
    x = pts[0][0]
    for _ in range(3):  # abscissas off the curve
        x += 1
        while wire.sqrt_mod(x ** 3 + 1, bls.Q) is not None:
            x += 1
        rejected.append(x.to_bytes(48, "little").hex()); statuses.append(2)
    rejected.append(bls.Q.to_bytes(48, "little").hex()); statuses.append(1)                     # x = q
    rejected.append((pts[0][0] | (1 << 380)).to_bytes(48, "little").hex()); statuses.append(1)   # a stray bit above 2^377
    rejected.append((bytes(47) + b"\xc0").hex()); statuses.append(1)                             # infinity + sign flag
    rejected.append((b"\x01" + bytes(46) + b"\x40").hex()); statuses.append(1)                   # infinity with x != 0
    x = 1
    while True:                                                                                  # on the curve, outside G1
        y = wire.sqrt_mod(x ** 3 + 1, bls.Q)
        if y is not None and not wire.in_subgroup((x, y), C):
            break
        x += 1
    rejected.append(wire.compress_generic((x, y), C).hex()); statuses.append(3)
    out["rejected"], out["rejected_statuses"] = rejected, statuses
    for b in rejected:
        try:
            wire.decompress_generic(bytes.fromhex(b), C)
            raise AssertionError("the oracle accepted " + b)
        except ValueError:
            pass
    with open(os.path.join(ROOT, "tests", "golden", "bls12_377_wire_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
