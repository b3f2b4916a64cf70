"""Oracle-level checks of the wire format (oracle/py/wire.py): round trips,
flag semantics, rejection of malformed encodings, and the committed golden vectors."""
import json
import os

import pytest

from oracle.py import stark, wire
from _util import chain_points

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "wire_vectors.json")))
h = bytes.fromhex


def test_round_trip_and_flags():
    _, _, pts, _ = chain_points(40, 5)
    for p in pts + [stark.INF, stark.neg(pts[0])]:
        c = wire.compress(p)
        assert len(c) == 32 and wire.decompress(c) == p
    p = pts[3]
    c, cn = wire.compress(p), wire.compress(stark.neg(p))
    assert c[:31] == cn[:31] and (c[31] ^ cn[31]) == 0x80          # same x, opposite "larger" flag
    assert wire.compress(stark.INF) == bytes(31) + b"\x40"
    deck = [(pts[2 * i], pts[2 * i + 1]) for i in range(5)] + [(stark.INF, pts[0])]
    ser = wire.deck_serialize(deck)
    assert len(ser) == 8 + 64 * 6 and wire.deck_deserialize(ser) == deck


def test_malformed_encodings_are_rejected():
    _, _, pts, _ = chain_points(2, 6)
    good = bytearray(wire.compress(pts[0]))
    # an x that is not on the curve: walk until x^3 + x + b is a non-residue
    x = pts[0][0]
    while stark.fq_sqrt((x ** 3 + x + stark.B) % stark.P) is not None:
        x += 1
    with pytest.raises(ValueError):
        wire.decompress(stark.fe_to_bytes(x))
    with pytest.raises(ValueError):
        wire.decompress(stark.fe_to_bytes(stark.P)[:31] + bytes([stark.fe_to_bytes(stark.P)[31]]))   # x = p, not canonical
    with pytest.raises(ValueError):
        wire.decompress(bytes(31) + b"\xc0")                        # infinity with the sign flag
    with pytest.raises(ValueError):
        wire.decompress(b"\x01" + bytes(30) + b"\x40")              # infinity with x != 0
    with pytest.raises(ValueError):
        wire.deck_deserialize((3).to_bytes(8, "little") + bytes(good) * 4)


def test_golden_vectors():
    for fx in GOLD["points"]:
        p = stark.point_from_bytes64(h(fx["point"]))
        assert wire.compress(p).hex() == fx["compressed"] and wire.decompress(h(fx["compressed"])) == p
    deck = wire.deck_deserialize(h(GOLD["deck_serialized"]))
    assert b"".join(stark.point_to_bytes64(a) + stark.point_to_bytes64(b) for a, b in deck).hex() == GOLD["deck"]
    for bad in GOLD["rejected"]:
        with pytest.raises(ValueError):
            wire.decompress(h(bad))


def test_c_oracle_matches_python():
    from oracle import c_oracle
    co = c_oracle.COracle(threads=2)
    pts = b"".join(h(fx["point"]) for fx in GOLD["points"])
    comp = b"".join(h(fx["compressed"]) for fx in GOLD["points"])
    assert co.points_compress(pts) == comp
    out, st = co.points_decompress(comp + b"".join(h(b) for b in GOLD["rejected"]))
    assert out[:len(pts)] == pts and st[:len(GOLD["points"])] == [0] * len(GOLD["points"])
    assert st[len(GOLD["points"]):] == [2, 2, 2, 1, 1, 1]
