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


def test_library_serialisers_without_a_gpu(pkg):
    """mp_points_compress / mp_deck_serialize / mp_proof_serialize are host byte handling (csrc/wire_host.hpp): the
    built library itself, no context, against the golden vectors and the oracle's encoder."""
    import ctypes
    import json as _json
    import os as _os
    from oracle.py import stark as _stark, wire as _wire, bayer_groth as _bg
    here = _os.path.dirname(__file__)
    gold = _json.load(open(_os.path.join(here, "golden", "wire_vectors.json")))
    lib = pkg.lib
    pts = b"".join(bytes.fromhex(fx["point"]) for fx in gold["points"])
    comp = b"".join(bytes.fromhex(fx["compressed"]) for fx in gold["points"])
    out = ctypes.create_string_buffer(len(comp))
    assert lib.mp_points_compress(pts, len(pts) // 64, out) == 0 and out.raw == comp
    deck = bytes.fromhex(gold["deck"])
    out = ctypes.create_string_buffer(lib.mp_deck_serialized_len(len(deck) // 128))
    assert lib.mp_deck_serialize(deck, len(deck) // 128, out) == 0 and out.raw == bytes.fromhex(gold["deck_serialized"])
    shuf = _json.load(open(_os.path.join(here, "golden", "oracle_vectors.json")))["shuffle"]
    for fx in shuf[:2]:
        m, n, proof = fx["m"], fx["n"], bytes.fromhex(fx["proof"])
        assert lib.mp_proof_serialized_len(m, n) == (11 * m + 8) * 32 + (5 * n + 9) * 32
        out = ctypes.create_string_buffer(lib.mp_proof_serialized_len(m, n))
        assert lib.mp_proof_serialize(m, n, proof, out) == 0
        # every point of the flat proof compressed in place, scalars copied: re-derive with the oracle's encoder
        runs = [(True, 5 * m + 4), (False, 2 * n + 3), (True, 3), (False, 2 * n + 2), (True, 6 * m + 1), (False, n + 4)]
        want, pos = b"", 0
        for is_pts, count in runs:
            for _ in range(count):
                if is_pts:
                    want += _wire.compress(_stark.point_from_bytes64(proof[pos:pos + 64]))
                    pos += 64
                else:
                    want += proof[pos:pos + 32]
                    pos += 32
        assert out.raw == want
