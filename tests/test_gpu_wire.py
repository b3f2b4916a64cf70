"""GPU parity of the wire format through the C ABI: decompression against
the committed golden vectors and the C oracle (bit-exact, including the rejected encodings and their
status codes), deck / proof serialisation round trips that still verify, and a 2^17-point round trip."""
import json
import os

import numpy as np
import pytest

from oracle import c_oracle
from oracle.py import stark, wire
from _util import pb

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "wire_vectors.json")))
SHUF = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.json")))["shuffle"]
h = bytes.fromhex
G64 = pb(stark.G)


def test_golden_points_deck_and_rejections(ctx, pkg):
    pts = b"".join(h(fx["point"]) for fx in GOLD["points"])
    comp = b"".join(h(fx["compressed"]) for fx in GOLD["points"])
    assert ctx.points_compress(pts) == comp
    assert ctx.points_decompress(comp) == pts
    assert ctx.launches > 0
    assert ctx.deck_serialize(h(GOLD["deck"])) == h(GOLD["deck_serialized"])
    assert ctx.deck_deserialize(h(GOLD["deck_serialized"])) == h(GOLD["deck"])
    bad = comp + b"".join(h(b) for b in GOLD["rejected"])
    out, st, rc = ctx.points_decompress(bad, want_statuses=True)
    k = len(GOLD["points"])
    assert rc == -3 and st[:k] == [0] * k and st[k:] == [2, 2, 2, 1, 1, 1]
    assert out[:len(pts)] == pts and out[len(pts):] == bytes(64 * 6)
    with pytest.raises(pkg.MpError):
        ctx.points_decompress(bad)
    with pytest.raises(pkg.MpError):                       # length prefix says 9 cards, buffer holds 8
        ctx.deck_deserialize((9).to_bytes(8, "little") + h(GOLD["deck_serialized"])[8:])
    assert ctx.points_decompress(b"") == b"" and ctx.deck_deserialize(bytes(8)) == b""


@pytest.mark.parametrize("idx", [0, 2, 3])
def test_serialised_proof_and_decks_still_verify(ctx, idx):
    fx = SHUF[idx]
    m, n = fx["m"], fx["n"]
    ctx.set_params(m, n, h(fx["enc_g"]), h(fx["ck_g"]), h(fx["ck_h"]), h(fx["ghat"]))
    wire_proof = ctx.proof_serialize(m, n, h(fx["proof"]))
    assert len(wire_proof) == (11 * m + 8) * 32 + (5 * n + 9) * 32
    proof = ctx.proof_deserialize(m, n, wire_proof)
    assert proof == h(fx["proof"])
    deck = ctx.deck_deserialize(ctx.deck_serialize(h(fx["deck"])))
    deck2 = ctx.deck_deserialize(ctx.deck_serialize(h(fx["deck2"])))
    assert deck == h(fx["deck"]) and deck2 == h(fx["deck2"])
    assert ctx.verify_shuffle(h(fx["pk"]), deck, deck2, proof) == 0
    # the oracle's encoder agrees with the library's on the deck
    cards = [(stark.point_from_bytes64(deck[128 * i:128 * i + 64]), stark.point_from_bytes64(deck[128 * i + 64:128 * i + 128]))
             for i in range(m * n)]
    assert wire.deck_serialize(cards) == ctx.deck_serialize(deck)


def test_decompress_2p17_points_and_c_oracle_sample(ctx):
    n = 1 << 17
    rng = np.random.default_rng(21)
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[:, 31] &= 0x07
    pts = ctx.dbg_scalar_mul(G64 * n, a.tobytes())
    comp = ctx.points_compress(pts)
    assert ctx.points_decompress(comp) == pts
    co = c_oracle.COracle(threads=4)
    assert co.points_compress(pts[:64 * 300]) == comp[:32 * 300]
    out, st = co.points_decompress(comp[:32 * 300])
    assert out == pts[:64 * 300] and st == [0] * 300
    # random 32-byte strings: about half are abscissas of curve points; statuses must match the oracle's
    junk = rng.integers(0, 256, size=(400, 32), dtype=np.uint8)
    junk[:, 31] &= 0x87                                         # keep x below 2^251, random sign flag
    got, gst, rc = ctx.points_decompress(junk.tobytes(), want_statuses=True)
    want, wst = co.points_decompress(junk.tobytes())
    assert gst == wst and got == want and 100 < wst.count(0) < 300 and rc == -3
