"""GPU parity of the batched sigma protocols either side of the shuffle
through the C ABI: byte-exact against the committed golden vectors and the C oracle, the reference's
negative cases (masking.rs:96-105, remasking.rs:103-112, reveal.rs:73-82, tests.rs:72-77), edge
inputs (identity cards, zero / maximal scalars), and a 2^16-item round trip."""
import json
import os

import numpy as np
import pytest

from oracle import c_oracle
from oracle.py import stark
from _util import b32, pb

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sigma_vectors.json")))
SHUF = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.json")))["shuffle"][0]
h = bytes.fromhex
G64 = pb(stark.G)
cat = lambda key, rows: b"".join(h(r[key]) for r in rows)
le = lambda key, rows: b"".join(int(r[key], 16).to_bytes(32, "little") for r in rows)


@pytest.fixture
def sctx(ctx):
    # any commitment key will do: the sigma protocols only use the ElGamal generator g
    ctx.set_params(SHUF["m"], SHUF["n"], h(GOLD["g"]), h(SHUF["ck_g"]), h(SHUF["ck_h"]), h(SHUF["ghat"]))
    return ctx


def test_mask_remask_golden_and_negative_cases(sctx):
    shared = h(GOLD["shared_key"])
    M, R = GOLD["mask"], GOLD["remask"]
    masked, proofs = sctx.mask_batch(shared, cat("card", M), le("r", M), le("omega", M))
    assert masked == cat("masked", M) and proofs == cat("proof", M)
    assert sctx.launches > 0
    assert sctx.verify_mask_batch(shared, cat("card", M), masked, proofs) == [0] * len(M)
    bad = bytearray(proofs); bad[160 + 130] ^= 1
    assert sctx.verify_mask_batch(shared, cat("card", M), masked, bytes(bad)) == [0, 5] + [0] * (len(M) - 2)
    swapped = masked[128:256] + masked[:128] + masked[256:]
    st = sctx.verify_mask_batch(shared, cat("card", M), swapped, proofs)
    assert st[:2] == [5, 5] and sctx.status_string(st[0]) == "Chaum-Pedersen"
    big = bytearray(proofs); big[128:160] = (stark.N + 5).to_bytes(32, "little")   # non-canonical response
    assert sctx.verify_mask_batch(shared, cat("card", M), masked, bytes(big))[0] == 5
    out, rproofs = sctx.remask_prove_batch(shared, cat("original", R), le("alpha", R), le("omega", R))
    assert out == cat("remasked", R) and rproofs == cat("proof", R)
    assert sctx.verify_remask_batch(shared, cat("original", R), out, rproofs) == [0] * len(R)
    assert sctx.verify_remask_batch(shared, cat("original", R), out[128:] + out[:128], rproofs).count(5) >= len(R) - 1


def test_reveal_and_key_ownership_golden_and_negative_cases(sctx):
    V, K = GOLD["reveal"], GOLD["key_ownership"]
    for fx in V:
        sk, om = int(fx["sk"], 16).to_bytes(32, "little"), int(fx["omega"], 16).to_bytes(32, "little")
        tok, pf = sctx.reveal_batch(sk, h(fx["pk"]), h(fx["masked"]), om)
        assert tok == h(fx["token"]) and pf == h(fx["proof"])
        assert sctx.verify_reveal_batch(h(fx["pk"]), tok, h(fx["masked"]), pf) == [0]
        assert sctx.verify_reveal_batch(h(fx["pk"]), h(V[0]["pk"]), h(fx["masked"]), pf) == [5]
    infos = [h(r["info"]) for r in K]
    kp = sctx.key_ownership_prove_batch(cat("pk", K), le("sk", K), infos, le("omega", K))
    assert kp == cat("proof", K)
    assert sctx.key_ownership_verify_batch(cat("pk", K), infos, kp) == [0] * len(K)
    st = sctx.key_ownership_verify_batch(cat("pk", K), infos[::-1], kp)
    assert st == [6, 0, 6] and sctx.status_string(6) == "Schnorr Identification"


def test_malformed_items_fail_alone(sctx):
    """A point off the curve (or a non-canonical coordinate) in ONE item of a verifier batch gives that item status 7
    ("malformed", what the reference's deserialiser would refuse) and leaves the verdicts of the others untouched; a
    bad key -- shared by the whole call -- fails the call."""
    shared = h(GOLD["shared_key"])
    M, R, V, K = GOLD["mask"], GOLD["remask"], GOLD["reveal"], GOLD["key_ownership"]
    cards, masked, proofs = cat("card", M), cat("masked", M), cat("proof", M)
    n = len(M)
    for buf_idx, off in ((0, 64 * 2 + 5), (1, 128 * 3 + 70), (2, 160 * 1 + 3), (2, 160 * 4 + 64 + 9)):
        bufs = [bytearray(cards), bytearray(masked), bytearray(proofs)]
        bufs[buf_idx][off] ^= 1
        item = off // (64, 128, 160)[buf_idx]
        st = sctx.verify_mask_batch(shared, bytes(bufs[0]), bytes(bufs[1]), bytes(bufs[2]))
        assert st == [7 if i == item else 0 for i in range(n)], (buf_idx, off)
    big_x = bytearray(masked); big_x[128 * 2:128 * 2 + 32] = (stark.P + 1).to_bytes(32, "little")   # x >= p
    assert sctx.verify_mask_batch(shared, cards, bytes(big_x), proofs)[2] == 7
    orig, out, rp = cat("original", R), cat("remasked", R), cat("proof", R)
    bad = bytearray(orig); bad[128 * 1 + 64 + 1] ^= 1
    assert sctx.verify_remask_batch(shared, bytes(bad), out, rp) == [7 if i == 1 else 0 for i in range(len(R))]
    fx = V[0]
    tok = bytearray(h(fx["token"])); tok[40] ^= 1
    assert sctx.verify_reveal_batch(h(fx["pk"]), bytes(tok), h(fx["masked"]), h(fx["proof"])) == [7]
    kp = bytearray(cat("proof", K)); kp[96 * 2 + 7] ^= 1
    assert sctx.key_ownership_verify_batch(cat("pk", K), [h(r["info"]) for r in K], bytes(kp)) == [0, 0, 7]
    badkey = bytearray(shared); badkey[3] ^= 1
    with pytest.raises(Exception) as e:
        sctx.verify_mask_batch(bytes(badkey), cards, masked, proofs)
    assert e.value.code == -3


def rand_scalars(rng, k):
    a = rng.integers(0, 256, size=(k, 32), dtype=np.uint8)
    a[:, 31] &= 0x07
    return a.tobytes()


def test_batch_of_500_vs_c_oracle_with_edge_items(sctx):
    n = 500
    rng = np.random.default_rng(12)
    co = c_oracle.COracle(threads=4)
    sks = rand_scalars(rng, 3)
    pks = sctx.dbg_scalar_mul(G64 * 3, sks)
    shared = sctx.dbg_point_add(sctx.dbg_point_add(pks[:64], pks[64:128]), pks[128:])
    cards = bytearray(sctx.dbg_scalar_mul(G64 * n, rand_scalars(rng, n)))
    r, om = bytearray(rand_scalars(rng, n)), bytearray(rand_scalars(rng, n))
    cards[0:64] = bytes(64)                                   # identity card
    r[32:64] = bytes(32)                                      # r = 0
    r[64:96] = (stark.N - 1).to_bytes(32, "little")           # r = n - 1
    om[96:128] = bytes(32)                                    # omega = 0
    cards[256:320] = cards[192:256]                           # duplicate card
    cards, r, om = bytes(cards), bytes(r), bytes(om)
    masked, proofs = sctx.mask_batch(shared, cards, r, om, host_threads=3)
    want_m, want_p = co.mask_batch(G64, shared, cards, r, om)
    assert masked == want_m and proofs == want_p
    assert sctx.verify_mask_batch(shared, cards, masked, proofs) == [0] * n == co.verify_mask_batch(G64, shared, cards, masked, proofs)
    alpha, om2 = rand_scalars(rng, n), rand_scalars(rng, n)
    out, rproofs = sctx.remask_prove_batch(shared, masked, alpha, om2)
    assert (out, rproofs) == co.remask_prove_batch(G64, shared, masked, alpha, om2)
    assert sctx.verify_remask_batch(shared, masked, out, rproofs) == [0] * n
    om3 = rand_scalars(rng, n)
    tok, tproofs = sctx.reveal_batch(sks[:32], pks[:64], out, om3)
    assert (tok, tproofs) == co.reveal_batch(G64, sks[:32], pks[:64], out, om3)
    assert sctx.verify_reveal_batch(pks[:64], tok, out, tproofs) == [0] * n
    # tamper a scattering of items: statuses must single them out exactly like the oracle
    bad = bytearray(tproofs)
    hit = sorted(rng.choice(n, size=17, replace=False).tolist())
    for i in hit:
        bad[160 * i + 128 + int(rng.integers(0, 16))] ^= 1 << int(rng.integers(0, 8))   # response scalar, stays canonical
    got = sctx.verify_reveal_batch(pks[:64], tok, out, bytes(bad))
    assert got == co.verify_reveal_batch(G64, pks[:64], tok, out, bytes(bad))
    assert [i for i, s in enumerate(got) if s] == hit
    # a proof point swapped for another valid curve point
    bad = bytearray(tproofs); bad[160 * 7:160 * 7 + 64] = tproofs[160 * 8:160 * 8 + 64]
    assert sctx.verify_reveal_batch(pks[:64], tok, out, bytes(bad)) == co.verify_reveal_batch(G64, pks[:64], tok, out, bytes(bad))
    # off-curve input is refused
    offc = bytearray(cards); offc[70] ^= 1
    with pytest.raises(Exception) as ei:
        sctx.mask_batch(shared, bytes(offc), r, om)
    assert ei.value.code == -3
    assert sctx.mask_batch(shared, b"", b"", b"") == (b"", b"")          # empty batch


def test_full_deck_round_trip_2p16(sctx):
    """65 536 cards: mask -> verify, remask -> verify, reveal -> verify, and unmask semantics on a sample."""
    n = 1 << 16
    rng = np.random.default_rng(13)
    sk = rand_scalars(rng, 1)
    pk = sctx.dbg_scalar_mul(G64, sk)
    cards = sctx.dbg_scalar_mul(G64 * n, rand_scalars(rng, n))
    r = rand_scalars(rng, n)
    masked, proofs = sctx.mask_batch(pk, cards, r, rand_scalars(rng, n))
    assert sctx.verify_mask_batch(pk, cards, masked, proofs) == [0] * n
    out, rproofs = sctx.remask_prove_batch(pk, masked, rand_scalars(rng, n), rand_scalars(rng, n))
    assert sctx.verify_remask_batch(pk, masked, out, rproofs) == [0] * n
    tok, tproofs = sctx.reveal_batch(sk, pk, out, rand_scalars(rng, n))
    assert sctx.verify_reveal_batch(pk, tok, out, tproofs) == [0] * n
    # single-player unmask (mod.rs:356-378): c2 - token == card
    for i in (0, 777, n - 1):
        neg = bytearray(tok[64 * i:64 * i + 64])
        y = (stark.P - int.from_bytes(neg[32:], "little")) % stark.P
        neg[32:] = y.to_bytes(32, "little")
        assert sctx.dbg_point_add(out[128 * i + 64:128 * i + 128], bytes(neg)) == cards[64 * i:64 * i + 64]
    # one wrong statement among 65 536 is found
    wrong = bytearray(out); wrong[128 * 4242:128 * 4243] = out[128 * 4243:128 * 4244]
    st = sctx.verify_remask_batch(pk, masked, bytes(wrong), rproofs)
    assert st[4242] == 5 and st.count(0) == n - 1


def test_scalars_above_the_group_order_act_modulo_n(sctx):
    """Prover scalars are the caller's to reduce, but any 256-bit value must still act as its residue:
    the fixed-base tables cover all 32 bytes and the variable-base window recoding carries out of the
    top nibble."""
    rng = np.random.default_rng(14)
    pk = sctx.dbg_scalar_mul(G64, rand_scalars(rng, 1))
    card = sctx.dbg_scalar_mul(G64, rand_scalars(rng, 1))
    for big in (2 ** 256 - 1, 2 ** 255 + 12345, stark.N, stark.N + 1, 0xF << 252):
        kb = big.to_bytes(32, "little")
        want_c1 = pb(stark.mul(stark.G, big % stark.N))
        masked, _ = sctx.mask_batch(pk, card, kb, b32(7))
        assert masked[:64] == want_c1                                   # fixed base
        c1 = stark.point_from_bytes64(masked[:64]) if big % stark.N else stark.G
        deck = pb(c1) + card
        tok, _ = sctx.reveal_batch(kb, pk, deck, b32(9))                # variable base: token = sk * c1
        assert tok == pb(stark.mul(c1, big % stark.N))
