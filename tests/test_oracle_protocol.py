"""Oracle-level restatement of the reference's only hot-path test, `test_shuffle`
(reference barnett-smart-card-protocol/src/discrete_log_cards/tests.rs:175-227): prove ->
verify == Ok; a wrong output deck fails with "Hadamard Product (5.1)".  Plus tampering cases
for every sub-argument and the committed golden fixtures."""
import copy
import json
import os

import pytest

from oracle.py import stark, bayer_groth as bg
from _util import instance, pb, b32, chain_points

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.json")))


@pytest.fixture(scope="module")
def small():
    pp, pk, deck, perm, rho, rnd = instance(3, 4, 2)
    deck2, proof = bg.shuffle_and_remask(pp, pk, deck, rho, perm, rnd)
    return pp, pk, deck, perm, rho, rnd, deck2, proof


def test_round_trip_and_remask_semantics(small):
    pp, pk, deck, perm, rho, rnd, deck2, proof = small
    assert bg.shuffle_verify(pp, pk, deck, deck2, proof) == bg.OK
    # remasking.rs:9-22 / mod.rs:388-395: out[i] = deck[mapping[i]] + (rho_i*g, rho_i*pk)
    for i in range(len(deck)):
        src = deck[perm[i]]
        assert deck2[i] == (stark.add(src[0], stark.mul(pp.enc_g, rho[i])), stark.add(src[1], stark.mul(pk, rho[i])))


def test_wrong_output_deck_fails_in_hadamard(small):
    # tests.rs:213-226: a fresh random output deck => Err("Hadamard Product (5.1)")
    pp, pk, deck, perm, rho, rnd, deck2, proof = small
    _, _, pts, _ = chain_points(2 * len(deck), 99)
    wrong = [(pts[2 * i], pts[2 * i + 1]) for i in range(len(deck))]
    st = bg.shuffle_verify(pp, pk, deck, wrong, proof)
    assert st == bg.ERR_HADAMARD and bg.ERR_STRINGS[st] == "Hadamard Product (5.1)"


def test_tampering_is_caught_by_the_right_sub_argument(small):
    pp, pk, deck, perm, rho, rnd, deck2, proof = small
    p = copy.deepcopy(proof)
    p["product"]["hadamard"]["zero"]["t"] = (p["product"]["hadamard"]["zero"]["t"] + 1) % stark.N
    assert bg.shuffle_verify(pp, pk, deck, deck2, p) == bg.ERR_ZERO
    p = copy.deepcopy(proof)
    p["product"]["svp"]["r"] = (p["product"]["svp"]["r"] + 1) % stark.N
    assert bg.shuffle_verify(pp, pk, deck, deck2, p) == bg.ERR_SVP
    p = copy.deepcopy(proof)
    p["multiexp"]["tau"] = (p["multiexp"]["tau"] + 1) % stark.N
    assert bg.shuffle_verify(pp, pk, deck, deck2, p) == bg.ERR_MULTIEXP
    # a shuffled deck that is a valid remask under a DIFFERENT permutation fails too
    perm2 = perm[1:] + perm[:1]
    other = bg.shuffle_and_remask_deck(pp, pk, deck, rho, perm2)
    assert bg.shuffle_verify(pp, pk, deck, other, proof) != bg.OK


def test_proof_byte_layout_roundtrip(small):
    pp, pk, deck, perm, rho, rnd, deck2, proof = small
    buf = bg.proof_to_bytes(proof)
    assert len(buf) == bg.proof_len(pp.m, pp.n) == (11 * pp.m + 8) * 64 + (5 * pp.n + 9) * 32
    assert bg.proof_to_bytes(bg.proof_from_bytes(buf, pp.m, pp.n)) == buf


def _load(fx):
    m, n = fx["m"], fx["n"]
    P = lambda h: stark.point_from_bytes64(bytes.fromhex(h))
    ck = bytes.fromhex(fx["ck_g"])
    pp = bg.Params(m, n, P(fx["enc_g"]), [stark.point_from_bytes64(ck[64 * i:64 * i + 64]) for i in range(n)],
                   P(fx["ck_h"]), P(fx["ghat"]))
    dk = lambda h: [(stark.point_from_bytes64(h[128 * i:128 * i + 64]), stark.point_from_bytes64(h[128 * i + 64:128 * i + 128]))
                    for i in range(m * n)]
    sc = lambda h: [int.from_bytes(h[32 * i:32 * i + 32], "little") for i in range(len(h) // 32)]
    return (pp, P(fx["pk"]), dk(bytes.fromhex(fx["deck"])), fx["perm"], sc(bytes.fromhex(fx["rho"])),
            sc(bytes.fromhex(fx["rand"])), dk(bytes.fromhex(fx["deck2"])), bytes.fromhex(fx["proof"]))


@pytest.mark.parametrize("idx", [0, 1])
def test_golden_prove_is_reproduced(idx):
    fx = GOLD["shuffle"][idx]
    pp, pk, deck, perm, rho, rnd, deck2, proof = _load(fx)
    d2, pf = bg.shuffle_and_remask(pp, pk, deck, rho, perm, rnd)
    assert d2 == deck2 and bg.proof_to_bytes(pf) == proof


@pytest.mark.parametrize("idx", [2, 3])
def test_golden_52_card_proofs_verify(idx):
    # (4,13) = tests.rs:178-179 and (2,26) = examples/round.rs:229-230
    fx = GOLD["shuffle"][idx]
    pp, pk, deck, perm, rho, rnd, deck2, proof = _load(fx)
    assert pp.m * pp.n == 52
    assert bg.shuffle_verify(pp, pk, deck, deck2, bg.proof_from_bytes(proof, pp.m, pp.n)) == bg.OK


def test_golden_msm():
    for fx in GOLD["msm"]:
        n = fx["n"]
        pts = bytes.fromhex(fx["points"])
        ks = bytes.fromhex(fx["scalars"])
        got = stark.msm([stark.point_from_bytes64(pts[64 * i:64 * i + 64]) for i in range(n)],
                        [int.from_bytes(ks[32 * i:32 * i + 32], "little") for i in range(n)])
        assert pb(got).hex() == fx["result"]
