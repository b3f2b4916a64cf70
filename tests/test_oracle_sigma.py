"""Oracle-level restatement of the reference's sigma-protocol tests:
masking.rs:64-107 `test_verify_masking`, remasking.rs:65-114 `test_verify_remasking`,
reveal.rs:43-84 `test_verify_reveal`, tests.rs:48-78 `generate_and_verify_key`, tests.rs:80-123
`aggregate_keys`, tests.rs:125-173 `test_unmask`: prove -> verify == Ok; a wrong statement fails
with "Chaum-Pedersen" / "Schnorr Identification".  Plus the committed golden fixtures."""
import json
import os

from oracle.py import stark, sigma
from _util import chain_points

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sigma_vectors.json")))
h = bytes.fromhex


def setup(seed=21, players=3):
    s0, s1, pts, st = chain_points(8, seed)
    g = stark.G
    sks = [st.scalar() for _ in range(players)]
    pks = [stark.mul(g, sk) for sk in sks]
    shared = stark.INF
    for pk in pks:
        shared = stark.add(shared, pk)
    return g, sks, pks, shared, pts, st


def test_key_ownership_and_aggregate():
    g, sks, pks, shared, pts, st = setup()
    for i, (sk, pk) in enumerate(zip(sks, pks)):
        info = b"player %d" % i
        proof = sigma.prove_key_ownership(g, pk, sk, info, st.scalar())
        assert sigma.verify_key_ownership(g, pk, info, proof) == sigma.OK
        # tests.rs:72-77: a proof made with another secret key fails
        bad = sigma.prove_key_ownership(g, pk, (sk + 1) % stark.N, info, st.scalar())
        stt = sigma.verify_key_ownership(g, pk, info, bad)
        assert stt == sigma.ERR_SCHNORR and sigma.ERR_STRINGS[stt] == "Schnorr Identification"
        # the public info is bound into the transcript seed (mod.rs:139-140)
        assert sigma.verify_key_ownership(g, pk, b"someone else", proof) == sigma.ERR_SCHNORR
    assert shared == stark.mul(g, sum(sks) % stark.N)  # tests.rs:104-107


def test_mask_remask_reveal_round_trip_and_negative_cases():
    g, sks, pks, shared, pts, st = setup()
    card, other = pts[0], pts[1]
    r = st.scalar()
    masked, proof = sigma.mask(g, shared, card, r, st.scalar())
    assert masked == (stark.mul(g, r), stark.add(card, stark.mul(shared, r)))  # masking.rs:10-20
    assert sigma.verify_mask(g, shared, card, masked, proof) == sigma.OK
    wrong = (pts[2], pts[3])
    stt = sigma.verify_mask(g, shared, card, wrong, proof)  # masking.rs:96-105
    assert stt == sigma.ERR_CHAUM_PEDERSEN and sigma.ERR_STRINGS[stt] == "Chaum-Pedersen"
    assert sigma.verify_mask(g, shared, other, masked, proof) == sigma.ERR_CHAUM_PEDERSEN

    alpha = st.scalar()
    remasked, rproof = sigma.remask(g, shared, masked, alpha, st.scalar())
    assert remasked == (stark.mul(g, (r + alpha) % stark.N), stark.add(card, stark.mul(shared, (r + alpha) % stark.N)))
    assert sigma.verify_remask(g, shared, masked, remasked, rproof) == sigma.OK
    assert sigma.verify_remask(g, shared, masked, wrong, rproof) == sigma.ERR_CHAUM_PEDERSEN  # remasking.rs:103-112

    # reveal.rs:43-84 and tests.rs:125-173: tokens of all players unmask the card
    acc = stark.INF
    for sk, pk in zip(sks, pks):
        token, tproof = sigma.compute_reveal_token(g, sk, pk, remasked, st.scalar())
        assert sigma.verify_reveal(g, pk, token, remasked, tproof) == sigma.OK
        assert sigma.verify_reveal(g, pk, pts[4], remasked, tproof) == sigma.ERR_CHAUM_PEDERSEN  # reveal.rs:73-82
        acc = stark.add(acc, token)
    assert stark.sub(remasked[1], acc) == card  # unmask, mod.rs:356-378


def test_identity_card_and_zero_scalars():
    g, sks, pks, shared, pts, st = setup(seed=22)
    masked, proof = sigma.mask(g, shared, stark.INF, 0, 0)  # everything degenerates to the identity
    assert masked == (stark.INF, stark.INF)
    assert sigma.verify_mask(g, shared, stark.INF, masked, proof) == sigma.OK
    remasked, rproof = sigma.remask(g, shared, masked, stark.N - 1, st.scalar())
    assert sigma.verify_remask(g, shared, masked, remasked, rproof) == sigma.OK


def test_golden_fixtures():
    g = stark.point_from_bytes64(h(GOLD["g"]))
    shared = stark.point_from_bytes64(h(GOLD["shared_key"]))
    P = stark.point_from_bytes64
    for fx in GOLD["mask"]:
        card, r, omega = P(h(fx["card"])), int(fx["r"], 16), int(fx["omega"], 16)
        masked, proof = sigma.mask(g, shared, card, r, omega)
        assert (stark.point_to_bytes64(masked[0]) + stark.point_to_bytes64(masked[1])).hex() == fx["masked"]
        assert sigma.cp_proof_bytes(proof).hex() == fx["proof"]
        assert sigma.verify_mask(g, shared, card, masked, sigma.cp_proof_from_bytes(h(fx["proof"]))) == sigma.OK
    for fx in GOLD["remask"]:
        orig = (P(h(fx["original"])[:64]), P(h(fx["original"])[64:]))
        remasked, proof = sigma.remask(g, shared, orig, int(fx["alpha"], 16), int(fx["omega"], 16))
        assert (stark.point_to_bytes64(remasked[0]) + stark.point_to_bytes64(remasked[1])).hex() == fx["remasked"]
        assert sigma.cp_proof_bytes(proof).hex() == fx["proof"]
    for fx in GOLD["reveal"]:
        masked = (P(h(fx["masked"])[:64]), P(h(fx["masked"])[64:]))
        sk, pk = int(fx["sk"], 16), P(h(fx["pk"]))
        token, proof = sigma.compute_reveal_token(g, sk, pk, masked, int(fx["omega"], 16))
        assert stark.point_to_bytes64(token).hex() == fx["token"] and sigma.cp_proof_bytes(proof).hex() == fx["proof"]
        assert sigma.verify_reveal(g, pk, token, masked, proof) == sigma.OK
    for fx in GOLD["key_ownership"]:
        sk, pk, info = int(fx["sk"], 16), P(h(fx["pk"])), h(fx["info"])
        proof = sigma.prove_key_ownership(g, pk, sk, info, int(fx["omega"], 16))
        assert sigma.schnorr_proof_bytes(proof).hex() == fx["proof"]
        assert sigma.verify_key_ownership(g, pk, info, proof) == sigma.OK


def test_c_oracle_matches_golden_and_python():
    """oracle/c (the CPU baseline of the batched entry points) == oracle/py byte for byte."""
    from oracle import c_oracle
    co = c_oracle.COracle(threads=2)
    g, shared = h(GOLD["g"]), h(GOLD["shared_key"])
    cat = lambda key, rows: b"".join(h(r[key]) for r in rows)
    M, R, V, K = GOLD["mask"], GOLD["remask"], GOLD["reveal"], GOLD["key_ownership"]
    masked, proofs = co.mask_batch(g, shared, cat("card", M), b"".join(int(r["r"], 16).to_bytes(32, "little") for r in M),
                                   b"".join(int(r["omega"], 16).to_bytes(32, "little") for r in M))
    assert masked == cat("masked", M) and proofs == cat("proof", M)
    assert co.verify_mask_batch(g, shared, cat("card", M), masked, proofs) == [0] * len(M)
    bad = bytearray(proofs); bad[160 + 130] ^= 1                       # response scalar of proof 1
    swapped = masked[128:256] + masked[:128] + masked[256:]             # statements 0 and 1 exchanged
    assert co.verify_mask_batch(g, shared, cat("card", M), masked, bytes(bad)) == [0, 5] + [0] * (len(M) - 2)
    assert co.verify_mask_batch(g, shared, cat("card", M), swapped, proofs)[:2] == [5, 5]
    le = lambda key, rows: b"".join(int(r[key], 16).to_bytes(32, "little") for r in rows)
    out, rproofs = co.remask_prove_batch(g, shared, cat("original", R), le("alpha", R), le("omega", R))
    assert out == cat("remasked", R) and rproofs == cat("proof", R)
    assert co.verify_remask_batch(g, shared, cat("original", R), out, rproofs) == [0] * len(R)
    for fx in V:                                                        # one player per call
        tok, pf = co.reveal_batch(g, int(fx["sk"], 16).to_bytes(32, "little"), h(fx["pk"]), h(fx["masked"]),
                                  int(fx["omega"], 16).to_bytes(32, "little"))
        assert tok == h(fx["token"]) and pf == h(fx["proof"])
        assert co.verify_reveal_batch(g, h(fx["pk"]), tok, h(fx["masked"]), pf) == [0]
        assert co.verify_reveal_batch(g, h(fx["pk"]), h(V[0]["pk"]), h(fx["masked"]), pf) == [5]
    infos = [h(r["info"]) for r in K]
    kp = co.key_ownership_prove_batch(g, cat("pk", K), le("sk", K), infos, le("omega", K))
    assert kp == cat("proof", K)
    assert co.key_ownership_verify_batch(g, cat("pk", K), infos, kp) == [0] * len(K)
    assert co.key_ownership_verify_batch(g, cat("pk", K), infos[::-1], kp) == [6, 0, 6]
