"""BASELINE configs 1-2 end to end: the reference's `examples/round.rs` (m = 2, n = 26, four players)
replayed through the C ABI -- key-ownership proofs and the aggregate key (round.rs:237-250), 52 cards
masked with r = 1 (round.rs:253-256), four chained shuffle_and_remask + verify_shuffle (round.rs:268-350),
reveal tokens with their proofs and the opening of four cards (round.rs:360-428) -- with the oracle
checking the bytes of every step (oracle/c for the shuffles, oracle/py for the sigma protocols)."""
import pytest

from oracle import c_oracle
from oracle.py import stark, sigma, bayer_groth as bg
from _util import chain_points, b32, pb

pytestmark = pytest.mark.gpu


def test_round_example_end_to_end(ctx):
    m, n = 2, 26
    N = m * n
    s0, s1, pts, st = chain_points(n + 2 + N, 2024)
    g = stark.G
    ck_g, ck_h, ghat, cards = pts[:n], pts[n], pts[n + 1], pts[n + 2:]
    G64, enc = pb(g), (b"".join(map(pb, ck_g)), pb(ck_h), pb(ghat))
    ctx.set_params(m, n, G64, *enc)
    co = c_oracle.COracle(msm_mode=1)
    names = [b"Andrija", b"Kobi", b"Nico", b"Tom"]

    # ---- players: keygen + Schnorr key-ownership proofs; everyone checks and aggregates (round.rs:237-250)
    sks = [st.scalar() for _ in names]
    pks = [stark.mul(g, sk) for sk in sks]
    pks_b, sks_b = b"".join(map(pb, pks)), b"".join(map(b32, sks))
    om = [st.scalar() for _ in names]
    kproofs = ctx.key_ownership_prove_batch(pks_b, sks_b, names, b"".join(map(b32, om)))
    for i in range(4):
        assert kproofs[96 * i:96 * i + 96] == sigma.schnorr_proof_bytes(sigma.prove_key_ownership(g, pks[i], sks[i], names[i], om[i]))
    assert ctx.key_ownership_verify_batch(pks_b, names, kproofs) == [0] * 4
    joint = stark.INF
    for pk in pks:
        joint = stark.add(joint, pk)
    joint_b = pb(joint)

    # ---- the initial deck: every card masked with r = 1, proofs checked by everyone (round.rs:253-256)
    cards_b, ones = b"".join(map(pb, cards)), b32(1) * N
    om = [st.scalar() for _ in range(N)]
    deck_b, mproofs = ctx.mask_batch(joint_b, cards_b, ones, b"".join(map(b32, om)))
    for i in (0, 17, N - 1):
        masked, proof = sigma.mask(g, joint, cards[i], 1, om[i])
        assert deck_b[128 * i:128 * i + 128] == pb(masked[0]) + pb(masked[1]) and mproofs[160 * i:160 * i + 160] == sigma.cp_proof_bytes(proof)
    assert ctx.verify_mask_batch(joint_b, cards_b, deck_b, mproofs) == [0] * N

    # ---- four chained shuffles, each verified (round.rs:268-350)
    perms = []
    for player in range(4):
        perm = st.permutation(N)
        rho = b"".join(b32(st.scalar()) for _ in range(N))
        rnd = b"".join(b32(st.scalar()) for _ in range(bg.prover_randomness_len(m, n)))
        deck2, proof = ctx.shuffle_and_remask(joint_b, deck_b, perm, rho, rnd)
        assert deck2 == co.remask(G64, joint_b, deck_b, perm, rho)
        assert proof == co.prove(m, n, G64, *enc, joint_b, deck_b, deck2, perm, rho, rnd)
        assert ctx.verify_shuffle(joint_b, deck_b, deck2, proof) == 0
        assert co.verify(m, n, G64, *enc, joint_b, deck_b, deck2, proof) == 0
        assert ctx.verify_shuffle(joint_b, deck2, deck_b, proof) != 0      # the statement is ordered
        deck_b, perms = deck2, perms + [perm]

    # ---- the round: cards 0..3 are dealt; every player publishes reveal tokens for them (round.rs:360-428)
    dealt = deck_b[:128 * 4]
    tokens = []
    for p in range(4):
        om = [st.scalar() for _ in range(4)]
        tok, tproofs = ctx.reveal_batch(b32(sks[p]), pb(pks[p]), dealt, b"".join(map(b32, om)))
        c1 = stark.point_from_bytes64(dealt[128:192])
        want_tok, want_proof = sigma.compute_reveal_token(g, sks[p], pks[p], (c1, None), om[1])
        assert tok[64:128] == pb(want_tok) and tproofs[160:320] == sigma.cp_proof_bytes(want_proof)
        assert ctx.verify_reveal_batch(pb(pks[p]), tok, dealt, tproofs) == [0] * 4
        other = pb(pks[(p + 1) % 4])
        assert ctx.verify_reveal_batch(other, tok, dealt, tproofs) == [5] * 4   # someone else's key: "Chaum-Pedersen"
        tokens.append(tok)
    # unmask (mod.rs:356-378): card = c2 - sum of tokens; trace it back through the four permutations
    src = list(range(N))
    for perm in perms:
        src = [src[perm[i]] for i in range(N)]
    for k in range(4):
        acc = stark.INF
        for p in range(4):
            acc = stark.add(acc, stark.point_from_bytes64(tokens[p][64 * k:64 * k + 64]))
        c2 = stark.point_from_bytes64(dealt[128 * k + 64:128 * k + 128])
        assert stark.sub(c2, acc) == cards[src[k]]
