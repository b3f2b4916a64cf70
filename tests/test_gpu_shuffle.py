"""GPU parity, protocol level: `shuffle_and_remask` / `verify_shuffle` through the C ABI against
the oracle -- the reference's `test_shuffle` (tests.rs:175-227) plus byte-exact proof parity on
identical decks, permutations and randomness."""
import json
import os
import random

import pytest

from oracle import c_oracle
from oracle.py import stark, bayer_groth as bg
from _util import chain_points, instance, pb, b32

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.json")))
h = bytes.fromhex


def setup_ctx(ctx, fx):
    ctx.set_params(fx["m"], fx["n"], h(fx["enc_g"]), h(fx["ck_g"]), h(fx["ck_h"]), h(fx["ghat"]))


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_verify_accepts_golden_proofs(ctx, idx):
    fx = GOLD["shuffle"][idx]
    setup_ctx(ctx, fx)
    assert ctx.verify_shuffle(h(fx["pk"]), h(fx["deck"]), h(fx["deck2"]), h(fx["proof"])) == 0
    assert ctx.launches > 0


@pytest.fixture(params=["host_scalars", "device_scalars"])
def scalar_path(request, monkeypatch):
    """Small decks default to the host-scalar lockstep prover; MP_SMALL_DECK_MAX=0 forces the
    device scalar kernels (the 2^16-card implementation) at the same sizes."""
    if request.param == "device_scalars":
        monkeypatch.setenv("MP_SMALL_DECK_MAX", "0")
    return request.param


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_prove_is_byte_exact_vs_golden(ctx, idx, scalar_path):
    fx = GOLD["shuffle"][idx]
    setup_ctx(ctx, fx)
    deck2, proof = ctx.shuffle_and_remask(h(fx["pk"]), h(fx["deck"]), fx["perm"], h(fx["rho"]), h(fx["rand"]))
    assert deck2.hex() == fx["deck2"]
    assert proof.hex() == fx["proof"]
    assert ctx.verify_shuffle(h(fx["pk"]), h(fx["deck"]), deck2, proof) == 0


def test_reference_test_shuffle_negative_case(ctx):
    # tests.rs:213-226: verifying against a fresh random output deck => "Hadamard Product (5.1)"
    fx = GOLD["shuffle"][2]
    setup_ctx(ctx, fx)
    _, _, pts, _ = chain_points(104, 99)
    wrong = b"".join(pb(p) for p in pts)
    st = ctx.verify_shuffle(h(fx["pk"]), h(fx["deck"]), wrong, h(fx["proof"]))
    assert st == 1 and ctx.status_string(st) == "Hadamard Product (5.1)"


def test_tampering_matches_oracle_verdicts(ctx):
    fx = GOLD["shuffle"][2]
    setup_ctx(ctx, fx)
    co = c_oracle.COracle()
    m, n = fx["m"], fx["n"]
    proof = h(fx["proof"])
    rnd = random.Random(11)
    args = (m, n, h(fx["enc_g"]), h(fx["ck_g"]), h(fx["ck_h"]), h(fx["ghat"]), h(fx["pk"]))
    # flip one bit in every scalar of the proof in turn (scalars stay canonical: low bits only)
    npts = 11 * m + 8
    offsets = []
    pos = 0
    layout = [("P", 2 * m + 1 + m + 2 * m + 3), ("F", 2 * n + 3), ("P", 3), ("F", 2 * n + 2), ("P", 2 * m + 1 + 4 * m), ("F", n + 4)]
    for kind, cnt in layout:
        for _ in range(cnt):
            if kind == "F":
                offsets.append(pos)
            pos += 64 if kind == "P" else 32
    assert pos == len(proof)
    seen = set()
    for off in offsets:
        p2 = bytearray(proof)
        p2[off] ^= 1 << rnd.randrange(8)
        want = co.verify(*args, h(fx["deck"]), h(fx["deck2"]), bytes(p2))
        got = ctx.verify_shuffle(h(fx["pk"]), h(fx["deck"]), h(fx["deck2"]), bytes(p2))
        assert got == want != 0, off
        seen.add(got)
    assert seen == {2, 3, 4}
    # swap two proof points (still on the curve): c_A[0] <-> c_A[1]
    p2 = bytearray(proof)
    p2[0:64], p2[64:128] = proof[64:128], proof[0:64]
    want = co.verify(*args, h(fx["deck"]), h(fx["deck2"]), bytes(p2))
    assert ctx.verify_shuffle(h(fx["pk"]), h(fx["deck"]), h(fx["deck2"]), bytes(p2)) == want != 0
    # a different valid remask of the same deck does not verify under this proof
    perm2 = fx["perm"][1:] + fx["perm"][:1]
    other = ctx.remask(h(fx["pk"]), h(fx["deck"]), perm2, h(fx["rho"]))
    want = co.verify(*args, h(fx["deck"]), other, proof)
    assert ctx.verify_shuffle(h(fx["pk"]), h(fx["deck"]), other, proof) == want != 0


def test_off_curve_inputs_are_rejected(ctx, pkg):
    fx = GOLD["shuffle"][0]
    setup_ctx(ctx, fx)
    deck = bytearray(h(fx["deck"]))
    deck[5] ^= 1
    with pytest.raises(pkg.MpError) as e:
        ctx.verify_shuffle(h(fx["pk"]), bytes(deck), h(fx["deck2"]), h(fx["proof"]))
    assert e.value.code == -3
    proof = bytearray(h(fx["proof"]))
    proof[3] ^= 1
    with pytest.raises(pkg.MpError):
        ctx.verify_shuffle(h(fx["pk"]), h(fx["deck"]), h(fx["deck2"]), bytes(proof))
    with pytest.raises(pkg.MpError):
        ctx.set_params(2, 2, b32(1) + b32(2), h(fx["ck_g"])[:128], h(fx["ck_h"]), h(fx["ghat"]))
    setup_ctx(ctx, fx)


def test_remask_and_commit_primitives(ctx):
    fx = GOLD["shuffle"][1]
    setup_ctx(ctx, fx)
    co = c_oracle.COracle()
    m, n = fx["m"], fx["n"]
    assert ctx.remask(h(fx["pk"]), h(fx["deck"]), fx["perm"], h(fx["rho"])).hex() == fx["deck2"]
    rnd = random.Random(3)
    for length in (n, n - 1, 1, 0):
        k = 5
        vals = [[rnd.randrange(stark.N) for _ in range(length)] for _ in range(k)]
        vals[0] = [0] * length
        blinds = [rnd.randrange(stark.N) for _ in range(k)]
        blinds[1] = 0
        got = ctx.commit_batch(b"".join(b32(v) for row in vals for v in row), b"".join(map(b32, blinds)), length)
        for i in range(k):
            want = co.commit(n, h(fx["ck_g"]), h(fx["ck_h"]), b"".join(map(b32, vals[i])), b32(blinds[i]))
            assert got[64 * i:64 * i + 64] == want, (length, i)


@pytest.mark.parametrize("m,n,seed", [(2, 2, 5), (5, 3, 6), (8, 8, 7), (16, 32, 8)])
def test_round_trip_other_shapes_vs_c_oracle(ctx, m, n, seed, scalar_path):
    """Byte-exact against the C oracle at sizes the Python oracle would take minutes for."""
    pp, pk, deck, perm, rho, rnd = instance(m, n, seed)
    co = c_oracle.COracle(msm_mode=1)
    enc_g, ck_g, ck_h, ghat = pb(pp.enc_g), b"".join(map(pb, pp.ck_g)), pb(pp.ck_h), pb(pp.ghat)
    deck_b = b"".join(pb(c[0]) + pb(c[1]) for c in deck)
    rho_b, rnd_b = b"".join(map(b32, rho)), b"".join(map(b32, rnd))
    ctx.set_params(m, n, enc_g, ck_g, ck_h, ghat)
    deck2, proof = ctx.shuffle_and_remask(pb(pk), deck_b, perm, rho_b, rnd_b)
    assert deck2 == co.remask(enc_g, pb(pk), deck_b, perm, rho_b)
    assert proof == co.prove(m, n, enc_g, ck_g, ck_h, ghat, pb(pk), deck_b, deck2, perm, rho_b, rnd_b)
    assert ctx.verify_shuffle(pb(pk), deck_b, deck2, proof) == 0
    assert co.verify(m, n, enc_g, ck_g, ck_h, ghat, pb(pk), deck_b, deck2, proof) == 0


def test_identity_and_duplicate_cards(ctx, scalar_path):
    """Edge decks: identity ciphertext components, duplicated cards, rho = 0 and rho = n-1."""
    m, n = 3, 4
    pp, pk, deck, perm, rho, rnd = instance(m, n, 9)
    deck[0] = (None, None)
    deck[1] = (deck[2][0], None)
    deck[5] = deck[4]
    rho[0], rho[1] = 0, stark.N - 1
    co = c_oracle.COracle(msm_mode=1)
    enc_g, ck_g, ck_h, ghat = pb(pp.enc_g), b"".join(map(pb, pp.ck_g)), pb(pp.ck_h), pb(pp.ghat)
    deck_b = b"".join(pb(c[0]) + pb(c[1]) for c in deck)
    rho_b, rnd_b = b"".join(map(b32, rho)), b"".join(map(b32, rnd))
    ctx.set_params(m, n, enc_g, ck_g, ck_h, ghat)
    deck2, proof = ctx.shuffle_and_remask(pb(pk), deck_b, perm, rho_b, rnd_b)
    assert deck2 == co.remask(enc_g, pb(pk), deck_b, perm, rho_b)
    assert proof == co.prove(m, n, enc_g, ck_g, ck_h, ghat, pb(pk), deck_b, deck2, perm, rho_b, rnd_b)
    assert ctx.verify_shuffle(pb(pk), deck_b, deck2, proof) == 0


@pytest.mark.parametrize("m,n,seed", [(2, 2, 5), (3, 4, 9), (5, 3, 6), (8, 8, 7), (13, 5, 10), (16, 32, 8)])
@pytest.mark.parametrize("form", ["schoolbook", "karatsuba"])
def test_diagonal_products_both_forms_vs_c_oracle(ctx, m, n, seed, form, monkeypatch):
    """The prover's E_k (multi-exponentiation argument) have two device implementations: m(m+1) row
    products over the pre-shifted deck table, and Karatsuba on the row index (csrc/diag.cu, the
    default from m ~ 16).  Both must give the oracle's bytes, including decks with identity
    components and duplicated cards (leaf rows then hit the doubling / cancellation branches) and
    row counts that are not powers of two."""
    monkeypatch.setenv("MP_SMALL_DECK_MAX", "0")
    monkeypatch.setenv("MP_DIAG_KARATSUBA", "1" if form == "karatsuba" else "0")
    pp, pk, deck, perm, rho, rnd = instance(m, n, seed)
    if (m, n) == (3, 4):
        # shuffled-deck coincidences inside one column (the rows a leaf adds up): positions 0 and 4
        # hold the SAME ciphertext, positions 1 and 9 opposite ones, position 2 the identity
        deck[perm[4]] = deck[perm[0]]
        rho[4] = rho[0]
        deck[perm[9]] = (stark.neg(deck[perm[1]][0]), stark.neg(deck[perm[1]][1]))
        rho[9] = (stark.N - rho[1]) % stark.N
        deck[perm[2]] = (None, None)
        rho[2] = 0
    co = c_oracle.COracle(msm_mode=1)
    enc_g, ck_g, ck_h, ghat = pb(pp.enc_g), b"".join(map(pb, pp.ck_g)), pb(pp.ck_h), pb(pp.ghat)
    deck_b = b"".join(pb(c[0]) + pb(c[1]) for c in deck)
    rho_b, rnd_b = b"".join(map(b32, rho)), b"".join(map(b32, rnd))
    ctx.set_params(m, n, enc_g, ck_g, ck_h, ghat)
    deck2, proof = ctx.shuffle_and_remask(pb(pk), deck_b, perm, rho_b, rnd_b)
    assert deck2 == co.remask(enc_g, pb(pk), deck_b, perm, rho_b)
    assert proof == co.prove(m, n, enc_g, ck_g, ck_h, ghat, pb(pk), deck_b, deck2, perm, rho_b, rnd_b)
    assert ctx.verify_shuffle(pb(pk), deck_b, deck2, proof) == 0


def _batch(m, n, seeds):
    co = c_oracle.COracle(msm_mode=1)
    pp0 = None
    decks = decks2 = proofs = b""
    for s in seeds:
        pp, pk, deck, perm, rho, rnd = instance(m, n, s)
        if pp0 is None:
            pp0, pk0 = pp, pk
        enc_g, ck_g, ck_h, ghat = pb(pp0.enc_g), b"".join(map(pb, pp0.ck_g)), pb(pp0.ck_h), pb(pp0.ghat)
        deck_b = b"".join(pb(c[0]) + pb(c[1]) for c in deck)
        rho_b, rnd_b = b"".join(map(b32, rho)), b"".join(map(b32, rnd))
        d2 = co.remask(enc_g, pb(pk0), deck_b, perm, rho_b)
        pf = co.prove(m, n, enc_g, ck_g, ck_h, ghat, pb(pk0), deck_b, d2, perm, rho_b, rnd_b)
        decks += deck_b; decks2 += d2; proofs += pf
    return (enc_g, ck_g, ck_h, ghat, pb(pk0)), decks, decks2, proofs


@pytest.mark.parametrize("m,n,B", [(3, 4, 7), (4, 13, 5)])
def test_verify_batch_matches_single_and_oracle(ctx, m, n, B):
    (enc_g, ck_g, ck_h, ghat, pk), decks, decks2, proofs = _batch(m, n, list(range(30, 30 + B)))
    ctx.set_params(m, n, enc_g, ck_g, ck_h, ghat)
    plen, dlen = len(proofs) // B, 128 * m * n
    assert ctx.verify_shuffle_batch(pk, decks, decks2, proofs) == [0] * B
    # tamper: proof 1 gets a flipped response scalar, proof 2 the (valid) shuffled deck of proof 3
    bad = bytearray(proofs)
    bad[plen * 2 - 1 - 32 * 3] ^= 1          # multi-exp response r of proof 1
    bad_decks2 = bytearray(decks2)
    bad_decks2[2 * dlen:3 * dlen] = decks2[3 * dlen:4 * dlen]
    got = ctx.verify_shuffle_batch(pk, decks, bytes(bad_decks2), bytes(bad), host_threads=3)
    co = c_oracle.COracle(msm_mode=1)
    want = [co.verify(m, n, enc_g, ck_g, ck_h, ghat, pk, decks[i * dlen:(i + 1) * dlen],
                      bytes(bad_decks2[i * dlen:(i + 1) * dlen]), bytes(bad[i * plen:(i + 1) * plen])) for i in range(B)]
    assert got == want and want[1] == 4 and want[2] == 1 and want.count(0) == B - 2
    single = [ctx.verify_shuffle(pk, decks[i * dlen:(i + 1) * dlen], bytes(bad_decks2[i * dlen:(i + 1) * dlen]),
                                 bytes(bad[i * plen:(i + 1) * plen])) for i in range(B)]
    assert single == want


@pytest.mark.parametrize("lanes", ["0", "1"])
def test_large_deck_batch_verifier_with_shared_statement_hashes(ctx, lanes, monkeypatch):
    """The large-deck batch verifier (worker contexts running the single-proof verifier; forced at a small size through
    MP_SMALL_DECK_MAX) with each worker hashing its own statement and with the statements of up to eight proofs hashed
    together (multi-stream Blake2s): the oracle's verdicts either way, valid and tampered proofs, 10 proofs = lane groups
    of 8 + 2."""
    monkeypatch.setenv("MP_SMALL_DECK_MAX", "0")
    monkeypatch.setenv("MP_HASH_LANES", lanes)
    m, n, B = 3, 4, 10
    (enc_g, ck_g, ck_h, ghat, pk), decks, decks2, proofs = _batch(m, n, list(range(60, 60 + B)))
    ctx.set_params(m, n, enc_g, ck_g, ck_h, ghat)
    plen, dlen = len(proofs) // B, 128 * m * n
    assert ctx.verify_shuffle_batch(pk, decks, decks2, proofs, host_threads=4) == [0] * B
    bad = bytearray(proofs)
    bad[plen * 2 - 1 - 32 * 3] ^= 1          # multi-exp response r of proof 1
    o9 = plen * 9                            # proof 9 (second lane group): the first two points of c_A swapped -- still
    bad[o9:o9 + 64], bad[o9 + 64:o9 + 128] = proofs[o9 + 64:o9 + 128], proofs[o9:o9 + 64]   # curve points, another first challenge
    bad_decks2 = bytearray(decks2)
    bad_decks2[4 * dlen:5 * dlen] = decks2[5 * dlen:6 * dlen]
    got = ctx.verify_shuffle_batch(pk, decks, bytes(bad_decks2), bytes(bad), host_threads=4)
    co = c_oracle.COracle(msm_mode=1)
    want = [co.verify(m, n, enc_g, ck_g, ck_h, ghat, pk, decks[i * dlen:(i + 1) * dlen],
                      bytes(bad_decks2[i * dlen:(i + 1) * dlen]), bytes(bad[i * plen:(i + 1) * plen])) for i in range(B)]
    assert got == want
    assert got[1] == 4 and got[4] == 1 and got[9] != 0 and got.count(0) == B - 3


def test_batches_split_over_two_worker_contexts(ctx):
    """Batches of >= 256 small decks are cut into chunks that two worker contexts work through concurrently (one chunk
    in its host phases while the other is on the device; csrc/shuffle_internal.cuh run_chunks).  Proofs equal the
    single-call proofs, one of them the C oracle's; the batched verifier (chunked the same way) returns the oracle's
    verdicts with a tampered proof in each chunk."""
    m, n, B = 2, 3, 260
    N = m * n
    co = c_oracle.COracle(msm_mode=1)
    pp0, pk0, deck, perm, rho, rnd = instance(m, n, 77)
    enc_g, ck_g, ck_h, ghat, pk = pb(pp0.enc_g), b"".join(map(pb, pp0.ck_g)), pb(pp0.ck_h), pb(pp0.ghat), pb(pk0)
    deck_b = b"".join(pb(c[0]) + pb(c[1]) for c in deck)
    rho_b, rnd_b = b"".join(map(b32, rho)), b"".join(map(b32, rnd))
    perms, rhos, rands = [], b"", b""
    for i in range(B):
        k = i % N
        perms += perm[k:] + perm[:k]
        rhos += rho_b[32 * k:] + rho_b[:32 * k]
        j = 32 * (i % (len(rnd_b) // 32))
        rands += rnd_b[j:] + rnd_b[:j]
    ctx.set_params(m, n, enc_g, ck_g, ck_h, ghat)
    out_decks, proofs = ctx.shuffle_and_remask_batch(pk, deck_b * B, perms, rhos, rands, host_threads=4)
    dl, pl, rl = len(deck_b), len(proofs) // B, len(rnd_b)
    for i in (0, 129, 130, 259):
        d, p = ctx.shuffle_and_remask(pk, deck_b, perms[N * i:N * (i + 1)], rhos[32 * N * i:32 * N * (i + 1)], rands[rl * i:rl * (i + 1)])
        assert d == out_decks[dl * i:dl * (i + 1)] and p == proofs[pl * i:pl * (i + 1)], i
    i = 130
    want = co.prove(m, n, enc_g, ck_g, ck_h, ghat, pk, deck_b, out_decks[dl * i:dl * (i + 1)], perms[N * i:N * (i + 1)],
                    rhos[32 * N * i:32 * N * (i + 1)], rands[rl * i:rl * (i + 1)])
    assert want == proofs[pl * i:pl * (i + 1)]
    assert ctx.verify_shuffle_batch(pk, deck_b * B, out_decks, proofs, host_threads=4) == [0] * B
    bad = bytearray(proofs)
    tampered = (3, 129, 130, 259)
    for i in tampered:
        bad[pl * (i + 1) - 1 - 32 * 3] ^= 1        # multi-exp response r
    got = ctx.verify_shuffle_batch(pk, deck_b * B, out_decks, bytes(bad), host_threads=4)
    for i in range(B):
        assert got[i] == (0 if i not in tampered else
                          co.verify(m, n, enc_g, ck_g, ck_h, ghat, pk, deck_b, out_decks[dl * i:dl * (i + 1)], bytes(bad[pl * i:pl * (i + 1)]))), i
    assert all(got[i] != 0 for i in tampered)


@pytest.mark.parametrize("m,n,B,workers", [(3, 4, 9, False), (3, 4, 9, True), (3, 4, 11, "lanes"), (4, 13, 6, False), (2, 2, 3, False)])
def test_prove_batch_is_byte_identical_to_single_calls(ctx, m, n, B, workers, monkeypatch):
    # two implementations behind mp_shuffle_and_remask_batch: the lockstep prover (default for
    # small decks) and concurrent worker contexts (large decks, or forced by MP_BATCH_WORKERS) -- the latter with
    # each worker hashing its own statement, or ("lanes") with the statement heads of up to eight decks hashed
    # together by the multi-stream Blake2s (csrc/shuffle_internal.cuh StatementHashes: 11 decks = lane groups of 8 + 3)
    if workers:
        monkeypatch.setenv("MP_BATCH_WORKERS", "1")
        monkeypatch.setenv("MP_HASH_LANES", "1" if workers == "lanes" else "0")
    co = c_oracle.COracle(msm_mode=1)
    pp0, pk0, *_ = instance(m, n, 50)
    enc_g, ck_g, ck_h, ghat, pk = pb(pp0.enc_g), b"".join(map(pb, pp0.ck_g)), pb(pp0.ck_h), pb(pp0.ghat), pb(pk0)
    decks = rhos = rands = b""
    perms = []
    for s in range(50, 50 + B):
        _, _, deck, perm, rho, rnd = instance(m, n, s)
        decks += b"".join(pb(c[0]) + pb(c[1]) for c in deck)
        perms += perm
        rhos += b"".join(map(b32, rho))
        rands += b"".join(map(b32, rnd))
    ctx.set_params(m, n, enc_g, ck_g, ck_h, ghat)
    out_decks, proofs = ctx.shuffle_and_remask_batch(pk, decks, perms, rhos, rands, host_threads=4)
    assert ctx.launches > 0
    N, plen, rl = m * n, len(proofs) // B, len(rands) // B
    for i in range(B):
        d, p = ctx.shuffle_and_remask(pk, decks[128 * N * i:128 * N * (i + 1)], perms[N * i:N * (i + 1)],
                                      rhos[32 * N * i:32 * N * (i + 1)], rands[rl * i:rl * (i + 1)])
        assert d == out_decks[128 * N * i:128 * N * (i + 1)] and p == proofs[plen * i:plen * (i + 1)]
    i = B - 1
    want = co.prove(m, n, enc_g, ck_g, ck_h, ghat, pk, decks[128 * N * i:128 * N * (i + 1)],
                    out_decks[128 * N * i:128 * N * (i + 1)], perms[N * i:N * (i + 1)],
                    rhos[32 * N * i:32 * N * (i + 1)], rands[rl * i:rl * (i + 1)])
    assert want == proofs[plen * i:plen * (i + 1)]
    assert ctx.verify_shuffle_batch(pk, decks, out_decks, proofs) == [0] * B
    # a permutation entry out of range is a usage error, not a crash
    import pytest as _pt
    bad = list(perms)
    bad[5] = 10 ** 6
    with _pt.raises(Exception):
        ctx.shuffle_and_remask_batch(pk, decks, bad, rhos, rands, host_threads=2)


def _scalar_offsets(m, n):
    """byte offsets of the 5n + 9 scalars of the flat proof layout"""
    f1 = 64 * (5 * m + 4)
    f2 = f1 + 32 * (2 * n + 3) + 64 * 3
    f3 = f2 + 32 * (2 * n + 2) + 64 * (6 * m + 1)
    return [f1 + 32 * k for k in range(2 * n + 3)] + [f2 + 32 * k for k in range(2 * n + 2)] + [f3 + 32 * k for k in range(n + 4)]


def test_non_canonical_proof_scalars_are_rejected(ctx, pkg):
    """s and s + order must not both verify (proof malleability): ark-serialize's CanonicalDeserialize rejects a
    scalar >= the group order before the reference's verifier runs; here the C ABI returns MP_ERR_NOT_CANONICAL,
    as the oracle does, for EVERY scalar position of the proof."""
    fx = GOLD["shuffle"][2]
    setup_ctx(ctx, fx)
    m, n = fx["m"], fx["n"]
    co = c_oracle.COracle()
    args = (m, n, h(fx["enc_g"]), h(fx["ck_g"]), h(fx["ck_h"]), h(fx["ghat"]), h(fx["pk"]))
    proof = h(fx["proof"])
    offs = _scalar_offsets(m, n)
    assert len(offs) == 5 * n + 9 and offs[-1] + 32 == len(proof)
    for off in offs:
        s = int.from_bytes(proof[off:off + 32], "little")
        assert s < stark.N
        if s + stark.N >= 1 << 256:
            continue
        p2 = proof[:off] + (s + stark.N).to_bytes(32, "little") + proof[off + 32:]
        assert co.verify(*args, h(fx["deck"]), h(fx["deck2"]), p2) == -5
        with pytest.raises(pkg.MpError) as e:
            ctx.verify_shuffle(h(fx["pk"]), h(fx["deck"]), h(fx["deck2"]), p2)
        assert e.value.code == -5, off
    # the Python oracle agrees
    with pytest.raises(bg.NonCanonicalScalar):
        bg.proof_from_bytes(p2, m, n)
    # the repository's proof container refuses the same bytes at deserialisation
    with pytest.raises(pkg.MpError) as e:
        ctx.proof_deserialize(m, n, ctx.proof_serialize(m, n, p2))
    assert e.value.code == -5
    assert ctx.proof_deserialize(m, n, ctx.proof_serialize(m, n, proof)) == proof


@pytest.mark.parametrize("m,n,B", [(3, 4, 6)])
def test_verify_batch_reports_malformed_items_individually(ctx, pkg, m, n, B, monkeypatch):
    """One malformed proof (a scalar >= the order; a point off the curve) must not block the other players' proofs:
    statuses[i] = MP_VERIFY_MALFORMED (7) for that item, the rest of the batch is verified as usual -- in the
    lockstep implementation and on the worker-context path of large decks."""
    (enc_g, ck_g, ck_h, ghat, pk), decks, decks2, proofs = _batch(m, n, list(range(70, 70 + B)))
    ctx.set_params(m, n, enc_g, ck_g, ck_h, ghat)
    plen, dlen = len(proofs) // B, 128 * m * n
    bad = bytearray(proofs)
    off = plen * 1 + _scalar_offsets(m, n)[3]                       # item 1: scalar + order
    s = int.from_bytes(bad[off:off + 32], "little")
    bad[off:off + 32] = (s + stark.N).to_bytes(32, "little")
    bad[plen * 3 + 5] ^= 1                                          # item 3: proof point off the curve
    bad[plen * 4 - 1 - 32 * 3] ^= 1                                 # item 3 again (irrelevant) ...
    bad_decks2 = bytearray(decks2)
    bad_decks2[4 * dlen + 70] ^= 1                                  # item 4: a shuffled-deck point off the curve
    bad[plen * 6 - 1 - 32 * 3] ^= 1                                 # item 5: well-formed but wrong -> multi-exp (4)
    want = [0, 7, 0, 7, 7, 4]
    assert ctx.verify_shuffle_batch(pk, decks, bytes(bad_decks2), bytes(bad), host_threads=2) == want
    assert pkg.lib.mp_verify_status_string(7).startswith(b"malformed")
    monkeypatch.setenv("MP_SMALL_DECK_MAX", "0")                    # the large-deck path: single-proof verifier on worker contexts
    assert ctx.verify_shuffle_batch(pk, decks, bytes(bad_decks2), bytes(bad), host_threads=2) == want
    monkeypatch.delenv("MP_SMALL_DECK_MAX")
    # the single-proof entry points name the defect
    for i, code in ((1, -5), (3, -3), (4, -3)):
        with pytest.raises(pkg.MpError) as e:
            ctx.verify_shuffle(pk, decks[i * dlen:(i + 1) * dlen], bytes(bad_decks2[i * dlen:(i + 1) * dlen]), bytes(bad[i * plen:(i + 1) * plen]))
        assert e.value.code == code


def test_remask_flags_do_not_hide_each_other(ctx, pkg):
    """An out-of-range permutation entry and an off-curve public key in the same call: the off-curve key must be
    reported (and must not be cached as a valid table); then the same bad key alone; then a good key."""
    fx = GOLD["shuffle"][0]
    setup_ctx(ctx, fx)
    N = fx["m"] * fx["n"]
    bad_pk = bytearray(h(fx["pk"]))
    bad_pk[3] ^= 1
    bad_perm = list(fx["perm"])
    bad_perm[0] = N + 5
    with pytest.raises(pkg.MpError) as e:
        ctx.remask(bytes(bad_pk), h(fx["deck"]), bad_perm, h(fx["rho"]))
    assert e.value.code == -3
    with pytest.raises(pkg.MpError) as e:      # the garbage table of the rejected key must not have been cached
        ctx.remask(bytes(bad_pk), h(fx["deck"]), fx["perm"], h(fx["rho"]))
    assert e.value.code == -3
    with pytest.raises(pkg.MpError) as e:
        ctx.remask(h(fx["pk"]), h(fx["deck"]), bad_perm, h(fx["rho"]))
    assert e.value.code == -1
    assert ctx.remask(h(fx["pk"]), h(fx["deck"]), fx["perm"], h(fx["rho"])).hex() == fx["deck2"]
