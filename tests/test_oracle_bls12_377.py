"""The BLS12-377 oracle against mathematics (curve-family identities, group laws) and against its committed
golden fixtures (tests/golden/bls12_377_vectors.json)."""
import json
import os
import random

import pytest

from sympy import isprime

from oracle.py import bls12_377 as bls
from _util_bls12_377 import chain_points, pb, b32

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bls12_377_vectors.json")))


def test_curve_constants():
    x = bls.X_PARAM
    # BLS12 family: r = x^4 - x^2 + 1, q = (x - 1)^2 r / 3 + x; G1: y^2 = x^3 + 1 of order h * r
    assert bls.N == x ** 4 - x ** 2 + 1 and bls.P == (x - 1) ** 2 * bls.N // 3 + x
    assert isprime(bls.P) and isprime(bls.N)
    assert bls.P.bit_length() == 377 and bls.N.bit_length() == 253
    assert bls.COFACTOR == (x - 1) ** 2 // 3
    # the trace of Frobenius t = x + 1 gives #E(F_q) = q + 1 - t = h * r
    assert bls.P + 1 - (x + 1) == bls.COFACTOR * bls.N
    assert bls.is_on_curve(bls.G) and bls.mul(bls.G, bls.N) is None
    assert bls.CURVE.mul.__self__.N == bls.N


def test_group_laws():
    rnd = random.Random(1)
    a, b, c = (bls.mul(bls.G, rnd.randrange(1, bls.N)) for _ in range(3))
    assert bls.add(a, b) == bls.add(b, a)
    assert bls.add(bls.add(a, b), c) == bls.add(a, bls.add(b, c))
    assert bls.add(a, bls.neg(a)) is None and bls.add(a, None) == a
    k1, k2 = rnd.randrange(bls.N), rnd.randrange(bls.N)
    assert bls.add(bls.mul(a, k1), bls.mul(a, k2)) == bls.mul(a, (k1 + k2) % bls.N)
    assert bls.mul(bls.mul(a, k1), k2) == bls.mul(a, k1 * k2 % bls.N)
    assert bls.point_from_bytes(pb(a)) == a and bls.point_from_bytes(pb(None)) is None


def test_chain_points_have_known_logs():
    s0, s1, pts, st = chain_points(6, 9)
    for i, p in enumerate(pts):
        assert p == bls.mul(bls.G, (s0 + i * s1) % bls.N)


def test_golden_vectors():
    assert pb(bls.G).hex() == GOLD["generator"]
    for k, v in GOLD["multiples"].items():
        assert pb(bls.mul(bls.G, int(k))).hex() == v
    for fx in GOLD["msm"]:
        nc = fx["ncomp"]
        pts = [bls.point_from_bytes(bytes.fromhex(fx["points"])[96 * i:96 * i + 96]) for i in range(fx["n"] * nc)]
        ks = [int.from_bytes(bytes.fromhex(fx["scalars"])[32 * i:32 * i + 32], "little") for i in range(fx["n"])]
        assert b"".join(pb(bls.msm(pts[c::nc], ks)) for c in range(nc)).hex() == fx["result"]
    for fx in GOLD["pedersen"]:
        L, k = fx["len"], fx["k"]
        ck = [bls.point_from_bytes(bytes.fromhex(fx["ck"])[96 * i:96 * i + 96]) for i in range(L + 1)]
        vals = bytes.fromhex(fx["values"])
        blinds = bytes.fromhex(fx["blinds"])
        out = b""
        for j in range(k):
            v = [int.from_bytes(vals[32 * (j * L + i):32 * (j * L + i) + 32], "little") for i in range(L)]
            r = int.from_bytes(blinds[32 * j:32 * j + 32], "little")
            out += pb(bls.add(bls.mul(ck[0], r), bls.msm(ck[1:], v)))
        assert out.hex() == fx["result"]


def test_c_restatement_matches_python_oracle():
    """oracle/c/bls12_377.c (6 x u64 CIOS field, Jacobian a = 0, double-and-add / ark-style Pippenger) against
    the big-int oracle and the golden vectors, byte for byte."""
    from oracle import c_oracle
    co = c_oracle.COracleBls12_377()
    rnd = random.Random(2)
    for _ in range(200):
        a, b = rnd.randrange(bls.P), rnd.randrange(bls.P)
        assert co.fq_mul(bls.fe_to_bytes(a), bls.fe_to_bytes(b)) == bls.fe_to_bytes(a * b % bls.P)
    for fx in GOLD["msm"]:
        for mode in (0, 1):
            assert co.msm(bytes.fromhex(fx["points"]), bytes.fromhex(fx["scalars"]), fx["ncomp"], mode).hex() == fx["result"]
    s0, s1, pts, st = chain_points(300, 12)
    ks = [st.scalar() for _ in range(300)]
    ks[:4] = [0, 1, bls.N - 1, bls.N + 5]  # zero / one shortcuts of the ark loop, unreduced input
    e = sum(k * (s0 + i * s1) for i, k in enumerate(ks)) % bls.N
    want = pb(bls.mul(bls.G, e))
    pbytes, kbytes = b"".join(map(pb, pts)), b"".join(k.to_bytes(32, "little") for k in ks)
    assert co.msm(pbytes, kbytes, 1, 1) == want
    co.set_threads(4)
    assert co.msm(pbytes, kbytes, 1, 1) == want and co.msm(pbytes[:96 * 40], kbytes[:32 * 40], 1, 0) == pb(
        bls.mul(bls.G, sum(k * (s0 + i * s1) for i, k in enumerate(ks[:40])) % bls.N))
    # identity points and cancelling pairs
    p = pts[0]
    assert co.msm(pb(p) + pb(bls.neg(p)) + pb(None), b32(7) + b32(7) + b32(9), 1, 0) == bytes(96)


def test_shuffle_protocol_over_bls12_377():
    """The protocol oracle instantiated over the second curve (`bayer_groth.curve`): the committed fixture
    re-derives byte for byte, verifies, and tampering is caught; the Stark instantiation is untouched afterwards."""
    import copy
    from oracle.py import bayer_groth as bg, stark
    from _util_bls12_377 import instance
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bls12_377_shuffle_vectors.json")))
    fx = gold["shuffle"][0]
    with bg.curve("bls12_377"):
        pp, pk, deck, perm, rho, rnd = instance(fx["m"], fx["n"], fx["seed"])
        deck2, proof = bg.shuffle_and_remask(pp, pk, deck, rho, perm, rnd)
        buf = bg.proof_to_bytes(proof)
        assert buf.hex() == fx["proof"] and len(buf) == bg.proof_len(fx["m"], fx["n"]) == (11 * fx["m"] + 8) * 96 + (5 * fx["n"] + 9) * 32
        assert b"".join(pb(c[0]) + pb(c[1]) for c in deck2).hex() == fx["deck2"]
        assert bg.shuffle_verify(pp, pk, deck, deck2, bg.proof_from_bytes(buf, fx["m"], fx["n"])) == bg.OK
        bad = copy.deepcopy(proof)
        bad["product"]["hadamard"]["zero"]["t"] = (bad["product"]["hadamard"]["zero"]["t"] + 1) % bls.N
        assert bg.shuffle_verify(pp, pk, deck, deck2, bad) == bg.ERR_ZERO
        assert bg.shuffle_verify(pp, pk, deck, deck2[1:] + deck2[:1], proof) != bg.OK
    assert bg.stark is stark and bg.Q == stark.N and bg.POINT_BYTES == 64


def test_wire_serialisation_of_the_second_curve(pkg):
    """Serialising half of the wire format over BLS12-377 (csrc/wire_host.hpp through mp377_points_compress /
    mp377_deck_serialize / mp377_proof_serialize: host byte handling, no GPU) against oracle/py/wire.py; and the one
    thing the reference itself says about proof sizes -- examples/parameter_selection.rs:10 "notice how proof size
    hits a minimum at m=10, n=30" among its five splits of 300 cards."""
    from oracle.py import wire
    b377 = pkg.bls12_377
    rnd = random.Random(8)
    pts = [bls.mul(bls.G, rnd.randrange(1, bls.N)) for _ in range(12)] + [None]
    pts += [bls.neg(p) for p in pts[:6]]
    flat = b"".join(map(pb, pts))
    want = b"".join(wire.compress_generic(p, bls.CURVE) for p in pts)
    assert b377.points_compress(flat) == want and len(want) == 48 * len(pts)
    # y and -y differ exactly in the "larger" flag
    for p in pts[:6]:
        a, b = wire.compress_generic(p, bls.CURVE), wire.compress_generic(bls.neg(p), bls.CURVE)
        assert a[:47] == b[:47] and (a[47] ^ b[47]) == 0x80
    deck = [(pts[2 * i], pts[2 * i + 1]) for i in range(6)]
    deck_bytes = b"".join(pb(c[0]) + pb(c[1]) for c in deck)
    assert b377.deck_serialize(deck_bytes) == wire.deck_serialize_generic(deck, bls.CURVE)
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bls12_377_shuffle_vectors.json")))
    for fx in gold["shuffle"]:
        proof = bytes.fromhex(fx["proof"])
        ser = b377.proof_serialize(fx["m"], fx["n"], proof)
        assert ser == wire.proof_serialize_generic(proof, fx["m"], fx["n"], bls.CURVE)
        assert len(ser) == (11 * fx["m"] + 8) * 48 + (5 * fx["n"] + 9) * 32 == pkg.lib.mp377_proof_serialized_len(fx["m"], fx["n"])
    sizes = {(m, n): pkg.lib.mp377_proof_serialized_len(m, n) for m, n in [(2, 150), (6, 50), (10, 30), (12, 25), (30, 10)]}
    assert min(sizes, key=sizes.get) == (10, 30)


def test_g1_membership_test_by_endomorphism():
    """The subgroup test the GPU verifier applies to untrusted BLS12-377 points (csrc/capi_bls12_377.cu
    k377_subgroup_check; the one ark-bls12-377 uses for CanonicalDeserialize): phi(P) == -[u^2]P with
    phi(x, y) = (beta x, y).  Big-int check of the constants and of the test's behaviour on subgroup points, random
    curve points, pure cofactor-torsion points and G1 + torsion."""
    import random
    Q, R = bls.Q, bls.N
    u = 0x8508c00000000001
    assert u == bls.X_PARAM and (u * u).bit_length() == 127 and bin(u * u).count("1") == 22
    assert u * u == 0x452217cc900000010a11800000000001
    beta = 0x1ae3a4617c510eabc8756ba8f8c524eb8882a75cc9bc8e359064ee822fb5bffd1e945779fffffffffffffffffffffff
    assert beta != 1 and pow(beta, 3, Q) == 1
    # the Montgomery constant in the kernel (beta * 2^384 mod q, little-endian words)
    words = [0x5a7b8727, 0x2c766f92, 0x253d58b5, 0x03d7f6b0, 0xec122131, 0x838ec0de, 0xf658bb10, 0xbd5eb3e9, 0x6ed3e52e,
             0x6942bd12, 0xdd04ed6a, 0x01673786]
    assert sum(w << (32 * i) for i, w in enumerate(words)) == beta * (1 << 384) % Q

    def smul(P, k):  # plain double-and-add (bls.mul reduces the scalar mod r, which is wrong outside G1)
        acc = None
        for bit in bin(k)[2:]:
            acc = bls.add(acc, acc)
            if bit == "1":
                acc = bls.add(acc, P)
        return acc

    def member(P):
        if P is None:
            return True
        T = smul(P, u * u)
        return T is not None and (beta * P[0] % Q, P[1]) == bls.neg(T)

    rnd = random.Random(3)
    assert all(member(bls.mul(bls.G, rnd.randrange(1, R))) for _ in range(4))
    found = 0
    while found < 3:
        x = rnd.randrange(Q)
        y2 = (x * x * x + 1) % Q
        if pow(y2, (Q - 1) // 2, Q) != 1:
            continue
        # square root by Tonelli-Shanks (q - 1 = 2^46 * odd)
        q, s = Q - 1, 0
        while q % 2 == 0:
            q //= 2
            s += 1
        z = 2
        while pow(z, (Q - 1) // 2, Q) != Q - 1:
            z += 1
        m, c, t, r = s, pow(z, q, Q), pow(y2, q, Q), pow(y2, (q + 1) // 2, Q)
        while t != 1:
            i, tt = 0, t
            while tt != 1:
                tt = tt * tt % Q
                i += 1
            b = pow(c, 1 << (m - i - 1), Q)
            m, c = i, b * b % Q
            t, r = t * c % Q, r * b % Q
        P = (x, r)
        assert bls.is_on_curve(P) and not member(P)
        T = smul(P, R)
        if T is not None:
            assert smul(T, bls.COFACTOR) is None and not member(T) and not member(bls.add(bls.G, T))
            found += 1


def test_sigma_golden_vectors_are_the_oracles():
    """tests/golden/bls12_377_sigma_vectors.json (what the GPU is compared with) re-derived from oracle/py/sigma.py over
    this curve: every mask / reveal / key-ownership proof, verified by the oracle, and the reference's negative cases
    (masking.rs:96-105, reveal.rs:73-82, tests.rs:72-77)."""
    import json, os
    from oracle.py import sigma
    from _util_bls12_377 import pb
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bls12_377_sigma_vectors.json")))
    hx = bytes.fromhex
    with sigma.curve("bls12_377"):
        g, shared = bls.point_from_bytes(hx(gold["g"])), bls.point_from_bytes(hx(gold["shared_key"]))
        assert g == bls.G
        for fx in gold["mask"]:
            card = bls.point_from_bytes(hx(fx["card"]))
            masked, proof = sigma.mask(g, shared, card, int(fx["r"], 16), int(fx["omega"], 16))
            assert (pb(masked[0]) + pb(masked[1])).hex() == fx["masked"] and sigma.cp_proof_bytes(proof).hex() == fx["proof"]
            assert sigma.verify_mask(g, shared, card, masked, proof) == sigma.OK
            assert sigma.verify_mask(g, shared, bls.add(card, g), masked, proof) == sigma.ERR_CHAUM_PEDERSEN
        for fx in gold["reveal"]:
            masked = (bls.point_from_bytes(hx(fx["masked"])[:96]), bls.point_from_bytes(hx(fx["masked"])[96:]))
            pk = bls.point_from_bytes(hx(fx["pk"]))
            token, proof = sigma.compute_reveal_token(g, int(fx["sk"], 16), pk, masked, int(fx["omega"], 16))
            assert pb(token).hex() == fx["token"] and sigma.cp_proof_bytes(proof).hex() == fx["proof"]
            assert sigma.verify_reveal(g, pk, token, masked, proof) == sigma.OK
            assert sigma.verify_reveal(g, pk, bls.add(token, g), masked, proof) == sigma.ERR_CHAUM_PEDERSEN
        for fx in gold["key_ownership"]:
            pk = bls.point_from_bytes(hx(fx["pk"]))
            proof = sigma.prove_key_ownership(g, pk, int(fx["sk"], 16), hx(fx["info"]), int(fx["omega"], 16))
            assert sigma.schnorr_proof_bytes(proof).hex() == fx["proof"]
            assert sigma.verify_key_ownership(g, pk, hx(fx["info"]), proof) == sigma.OK
            assert sigma.verify_key_ownership(g, pk, hx(fx["info"]) + b"x", proof) == sigma.ERR_SCHNORR
    # the context manager restored the Stark curve
    from oracle.py import stark
    assert sigma.Q == stark.N


def test_wire_golden_vectors_are_the_oracles():
    """tests/golden/bls12_377_wire_vectors.json against oracle/py/wire.py: every point decompresses to itself, the deck
    round-trips, every rejected encoding is refused -- including the curve point outside G1, which only the subgroup
    test catches."""
    import json, os
    from oracle.py import wire
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bls12_377_wire_vectors.json")))
    hx = bytes.fromhex
    for fx in gold["points"]:
        p = bls.point_from_bytes(hx(fx["point"]))
        assert wire.compress_generic(p, bls.CURVE) == hx(fx["compressed"])
        assert wire.decompress_generic(hx(fx["compressed"]), bls.CURVE) == p
    deck = wire.deck_deserialize_generic(hx(gold["deck_serialized"]), bls.CURVE)
    assert b"".join(pb(a) + pb(b) for a, b in deck) == hx(gold["deck"])
    for enc, st in zip(gold["rejected"], gold["rejected_statuses"]):
        with pytest.raises(ValueError):
            wire.decompress_generic(hx(enc), bls.CURVE)
        if st == 3:
            assert wire.decompress_generic(hx(enc), bls.CURVE, subgroup_check=False) is not None
