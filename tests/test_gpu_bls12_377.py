"""GPU parity, second curve: the BLS12-377 G1 group layer -- 12-limb device field
arithmetic, XYZZ group law with a = 0, the Pippenger MSM (1 and 2 components) and fixed-base batched
Pedersen commitments -- called through the C ABI (include/mpshuffle_bls12_377.h), bit-exact against the
big-int oracle and the committed golden vectors."""
import json
import os
import random

import pytest

from oracle.py import bls12_377 as bls
from _util_bls12_377 import Stream, chain_points, scalars, pb, b32

pytestmark = pytest.mark.gpu
Q, R = bls.P, bls.N
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bls12_377_vectors.json")))
rnd = random.Random(5)
PTS = [bls.mul(bls.G, rnd.randrange(1, R)) for _ in range(12)]
b48 = bls.fe_to_bytes


@pytest.fixture(scope="module")
def ctx377(pkg):
    c = pkg.bls12_377.Context(0)  # raises without a CUDA device: no CPU fallback
    yield c
    c.close()


def test_fq_mul(ctx377):
    # the PTX carry chains (fq_mul_wide over mp_mad6, the parity-split word-serial reduction)
    Rinv = pow(1 << 384, -1, Q)
    n = 8192
    a = [rnd.randrange(0, 5 * Q) for _ in range(n)]
    b = [rnd.randrange(0, 6 * Q) for _ in range(n)]
    a[:8] = [0, 1, Q - 1, Q, 2 * Q, 5 * Q - 1, (1 << 380) - 1, 15 * Q]
    b[:8] = [0, 1, Q - 1, Q, 6 * Q - 1, 2 * Q, 1, 2 * Q - 1]
    # words of all ones / all zeros push carries through every position of the chains
    for k in range(11):
        a[8 + k] = int("ffffffff" * (k + 1), 16) << (32 * ((3 * k) % (11 - k) if k < 10 else 0))
        a[8 + k] %= 8 * Q
        b[8 + k] = (1 << (32 * (k + 1))) - 1
    out = ctx377.dbg_fq_mul(b"".join(map(b48, a)), b"".join(map(b48, b)))
    for i in range(n):
        r = int.from_bytes(out[48 * i:48 * i + 48], "little")
        assert r < 2 * Q and r % Q == a[i] * b[i] * Rinv % Q, i


def test_point_add_complete(ctx377):
    ps, qs, want = [], [], []
    for a in PTS[:6]:
        for b in PTS[6:]:
            ps.append(a); qs.append(b); want.append(bls.add(a, b))
        for b in (a, bls.neg(a), None):
            ps.append(a); qs.append(b); want.append(bls.add(a, b))
        ps.append(None); qs.append(a); want.append(a)
    out = ctx377.dbg_point_add(b"".join(map(pb, ps)), b"".join(map(pb, qs)))
    for i, w in enumerate(want):
        assert out[96 * i:96 * i + 96] == pb(w), i


def test_scalar_mul(ctx377):
    ks = [0, 1, 2, R - 1, R, R + 1] + [rnd.randrange(0, 1 << 256) for _ in range(10)]
    ps = [PTS[i % 12] for i in range(len(ks))]
    out = ctx377.dbg_scalar_mul(b"".join(map(pb, ps)), b"".join(k.to_bytes(32, "little") for k in ks))
    for i, (p, k) in enumerate(zip(ps, ks)):
        assert out[96 * i:96 * i + 96] == pb(bls.mul(p, k)), i


def msm_case(ctx377, n, c, seed=1, kind="uniform"):
    s0, s1, pts, st = chain_points(n, seed)
    ks = scalars(st, n, kind)
    e = sum(k * (s0 + i * s1) for i, k in enumerate(ks)) % R
    got = ctx377.msm_g1(b"".join(map(pb, pts)), b"".join(map(b32, ks)), c)
    assert got == pb(bls.mul(bls.G, e)), (n, c, kind)


@pytest.mark.parametrize("n,c", [(1, 4), (2, 4), (7, 4), (33, 5), (100, 6), (300, 0), (300, 8), (1000, 9), (1000, 0)])
def test_msm_small(ctx377, n, c):
    msm_case(ctx377, n, c)


@pytest.mark.parametrize("kind", ["zero", "max", "small", "same"])
def test_msm_scalar_edges(ctx377, kind):
    # "max" = r - 1 has bit 252 set: the signed-digit recoding needs the 254th bit (kScalarBits)
    for c in (6, 11, 0):
        msm_case(ctx377, 200, c, kind=kind)


def test_msm_empty(ctx377):
    assert ctx377.msm_g1(b"", b"", 0) == bytes(96)


def test_msm_golden(ctx377):
    for fx in GOLD["msm"]:
        fn = ctx377.msm_g1 if fx["ncomp"] == 1 else ctx377.ct_msm
        assert fn(bytes.fromhex(fx["points"]), bytes.fromhex(fx["scalars"]), 0).hex() == fx["result"]
    assert ctx377.msm_g1(bytes.fromhex(GOLD["generator"]), b32(R - 1), 0).hex() == GOLD["multiples"][str(R - 1)]


def test_msm_point_edges(ctx377):
    st = Stream(3)
    Pt = PTS[0]
    n = 150
    ks = [st.scalar() for _ in range(n)]
    for name, pts in [("equal", [Pt] * n), ("pm", [Pt if i % 2 == 0 else bls.neg(Pt) for i in range(n)]),
                      ("ident", [None if i % 3 == 0 else PTS[i % 12] for i in range(n)])]:
        want = pb(bls.msm(pts, ks))
        for c in (4, 7, 0):
            assert ctx377.msm_g1(b"".join(map(pb, pts)), b"".join(map(b32, ks)), c) == want, (name, c)
    # one heavy bucket spanning many accumulate chunks
    assert ctx377.msm_g1(pb(Pt) * 500, b32(12345) * 500, 8) == pb(bls.mul(Pt, 12345 * 500))


def test_ct_msm(ctx377):
    s0, s1, pts, st = chain_points(400, 9)
    n = 200
    ks = [st.scalar() for _ in range(n)]
    deck = b"".join(pb(pts[2 * i]) + pb(pts[2 * i + 1]) for i in range(n))
    e1 = sum(k * (s0 + (2 * i) * s1) for i, k in enumerate(ks)) % R
    e2 = sum(k * (s0 + (2 * i + 1) * s1) for i, k in enumerate(ks)) % R
    for c in (5, 0):
        got = ctx377.ct_msm(deck, b"".join(map(b32, ks)), c)
        assert got == pb(bls.mul(bls.G, e1)) + pb(bls.mul(bls.G, e2)), c


@pytest.mark.parametrize("n,c", [(4096, 0), (4096, 12), (65536, 0), (65536, 16)])
def test_msm_mid(ctx377, n, c):
    msm_case(ctx377, n, c, seed=2)


def test_rejects_bad_points(ctx377, pkg):
    with pytest.raises(pkg.MpError) as e:  # not on the curve
        ctx377.msm_g1(b48(5) + b48(7), b32(3), 4)
    assert e.value.code == -3
    g = bls.G
    with pytest.raises(pkg.MpError) as e:  # non-canonical alias x + q of a curve point
        ctx377.msm_g1(b48(g[0] + Q) + b48(g[1]), b32(3), 4)
    assert e.value.code == -3


def test_pedersen_commit_batch(ctx377, pkg):
    fx = GOLD["pedersen"][0]
    ctx377.set_commit_key(bytes.fromhex(fx["ck"]))
    assert ctx377.pedersen_commit_batch(bytes.fromhex(fx["values"]), bytes.fromhex(fx["blinds"]), fx["len"]).hex() == fx["result"]
    # the reference benchmark's key lengths (parameter_selection.rs:42-43: n in 150, 50, 30, 25, 10), short rows
    s0, s1, ck, st = chain_points(151, 11)
    ctx377.set_commit_key(b"".join(map(pb, ck)))
    logs = [(s0 + i * s1) % R for i in range(151)]
    for length, k in [(150, 4), (30, 7), (1, 3), (0, 2)]:
        vals = [[st.scalar() for _ in range(length)] for _ in range(k)]
        vals[0] = [0] * length
        blinds = [st.scalar() for _ in range(k)]
        got = ctx377.pedersen_commit_batch(b"".join(b32(x) for v in vals for x in v), b"".join(map(b32, blinds)), length)
        for j in range(k):
            e = (blinds[j] * logs[0] + sum(v * l for v, l in zip(vals[j], logs[1:]))) % R
            assert got[96 * j:96 * j + 96] == pb(bls.mul(bls.G, e)), (length, j)
    with pytest.raises(pkg.MpError):
        ctx377.pedersen_commit_batch(b32(1) * 151, b32(1), 151)  # longer than the key


@pytest.mark.parametrize("c,world", [(8, 3), (13, 2), (16, 8), (16, 5)])
def test_msm_window_range_split(ctx377, pkg, c, world):
    """Window-range split (multi-GPU path of SURVEY 8(e)) emulated on one GPU: per-'rank' partials + fold."""
    import torch
    n = 500
    s0, s1, pts, st = chain_points(n, 21)
    ks = scalars(st, n, "uniform")
    ks[0], ks[1] = 0, R - 1
    dev = torch.device("cuda:0")
    d_pts = torch.frombuffer(bytearray(b"".join(map(pb, pts))), dtype=torch.uint8).to(dev)
    d_sc = torch.frombuffer(bytearray(b"".join(map(b32, ks))), dtype=torch.uint8).to(dev)
    d_out = torch.zeros(96, dtype=torch.uint8, device=dev)
    W = ctx377.msm_num_windows(c)
    assert W == (254 + c - 1) // c
    points, scs = b"", b""
    for r, s in pkg.dist.fold_scalars(c, W, world):
        b, e = pkg.dist.window_range(W, r, world)
        ctx377.msm_g1_windows_device(d_pts.data_ptr(), d_sc.data_ptr(), n, d_out.data_ptr(), c, b, e - b)
        ctx377.sync()
        points += bytes(d_out.cpu().numpy().tobytes())
        scs += s
    got = ctx377.msm_g1(points, scs, 0)
    e = sum(k * (s0 + i * s1) for i, k in enumerate(ks)) % R
    assert got == pb(bls.mul(bls.G, e))


def test_msm_jobs_diagonal_products(ctx377, pkg):
    """Batched MSM jobs over shared arrays, in the shape of the multi-exponentiation argument's diagonal products:
    every (deck row i, scalar row j) pair of an (m = 4, n = 25) instance -- the reference benchmark's 300-card deck is
    (12, 25) -- as one call, 2 components; plus ragged / empty / overlapping jobs."""
    m, n = 4, 25
    s0, s1, pts, st = chain_points(2 * m * n, 17)
    logs = [(s0 + i * s1) % R for i in range(2 * m * n)]
    ks = [st.scalar() for _ in range((m + 1) * n)]
    deck = b"".join(map(pb, pts))
    jobs = [(j * n, i * n, n) for i in range(m) for j in range(m + 1)]
    out = ctx377.msm_jobs(deck, b"".join(map(b32, ks)), jobs, ncomp=2)
    for q, (so, po, ln) in enumerate(jobs):
        for comp in range(2):
            e = sum(ks[so + t] * logs[2 * (po + t) + comp] for t in range(ln)) % R
            assert out[192 * q + 96 * comp:192 * q + 96 * comp + 96] == pb(bls.mul(bls.G, e)), (q, comp)
    assert ctx377.launches > 0
    jobs = [(0, 0, 1), (3, 7, 0), (5, 2, 40), (0, 0, 100), (10, 150, 50)]
    out = ctx377.msm_jobs(deck, b"".join(map(b32, ks)), jobs, ncomp=1, window_bits=7)
    for q, (so, po, ln) in enumerate(jobs):
        e = sum(ks[so + t] * logs[po + t] for t in range(ln)) % R
        assert out[96 * q:96 * q + 96] == pb(bls.mul(bls.G, e)), q
    with pytest.raises(pkg.MpError):
        ctx377.msm_jobs(deck, b"".join(map(b32, ks)), [(0, 190, 20)], ncomp=1)  # reaches past the points


def test_shuffle_verifier_group_work_on_gpu(ctx377, monkeypatch):
    """The NEXT row in miniature: `verify_shuffle` over BLS12-377 with the oracle's host logic (transcript, scalar
    algebra, check order) and every group computation of the verifier -- Pedersen commitments, the ciphertext
    MSMs, the point linear combinations -- executed by the GPU group layer through the C ABI.  Valid proofs must
    verify, tampered ones must fail in the same sub-argument as with the pure big-int oracle."""
    import copy
    from oracle.py import bayer_groth as bg
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bls12_377_shuffle_vectors.json")))
    pt = bls.point_from_bytes
    for fx in gold["shuffle"]:
        m, n = fx["m"], fx["n"]
        N = m * n
        raw = {k: bytes.fromhex(fx[k]) for k in ("enc_g", "ck_g", "ck_h", "ghat", "pk", "deck", "deck2", "proof")}
        with bg.curve("bls12_377") as grp:
            pp = bg.Params(m, n, pt(raw["enc_g"]), [pt(raw["ck_g"][96 * i:96 * i + 96]) for i in range(n)],
                           pt(raw["ck_h"]), pt(raw["ghat"]))
            pk = pt(raw["pk"])
            deck = [(pt(raw["deck"][192 * i:192 * i + 96]), pt(raw["deck"][192 * i + 96:192 * i + 192])) for i in range(N)]
            deck2 = [(pt(raw["deck2"][192 * i:192 * i + 96]), pt(raw["deck2"][192 * i + 96:192 * i + 192])) for i in range(N)]
            proof = bg.proof_from_bytes(raw["proof"], m, n)
            # what the pure oracle says, before anything is patched
            bad = copy.deepcopy(proof)
            bad["multiexp"]["tau"] = (bad["multiexp"]["tau"] + 1) % R
            bad2 = copy.deepcopy(proof)
            bad2["product"]["svp"]["r"] = (bad2["product"]["svp"]["r"] + 1) % R
            wrong_deck = deck2[1:] + deck2[:1]
            want = [bg.shuffle_verify(pp, pk, deck, deck2, proof), bg.shuffle_verify(pp, pk, deck, deck2, bad),
                    bg.shuffle_verify(pp, pk, deck, deck2, bad2), bg.shuffle_verify(pp, pk, deck, wrong_deck, proof)]
            assert want[0] == bg.OK and want[1] == bg.ERR_MULTIEXP and want[2] == bg.ERR_SVP and want[3] != bg.OK
            # route the group work through the GPU
            calls = {"msm": 0, "ct_msm": 0, "commit": 0}
            ctx377.set_commit_key(raw["ck_h"] + raw["ck_g"])

            def gpu_msm(points, scalars):
                calls["msm"] += 1
                return pt(ctx377.msm_g1(b"".join(map(pb, points)), b"".join(b32(k % R) for k in scalars)))

            def gpu_ct_msm(cts, scalars):
                calls["ct_msm"] += 1
                cts, scalars = list(cts), list(scalars)
                out = ctx377.ct_msm(b"".join(pb(c[0]) + pb(c[1]) for c in cts), b"".join(b32(k % R) for k in scalars))
                return (pt(out[:96]), pt(out[96:]))

            def gpu_commit(pp_, values, r):
                calls["commit"] += 1
                return pt(ctx377.pedersen_commit_batch(b"".join(b32(v % R) for v in values), b32(r % R), len(values)))

            monkeypatch.setattr(grp, "msm", gpu_msm)
            monkeypatch.setattr(bg, "ct_msm", gpu_ct_msm)
            monkeypatch.setattr(bg, "commit", gpu_commit)
            got = [bg.shuffle_verify(pp, pk, deck, deck2, proof), bg.shuffle_verify(pp, pk, deck, deck2, bad),
                   bg.shuffle_verify(pp, pk, deck, deck2, bad2), bg.shuffle_verify(pp, pk, deck, wrong_deck, proof)]
            monkeypatch.undo()
        assert got == want, (m, n, got, want)
        assert calls["ct_msm"] >= 2 and calls["commit"] >= 3 and calls["msm"] >= 1, calls


def test_verify_shuffle_c_abi(ctx377, pkg):
    """`mp377_shuffle_verify`: BarnettSmartProtocol::verify_shuffle over BLS12-377 through the C ABI (host half =
    csrc/shuffle_host.hpp, group work = the batched MSM).  Valid fixtures verify; a rotated output deck, flipped
    proof bytes in each sub-argument and a proof for another statement get the verdict of the big-int oracle --
    the reference's negative test expects "Hadamard Product (5.1)" for a wrong deck (tests.rs:213-226)."""
    from oracle.py import bayer_groth as bg
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bls12_377_shuffle_vectors.json")))
    pt = bls.point_from_bytes
    for fx in gold["shuffle"]:
        m, n = fx["m"], fx["n"]
        N = m * n
        raw = {k: bytes.fromhex(fx[k]) for k in ("enc_g", "ck_g", "ck_h", "ghat", "pk", "deck", "deck2", "proof")}
        args = (m, n, raw["enc_g"], raw["ck_g"], raw["ck_h"], raw["ghat"], raw["pk"], raw["deck"])
        assert pkg.lib.mp377_proof_len(m, n) == len(raw["proof"])
        assert ctx377.verify_shuffle(*args, raw["deck2"], raw["proof"]) == 0
        assert ctx377.launches > 0
        cases = [(raw["deck2"][192:] + raw["deck2"][:192], raw["proof"])]
        plen = len(raw["proof"])
        for off in (plen - 32 * 4, plen - 32 * 5,                            # multi-exp r, multi-exp a_n (low bytes: stay canonical)
                    (11 * m + 8) * 96 - 96 * (4 * m + 2 * m + 1) - 32 * (2 * n + 2) - 40):  # inside the SVP block
            p2 = bytearray(raw["proof"])
            p2[off] ^= 1
            cases.append((raw["deck2"], bytes(p2)))
        with bg.curve("bls12_377"):
            pp = bg.Params(m, n, pt(raw["enc_g"]), [pt(raw["ck_g"][96 * i:96 * i + 96]) for i in range(n)],
                           pt(raw["ck_h"]), pt(raw["ghat"]))
            pk = pt(raw["pk"])
            cts = lambda b: [(pt(b[192 * i:192 * i + 96]), pt(b[192 * i + 96:192 * i + 192])) for i in range(N)]
            deck = cts(raw["deck"])
            for d2, pf in cases:
                try:
                    parsed = bg.proof_from_bytes(pf, m, n)
                    if not all(bls.is_on_curve(q) for q in parsed["c_A"] + parsed["c_B"]):
                        continue
                    want = bg.shuffle_verify(pp, pk, deck, cts(d2), parsed)
                except Exception:
                    continue  # the flipped byte broke a point encoding: the C ABI reports MP_ERR_NOT_ON_CURVE instead
                try:
                    got = ctx377.verify_shuffle(*args, d2, pf)
                except pkg.MpError as e:
                    assert e.code == -3
                    continue
                assert got == want != 0, (m, n, got, want)
        assert pkg.lib.mp_verify_status_string(1) == b"Hadamard Product (5.1)"
    # the wrong-deck case is the reference's own negative test and must be caught as a verification failure
    fx = gold["shuffle"][1]
    raw = {k: bytes.fromhex(fx[k]) for k in ("enc_g", "ck_g", "ck_h", "ghat", "pk", "deck", "deck2", "proof")}
    st = ctx377.verify_shuffle(fx["m"], fx["n"], raw["enc_g"], raw["ck_g"], raw["ck_h"], raw["ghat"], raw["pk"], raw["deck"],
                               raw["deck2"][192:] + raw["deck2"][:192], raw["proof"])
    assert st > 0


def test_kernels_were_launched(ctx377):
    msm_case(ctx377, 64, 0)
    assert ctx377.launches > 0


def test_msm_2p18_linearity(ctx377):
    """Full-size property (no oracle at this size): MSM(P, k) + MSM(P, k') = MSM(P, k + k')."""
    import torch
    n = 1 << 18
    dev = torch.device("cuda:0")
    s0, s1, pts, st = chain_points(512, 31)
    base = torch.frombuffer(bytearray(b"".join(map(pb, pts))), dtype=torch.uint8).reshape(512, 96)
    idx = torch.arange(n) % 512
    d_pts = base[idx].contiguous().to(dev)
    g = torch.Generator().manual_seed(7)
    k1 = torch.randint(0, 256, (n, 32), dtype=torch.uint8, generator=g)
    k2 = torch.randint(0, 256, (n, 32), dtype=torch.uint8, generator=g)
    k1[:, 31] &= 0x07  # < 2^251 < r, and k1 + k2 < 2^252 < r: no reduction needed for the sum
    k2[:, 31] &= 0x07
    a = k1.to(torch.int32).reshape(n, 32)
    b = k2.to(torch.int32).reshape(n, 32)
    s = a + b
    carry = torch.zeros(n, dtype=torch.int32)
    out = torch.zeros(n, 32, dtype=torch.uint8)
    for j in range(32):
        t = s[:, j] + carry
        out[:, j] = (t & 0xFF).to(torch.uint8)
        carry = t >> 8
    res = []
    d_out = torch.zeros(96, dtype=torch.uint8, device=dev)
    for ks in (k1, k2, out):
        d_k = ks.contiguous().to(dev)
        ctx377.msm_g1_device(d_pts.data_ptr(), d_k.data_ptr(), n, d_out.data_ptr(), 0)
        ctx377.sync()
        res.append(bls.point_from_bytes(bytes(d_out.cpu().numpy())))
    assert bls.is_on_curve(res[0]) and res[0] is not None
    assert bls.add(res[0], res[1]) == res[2]


SHUF = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bls12_377_shuffle_vectors.json")))


def _torsion_point():
    """a point of E(F_q) outside G1: r * (random curve point) is a non-trivial cofactor-torsion point"""
    rr = random.Random(77)
    while True:
        x = rr.randrange(bls.Q)
        y2 = (x * x * x + 1) % bls.Q
        if pow(y2, (bls.Q - 1) // 2, bls.Q) != 1:
            continue
        # q = 1 mod 4: Tonelli-Shanks
        q, s2 = bls.Q - 1, 0
        while q % 2 == 0:
            q //= 2; s2 += 1
        z = 2
        while pow(z, (bls.Q - 1) // 2, bls.Q) != bls.Q - 1:
            z += 1
        mm, c, t, r_ = s2, pow(z, q, bls.Q), pow(y2, q, bls.Q), pow(y2, (q + 1) // 2, bls.Q)
        while t != 1:
            i, tt = 0, t
            while tt != 1:
                tt = tt * tt % bls.Q; i += 1
            b = pow(c, 1 << (mm - i - 1), bls.Q)
            mm, c = i, b * b % bls.Q
            t, r_ = t * c % bls.Q, r_ * b % bls.Q
        P = (x, r_)
        assert bls.is_on_curve(P)
        acc = None
        for bit in bin(bls.N)[2:]:          # plain double-and-add: bls.mul would reduce the scalar mod r
            acc = bls.add(acc, acc)
            if bit == "1":
                acc = bls.add(acc, P)
        if acc is not None:
            return P, acc


def test_subgroup_check(ctx377, pkg):
    """G1 membership (the check ark-ec's CanonicalDeserialize performs on the reference's side): subgroup points and
    the identity pass; a random curve point, a pure cofactor-torsion point and G1 + torsion are on the curve but
    rejected; an off-curve point is reported as such."""
    P, T = _torsion_point()
    good = PTS[:5] + [None]
    bad = [P, T, bls.add(PTS[0], T)]
    off = bytearray(pb(PTS[1])); off[7] ^= 1
    rc, st = ctx377.subgroup_check(b"".join(map(pb, good)))
    assert rc == 0 and st == [0] * 6
    rc, st = ctx377.subgroup_check(b"".join(map(pb, good + bad)))
    assert rc == -6 and st == [0] * 6 + [2, 2, 2]
    rc, st = ctx377.subgroup_check(b"".join(map(pb, good + bad)) + bytes(off))
    assert rc == -3 and st == [0] * 6 + [2, 2, 2, 1]
    # the generic MSM entry accepts the torsion point (on the curve) -- membership is the verifier's job
    assert len(ctx377.msm_g1(pb(T), b32(5))) == 96


def test_shuffle_verify_rejects_torsion_and_non_canonical_inputs(ctx377, pkg):
    """mp377_shuffle_verify validates untrusted inputs like the reference's deserialiser: a deck or proof point with a
    cofactor-torsion component -> MP_ERR_NOT_IN_SUBGROUP; a proof scalar >= r -> MP_ERR_NOT_CANONICAL."""
    fx = SHUF["shuffle"][0]
    hx = bytes.fromhex
    m, n = fx["m"], fx["n"]
    args = [m, n, hx(fx["enc_g"]), hx(fx["ck_g"]), hx(fx["ck_h"]), hx(fx["ghat"]), hx(fx["pk"]), hx(fx["deck"]), hx(fx["deck2"]), hx(fx["proof"])]
    assert ctx377.verify_shuffle(*args) == 0
    _, T = _torsion_point()
    for which, off in ((8, 96), (9, 0), (6, 0)):          # a shuffled-deck point, the first proof point, the public key
        buf = bytearray(args[which])
        pt = bls.point_from_bytes(bytes(buf[off:off + 96]))
        buf[off:off + 96] = pb(bls.add(pt, T))
        a2 = list(args); a2[which] = bytes(buf)
        with pytest.raises(pkg.MpError) as e:
            ctx377.verify_shuffle(*a2)
        assert e.value.code == -6, which
    proof = args[9]
    off = 96 * (5 * m + 4) + 32 * 2
    s = int.from_bytes(proof[off:off + 32], "little")
    a2 = list(args); a2[9] = proof[:off] + (s + bls.N).to_bytes(32, "little") + proof[off + 32:]
    with pytest.raises(pkg.MpError) as e:
        ctx377.verify_shuffle(*a2)
    assert e.value.code == -5


def _raw(fx):
    return {k: bytes.fromhex(fx[k]) for k in ("enc_g", "ck_g", "ck_h", "ghat", "pk", "deck", "deck2", "proof", "rho", "rand")}


@pytest.mark.parametrize("idx", [0, 1])
def test_shuffle_and_remask_is_byte_exact_vs_golden(ctx377, idx):
    """`mp377_shuffle_and_remask`: BarnettSmartProtocol::shuffle_and_remask over BLS12-377 (the call the reference's
    benchmark harness times, examples/parameter_selection.rs:78-96).  Shuffled deck and proof are byte-identical to the
    oracle's (`bayer_groth.curve("bls12_377")`) on identical decks, permutations and randomness; the proof verifies."""
    fx = SHUF["shuffle"][idx]
    r = _raw(fx)
    m, n = fx["m"], fx["n"]
    ctx377.set_params(m, n, r["enc_g"], r["ck_g"], r["ck_h"], r["ghat"])
    assert ctx377.remask(r["pk"], r["deck"], fx["perm"], r["rho"]) == r["deck2"]
    deck2, proof = ctx377.shuffle_and_remask(r["pk"], r["deck"], fx["perm"], r["rho"], r["rand"])
    assert ctx377.launches > 0
    assert deck2 == r["deck2"]
    assert proof == r["proof"]
    assert ctx377.verify_shuffle(m, n, r["enc_g"], r["ck_g"], r["ck_h"], r["ghat"], r["pk"], r["deck"], deck2, proof) == 0


def test_shuffle_and_remask_300_cards_reference_benchmark_shape(ctx377):
    """The reference's own benchmark: a 300-card deck over BLS12-377 G1 (examples/parameter_selection.rs:31-43), at
    (m, n) = (10, 30) -- its proof-size optimum -- and (30, 10).  Inputs regenerated from the seed; proof bytes against
    the committed oracle proofs (tests/golden/make_bls12_377_shuffle_300_golden.py); round trip through the verifier;
    a batch of two decks equals two single calls."""
    import hashlib
    from _util_bls12_377 import instance
    from oracle.py import bayer_groth as bg
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bls12_377_shuffle_300_vectors.json")))
    for fx in gold["shuffle"]:
        m, n = fx["m"], fx["n"]
        with bg.curve("bls12_377"):
            pp, pk, deck, perm, rho, rnd = instance(m, n, fx["seed"])
        enc_g, ck_g, ck_h, ghat = pb(pp.enc_g), b"".join(map(pb, pp.ck_g)), pb(pp.ck_h), pb(pp.ghat)
        deck_b = b"".join(pb(c[0]) + pb(c[1]) for c in deck)
        rho_b, rnd_b = b"".join(map(b32, rho)), b"".join(map(b32, rnd))
        ctx377.set_params(m, n, enc_g, ck_g, ck_h, ghat)
        deck2, proof = ctx377.shuffle_and_remask(pb(pk), deck_b, perm, rho_b, rnd_b)
        assert hashlib.sha256(deck2).hexdigest() == fx["deck2_sha256"]
        assert proof.hex() == fx["proof"]
        assert ctx377.verify_shuffle(m, n, enc_g, ck_g, ck_h, ghat, pb(pk), deck_b, deck2, proof) == 0
        wrong = deck2[192:] + deck2[:192]
        assert ctx377.verify_shuffle(m, n, enc_g, ck_g, ck_h, ghat, pb(pk), deck_b, wrong, proof) == 1   # "Hadamard Product (5.1)"
        if (m, n) == (10, 30):
            perm2 = perm[1:] + perm[:1]
            d2b, pfb = ctx377.shuffle_and_remask_batch(pb(pk), deck_b * 2, perm + perm2, rho_b * 2, rnd_b * 2, host_threads=2)
            assert d2b[:len(deck2)] == deck2 and pfb[:len(proof)] == proof
            d2c, pfc = ctx377.shuffle_and_remask(pb(pk), deck_b, perm2, rho_b, rnd_b)
            assert d2b[len(deck2):] == d2c and pfb[len(proof):] == pfc
            assert ctx377.verify_shuffle(m, n, enc_g, ck_g, ck_h, ghat, pb(pk), deck_b, d2c, pfc) == 0


def test_batch_split_over_two_worker_contexts_equals_single_calls(ctx377):
    """A batch large enough to be cut into chunks that two worker contexts prove concurrently (host phases of one
    chunk beside the device phases of the other, csrc/shuffle_internal.cuh run_chunks): every proof equals the
    single-call proof for the same inputs, which the golden test above pins to the oracle."""
    fx = SHUF["shuffle"][0]
    r = _raw(fx)
    m, n, B = fx["m"], fx["n"], 260
    N = m * n
    ctx377.set_params(m, n, r["enc_g"], r["ck_g"], r["ck_h"], r["ghat"])
    perms = []
    for i in range(B):
        k = i % N
        perms += fx["perm"][k:] + fx["perm"][:k]
    rhos = b"".join(r["rho"][32 * (i % N):] + r["rho"][:32 * (i % N)] for i in range(B))
    d2b, pfb = ctx377.shuffle_and_remask_batch(r["pk"], r["deck"] * B, perms, rhos, r["rand"] * B, host_threads=4)
    dl, pl = len(r["deck"]), len(r["proof"])
    assert d2b[:dl] == r["deck2"] and pfb[:pl] == r["proof"]
    for i in (1, 129, 130, 259):
        d, p = ctx377.shuffle_and_remask(r["pk"], r["deck"], perms[N * i:N * (i + 1)], rhos[32 * N * i:32 * N * (i + 1)], r["rand"])
        assert d == d2b[dl * i:dl * (i + 1)] and p == pfb[pl * i:pl * (i + 1)], i
        assert ctx377.verify_shuffle(m, n, r["enc_g"], r["ck_g"], r["ck_h"], r["ghat"], r["pk"], r["deck"], d, p) == 0


SIG = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bls12_377_sigma_vectors.json")))


@pytest.fixture
def sctx377(ctx377):
    fx = SHUF["shuffle"][0]
    r = _raw(fx)    # any commitment key will do: the sigma protocols only use the ElGamal generator g
    ctx377.set_params(fx["m"], fx["n"], bytes.fromhex(SIG["g"]), r["ck_g"], r["ck_h"], r["ghat"])
    return ctx377


def test_sigma_mask_remask_golden_and_negative_cases(sctx377):
    """`mp377_mask_batch` ... `mp377_verify_remask_batch` (BarnettSmartProtocol::mask / verify_mask / remask /
    verify_remask over BLS12-377; reference mod.rs:182-299): bytes against the oracle's golden vectors
    (tests/golden/make_bls12_377_sigma_golden.py: identity card, zero and maximal scalars included), the reference's
    negative cases (masking.rs:96-105, remasking.rs:103-112), a non-canonical response."""
    from oracle.py import bls12_377 as bls
    h = bytes.fromhex
    cat = lambda key, rows: b"".join(h(r[key]) for r in rows)
    le = lambda key, rows: b"".join(int(r[key], 16).to_bytes(32, "little") for r in rows)
    shared = h(SIG["shared_key"])
    M, R = SIG["mask"], SIG["remask"]
    masked, proofs = sctx377.mask_batch(shared, cat("card", M), le("r", M), le("omega", M))
    assert masked == cat("masked", M) and proofs == cat("proof", M)
    assert sctx377.launches > 0
    assert sctx377.verify_mask_batch(shared, cat("card", M), masked, proofs) == [0] * len(M)
    bad = bytearray(proofs); bad[224 + 194] ^= 1          # response scalar of proof 1
    assert sctx377.verify_mask_batch(shared, cat("card", M), masked, bytes(bad)) == [0, 5] + [0] * (len(M) - 2)
    swapped = masked[192:384] + masked[:192] + masked[384:]
    assert sctx377.verify_mask_batch(shared, cat("card", M), swapped, proofs)[:2] == [5, 5]
    big = bytearray(proofs); big[192:224] = (bls.N + 5).to_bytes(32, "little")   # non-canonical response
    assert sctx377.verify_mask_batch(shared, cat("card", M), masked, bytes(big))[0] == 5
    out, rproofs = sctx377.remask_prove_batch(shared, cat("original", R), le("alpha", R), le("omega", R))
    assert out == cat("remasked", R) and rproofs == cat("proof", R)
    assert sctx377.verify_remask_batch(shared, cat("original", R), out, rproofs) == [0] * len(R)
    assert sctx377.verify_remask_batch(shared, cat("original", R), out[192:] + out[:192], rproofs).count(5) >= len(R) - 1


def test_sigma_reveal_and_key_ownership_golden_and_negative_cases(sctx377, pkg):
    """compute_reveal_token / verify_reveal / prove_key_ownership / verify_key_ownership over BLS12-377 (reference
    mod.rs:132-165, 301-354; negative cases reveal.rs:73-82, tests.rs:72-77), and the G1 membership test the
    verifiers apply to what they are handed."""
    h = bytes.fromhex
    cat = lambda key, rows: b"".join(h(r[key]) for r in rows)
    le = lambda key, rows: b"".join(int(r[key], 16).to_bytes(32, "little") for r in rows)
    V, K = SIG["reveal"], SIG["key_ownership"]
    for fx in V:
        sk, om = int(fx["sk"], 16).to_bytes(32, "little"), int(fx["omega"], 16).to_bytes(32, "little")
        tok, pf = sctx377.reveal_batch(sk, h(fx["pk"]), h(fx["masked"]), om)
        assert tok == h(fx["token"]) and pf == h(fx["proof"])
        assert sctx377.verify_reveal_batch(h(fx["pk"]), tok, h(fx["masked"]), pf) == [0]
        assert sctx377.verify_reveal_batch(h(fx["pk"]), h(V[0]["pk"]), h(fx["masked"]), pf) == [5]
    infos = [h(r["info"]) for r in K]
    kp = sctx377.key_ownership_prove_batch(cat("pk", K), le("sk", K), infos, le("omega", K))
    assert kp == cat("proof", K)
    assert sctx377.key_ownership_verify_batch(cat("pk", K), infos, kp) == [0] * len(K)
    assert sctx377.key_ownership_verify_batch(cat("pk", K), infos[::-1], kp) == [6, 0, 6]
    # what ark-serialize would refuse at deserialisation fails per ITEM (status 7, "malformed"), not per call: a token /
    # a Schnorr commitment on the curve but outside G1, a token off the curve; a bad KEY (call-level) fails the call
    _, T = _torsion_point()
    toks = b"".join(h(fx["token"]) for fx in V[:3])
    maskeds, pfs = b"".join(h(fx["masked"]) for fx in V[:3]), b"".join(h(fx["proof"]) for fx in V[:3])
    pk0 = h(V[0]["pk"])
    want = [0 if fx["pk"] == V[0]["pk"] else 5 for fx in V[:3]]
    assert sctx377.verify_reveal_batch(pk0, toks, maskeds, pfs) == want
    tok_bad = toks[:96] + pb(bls.add(bls.point_from_bytes(toks[96:192]), T)) + toks[192:]
    assert sctx377.verify_reveal_batch(pk0, tok_bad, maskeds, pfs) == [want[0], 7, want[2]]
    off_curve = bytearray(toks); off_curve[2 * 96 + 50] ^= 1
    assert sctx377.verify_reveal_batch(pk0, bytes(off_curve), maskeds, pfs) == [want[0], want[1], 7]
    with pytest.raises(pkg.MpError) as e:
        sctx377.verify_reveal_batch(pb(bls.add(bls.point_from_bytes(pk0), T)), toks, maskeds, pfs)
    assert e.value.code == -6
    kp_bad = pb(T) + kp[96:]        # commitment of the first Schnorr proof replaced by a torsion point
    assert sctx377.key_ownership_verify_batch(cat("pk", K), infos, kp_bad) == [7, 0, 0]


def test_sigma_batch_round_trip(sctx377):
    """A batch of 300 cards (the reference benchmark's deck size) masked, remasked and revealed; every proof verifies,
    one tampered item per call is the only one refused."""
    import numpy as np
    rng = np.random.default_rng(377)
    def scal(k):
        a = rng.integers(0, 256, size=(k, 32), dtype=np.uint8); a[:, 31] &= 0x0f
        return a.tobytes()
    n = 300
    g = bytes.fromhex(SIG["g"])
    sk = scal(1)
    pk = sctx377.dbg_scalar_mul(g, sk)
    cards = sctx377.dbg_scalar_mul(g * n, scal(n))
    masked, p1 = sctx377.mask_batch(pk, cards, scal(n), scal(n))
    assert sctx377.verify_mask_batch(pk, cards, masked, p1) == [0] * n
    out, p2 = sctx377.remask_prove_batch(pk, masked, scal(n), scal(n))
    assert sctx377.verify_remask_batch(pk, masked, out, p2) == [0] * n
    tok, p3 = sctx377.reveal_batch(sk, pk, out, scal(n))
    assert sctx377.verify_reveal_batch(pk, tok, out, p3) == [0] * n
    bad = bytearray(p3); bad[224 * 7 + 200] ^= 1
    st = sctx377.verify_reveal_batch(pk, tok, out, bytes(bad))
    assert st[7] == 5 and st.count(0) == n - 1


WIRE = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bls12_377_wire_vectors.json")))


def test_wire_decompress_golden_deck_and_rejections(ctx377, pkg):
    """Deserialising half of the wire format over BLS12-377 (`mp377_points_decompress`, `mp377_deck_deserialize`;
    ark-serialize's CanonicalDeserialize behind every type of reference src/lib.rs:45-71): square roots in F_q and the
    G1 membership test on the GPU, against the oracle's golden vectors, with the status code of every rejected
    encoding (off-curve abscissas, x >= q, stray bits, bad infinity encodings, a curve point outside G1)."""
    from mental_poker_b200 import bls12_377 as b377
    h = bytes.fromhex
    pts = b"".join(h(fx["point"]) for fx in WIRE["points"])
    comp = b"".join(h(fx["compressed"]) for fx in WIRE["points"])
    assert b377.points_compress(pts) == comp
    assert ctx377.points_decompress(comp) == pts
    assert ctx377.launches > 0
    assert ctx377.deck_deserialize(h(WIRE["deck_serialized"])) == h(WIRE["deck"])
    k = len(WIRE["points"])
    curve_only = [i for i, s in enumerate(WIRE["rejected_statuses"]) if s != 3]
    bad = comp + b"".join(h(WIRE["rejected"][i]) for i in curve_only)
    out, st, rc = ctx377.points_decompress(bad, want_statuses=True)
    assert rc == -3 and st[:k] == [0] * k and st[k:] == [WIRE["rejected_statuses"][i] for i in curve_only]
    assert out[:len(pts)] == pts and out[len(pts):] == bytes(96 * len(curve_only))
    torsion = [h(WIRE["rejected"][i]) for i, s in enumerate(WIRE["rejected_statuses"]) if s == 3]
    out, st, rc = ctx377.points_decompress(comp + torsion[0], want_statuses=True)
    assert rc == -6 and st == [0] * k + [3] and out == pts + bytes(96)
    with pytest.raises(pkg.MpError):
        ctx377.points_decompress(bad)
    with pytest.raises(pkg.MpError):                       # length prefix says 9 cards, buffer holds 8
        ctx377.deck_deserialize((9).to_bytes(8, "little") + h(WIRE["deck_serialized"])[8:])
    assert ctx377.points_decompress(b"") == b"" and ctx377.deck_deserialize(bytes(8)) == b""


def test_wire_round_trip_of_proof_and_decks_still_verifies(ctx377):
    """serialize -> deserialize of the golden shuffle proof and both decks gives back the same bytes, which verify;
    a non-canonical scalar in the serialised proof is refused at deserialisation; 4 096 random points round-trip."""
    from mental_poker_b200 import bls12_377 as b377
    import numpy as np
    fx = SHUF["shuffle"][0]
    r = _raw(fx)
    m, n = fx["m"], fx["n"]
    ser = b377.proof_serialize(m, n, r["proof"])
    assert len(ser) == (11 * m + 8) * 48 + (5 * n + 9) * 32
    proof = ctx377.proof_deserialize(m, n, ser)
    assert proof == r["proof"]
    deck = ctx377.deck_deserialize(b377.deck_serialize(r["deck"]))
    deck2 = ctx377.deck_deserialize(b377.deck_serialize(r["deck2"]))
    assert deck == r["deck"] and deck2 == r["deck2"]
    assert ctx377.verify_shuffle(m, n, r["enc_g"], r["ck_g"], r["ck_h"], r["ghat"], r["pk"], deck, deck2, proof) == 0
    off = (5 * m + 4) * 48                                    # first scalar of the serialised proof
    s = int.from_bytes(ser[off:off + 32], "little")
    big = ser[:off] + (s + bls.N).to_bytes(32, "little") + ser[off + 32:]
    with pytest.raises(Exception) as e:
        ctx377.proof_deserialize(m, n, big)
    assert e.value.code == -5
    rng = np.random.default_rng(48)
    k = 4096
    a = rng.integers(0, 256, size=(k, 32), dtype=np.uint8); a[:, 31] &= 0x0f
    pts = ctx377.dbg_scalar_mul(pb(bls.G) * k, a.tobytes())
    assert ctx377.points_decompress(b377.points_compress(pts)) == pts


def test_prover_usage_errors(ctx377, pkg):
    fresh = pkg.bls12_377.Context(0)
    buf = bytes(96)
    assert pkg.lib.mp377_shuffle_and_remask(fresh.h, buf, buf, None, buf, buf, None, None) == -4   # MP_ERR_NO_PARAMS
    fx = SHUF["shuffle"][0]
    r = _raw(fx)
    bad = bytearray(r["ck_h"]); bad[5] ^= 1
    with pytest.raises(pkg.MpError) as e:
        fresh.set_params(fx["m"], fx["n"], r["enc_g"], r["ck_g"], bytes(bad), r["ghat"])
    assert e.value.code == -3
    fresh.set_params(fx["m"], fx["n"], r["enc_g"], r["ck_g"], r["ck_h"], r["ghat"])
    perm = list(fx["perm"]); perm[0] = 99
    with pytest.raises(pkg.MpError):
        fresh.shuffle_and_remask(r["pk"], r["deck"], perm, r["rho"], r["rand"])
    # decks above the lockstep prover's limit (8 192 cards) are refused on this curve, not mis-proved
    m, n = 64, 256
    pts = fresh.dbg_scalar_mul(pb(bls.G) * (n + 3), b"".join(b32(rnd.randrange(1, R)) for _ in range(n + 3)))
    fresh.set_params(m, n, pb(bls.G), pts[:96 * n], pts[96 * n:96 * (n + 1)], pts[96 * (n + 1):96 * (n + 2)])
    N = m * n
    with pytest.raises(pkg.MpError) as e:
        fresh.shuffle_and_remask(pts[96 * (n + 2):], pts[:96] * (2 * N), list(range(N)), bytes(32 * N), bytes(32 * (11 * m + 5 * n)))
    assert e.value.code == -1 and b"not supported" in pkg.lib.mp377_last_error_string(fresh.h)
    fresh.close()
