"""GPU parity, primitives: device field / group arithmetic and the Pippenger MSM, called
through the C ABI, bit-exact against the oracle."""
import json
import os
import random

import pytest

from oracle.py import stark
from oracle.py.transcript import SeededStream
from _util import chain_points, scalars, b32, pb

pytestmark = pytest.mark.gpu
P, N = stark.P, stark.N
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.json")))
rnd = random.Random(5)
PTS = [stark.mul(stark.G, rnd.randrange(1, N)) for _ in range(16)]


def test_fq_mul(ctx):
    Rinv = pow(1 << 256, -1, P)
    n = 4096
    a = [rnd.randrange(0, 5 * P) for _ in range(n)]
    b = [rnd.randrange(0, 6 * P) for _ in range(n)]
    a[:6] = [0, 1, P - 1, P, 2 * P, 5 * P - 1]
    b[:6] = [0, 1, P - 1, P, 6 * P - 1, 2 * P]
    out = ctx.dbg_fq_mul(b"".join(map(b32, a)), b"".join(map(b32, b)))
    for i in range(n):
        r = int.from_bytes(out[32 * i:32 * i + 32], "little")
        assert r < 2 * P and r % P == a[i] * b[i] * Rinv % P, i


def test_point_add_complete(ctx):
    ps, qs, want = [], [], []
    for a in PTS[:8]:
        for b in PTS[8:]:
            ps.append(a); qs.append(b); want.append(stark.add(a, b))
        for b in (a, stark.neg(a), None):
            ps.append(a); qs.append(b); want.append(stark.add(a, b))
        ps.append(None); qs.append(a); want.append(a)
    out = ctx.dbg_point_add(b"".join(map(pb, ps)), b"".join(map(pb, qs)))
    for i, w in enumerate(want):
        assert out[64 * i:64 * i + 64] == pb(w), i


def test_scalar_mul(ctx):
    ks = [0, 1, 2, N - 1, N, N + 1] + [rnd.randrange(0, 1 << 256) for _ in range(10)]
    ps = [PTS[i % 16] for i in range(len(ks))]
    out = ctx.dbg_scalar_mul(b"".join(map(pb, ps)), b"".join(map(b32, ks)))
    for i, (p, k) in enumerate(zip(ps, ks)):
        assert out[64 * i:64 * i + 64] == pb(stark.mul(p, k)), i


def msm_case(ctx, n, c, seed=1, kind="uniform"):
    s0, s1, pts, st = chain_points(n, seed)
    ks = scalars(st, n, kind)
    e = sum(k * (s0 + i * s1) for i, k in enumerate(ks)) % N
    got = ctx.msm_g1(b"".join(map(pb, pts)), b"".join(map(b32, ks)), c)
    assert got == pb(stark.mul(stark.G, e)), (n, c, kind)


@pytest.mark.parametrize("n,c", [(1, 4), (2, 4), (7, 4), (33, 5), (100, 6), (300, 0), (300, 8), (1000, 9), (1000, 0)])
def test_msm_small(ctx, n, c):
    msm_case(ctx, n, c)


@pytest.mark.parametrize("kind", ["zero", "max", "small", "same"])
def test_msm_scalar_edges(ctx, kind):
    msm_case(ctx, 200, 6, kind=kind)
    msm_case(ctx, 200, 0, kind=kind)


def test_msm_empty(ctx):
    assert ctx.msm_g1(b"", b"", 0) == bytes(64)


def test_msm_golden(ctx):
    for fx in GOLD["msm"]:
        got = ctx.msm_g1(bytes.fromhex(fx["points"]), bytes.fromhex(fx["scalars"]), 0)
        assert got.hex() == fx["result"]


def test_msm_point_edges(ctx):
    # all points equal (doubling branch), P/-P alternating (cancellation), identity inputs
    st = SeededStream(3)
    Pt = PTS[0]
    n = 150
    ks = [st.scalar() for _ in range(n)]
    for name, pts in [("equal", [Pt] * n), ("pm", [Pt if i % 2 == 0 else stark.neg(Pt) for i in range(n)]),
                      ("ident", [None if i % 3 == 0 else PTS[i % 16] for i in range(n)])]:
        want = pb(stark.msm(pts, ks))
        for c in (4, 7, 0):
            assert ctx.msm_g1(b"".join(map(pb, pts)), b"".join(map(b32, ks)), c) == want, (name, c)
    # one heavy bucket spanning many accumulate chunks
    assert ctx.msm_g1(pb(Pt) * 500, b32(12345) * 500, 8) == pb(stark.mul(Pt, 12345 * 500))


def test_ct_msm(ctx):
    s0, s1, pts, st = chain_points(400, 9)
    n = 200
    ks = [st.scalar() for _ in range(n)]
    deck = b"".join(pb(pts[2 * i]) + pb(pts[2 * i + 1]) for i in range(n))
    e1 = sum(k * (s0 + (2 * i) * s1) for i, k in enumerate(ks)) % N
    e2 = sum(k * (s0 + (2 * i + 1) * s1) for i, k in enumerate(ks)) % N
    for c in (5, 0):
        got = ctx.ct_msm(deck, b"".join(map(b32, ks)), c)
        assert got == pb(stark.mul(stark.G, e1)) + pb(stark.mul(stark.G, e2)), c


@pytest.mark.parametrize("n,c", [(4096, 0), (4096, 12), (65536, 0), (65536, 16)])
def test_msm_mid(ctx, n, c):
    msm_case(ctx, n, c, seed=2)


def test_rejects_point_off_curve(ctx, pkg):
    with pytest.raises(pkg.MpError) as e:
        ctx.msm_g1(b32(5) + b32(7), b32(3), 4)
    assert e.value.code == -3


def test_kernels_were_launched(ctx):
    msm_case(ctx, 64, 0)
    assert ctx.launches > 0


@pytest.mark.parametrize("c,world", [(8, 1), (8, 3), (13, 2), (16, 8), (16, 5)])
def test_msm_window_range_split(ctx, pkg, c, world):
    """Window-range split (multi-GPU path) emulated on one GPU: per-'rank' partials + fold."""
    import torch
    n = 500
    s0, s1, pts, st = chain_points(n, 21)
    ks = scalars(st, n, "uniform")
    ks[0], ks[1] = 0, N - 1
    dev = torch.device("cuda:0")
    d_pts = torch.frombuffer(bytearray(b"".join(map(pb, pts))), dtype=torch.uint8).to(dev)
    d_sc = torch.frombuffer(bytearray(b"".join(map(b32, ks))), dtype=torch.uint8).to(dev)
    d_out = torch.zeros(64, dtype=torch.uint8, device=dev)
    W = pkg.lib.mp_msm_num_windows(c)
    points, scs = b"", b""
    for r, s in pkg.dist.fold_scalars(c, W, world):
        b, e = pkg.dist.window_range(W, r, world)
        ctx.msm_g1_windows_device(d_pts.data_ptr(), d_sc.data_ptr(), n, d_out.data_ptr(), c, b, e - b)
        ctx.sync()
        points += bytes(d_out.cpu().numpy().tobytes())
        scs += s
    got = ctx.msm_g1(points, scs, 0)
    e = sum(k * (s0 + i * s1) for i, k in enumerate(ks)) % N
    assert got == pb(stark.mul(stark.G, e))


def test_msm_jobs_diagonal_products(ctx, pkg):
    """K2 / `mp_msm_batch_shared_bases` of SURVEY.md section 8(b) on the Stark curve: batched MSM jobs over shared
    arrays in the shape of the multi-exponentiation argument's diagonal products -- every (deck row i, scalar row j)
    pair of an (m = 4, n = 13) instance, the reference's own test shape -- as one call with 2 components, against
    known discrete logs and the C oracle's ciphertext MSM; plus ragged / empty / overlapping jobs."""
    from oracle import c_oracle
    m, n = 4, 13
    s0, s1, pts, st = chain_points(2 * m * n, 17)
    logs = [(s0 + i * s1) % N for i in range(2 * m * n)]
    ks = [st.scalar() for _ in range((m + 1) * n)]
    deck = b"".join(map(pb, pts))
    kb = b"".join(map(b32, ks))
    jobs = [(j * n, i * n, n) for i in range(m) for j in range(m + 1)]
    out = ctx.msm_jobs(deck, kb, jobs, ncomp=2)
    co = c_oracle.COracle()
    for q, (so, po, ln) in enumerate(jobs):
        for comp in range(2):
            e = sum(ks[so + t] * logs[2 * (po + t) + comp] for t in range(ln)) % N
            assert out[128 * q + 64 * comp:128 * q + 64 * comp + 64] == pb(stark.mul(stark.G, e)), (q, comp)
        if q % 7 == 0:
            assert out[128 * q:128 * q + 128] == co.msm(deck[128 * po:128 * (po + ln)], kb[32 * so:32 * (so + ln)], 2, 0)
    assert ctx.launches > 0
    jobs = [(0, 0, 1), (3, 7, 0), (5, 2, 40), (0, 0, 60), (10, 50, 50)]
    out = ctx.msm_jobs(deck, kb, jobs, ncomp=1, window_bits=7)
    for q, (so, po, ln) in enumerate(jobs):
        e = sum(ks[so + t] * logs[po + t] for t in range(ln)) % N
        assert out[64 * q:64 * q + 64] == pb(stark.mul(stark.G, e)), q
    with pytest.raises(pkg.MpError):
        ctx.msm_jobs(deck, kb, [(0, 90, 20)], ncomp=1)   # reaches past the points
    assert ctx.msm_jobs(deck, kb, [], ncomp=2) == b""
