"""First GPU bring-up: primitives vs oracle, MSM vs oracle at several sizes, microbenches.
Run on the GPU box:  python scripts/gpu_check1.py"""
import os, sys, time, random, json, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
from oracle.py import stark
from oracle.py.transcript import SeededStream

P, N = stark.P, stark.N
b32 = stark.fe_to_bytes
pb = stark.point_to_bytes64
ctx = pkg.Context(0)
rnd = random.Random(5)
results = {}

def section(name, fn):
    t0 = time.time()
    try:
        fn()
        results[name] = "ok"
        print(f"[ok] {name} ({time.time()-t0:.1f}s)", flush=True)
    except Exception as e:
        results[name] = f"FAIL {e!r}"
        print(f"[FAIL] {name}: {e!r}", flush=True)
        traceback.print_exc()

def t_fq_mul():
    Rinv = pow(1 << 256, -1, P)
    n = 4096
    a = [rnd.randrange(0, 5 * P) for _ in range(n)]
    b = [rnd.randrange(0, 6 * P) for _ in range(n)]
    a[:6] = [0, 1, P - 1, P, 2 * P, 5 * P - 1]; b[:6] = [0, 1, P - 1, P, 6 * P - 1, 2 * P]
    out = ctx.dbg_fq_mul(b"".join(map(b32, a)), b"".join(map(b32, b)))
    for i in range(n):
        r = int.from_bytes(out[32 * i:32 * i + 32], "little")
        assert r < 2 * P and r % P == a[i] * b[i] * Rinv % P, (i, hex(a[i]), hex(b[i]), hex(r))

pts_small = [stark.mul(stark.G, rnd.randrange(1, N)) for _ in range(16)]

def t_point_add():
    ps, qs, want = [], [], []
    for a in pts_small[:8]:
        for b in pts_small[8:]:
            ps.append(a); qs.append(b); want.append(stark.add(a, b))
        for b in (a, stark.neg(a), None):
            ps.append(a); qs.append(b); want.append(stark.add(a, b))
        ps.append(None); qs.append(a); want.append(a)
    out = ctx.dbg_point_add(b"".join(map(pb, ps)), b"".join(map(pb, qs)))
    for i, w in enumerate(want):
        assert out[64 * i:64 * i + 64] == pb(w), i

def t_scalar_mul():
    ks = [0, 1, 2, N - 1, N, N + 1] + [rnd.randrange(0, 1 << 256) for _ in range(10)]
    ps = [pts_small[i % 16] for i in range(len(ks))]
    out = ctx.dbg_scalar_mul(b"".join(map(pb, ps)), b"".join(map(b32, ks)))
    for i, (p, k) in enumerate(zip(ps, ks)):
        assert out[64 * i:64 * i + 64] == pb(stark.mul(p, k)), i

def chain_points(n, seed):
    st = SeededStream(seed)
    s0, s1 = st.scalar(), st.scalar()
    cur, step = stark.mul(stark.G, s0), stark.mul(stark.G, s1)
    pts = []
    for _ in range(n):
        pts.append(cur)
        cur = stark.add(cur, step)
    return s0, s1, pts, st

def msm_case(n, c, seed=1, scal="uniform"):
    s0, s1, pts, st = chain_points(n, seed)
    if scal == "uniform": ks = [st.scalar() for _ in range(n)]
    elif scal == "zero": ks = [0] * n
    elif scal == "max": ks = [N - 1] * n
    elif scal == "small": ks = [st.below(1 << 16) for _ in range(n)]
    elif scal == "same": ks = [st.scalar()] * n
    e = sum(k * (s0 + i * s1) for i, k in enumerate(ks)) % N
    want = pb(stark.mul(stark.G, e))
    got = ctx.msm_g1(b"".join(map(pb, pts)), b"".join(map(b32, ks)), c)
    assert got == want, (n, c, scal, got.hex()[:32], want.hex()[:32])

def t_msm_small():
    for n, c in [(1, 4), (2, 4), (7, 4), (33, 5), (100, 6), (300, 0), (300, 8), (1000, 9), (1000, 0)]:
        msm_case(n, c)
    for scal in ["zero", "max", "small", "same"]:
        msm_case(200, 6, scal=scal)
        msm_case(200, 0, scal=scal)

def t_msm_dups():
    # all points equal (forces doubling branch), P/-P alternating, identity in inputs
    st = SeededStream(3)
    Pt = pts_small[0]
    n = 150
    ks = [st.scalar() for _ in range(n)]
    for name, pts in [("equal", [Pt] * n), ("pm", [Pt if i % 2 == 0 else stark.neg(Pt) for i in range(n)]),
                      ("ident", [None if i % 3 == 0 else pts_small[i % 16] for i in range(n)])]:
        want = pb(stark.msm(pts, ks))
        for c in (4, 7, 0):
            got = ctx.msm_g1(b"".join(map(pb, pts)), b"".join(map(b32, ks)), c)
            assert got == want, (name, c)
    # same scalar + same point -> heavy single bucket spanning many chunks
    pts = [Pt] * 500; ks = [12345] * 500
    assert ctx.msm_g1(b"".join(map(pb, pts)), b"".join(map(b32, ks)), 8) == pb(stark.mul(Pt, 12345 * 500))

def t_ct_msm():
    s0, s1, pts, st = chain_points(400, 9)
    n = 200
    ks = [st.scalar() for _ in range(n)]
    deck = b"".join(pb(pts[2 * i]) + pb(pts[2 * i + 1]) for i in range(n))
    e1 = sum(k * (s0 + (2 * i) * s1) for i, k in enumerate(ks)) % N
    e2 = sum(k * (s0 + (2 * i + 1) * s1) for i, k in enumerate(ks)) % N
    for c in (5, 0):
        got = ctx.ct_msm(deck, b"".join(map(b32, ks)), c)
        assert got == pb(stark.mul(stark.G, e1)) + pb(stark.mul(stark.G, e2)), c

def t_msm_mid():
    for n, c in [(4096, 0), (4096, 12), (65536, 0), (65536, 16)]:
        t0 = time.time(); msm_case(n, c, seed=2); print(f"   msm n={n} c={c} ok ({time.time()-t0:.1f}s)", flush=True)

def t_not_on_curve():
    bad = b32(5) + b32(7)
    try:
        ctx.msm_g1(bad, b32(3), 4)
    except pkg.MpError as e:
        assert e.code == -3
    else:
        raise AssertionError("off-curve point accepted")

section("fq_mul", t_fq_mul)
section("point_add", t_point_add)
section("scalar_mul", t_scalar_mul)
section("msm_small", t_msm_small)
section("msm_dups", t_msm_dups)
section("ct_msm", t_ct_msm)
section("not_on_curve", t_not_on_curve)
section("msm_mid", t_msm_mid)

# ---- microbenches
def bench():
    import torch
    names = {0: "imad_wide", 1: "imad_lo", 2: "fq_mul", 3: "madd"}
    for which, iters in [(0, 2000), (1, 2000), (2, 2000), (3, 500)]:
        ms, ops = ctx.dbg_bench(which, iters)
        print(f"   bench {names[which]}: {ms:.3f} ms, {ops/ms/1e6:.2f} Gop/s", flush=True)
        results["bench_" + names[which]] = ops / ms / 1e6
    # device-resident MSM timing
    dev = torch.device("cuda:0")
    stream = torch.cuda.ExternalStream(ctx.stream)
    for logn in (16, 18, 20):
        n = 1 << logn
        # points: reuse a small set of valid points tiled (timing only), scalars random
        s0, s1, pts, st = chain_points(1024, 11)
        base = torch.frombuffer(bytearray(b"".join(map(pb, pts))), dtype=torch.uint8).to(dev)
        bases = base.repeat(n // 1024).contiguous()
        scal = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev)
        scal[:, 31] &= 0x07  # < 2^251 < group order
        out = torch.zeros(64, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        for c in ([13, 14, 15, 16] if logn == 20 else [11, 12, 13, 14] if logn == 16 else [13, 14, 15]):
            for rep in range(3):
                with torch.cuda.stream(stream):
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    ctx.msm_g1_device(bases.data_ptr(), scal.data_ptr(), n, out.data_ptr(), c)
                    e1.record(stream)
                ctx.sync(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            adds = ctx.last_msm_ec_adds
            print(f"   msm 2^{logn} c={c}: {ms:.3f} ms  {adds/ms/1e6:.2f} G EC-adds/s (scheduled adds {adds})", flush=True)
            results[f"msm_2^{logn}_c{c}_ms"] = ms
section("bench", bench)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(results, open(os.path.join(ROOT, "gpurun_out", "check1.json"), "w"), indent=1)
print(json.dumps(results))
