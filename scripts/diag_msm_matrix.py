"""Diagnostic: MSM parity over a matrix of (curve, n, window, scalar kind); prints one line per failing case."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
pkg = g.load_package()
from oracle.py import stark, bls12_377 as bls
import _util as us, _util_bls12_377 as ub
curves = [("stark", pkg.Context(0), stark, us.chain_points, us.scalars, us.pb, us.b32),
          ("bls377", pkg.bls12_377.Context(0), bls, ub.chain_points, ub.scalars, ub.pb, ub.b32)]
bad = 0
for name, ctx, cv, chain, scal, pb, b32 in curves:
    for n in (1, 3, 200):
        s0, s1, pts, st = chain(n, 1)
        for kind in ("uniform", "max", "same", "small"):
            ks = scal(st, n, kind)
            e = sum(k * (s0 + i * s1) for i, k in enumerate(ks)) % cv.N
            want = pb(cv.mul(cv.G, e))
            for c in (4, 6, 7, 8, 9, 10, 11, 12, 13, 16):
                got = ctx.msm_g1(b"".join(map(pb, pts)), b"".join(map(b32, ks)), c)
                if got != want:
                    bad += 1
                    print("FAIL", name, n, kind, c, flush=True)
print("failures:", bad)
