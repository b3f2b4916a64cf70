import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
pkg = g.load_package()
from oracle.py import bls12_377 as bls
import _util_bls12_377 as ub
ctx = pkg.bls12_377.Context(0)
for n in (1, 3):
    s0, s1, pts, st = ub.chain_points(n, 1)
    ks = ub.scalars(st, n, "small")
    e = sum(k * (s0 + i * s1) for i, k in enumerate(ks)) % bls.N
    want = ub.pb(bls.mul(bls.G, e))
    for c in (10, 11, 12):
        got = ctx.msm_g1(b"".join(map(ub.pb, pts)), b"".join(map(ub.b32, ks)), c)
        print(n, c, "ok" if got == want else "FAIL", bls.is_on_curve(bls.point_from_bytes(got)) if got != bytes(96) else "identity", flush=True)
