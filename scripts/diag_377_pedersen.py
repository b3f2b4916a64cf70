import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
pkg = g.load_package()
from oracle.py import bls12_377 as bls
import _util_bls12_377 as ub
ctx = pkg.bls12_377.Context(0)
pb, b32 = ub.pb, ub.b32
s0, s1, ck, st = ub.chain_points(7, 5)
logs = [(s0 + i * s1) % bls.N for i in range(7)]
ctx.set_commit_key(b"".join(map(pb, ck)))
L = 6
def commit(blind, vals):
    return ctx.pedersen_commit_batch(b"".join(map(b32, vals)), b32(blind), L)
def want(blind, vals):
    e = (blind * logs[0] + sum(v * l for v, l in zip(vals, logs[1:]))) % bls.N
    return pb(bls.mul(bls.G, e))
cases = [("h", 1, [0] * 6), ("g1", 0, [1, 0, 0, 0, 0, 0]), ("g6", 0, [0, 0, 0, 0, 0, 1]), ("2h", 2, [0] * 6),
         ("2^4 h", 16, [0] * 6), ("2^8 h", 256, [0] * 6), ("2^16 h", 1 << 16, [0] * 6), ("2^100 h", 1 << 100, [0] * 6),
         ("2^250 h", 1 << 250, [0] * 6), ("h+g1", 1, [1, 0, 0, 0, 0, 0]), ("3h", 3, [0] * 6), ("rand", st.scalar(), [st.scalar() for _ in range(6)])]
for name, b, v in cases:
    got = commit(b, v)
    print(name, "ok" if got == want(b, v) else "FAIL", "window", ctx.last_msm_window if hasattr(ctx, "last_msm_window") else "", flush=True)
# the same linear combination through the variable-base MSM
b, v = cases[-1][1], cases[-1][2]
print("varbase", ctx.msm_g1(b"".join(map(pb, ck)), b"".join(map(b32, [b] + v)), 0) == want(b, v))
