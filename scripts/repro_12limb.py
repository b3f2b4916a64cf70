"""Repro of the two msm.cu kernels that misbehaved on the 12-limb (BLS12-377) build in round 1.

    python scripts/repro_12limb.py win     # variable-base MSM at c = 10..16 (per-window combine levels)
    MP_TABLE_TRICK={0,1,2,3} python scripts/repro_12limb.py table   # fixed-base tables: _each / k_table_normalise<0..2>

Every case has a known answer (chain points with known discrete logs), so no oracle library is needed on the
GPU box.  Prints one line per case and a final PASS/FAIL.  Driven by scripts/sanitize.sh."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
from oracle.py import bls12_377 as bls  # noqa: E402  (big-int group law, used only to state the expected answer)
import _util_bls12_377 as ub  # noqa: E402

which = sys.argv[1]
ctx = pkg.bls12_377.Context(0)
bad = 0
if which == "win":
    for n in (3, 200):
        s0, s1, pts, st = ub.chain_points(n, 1)
        for kind in ("small", "uniform"):
            ks = ub.scalars(st, n, kind)
            e = sum(k * (s0 + i * s1) for i, k in enumerate(ks)) % bls.N
            want = ub.pb(bls.mul(bls.G, e))
            for c in (10, 11, 12, 14, 16):
                got = ctx.msm_g1(b"".join(map(ub.pb, pts)), b"".join(map(ub.b32, ks)), c)
                ok = got == want
                bad += not ok
                print(f"win n={n} {kind} c={c}: {'ok' if ok else 'FAIL'}", flush=True)
else:
    s0, s1, ck, st = ub.chain_points(70, 5)
    logs = [(s0 + i * s1) % bls.N for i in range(70)]
    for L in (6, 69):
        ctx.set_commit_key(b"".join(map(ub.pb, ck[:L + 1])))
        for name, blind, vals in [("h", 1, [0] * L), ("g1", 0, [1] + [0] * (L - 1)), ("2^100 h", 1 << 100, [0] * L),
                                  ("rand", st.scalar(), [st.scalar() for _ in range(L)])]:
            got = ctx.pedersen_commit_batch(b"".join(map(ub.b32, vals)), ub.b32(blind), L)
            e = (blind * logs[0] + sum(v * l for v, l in zip(vals, logs[1:]))) % bls.N
            ok = got == ub.pb(bls.mul(bls.G, e))
            bad += not ok
            print(f"table L={L} {name}: {'ok' if ok else 'FAIL'}", flush=True)
print("PASS" if bad == 0 else f"FAIL ({bad} cases)")
ctx.close()
