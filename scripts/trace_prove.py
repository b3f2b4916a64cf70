"""One 2^16-card shuffle_and_remask with MP_TRACE=1 (host-side timestamps of the prover's sync points)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from oracle.py import stark
from _util import pb
G64 = pb(stark.G)
def rs(rng, k):
    a = rng.integers(0, 256, size=(k, 32), dtype=np.uint8); a[:, 31] &= 7; return a.tobytes()
m, n = int(sys.argv[1]) if len(sys.argv) > 1 else 128, int(sys.argv[2]) if len(sys.argv) > 2 else 512
Nc = m * n
rng = np.random.default_rng(5)
ctx = pkg.Context(0)
npts = (n + 3) + 2 * Nc
pts = ctx.dbg_scalar_mul(G64 * npts, rs(rng, npts))
P = lambda i: pts[64 * i:64 * (i + 1)]
ck_g, ck_h, ghat, pk, deck = pts[:64 * n], P(n), P(n + 1), P(n + 2), pts[64 * (n + 3):]
perm = [int(v) for v in rng.permutation(Nc)]
rho, rand = rs(rng, Nc), rs(rng, 11 * m + 5 * n)
ctx.set_params(m, n, G64, ck_g, ck_h, ghat)
for it in range(4):
    print(f"--- iteration {it}", file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    deck2, proof = ctx.shuffle_and_remask(pk, deck, perm, rho, rand)
    t1 = time.perf_counter()
    ok = ctx.verify_shuffle(pk, deck, deck2, proof)
    t2 = time.perf_counter()
    print(f"prove {1e3 * (t1 - t0):.2f} ms  verify {1e3 * (t2 - t1):.2f} ms  status {ok}", file=sys.stderr, flush=True)
