import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
pkg = g.load_package()
from _util import instance, pb, b32
m, n, seed = int(sys.argv[1]), int(sys.argv[2]), 6
pp, pk, deck, perm, rho, rnd = instance(m, n, seed)
ctx = pkg.Context(0)
enc_g, ck_g, ck_h, ghat = pb(pp.enc_g), b"".join(map(pb, pp.ck_g)), pb(pp.ck_h), pb(pp.ghat)
deck_b = b"".join(pb(c[0]) + pb(c[1]) for c in deck)
ctx.set_params(m, n, enc_g, ck_g, ck_h, ghat)
print("params ok", flush=True)
deck2, proof = ctx.shuffle_and_remask(pb(pk), deck_b, perm, b"".join(map(b32, rho)), b"".join(map(b32, rnd)))
print("prove ok", flush=True)
print("verify", ctx.verify_shuffle(pb(pk), deck_b, deck2, proof))
