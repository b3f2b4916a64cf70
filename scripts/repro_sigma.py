import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
pkg = g.load_package()
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "sigma_vectors.json")))
SHUF = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_vectors.json")))["shuffle"][0]
h = bytes.fromhex
cat = lambda key, rows: b"".join(h(r[key]) for r in rows)
le = lambda key, rows: b"".join(int(r[key], 16).to_bytes(32, "little") for r in rows)
ctx = pkg.Context(0)
ctx.set_params(SHUF["m"], SHUF["n"], h(GOLD["g"]), h(SHUF["ck_g"]), h(SHUF["ck_h"]), h(SHUF["ghat"]))
shared = h(GOLD["shared_key"]); M = GOLD["mask"]; R = GOLD["remask"]; V = GOLD["reveal"]; K = GOLD["key_ownership"]
masked, proofs = ctx.mask_batch(shared, cat("card", M), le("r", M), le("omega", M))
print("mask", masked == cat("masked", M), proofs == cat("proof", M), ctx.verify_mask_batch(shared, cat("card", M), masked, proofs), flush=True)
out, rp = ctx.remask_prove_batch(shared, cat("original", R), le("alpha", R), le("omega", R))
print("remask", out == cat("remasked", R), rp == cat("proof", R), ctx.verify_remask_batch(shared, cat("original", R), out, rp), flush=True)
fx = V[0]
tok, pf = ctx.reveal_batch(int(fx["sk"], 16).to_bytes(32, "little"), h(fx["pk"]), h(fx["masked"]), int(fx["omega"], 16).to_bytes(32, "little"))
print("reveal", tok == h(fx["token"]), pf == h(fx["proof"]), ctx.verify_reveal_batch(h(fx["pk"]), tok, h(fx["masked"]), pf), flush=True)
infos = [h(r["info"]) for r in K]
kp = ctx.key_ownership_prove_batch(cat("pk", K), le("sk", K), infos, le("omega", K))
print("schnorr", kp == cat("proof", K), ctx.key_ownership_verify_batch(cat("pk", K), infos, kp), flush=True)
