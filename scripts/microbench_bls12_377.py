"""Integer-pipe microbenchmarks of the 12-limb BLS12-377 field (fq_mul, XYZZ mixed addition)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
ctx = pkg.bls12_377.Context(0)
res = {}
for which, name, iters in [(0, "fq_mul", 1000), (1, "madd", 300)]:
    best = 0
    for rep in range(3):
        ms, ops = ctx.dbg_bench(which, iters)
        best = max(best, ops / ms / 1e6)
    res[name] = best
    print(f"bls12_377 {name:8s} {best:10.2f} G ops/s", flush=True)
res["units"] = "G ops/s over the whole chip (fq_mul = 12-limb Montgomery products; madd = XYZZ mixed additions)"
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "microbench_bls12_377.json"), "w"), indent=1)
