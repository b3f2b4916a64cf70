"""Integer-pipe microbenchmarks (roofline denominators measured on the box)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
ctx = pkg.Context(0)
names = {0: "imad_wide", 1: "imad_lo", 2: "fq_mul", 3: "madd", 4: "fq_sqr", 5: "imad_wide_carry_chain", 6: "imad_wide_plus_iadd_1to1",
         7: "dfma_rz", 8: "dfma_plus_imad_wide_1to1", 9: "dfma_plus_iadd_1to1"}
res = {}
for which, iters in [(0, 2000), (1, 2000), (2, 2000), (3, 500), (4, 2000), (5, 1000), (6, 2000), (7, 1000), (8, 1000), (9, 1000)]:
    best = 0
    for rep in range(3):
        ms, ops = ctx.dbg_bench(which, iters)
        best = max(best, ops / ms / 1e6)
    res[names[which]] = best
    print(f"{names[which]:28s} {best:10.2f} G ops/s", flush=True)
res["units"] = "G ops/s over the whole chip (imad* = instructions; dfma* = DFMA instructions; fq_mul / fq_sqr = field ops; madd = XYZZ mixed additions)"
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "microbench.json"), "w"), indent=1)
