"""Times the six batched sigma-protocol calls (csrc/sigma.cu, one k_lincomb launch each) at a deck's worth of items.
    gpurun -- 'for b in 0 1 2; do MP_LINCOMB_BLOCKS=$b python scripts/sigma_probe.py; done'
Under ncu:  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:k_lincomb python scripts/sigma_probe.py 1"""
import os, sys
sys.path.insert(0, ".")
import __graft_entry__ as g, bench
pkg = g.load_package(); ctx = pkg.Context(0)
inst = bench.make_instance(ctx, 4, 13, 1)
ctx.set_params(4, 13, inst["enc_g"], inst["ck_g"], inst["ck_h"], inst["ghat"])
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(reps):
    r = bench.sigma_bench(pkg, ctx, 65536, False)
print("MP_LINCOMB_BLOCKS", os.environ.get("MP_LINCOMB_BLOCKS"), {k: round(v) for k, v in r.items() if k.endswith("_per_s")}, flush=True)
