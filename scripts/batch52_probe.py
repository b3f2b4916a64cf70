import sys, os, torch
sys.path.insert(0, ".")
import __graft_entry__ as g, bench
pkg = g.load_package(); ctx = pkg.Context(0); dev = torch.device("cuda:0")
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
for ht in (16, 4, 2):
    r = bench.batch52_bench(pkg, ctx, torch, stream, 512, 0, ht)
    print("host_threads", ht, {k: round(v) for k, v in r.items() if k.endswith("_per_s")}, flush=True)
